"""Turn `ncu -i X.ncu-rep --page raw --csv` into the small `metric,value,unit` summaries kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/summarize_ncu.py /tmp/raw.csv <kernel-name-substring> [launch-index] [samples] > profiles/rNN_<kernel>_ncu_summary.csv

`samples` (how many samples the captured launch processed) is written as a `samples` row; bench.py uses it for the
algorithmic bytes of the same launch.
"""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tc", "sm__pipe_tensor",
        "sm__inst_executed_pipe_tc", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared",
        "l1tex__throughput.avg.pct", "lts__t_sector_hit_rate", "lts__throughput.avg.pct", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak", "gpu__dram_throughput.avg.pct", "sm__throughput.avg.pct",
        "smsp__average_warp", "launch__grid_size", "launch__block_size", "sm__cycles_active.avg", "dram__throughput.avg.pct")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pat = sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    hits = [r for r in rows[2:] if len(r) == len(hdr) and pat in r[name_col]]
    if not hits:
        sys.exit("no launch of a kernel matching %r" % pat)
    r = hits[min(which, len(hits) - 1)]
    print("metric,value,unit")
    if len(sys.argv) > 4:
        print("samples,%d," % int(sys.argv[4]))
    for k in ("Kernel Name", "Block Size", "Grid Size"):
        if k in hdr:
            print('%s,"%s",' % (k, r[hdr.index(k)]))
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP):
            print("%s,%s,%s" % (h, r[i].replace(",", ""), units[i]))


if __name__ == "__main__":
    main()
