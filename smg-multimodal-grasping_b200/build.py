"""Build csrc/libsmg_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(CSRC, "libsmg_b200.so")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "smg_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source with `-gencode arch=compute_100a,code=sm_100a -lineinfo`."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libsmg_b200.so")
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4), "NVCC=" + nvcc]
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libsmg_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
