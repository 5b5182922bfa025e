"""Repo contract checks that need no GPU: C-ABI symbols, no oracle in the product, loud failure without CUDA."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "smg-multimodal-grasping_b200")


def header_functions():
    src = open(os.path.join(ROOT, "include", "smg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = os.path.join(PKG, "csrc", "libsmg_b200.so")
    assert os.path.exists(lib)
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    exported = set(re.findall(r" T (smg_[a-z0-9_]+)", out))
    declared = header_functions()
    assert len(declared) >= 20
    missing = [f for f in declared if f not in exported]
    assert not missing, "declared in include/smg_b200.h but not exported: %s" % missing


def test_ctypes_binding_covers_header_and_loads():
    from smg_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_functions()
    lib = _lib.load()  # dlopen + symbol lookup only; no compute without a GPU
    assert lib.smg_version() >= 100


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt and f.endswith(".py"):
                    bad.append(f)
    assert not bad, "product files referencing the oracle: %s" % bad


def test_tensor_core_kernels_are_tcgen05():
    lib = os.path.join(PKG, "csrc", "libsmg_b200.so")
    sass = subprocess.check_output(["cuobjdump", "-sass", lib], text=True)
    assert "UTCHMMA" in sass or "UTCQMMA" in sass, "no tcgen05.mma in the SASS"
    assert "LDTM" in sass and "UBLKCP" in sass
    assert "UTMALDG" in sass, "no tensor-map TMA load (cp.async.bulk.tensor) in the SASS"
    assert "STTM" in sass, "no tcgen05.st (operands in tensor memory) in the SASS"
    assert "HMMA." not in sass.replace("UTCHMMA", ""), "legacy mma.sync found"


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smg_b200 import _lib, engine
    with pytest.raises(_lib.SmgError):
        engine.Engine(0, 2, 640)
    import smg_b200.models as models
    net = models.reinforcement_net(True)
    x = torch.zeros(1, 3, 640, 640)
    with pytest.raises(Exception):
        net.forward(x, x, 0, True, -1)


def test_state_dict_surface():
    import copy
    import torch
    import smg_b200.models as models
    torch.manual_seed(0)
    net = models.reinforcement_net(True)
    sd = net.state_dict()
    assert len(sd) == 2217
    for k in ("suction_depth_trunk.features.conv0.weight", "gs_depth_trunk.classifier.weight",
              "suctionnet_val.suction-val-conv1.weight", "gsnet_val.grasp-val-norm0.running_mean"):
        assert k in sd
    assert tuple(sd["graspnet_val.grasp-val-conv1.weight"].shape) == (1, 64, 20, 20)
    assert tuple(models.reactive_net(True).state_dict()["graspnet_val.grasp-val-conv1.weight"].shape) == (3, 64, 20, 20)
    c = copy.deepcopy(net)
    c.load_state_dict(sd)
    assert net.training and c.training and net.gnum_rotations == 1


def test_bn_channel_bookkeeping():
    import smg_b200.models as models
    from smg_b200 import engine
    net = models.reinforcement_net(True)
    mods = net._bn_modules(net.grasp_depth_trunk)
    assert len(mods) == 121 and sum(m.num_features for m in mods) == engine.TRUNK_BN_CHANNELS
    assert len(models._bn_counts(640)) == 121
    assert len(engine.trunk_param_list(net.grasp_depth_trunk)) == 362
    assert [tuple(p.shape) for p in engine.head_param_list(net.graspnet_val)] == \
        [(2048,), (2048,), (64, 2048, 1, 1), (64,), (64,), (1, 64, 20, 20)]


def test_running_stat_ema_weights_match_sequential_batchnorm():
    """_apply_running_stats (closed-form EMA over an ordered list of passes) == torch BatchNorm2d sequentially."""
    import torch
    import smg_b200.models as models
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm2d(4)
    bn.train()
    xs = [torch.randn(1, 4, 6, 6) * (i + 1) + i for i in range(3)]
    order = [0, 2, 1, 2]
    for s in order:
        bn(xs[s])
    mean = torch.stack([x.mean((0, 2, 3)) for x in xs])
    var = torch.stack([x.var((0, 2, 3), unbiased=False) for x in xs])
    k = len(order)
    w = torch.zeros(3, dtype=torch.float64)
    for i, s in enumerate(order):
        w[s] += 0.1 * 0.9 ** (k - 1 - i)
    rm = (w[:, None] * mean.double()).sum(0)
    rv = 0.9 ** k + (w[:, None] * var.double()).sum(0) * (36 / 35.0)
    assert torch.allclose(rm.float(), bn.running_mean, atol=1e-6)
    assert torch.allclose(rv.float(), bn.running_var, atol=1e-5)
    assert int(bn.num_batches_tracked) == k


def test_every_cuda_source_is_built():
    """Every .cu under csrc/ is compiled into the library: the Makefile globs the directory, and each source has a fresh
    object file next to the .so (a kernel file that is not linked would leave its launcher undefined only at load time
    on the GPU box)."""
    import glob
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smg-multimodal-grasping_b200", "csrc")
    mk = open(os.path.join(root, "Makefile")).read()
    assert "$(wildcard *.cu)" in mk and "$(wildcard *.cuh)" in mk, "sources and header prerequisites must be globbed"
    import __graft_entry__ as entry
    entry.build()
    on_disk = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(root, "*.cu")))
    objs = sorted(os.path.basename(p)[:-2] for p in glob.glob(os.path.join(root, "build", "*.o")))
    assert objs == on_disk, (sorted(set(on_disk) - set(objs)), sorted(set(objs) - set(on_disk)))
