"""Import the UNMODIFIED reference modules from /root/reference/code on CPU.

TEST INFRASTRUCTURE ONLY - used by tests/golden/make_golden.py in the build
container (where /root/reference exists) to generate golden vectors.  Nothing
that runs on the GPU box may import this module (there is no /root/reference
there).

Shims (all process-local, nothing on disk is touched), SURVEY.md Appendix A:
  1. stub `matplotlib`, `matplotlib.pyplot` (imported by code/models.py:11,
     code/trainer.py:12, code/utils.py:10; never called on this path) and
     `apex.amp` (opt_level "O0" == fp32 identity, code/trainer.py:101,350);
  2. `densenet121(pretrained=True)` -> `weights=None` (no network; the
     benchmark configs are random-init anyway), code/models.py:22-24,308-310;
  3. `.cuda()` -> identity so the CUDA-only forward (code/models.py:377-385)
     runs on CPU when the nets are built with use_cuda=True.
"""
import contextlib
import os
import sys
import types

REFERENCE_CODE = "/root/reference/code"
# build-time copy of the five importable reference modules (git-ignored, made by __graft_entry__.build() in the build
# container, travels to the GPU box with the snapshot): what bench.py's reference arms run there
VENDORED_CODE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "code")
VENDORED_FILES = ("models.py", "trainer.py", "utils.py", "NMS.py", "logger.py")


def code_dir():
    """The reference's code directory: the read-only tree in the build container, else the vendored copy."""
    if os.path.isdir(REFERENCE_CODE):
        return REFERENCE_CODE
    if all(os.path.exists(os.path.join(VENDORED_CODE, f)) for f in VENDORED_FILES):
        return VENDORED_CODE
    return None


def available():
    return code_dir() is not None


def vendor():
    """Copy the importable reference modules to baseline/_ref/code (build container only; never committed)."""
    if not os.path.isdir(REFERENCE_CODE):
        return False
    import shutil
    os.makedirs(VENDORED_CODE, exist_ok=True)
    for f in VENDORED_FILES:
        shutil.copyfile(os.path.join(REFERENCE_CODE, f), os.path.join(VENDORED_CODE, f))
    return True


def patched_trainer_module(mean, std):
    """The reference's trainer.py with its two NaN-producing literals (image_mean = image_std = 0,
    code/trainer.py:176-177) replaced, compiled in memory - nothing is written to disk."""
    src = open(os.path.join(code_dir(), "trainer.py")).read()
    assert "image_mean = [0.0, 0.0, 0.0]" in src and "image_std = [0.0, 0.0, 0.0]" in src
    src = src.replace("image_mean = [0.0, 0.0, 0.0]", "image_mean = [%r, %r, %r]" % (mean, mean, mean))
    src = src.replace("image_std = [0.0, 0.0, 0.0]", "image_std = [%r, %r, %r]" % (std, std, std))
    m = types.ModuleType("trainer_patched")
    exec(compile(src, "trainer_patched.py", "exec"), m.__dict__)
    return m


_ORIG_CUDA = {}


def set_cpu_mode(cpu):
    """cpu=True: `.cuda()` becomes the identity (tensors and modules); cpu=False: the original methods are restored.
    Process-global, so callers that mix the CPU arm and GPU work in one process switch it around each use."""
    import torch
    if not _ORIG_CUDA:
        _ORIG_CUDA["tensor"], _ORIG_CUDA["module"] = torch.Tensor.cuda, torch.nn.Module.cuda
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    else:
        torch.Tensor.cuda, torch.nn.Module.cuda = _ORIG_CUDA["tensor"], _ORIG_CUDA["module"]


def install(cpu=True):
    """Install the shims and return the imported reference modules.  cpu=False leaves `.cuda()` alone, so the
    reference runs on the GPU exactly as written (bench.py --impl reference-gpu)."""
    global REFERENCE_CODE
    if not available():
        raise RuntimeError("reference tree %s is not present" % REFERENCE_CODE)
    REFERENCE_CODE = code_dir()
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if "apex" not in sys.modules:
        apex, amp = types.ModuleType("apex"), types.ModuleType("apex.amp")
        amp.initialize = lambda model, opt, opt_level="O0": (model, opt)
        amp.scale_loss = contextlib.contextmanager(lambda loss, opt: (yield loss))
        apex.amp = amp
        sys.modules["apex"], sys.modules["apex.amp"] = apex, amp
    import torch
    import torchvision

    if not getattr(torchvision.models.densenet, "_smg_shimmed", False):
        _dn = torchvision.models.densenet.densenet121
        def _dn_noweights(pretrained=False, **kw):
            kw.pop("weights", None)
            return _dn(weights=None, **kw)

        torchvision.models.densenet.densenet121 = _dn_noweights
        torchvision.models.densenet._smg_shimmed = True
    set_cpu_mode(cpu)

    import importlib

    mods = {}
    for name in ("models", "utils", "NMS", "trainer"):
        mods[name] = importlib.import_module(name)
    return mods
