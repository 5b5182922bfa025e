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


def available():
    return os.path.isdir(REFERENCE_CODE)


def install():
    """Install the shims and return the imported reference modules."""
    if not available():
        raise RuntimeError("reference tree %s is not present" % REFERENCE_CODE)
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
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    import importlib

    mods = {}
    for name in ("models", "utils", "NMS", "trainer"):
        mods[name] = importlib.import_module(name)
    return mods
