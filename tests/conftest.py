import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
MEAN, STD = 0.01, 0.03  # harness substitution for the reference's 0/0 literals


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def scene_inputs():
    """The seeded synthetic scene used by the golden fixtures: (scene, mask0, pair01) 224x224 float64."""
    import smg_b200.synth as synth
    sc = synth.make_scene(1, num_objects=4, cluttered=False)
    return (sc["scene"], synth.masked_scene(sc["scene"], sc["masks"], [0]),
            synth.masked_scene(sc["scene"], sc["masks"], [0, 1]), sc)


@pytest.fixture(scope="session")
def rl_state_dict():
    """Random-init reinforcement_net weights, seed 0 (same RNG stream as the reference's constructor)."""
    import torch
    import smg_b200.models as models
    torch.manual_seed(0)
    net = models.reinforcement_net(True)
    return net.state_dict()


def check_fingerprint(t, fp, rtol, atol=0.0):
    """Compare a tensor with a make_golden.py fingerprint (shape, mean, absmean, sampled values)."""
    import numpy as np
    a = t.detach().cpu().double().numpy().ravel()
    assert list(t.shape) == fp["shape"]
    scale = max(abs(fp["max"]), abs(fp["min"]), 1e-30)
    val = a[np.asarray(fp["pos"])]
    err = np.abs(val - np.asarray(fp["val"])).max() / scale
    assert err <= rtol + atol, "sampled values differ: rel-to-range err %.3g" % err
    assert abs(a.mean() - fp["mean"]) <= (rtol + atol) * scale
    assert abs(np.abs(a).mean() - fp["absmean"]) <= (rtol + atol) * scale
    return err
