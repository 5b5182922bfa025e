"""GPU parity of the self-contained kernels (K1 prep/rotate, K9 argmax, K11 heightmap, K12 NMS)
through the C ABI, against the CPU oracle and the golden fixtures.  Integer / index / float64
work is required to be bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, MEAN, STD
from oracle import heightmap as ohm
from oracle import nms as onms
from oracle import qnet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from smg_b200 import engine
    return engine.get_engine(0, 18, 640, "fp32", owner="kernels")


def test_library_loaded_is_in_tree():
    from smg_b200 import _lib
    lib = _lib.load()
    assert lib.smg_version() >= 100
    assert os.path.samefile(os.path.dirname(_lib.LIB_PATH), os.path.join(os.path.dirname(GOLDEN_DIR), "..", "smg-multimodal-grasping_b200", "csrc"))


def test_prep_matches_oracle(eng, scene_inputs):
    scene, mask, pair, _ = scene_inputs
    hm = torch.from_numpy(np.stack([scene, mask, pair])).cuda()
    out = eng.prep(hm, MEAN, STD).cpu()
    for i, v in enumerate((scene, mask, pair)):
        assert torch.equal(out[i:i + 1], qnet.preprocess(v, MEAN, STD)), "prep differs for map %d" % i


def test_prep_empty_map(eng):
    out = eng.prep(torch.zeros(1, 224, 224, dtype=torch.float64, device="cuda"), MEAN, STD).cpu()
    assert torch.equal(out, qnet.preprocess(np.zeros((224, 224)), MEAN, STD))


@pytest.mark.parametrize("R", [1, 4, 16])
def test_rotate_index_exact(eng, R):
    H = 640
    bad = 0
    for r in range(R):
        got = eng.rotate_index_map(r, R).cpu().numpy()
        ref = qnet.rotate_index_map(H, r, R)
        bad += int((got != ref).sum())
    assert bad == 0, "rotation index mismatches vs torch-CPU arithmetic: %d pixels" % bad


def test_rotate_values_match_torch(eng, scene_inputs):
    x = qnet.preprocess(scene_inputs[0], MEAN, STD)
    rots = list(range(16))
    got = eng.rotate(x[0].cuda(), rots, 16).cpu()
    mism_cpu = 0
    for r in rots:
        ref = qnet.rotate_nearest(x, r, 16)
        mism_cpu += int((got[r:r + 1] != ref).sum())
    assert mism_cpu == 0, "rotated images differ from torch CPU grid_sample in %d values" % mism_cpu
    # informational: torch CUDA (cuBLAS bmm) may break .5 ties differently from torch CPU
    xc = x.cuda()
    mism_gpu = sum(int((got[r:r + 1].cuda() != qnet.rotate_nearest(xc, r, 16)).sum()) for r in rots)
    print("rotate: values differing from torch CUDA grid_sample over 16 rotations: %d" % mism_gpu)


def test_argmax_first_max_wins(eng):
    rs = np.random.RandomState(0)
    for n in (1, 7, 160, 1000):
        q = rs.randn(n).astype(np.float32)
        q[rs.randint(0, n)] = q.max()  # plant a tie
        val, idx = eng.argmax(torch.from_numpy(q))
        assert int(idx.item()) == int(np.argmax(q)) and float(val.item()) == float(q.max())


def test_heightmap_bit_exact(eng, golden):
    import smg_b200.synth as synth
    cam = synth.make_camera(golden["heightmap"]["camera_seed"])
    o224, o448, A = eng.heightmap(torch.from_numpy(cam["depth"]), cam["intrinsics"], cam["pose"])
    z = np.load(os.path.join(GOLDEN_DIR, "heightmap_seed3.npz"))
    d224, d448 = o224.cpu().numpy(), o448.cpu().numpy()
    assert np.array_equal(d224, z["depth224"]), "224 map: %d of 50176 values differ" % int((d224 != z["depth224"]).sum())
    assert np.array_equal(d448[::7], z["depth448_rows"])
    assert np.array_equal(A, z["A_htor"])
    # a second, different camera against the oracle
    cam = synth.make_camera(11, num_objects=4, cluttered=False)
    o224, o448, A = eng.heightmap(torch.from_numpy(cam["depth"]), cam["intrinsics"], cam["pose"])
    r224, r448, rA = ohm.get_heightmap_depth(cam["depth"], cam["intrinsics"], cam["pose"])
    assert np.array_equal(o224.cpu().numpy(), r224) and np.array_equal(o448.cpu().numpy(), r448) and np.array_equal(A, rA)


def test_heightmap_dropin_signature(golden):
    import hashlib
    import smg_b200.synth as synth
    from smg_b200 import utils
    cam = synth.make_camera(golden["heightmap"]["camera_seed"])
    out = utils.get_heightmap(cam["color"], cam["depth"], cam["intrinsics"], cam["pose"], synth.WORKSPACE_LIMITS, 0.002)
    assert len(out) == 5 and out[1].shape == (224, 224) and out[3].shape == (448, 448) and out[1].dtype == np.float64
    # colour maps: uint8, bit-identical to the reference's cv2.warpPerspective output of the same camera image
    assert out[0].shape == (224, 224, 3) and out[2].shape == (448, 448, 3) and out[0].dtype == np.uint8
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(out[0]) == golden["heightmap"]["color224_sha"] and sha(out[2]) == golden["heightmap"]["color448_sha"]
    # and a second image against the oracle restatement, including a saturated / constant one
    for img in (synth.make_camera(11)["color"], np.full((480, 640, 3), 255, np.uint8)):
        got = utils.get_heightmap(img, cam["depth"], cam["intrinsics"], cam["pose"], synth.WORKSPACE_LIMITS, 0.002)
        r224, r448 = ohm.get_heightmap_color(img)
        assert np.array_equal(got[0], r224) and np.array_equal(got[2], r448)


def test_nms_matches_reference_lists(eng, golden):
    import smg_b200.synth as synth
    from smg_b200 import NMS
    known = np.array([[[10, 10], [60, 60]], [[12, 12], [62, 62]], [[100, 100], [160, 150]], [[0, 0], [5, 5]],
                      [[0, 0], [200, 200]]], np.float32)
    for case in golden["nms"]:
        if case["kind"] == "known5":
            boxes, scores = known, np.ones(5)
        else:
            n = case["n"]
            boxes, scores = synth.make_boxes(case["seed"], n) if n else (np.zeros((0, 2, 2), np.float32), np.zeros(0))
        keep = NMS.py_cpu_nms(boxes, scores, 0.40, 224 * 224 / 60, 224 * 224 / 5)
        assert keep == case["keep"], case
    boxes, scores = synth.make_boxes(7, 1000)  # maximum supported size, against the oracle
    assert NMS.py_cpu_nms(boxes, scores, 0.40, 224 * 224 / 60, 224 * 224 / 5) == onms.nms(boxes, scores, 0.40, 224 * 224 / 60, 224 * 224 / 5)


def test_pe_oo_geometry_matches_reference():
    """utils.global_position / get_best_grasp_angle / get_best_suction_angle (csrc/geometry.cu, one device thread in fp64)
    against the outputs of the unmodified reference functions (tests/golden/golden_r02.json)."""
    import json
    import os
    import smg_b200.synth as synth
    from smg_b200 import utils
    with open(os.path.join(GOLDEN_DIR, "golden_r02.json")) as f:
        geo = json.load(f)["geometry"]
    n = 0
    for e in geo:
        cam = synth.make_camera(e["camera_seed"])
        A, K, P, depth = np.asarray(e["A_htor"]), cam["intrinsics"], cam["pose"], cam["depth"]
        box, cter = np.asarray(e["box_mask_cors"]), np.asarray(e["masks_cter"])
        for gp in e["global_position"]:
            assert np.allclose(utils.global_position(np.array(gp["pix"]), A, K, P, depth), gp["xyz"], rtol=0, atol=1e-12)
        for gr in e["grasp"]:
            c, ang, dist = utils.get_best_grasp_angle(gr["is_pe"], box, [gr["obj"], 0], A, K, P, depth)
            assert np.allclose(c, gr["center"], atol=1e-12)
            assert abs(ang - gr["angle"]) <= 1e-9 and abs(dist - gr["open_distance"]) <= 1e-12, gr
        for su in e["suction"]:
            c, ang = utils.get_best_suction_angle(su["is_oo"], e["K"], cter, box, [su["obj"], 0], A, K, P, depth)
            assert np.allclose(c, su["center"], atol=1e-12)
            assert abs(ang - su["angle"]) <= 1e-9, (e["scene_seed"], su, ang)    # whole degrees: any slip would be >= 0.017
            n += 1
    assert n == 40


def test_soft_mask_resize_and_detection_postprocess():
    """N2, weight-free half (code/masks.py:36-83): the 448 -> 224 bilinear(align_corners=True) resize of soft masks against
    torch's own F.interpolate on the CPU (the call the reference makes), and the score filter + box halving + NMS chain
    against a restatement with the oracle's NMS."""
    import torch.nn.functional as F
    import smg_b200.synth as synth
    from smg_b200 import masks as pmasks
    rs = np.random.RandomState(5)
    fp, _ = synth._footprints(np.random.RandomState(11), 6, True)
    soft448 = torch.from_numpy((fp * rs.uniform(0.6, 1.0, size=(6, 1, 1)) + rs.uniform(0, 0.02, size=fp.shape)).astype(np.float32))
    got = pmasks.resize_soft_masks(soft448[:, None])
    ref = F.interpolate(soft448[:, None], size=[224, 224], mode="bilinear", align_corners=True)[:, 0].numpy()
    assert got.shape == (6, 224, 224) and np.abs(got - ref).max() <= 1e-6
    boxes, _ = synth.make_boxes(2, 6)
    b4 = np.concatenate([boxes[:, 0], boxes[:, 1]], axis=1) * 2          # detector boxes live on the 448 grid
    scores = np.array([0.9, 0.8, 0.7, 0.005, 0.3, 0.001], np.float32)     # a failing score in the middle is kept (last passing wins)
    init, soft, hb, keep, number = pmasks.postprocess_detections(soft448[:, None], b4, scores, 0.01)
    want_keep = onms.nms((b4.reshape(-1, 2, 2) / 2)[:5], scores[:5], 0.40, 224 * 224 / 60, 224 * 224 / 5)
    assert keep == want_keep and number == len(want_keep)
    assert np.array_equal(init, (soft448.numpy() > 0.5)[want_keep]) and np.abs(soft - ref[want_keep]).max() <= 1e-6
    assert pmasks.postprocess_detections(soft448[:, None], b4, scores * 0, 0.01)[4] == 0
