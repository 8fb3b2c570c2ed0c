"""CPU: tools.scene_motion_tracking.camera_to_scene_motion (imported by the unchanged scripts/inference_video.py)
against the reference's own function where /root/reference is mounted, and against a golden vector generated from
it (tests/golden/scene_motion.npz, written by this file's `_make_case` + the reference) everywhere."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE, have_reference
from tools.scene_motion_tracking import camera_to_scene_motion, get_K_matrix


def _make_case(seed, T, H, W):
    rng = np.random.default_rng(seed)
    w2cs, c2ws = [], []
    for _ in range(T):
        ang = rng.normal(0, 0.05, 3)
        cx, sx, cy, sy, cz, sz = np.cos(ang[0]), np.sin(ang[0]), np.cos(ang[1]), np.sin(ang[1]), np.cos(ang[2]), np.sin(ang[2])
        R = (np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
             @ np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]))
        M = np.eye(4)
        M[:3, :3] = R
        M[:3, 3] = rng.normal(0, 2.0, 3)
        w2cs.append(M)
        c2ws.append(np.linalg.inv(M))
    K = np.array([500.0 + 20 * rng.random(), 480.0, 3.0, -2.0])
    depth = rng.random((H, W)).astype(np.float16)
    return w2cs, c2ws, K, depth


def _reference_fn():
    spec = importlib.util.spec_from_file_location("ref_scene_motion", os.path.join(REFERENCE, "tools", "scene_motion_tracking.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not have_reference(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("seed,T,H,W,istrain", [(0, 5, 12, 16, True), (1, 3, 9, 7, False), (2, 1, 4, 4, True)])
def test_equals_reference_function(seed, T, H, W, istrain):
    ref = _reference_fn()
    w2cs, c2ws, K, depth = _make_case(seed, T, H, W)
    want = ref.camera_to_scene_motion(w2cs, c2ws, K, depth, W, H, istrain=istrain)
    got = camera_to_scene_motion(w2cs, c2ws, K, depth, W, H, istrain=istrain)
    assert got.shape == want.shape == (T, 2, H, W)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-9)
    np.testing.assert_array_equal(get_K_matrix(K, T), ref.get_K_matrix(K, T))


def test_golden_vector_and_identity_cameras():
    path = os.path.join(GOLDEN, "scene_motion.npz")
    if have_reference() and not os.path.exists(path):                      # (re)generate from the reference
        w2cs, c2ws, K, depth = _make_case(7, 4, 10, 12)
        np.savez_compressed(path, flow=_reference_fn().camera_to_scene_motion(w2cs, c2ws, K, depth, 12, 10))
    z = np.load(path)
    w2cs, c2ws, K, depth = _make_case(7, 4, 10, 12)
    np.testing.assert_allclose(camera_to_scene_motion(w2cs, c2ws, K, depth, 12, 10), z["flow"], rtol=1e-10, atol=1e-9)
    # scripts/inference_video.py:149-189 with `null` cameras: identity poses, zero depth -> no motion at all
    eye = [np.eye(4)] * 3
    flow = camera_to_scene_motion(eye, eye, np.array([1000.0, 1000.0, 0.0, 0.0]), np.zeros((8, 8)), 8, 8)
    assert flow.shape == (3, 2, 8, 8) and not flow.any()
