"""`tools.scene_motion_tracking.camera_to_scene_motion` — imported by the unchanged scripts/inference_video.py
(:26, called at :187-189).  Condition preprocessing, not part of the accelerated path: plain numpy on the host.

Semantics follow /root/reference/tools/scene_motion_tracking.py:14-67 (checked against it on random cameras in
tests/test_scene_motion.py where the reference is mounted, and against a committed golden vector elsewhere):
a centred pixel grid is lifted to 3-D with z = 100 - 50 * depth, moved from camera t to camera t+1
(w2c[t+1] @ c2w[t]), projected with the pinhole intrinsics (fx, fy, cx, cy), and the displacement against the
un-moved projection is the 2-channel "scene motion" of frame t+1; frame 0 has none.  The flow is clipped to
mean +- 3 sigma; if any value is non-finite the whole flow stays zero."""
import numpy as np


def pinhole_matrix(K):
    """3 x 4 projection [[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0]] from K = (fx, fy, cx, cy)."""
    fx, fy, cx, cy = (float(v) for v in K[:4])
    return np.array([[fx, 0.0, cx, 0.0], [0.0, fy, cy, 0.0], [0.0, 0.0, 1.0, 0.0]])


def get_K_matrix(K, T):
    """The reference's helper: the projection repeated for T frames, [T, 3, 4]."""
    return np.repeat(pinhole_matrix(K)[None], T, axis=0)


def _project(P, pts):
    """pts [..., 4] homogeneous -> pixel coordinates [..., 2]."""
    img = pts @ P.T
    return img[..., :2] / img[..., 2:3]


def camera_to_scene_motion(w2cs, c2ws, K, depth_map, width, height, istrain=True):
    """w2cs / c2ws: T world-to-camera / camera-to-world 4x4 matrices; depth_map: height x width (any shape with
    that many values).  Returns float64 [T, 2, height, width].  `istrain` does not change the result (the
    reference's two branches assign the same clipped flow)."""
    T = len(w2cs)
    P = pinhole_matrix(K)
    xs = np.arange(-width // 2, width // 2, 1)
    ys = np.arange(-height // 2, height // 2, 1)
    gx, gy = np.meshgrid(xs, ys)
    # the depth arithmetic stays in the map's own dtype (the demo depth maps are fp16 .npy files and the reference
    # computes 100 - depth * 50 before anything is widened: at z ~ 100 an fp16 step is 0.06)
    z = (100 - np.asarray(depth_map).reshape(-1) * 50).astype(np.float64)
    pts = np.stack([gx.reshape(-1).astype(np.float64), gy.reshape(-1).astype(np.float64), z, np.ones_like(z)], axis=-1)
    base = _project(P, pts)                                           # [HW, 2], the same for every frame
    flow = np.zeros((T, 2, height, width))
    if T < 2:
        return flow
    w2c = np.stack([np.asarray(m, dtype=np.float64) for m in w2cs])
    c2w = np.stack([np.asarray(m, dtype=np.float64) for m in c2ws])
    step = w2c[1:] @ c2w[:-1]                                         # camera t -> camera t+1, [T-1, 4, 4]
    moved = np.einsum("tij,nj->tni", step, pts)                       # [T-1, HW, 4]
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = _project(P, moved) - base[None]                        # [T-1, HW, 2]
    disp = disp.transpose(0, 2, 1).reshape(T - 1, 2, height, width)
    if np.isfinite(disp).all():
        mu, sigma = disp.mean(), disp.std()
        flow[1:] = np.clip(disp, mu - 3.0 * sigma, mu + 3.0 * sigma)
    return flow
