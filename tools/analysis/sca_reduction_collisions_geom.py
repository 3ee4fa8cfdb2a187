"""Shared geometry helper of the tools/analysis scripts."""
import numpy as np
from bevformer_b200 import synthetic as syn


def project_pillars(w, metas):
    """Pillar anchors of every BEV query projected into every camera (float64 numpy; this tool only needs
    the geometry, so it does its own projection instead of borrowing the test oracle's)."""
    pc = syn.PC_RANGE
    xs = (np.arange(w.bev_w) + 0.5) / w.bev_w * (pc[3] - pc[0]) + pc[0]
    ys = (np.arange(w.bev_h) + 0.5) / w.bev_h * (pc[4] - pc[1]) + pc[1]
    zs = (np.linspace(0.5, 7.5, 4) / 8.0) * (pc[5] - pc[2]) + pc[2]
    X, Y = np.meshgrid(xs, ys)                                  # q = i * W + j
    pts = np.stack([np.broadcast_to(X.reshape(-1, 1), (X.size, 4)), np.broadcast_to(Y.reshape(-1, 1), (X.size, 4)),
                    np.broadcast_to(zs[None, :], (X.size, 4)), np.ones((X.size, 4))], -1)   # (Nq, D, 4)
    l2i = np.asarray(metas[0]["lidar2img"], dtype=np.float64)    # (cam, 4, 4)
    cam = np.einsum("cij,qdj->cqdi", l2i, pts)
    depth = cam[..., 2]
    xy = cam[..., :2] / np.maximum(depth, 1e-5)[..., None]
    h, wd = metas[0]["img_shape"][0][:2]
    xy = xy / np.array([wd, h])
    ok = (depth > 1e-5) & (xy[..., 0] > 0) & (xy[..., 0] < 1) & (xy[..., 1] > 0) & (xy[..., 1] < 1)
    return xy.astype(np.float32), ok


