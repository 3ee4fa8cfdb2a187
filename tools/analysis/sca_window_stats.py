"""Offline geometry statistics (CPU, numpy) for the register-window splat of the SCA sampler backward:
for groups of G tile-ordered in-view pairs and one (head, level, point) slice, how large is the bounding
box of the samples' top-left cells, how many distinct pixel rows are touched, and which share of the
slices fits a (WX x WY) window?  Same synthetic rig / ring offsets + 0.3 px jitter as
sca_reduction_collisions.py."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bevformer_b200 import synthetic as syn
from tools.analysis.sca_reduction_collisions_geom import project_pillars  # noqa

w = syn.WORKLOADS["base"]
metas = syn.make_img_metas(w, 1)
ref_cam, mask = project_pillars(w, metas)
vis = mask.any(-1)
sd = syn.make_state_dict(w)
bias = sd["layers.0.attentions.1.deformable_attention.sampling_offsets.bias"].view(8, 4, 8, 2).numpy()
tile = int(os.environ.get("TILE", 8))
for G in (32, 64):
    print(f"== G = {G} pairs per group, tile {tile}x{tile}")
    rng = np.random.default_rng(0)
    for l, (H, W) in enumerate(w.levels):
        ext_x, ext_y, ndist, nsamp = [], [], 0, 0
        for cam in range(6):
            q = np.nonzero(vis[cam])[0]
            qi, qj = q // w.bev_w, q % w.bev_w
            key = ((qi // tile) * ((w.bev_w + tile - 1) // tile) + qj // tile) * (tile * tile) + (qi % tile) * tile + qj % tile
            q = q[np.argsort(key, kind="stable")]
            R = len(q) // G * G
            q = q[:R]
            off = bias[:, l][None] + 0.3 * rng.standard_normal((R, 8, 8, 2)).astype(np.float32)
            anchors = ref_cam[cam][q][:, np.arange(8) % 4]
            loc = anchors[:, None, :, :] + off / np.array([W, H], np.float32)
            x = loc[..., 0] * W - 0.5; y = loc[..., 1] * H - 0.5
            valid = (x > -1) & (y > -1) & (x < W) & (y < H)
            x0 = np.floor(x); y0 = np.floor(y)
            # group view: (R/G, G, M, P)
            xv = np.where(valid, x0, np.nan).reshape(R // G, G, 8, 8)
            yv = np.where(valid, y0, np.nan).reshape(R // G, G, 8, 8)
            with np.errstate(all="ignore"):
                ex = np.nanmax(xv, 1) - np.nanmin(xv, 1) + 1      # extent in top-left cells
                ey = np.nanmax(yv, 1) - np.nanmin(yv, 1) + 1
            ok = ~np.isnan(ex)
            ext_x.append(ex[ok]); ext_y.append(ey[ok])
            cid = (y0 + 1) * (W + 2) + (x0 + 1)
            g = np.broadcast_to((np.arange(R) // G)[:, None, None], cid.shape)
            m = np.broadcast_to(np.arange(8)[None, :, None], cid.shape)
            p = np.broadcast_to(np.arange(8)[None, None, :], cid.shape)
            # distinct pixel rows (corners) per slice: expand the 4 corners
            keys = []
            for dx in (0, 1):
                for dy in (0, 1):
                    xx = x0 + dx; yy = y0 + dy
                    okc = valid & (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
                    k = ((g * 8 + m) * 8 + p) * (H * W) + (yy * W + xx)
                    keys.append(k[okc])
                    nsamp += okc.sum()
            ndist += len(np.unique(np.concatenate(keys)))
        ex = np.concatenate(ext_x); ey = np.concatenate(ext_y)
        fits = {f"{a}x{b}": float(((ex <= a) & (ey <= b)).mean()) for a, b in ((2, 2), (3, 3), (4, 4), (6, 4), (6, 6), (8, 4), (8, 8), (12, 6), (16, 8))}
        print(f"level {l} {H}x{W}: ext_x mean {ex.mean():.1f} p90 {np.percentile(ex, 90):.0f} | ext_y mean {ey.mean():.1f} p90 {np.percentile(ey, 90):.0f} | "
              f"row reductions after merge / before = {ndist / nsamp:.3f}")
        print("    slices fitting window (top-left extents):", {k: round(v, 2) for k, v in fits.items()})
