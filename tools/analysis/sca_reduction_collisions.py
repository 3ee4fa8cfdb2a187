"""Offline count (CPU, numpy): how many of the SCA sampler backward's row reductions could be merged
before they reach L2?  Uses the synthetic rig's real geometry (in-view pairs, 8x8-tile order, the
reference's ring offsets + 0.3 px jitter) and reports, per pyramid level, the number of distinct
(head, cell) rows touched by groups of G consecutive pairs against the number of row reductions.
Result (profiles/README.md): G = 8 -> 0.43, G = 64 -> 0.16, G = 256 -> 0.10 of today's reductions;
merging only samples whose top-left cell coincides (what a warp can do with match/shuffle) -> 0.69.
"""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bevformer_b200 import synthetic as syn
w = syn.WORKLOADS["base"]
metas = syn.make_img_metas(w, 1)
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


ref_cam, mask = project_pillars(w, metas)                       # (cam, Nq, D, 2), (cam, Nq, D)
vis = mask.any(-1)   # (cam, Nq)
sd = syn.make_state_dict(w)
bias = sd["layers.0.attentions.1.deformable_attention.sampling_offsets.bias"].view(8,4,8,2).numpy()  # (M,L,P,2)
rng = np.random.default_rng(0)
levels = w.levels
tile=8
tot = {l:0 for l in range(4)}; dist = {G:{l:0 for l in range(4)} for G in (8,64,256)}
for cam in range(6):
    q = np.nonzero(vis[cam])[0]
    qi, qj = q // w.bev_w, q % w.bev_w
    key = ((qi//tile)*((w.bev_w+tile-1)//tile) + qj//tile)*(tile*tile) + (qi%tile)*tile + qj%tile
    q = q[np.argsort(key, kind='stable')]
    R = len(q)
    for l,(H,W) in enumerate(levels):
        # loc (R, M, P, 2): anchor d = p % 4
        off = bias[:, l][None] + 0.3*rng.standard_normal((R,8,8,2)).astype(np.float32)   # (R,M,P,2)
        anchors = ref_cam[cam][q][:, np.arange(8)%4]          # (R,P,2)
        loc = anchors[:,None,:,:] + off/np.array([W,H],np.float32)
        x = loc[...,0]*W-0.5; y = loc[...,1]*H-0.5
        valid = (x>-1)&(y>-1)&(x<W)&(y<H)
        x0=np.floor(x).astype(np.int64); y0=np.floor(y).astype(np.int64)
        cells=[]
        for dx in (0,1):
            for dy in (0,1):
                xx=x0+dx; yy=y0+dy
                ok = valid&(xx>=0)&(xx<W)&(yy>=0)&(yy<H)
                m = np.broadcast_to(np.arange(8)[None,:,None], xx.shape)
                r = np.broadcast_to(np.arange(R)[:,None,None], xx.shape)
                cid = (m*H+yy)*W+xx
                cells.append((r[ok], cid[ok]))
        r = np.concatenate([c[0] for c in cells]); cid=np.concatenate([c[1] for c in cells])
        tot[l]+= len(r)
        for G in dist:
            k = (r//G)*(8*H*W)+cid
            dist[G][l]+= len(np.unique(k))
print("level: total row-reductions | distinct per group of G consecutive pairs (ratio)")
for l in range(4):
    print(l, levels[l], tot[l], {G: f"{dist[G][l]} ({dist[G][l]/tot[l]:.2f})" for G in dist})
T=sum(tot.values())
for G in dist: print("G",G,"overall ratio", sum(dist[G].values())/T)

# sample-level merge: rows of a group of G consecutive pairs (same head, level, point) whose top-left cell coincides
print("sample-level (same top-left cell) distinct ratio per level, G=8 and G=4:")
for G in (4, 8):
    tot_s = 0; dist_s = 0; per = []
    rng = np.random.default_rng(0)
    for cam in range(6):
        q = np.nonzero(vis[cam])[0]
        qi, qj = q // w.bev_w, q % w.bev_w
        key = ((qi//tile)*((w.bev_w+tile-1)//tile) + qj//tile)*(tile*tile) + (qi%tile)*tile + qj%tile
        q = q[np.argsort(key, kind='stable')]
        R = len(q)
        for l,(H,W) in enumerate(levels):
            off = bias[:, l][None] + 0.3*rng.standard_normal((R,8,8,2)).astype(np.float32)
            anchors = ref_cam[cam][q][:, np.arange(8)%4]
            loc = anchors[:,None,:,:] + off/np.array([W,H],np.float32)
            x = loc[...,0]*W-0.5; y = loc[...,1]*H-0.5
            valid = (x>-1)&(y>-1)&(x<W)&(y<H)
            x0=np.floor(x).astype(np.int64); y0=np.floor(y).astype(np.int64)
            m = np.broadcast_to(np.arange(8)[None,:,None], x0.shape)
            pp = np.broadcast_to(np.arange(8)[None,None,:], x0.shape)
            r = np.broadcast_to(np.arange(R)[:,None,None], x0.shape)
            cid = (((r//G)*8+m)*8+pp)*((H+2)*(W+2)) + (y0+1)*(W+2)+(x0+1)
            n = valid.sum(); d = len(np.unique(cid[valid]))
            tot_s += n; dist_s += d
            if cam == 0: per.append(round(d/n,2))
    print("G",G,"overall", round(dist_s/tot_s,3), "cam0 per level", per)
