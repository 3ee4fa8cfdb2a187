"""Generates the golden vectors under tests/golden/ from the REFERENCE'S OWN unmodified modules
(Oracle-R, oracle/mmcv_stub.py).  Run in the dev container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests and no golden vectors for this path, so these files are
what pins parity.  Inputs are never stored: every consumer regenerates them from the same seeded
generators (bevformer_b200/synthetic.py), which are deterministic on CPU.  Stored per case: a fixed
subset of rows of each result plus whole-tensor statistics, small enough to commit.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bevformer_b200 import synthetic as syn  # noqa: E402
from oracle import mmcv_stub  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def stats(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double()
    return np.array([t.sum().item(), t.abs().sum().item(), t.square().sum().item(),
                     t.abs().max().item()], dtype=np.float64)


def row_subset(n: int, k: int, seed: int = 7) -> np.ndarray:
    g = np.random.default_rng(seed)
    return np.sort(g.choice(n, size=min(k, n), replace=False)).astype(np.int64)


def fixed_projection(shape, seed=11, dtype=torch.float32) -> torch.Tensor:
    """The test loss is (out * R).sum() with this fixed R (a LayerNorm-terminated encoder has ~zero
    gradient under out.square().mean())."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype)


# ------------------------------------------------------------------------------------------------
def op_case(name, bs, levels, nq, heads, dim, pts, seed, dtype=torch.float64, loc_range=(-0.1, 1.1),
            value_scale=1.0, keep_rows=256):
    ref = mmcv_stub.load_reference_modules()
    v, ss, lsi, loc, attn = syn.make_msda_inputs(bs, levels, nq, heads, dim, pts, seed=seed,
                                                 dtype=dtype, loc_range=loc_range,
                                                 value_scale=value_scale)
    v.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    out = ref.msda_pytorch(v, ss, loc, attn)
    gout = fixed_projection(out.shape, dtype=dtype)
    (out * gout).sum().backward()
    rq = row_subset(nq, keep_rows)
    rs = row_subset(v.shape[1], keep_rows, seed=8)
    np.savez_compressed(
        os.path.join(OUT, f"msda_{name}.npz"),
        meta=np.array([bs, nq, heads, dim, pts, seed], dtype=np.int64),
        levels=np.array(levels, dtype=np.int64), loc_range=np.array(loc_range),
        value_scale=np.array(value_scale),
        rows_q=rq, rows_s=rs,
        out_rows=out.detach()[:, rq].float().numpy(), out_stats=stats(out),
        grad_value_rows=v.grad[:, rs].float().numpy(), grad_value_stats=stats(v.grad),
        grad_loc_rows=loc.grad[:, rq].float().numpy(), grad_loc_stats=stats(loc.grad),
        grad_attn_rows=attn.grad[:, rq].float().numpy(), grad_attn_stats=stats(attn.grad))
    print(f"msda_{name}: out {tuple(out.shape)} |max| {out.abs().max():.4f}")


def encoder_case(name, workload, bs=1, with_prev=True, seed=0, keep_rows=512, backward=True,
                 full_output=False):
    w = syn.WORKLOADS[workload]
    cfg = (mmcv_stub.load_reference_encoder_cfg(w.config_file) if w.config_file
           else syn.encoder_cfg(w))
    enc = mmcv_stub.build_reference_encoder(encoder_cfg=cfg).eval()   # eval(): dropout = identity
    enc.load_state_dict(syn.make_state_dict(w, seed=seed))
    inp = syn.make_encoder_inputs(w, bs=bs, seed=seed, with_prev=with_prev)
    if bs > 1:   # make the batch items differ
        g = torch.Generator().manual_seed(99)
        inp.feat = inp.feat + 0.5 * torch.randn(inp.feat.shape, generator=g)
        inp.bev_query = inp.bev_query + 0.1 * torch.randn(inp.bev_query.shape, generator=g)
    t0 = time.time()
    save = {}
    if backward:
        inp.bev_query.requires_grad_(True)
        inp.feat.requires_grad_(True)
        inp.bev_pos.requires_grad_(True)      # the learned positional encoding trains (every layer's TSA adds it)
    with torch.set_grad_enabled(backward):
        out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs())
    rq = row_subset(w.num_query, keep_rows)
    save.update(out_rows=out.detach()[:, rq].numpy(), out_stats=stats(out), rows_q=rq)
    if full_output:
        save["out_full"] = out.detach().numpy()
    if backward:
        proj = fixed_projection(out.shape)
        (out * proj).sum().backward()
        rs = row_subset(w.num_value, keep_rows, seed=8)
        save.update(rows_s=rs,
                    grad_query_rows=inp.bev_query.grad[rq].numpy(),
                    grad_query_stats=stats(inp.bev_query.grad),
                    grad_feat_rows=inp.feat.grad[:, rs].numpy(),
                    grad_feat_stats=stats(inp.feat.grad),
                    grad_pos_rows=inp.bev_pos.grad[rq].numpy(),
                    grad_pos_stats=stats(inp.bev_pos.grad))
        for k, p in enc.named_parameters():
            save["gstat:" + k] = stats(p.grad)
            if p.grad.numel() <= 1024:
                save["gfull:" + k] = p.grad.numpy()
            elif p.dim() == 2:
                save["grows:" + k] = p.grad[: min(4, p.shape[0])].numpy()
    np.savez_compressed(os.path.join(OUT, f"encoder_{name}.npz"),
                        meta=np.array([bs, int(with_prev), seed], dtype=np.int64), **save)
    print(f"encoder_{name}: out {tuple(out.shape)} in {time.time() - t0:.1f}s")


def grid_length_of(w):
    """BEV cell size that keeps the 102.4 m range at this BEV resolution."""
    return (0.512 * 200 / w.bev_h, 0.512 * 200 / w.bev_w)


def perception_case(name, workload, bs=2, with_prev=True, seed=0, keep_rows=512):
    """PerceptionTransformer.get_bev_features of the reference's own class (fp32, eval mode)."""
    w = syn.WORKLOADS[workload]
    PT = mmcv_stub.load_reference_transformer()
    cfg = (mmcv_stub.load_reference_encoder_cfg(w.config_file) if w.config_file
           else syn.encoder_cfg(w))
    m = PT(num_feature_levels=len(w.levels), num_cams=w.num_cams, encoder=cfg, decoder=None,
           embed_dims=w.embed_dims, rotate_center=[w.bev_h // 2, w.bev_w // 2]).eval()
    m.load_state_dict(syn.make_perception_state_dict(w, seed=seed))
    inp = syn.make_perception_inputs(w, bs=bs, seed=seed, with_prev=with_prev)
    for f in inp.mlvl_feats:
        f.requires_grad_(True)
    inp.bev_queries.requires_grad_(True)
    t0 = time.time()
    prev = None if inp.prev_bev is None else inp.prev_bev.clone()      # the reference rotates in place
    out = m.get_bev_features(inp.mlvl_feats, inp.bev_queries, w.bev_h, w.bev_w,
                             grid_length=grid_length_of(w), bev_pos=inp.bev_pos, prev_bev=prev,
                             img_metas=inp.img_metas)
    (out * fixed_projection(out.shape)).sum().backward()
    rq = row_subset(w.num_query, keep_rows)
    save = dict(out_rows=out.detach()[:, rq].numpy(), out_stats=stats(out), rows_q=rq,
                grad_queries_rows=inp.bev_queries.grad[rq].numpy(),
                grad_queries_stats=stats(inp.bev_queries.grad))
    for i, f in enumerate(inp.mlvl_feats):
        save[f"grad_feat{i}_stats"] = stats(f.grad)
        save[f"grad_feat{i}_slice"] = f.grad[:, :, :8, :2].numpy()
    for k, p in m.named_parameters():
        if k.startswith("encoder.") or p.grad is None:
            continue
        save["gstat:" + k] = stats(p.grad)
        save["gfull:" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(OUT, f"perception_{name}.npz"),
                        meta=np.array([bs, int(with_prev), seed], dtype=np.int64), **save)
    print(f"perception_{name}: out {tuple(out.shape)} in {time.time() - t0:.1f}s")


def main(which):
    if not mmcv_stub.reference_available():
        raise SystemExit("needs /root/reference (dev container only)")
    torch.manual_seed(0)
    cases = {
        # mmcv's upstream known-answer geometry: N=1, M=2, D=2, Lq=2, L=2, P=2
        "kat": lambda: op_case("kat", 1, [(6, 4), (3, 2)], 2, 2, 2, 2, seed=3, value_scale=0.01,
                               loc_range=(0.0, 1.0)),
        # same geometry with out-of-range locations and an odd head_dim (generic kernel)
        "kat_oob": lambda: op_case("kat_oob", 2, [(6, 4), (3, 2)], 5, 2, 30, 3, seed=4,
                                   loc_range=(-0.4, 1.4)),
        # BASELINE.json configs[0]: 1 cam, 1 level, 64x64 BEV, 4 points
        "config0": lambda: op_case("config0", 1, [(64, 64)], 4096, 8, 32, 4, seed=0),
        # a 4-level pyramid with SCA's point count
        "pyramid": lambda: op_case("pyramid", 2, [(29, 50), (15, 25), (8, 13), (4, 7)], 700, 8, 32,
                                   8, seed=5, loc_range=(-0.2, 1.2)),
        "enc_toy": lambda: encoder_case("toy", "toy", full_output=True),
        "enc_toy_bs2": lambda: encoder_case("toy_bs2", "toy", bs=2, full_output=True),
        "enc_toy_noprev": lambda: encoder_case("toy_noprev", "toy", with_prev=False,
                                               full_output=True),
        "enc_tiny": lambda: encoder_case("tiny", "tiny", keep_rows=256),
        "enc_tiny_noprev": lambda: encoder_case("tiny_noprev", "tiny", with_prev=False,
                                                backward=False),
        "enc_small": lambda: encoder_case("small", "small", keep_rows=256),
        "enc_small4": lambda: encoder_case("small4", "small4", keep_rows=256),
        "enc_base": lambda: encoder_case("base", "base", keep_rows=256),
        # PerceptionTransformer.get_bev_features: CAN-bus shift / MLP, prev_bev rotation, level + camera
        # embeddings, then the encoder
        "per_toy": lambda: perception_case("toy", "toy", bs=2),
        "per_toy_noprev": lambda: perception_case("toy_noprev", "toy", bs=1, with_prev=False),
        "per_tiny": lambda: perception_case("tiny", "tiny", bs=1, keep_rows=256),
    }
    for k in (which or cases):
        cases[k]()


if __name__ == "__main__":
    main(sys.argv[1:])
