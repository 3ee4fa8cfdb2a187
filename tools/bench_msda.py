"""Kernel-level microbenchmark of the sampler at BASELINE shapes (development tool, GPU only).

Times bevf_msda_forward/backward with CUDA events on the launching stream; every timed iteration is
preceded by an L2 flush (a 256 MB memset, larger than the 126 MB L2).  Reports achieved algorithmic
GB/s using SURVEY.md §8d's compulsory-byte formulas.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevformer_b200 import ops, synthetic as syn  # noqa: E402


def alg_bytes(B, S, Q, M, L, P, sv, so, bwd):
    C = M * 32
    fwd = B * S * C * sv + B * Q * M * L * P * 12 + B * Q * C * so
    if not bwd:
        return fwd
    return fwd + B * S * C * 4 + B * Q * M * L * P * 12


def rig_sca_inputs(dev):
    """SCA sampler inputs with the REAL geometry: in-view (camera, query) pairs of the synthetic rig,
    sampling points = projected pillar anchors + the reference's ring offsets (+ jitter)."""
    import numpy as np
    from bevformer_b200 import ops as _ops
    from bevformer_b200.plugin import ScaPlan
    w = syn.WORKLOADS["base"]
    metas = syn.make_img_metas(w, 1)
    l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in metas], dtype=np.float32)).to(dev)
    z = (torch.linspace(0.5, 7.5, 4) / 8.0).tolist()
    ref_cam, mask = _ops.point_sampling(l2i, syn.PC_RANGE, z, w.img_hw[0], w.img_hw[1], w.bev_h, w.bev_w)
    plan = ScaPlan.build(mask, ref_cam, (w.bev_h, w.bev_w) if os.environ.get("BENCH_TILE", "1") == "1" else None)
    sd = syn.make_state_dict(w)
    g = torch.Generator().manual_seed(0)
    nq, m, l, p = w.num_query, 8, 4, 8
    raw = torch.cat([sd["layers.0.attentions.1.deformable_attention.sampling_offsets.bias"].view(1, -1)
                     + 0.5 * torch.randn(nq, m * l * p * 2, generator=g),
                     torch.randn(nq, m * l * p, generator=g)], 1).to(dev).contiguous()
    ss = torch.tensor(w.levels, dtype=torch.int64, device=dev)
    lsi = torch.tensor(w.level_start, dtype=torch.int64, device=dev)
    loc, attn = _ops.sca_prep_forward(raw, plan.ref_cam, plan.pair_q, plan.pair_cam, ss, 1, nq, m, l, p)
    value = torch.randn(6, w.num_value, 8, 32, generator=g).to(dev)
    return value, ss, lsi, loc, attn, plan.row_map


def rig_tsa_rows_inputs(dev):
    """TSA as the encoder launches it: interleaved rows (b, q, frame) through the row-list entry points,
    with the 8x8-tile group order for the backward's merge."""
    from bevformer_b200.plugin.temporal_self_attention import TemporalSelfAttention
    w = syn.WORKLOADS["base"]
    sd = syn.make_state_dict(w)
    g = torch.Generator().manual_seed(0)
    nq, m, p = w.num_query, 8, 4
    from bevformer_b200 import ops as _ops
    raw = torch.cat([sd["layers.0.attentions.0.sampling_offsets.bias"].view(1, -1)
                     + 0.5 * torch.randn(nq, m * 2 * p * 2, generator=g),
                     torch.randn(nq, m * 2 * p, generator=g)], 1).to(dev).contiguous()
    ys, xs = torch.meshgrid(torch.arange(200.0), torch.arange(200.0), indexing="ij")
    ref = torch.stack([(xs + 0.5) / 200, (ys + 0.5) / 200], -1).reshape(1, nq, 1, 2)
    ref = torch.cat([ref + torch.tensor([0.01, -0.02]), ref], 0).to(dev).contiguous()
    ss = torch.tensor([[200, 200]], dtype=torch.int64, device=dev)
    lsi = torch.zeros(1, dtype=torch.int64, device=dev)
    loc, attn = _ops.tsa_prep_forward(raw, ref, ss, 1, nq, m, 1, p, True)
    value = torch.randn(2, nq, 8, 32, generator=g).to(dev)
    tsa = TemporalSelfAttention()
    return value, ss, lsi, loc, attn, tsa._frame_map(1, nq, torch.device(dev)), tsa._group_order(1, 200, 200, torch.device(dev))


def rig_tsa_inputs(dev):
    w = syn.WORKLOADS["base"]
    sd = syn.make_state_dict(w)
    g = torch.Generator().manual_seed(0)
    nq, m, p = w.num_query, 8, 4
    from bevformer_b200 import ops as _ops
    raw = torch.cat([sd["layers.0.attentions.0.sampling_offsets.bias"].view(1, -1)
                     + 0.5 * torch.randn(nq, m * 2 * p * 2, generator=g),
                     torch.randn(nq, m * 2 * p, generator=g)], 1).to(dev).contiguous()
    ys, xs = torch.meshgrid(torch.arange(200.0), torch.arange(200.0), indexing="ij")
    ref = torch.stack([(xs + 0.5) / 200, (ys + 0.5) / 200], -1).reshape(1, nq, 1, 2)
    ref = torch.cat([ref + torch.tensor([0.01, -0.02]), ref], 0).to(dev).contiguous()
    ss = torch.tensor([[200, 200]], dtype=torch.int64, device=dev)
    lsi = torch.zeros(1, dtype=torch.int64, device=dev)
    loc, attn = _ops.tsa_prep_forward(raw, ref, ss, 1, nq, m, 1, p)
    value = torch.randn(2, nq, 8, 32, generator=g).to(dev)
    return value, ss, lsi, loc, attn


def time_op(fn, iters, flush):
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(iters)]
    for i in range(3):
        fn()
    for s, e in ev:
        flush.zero_()
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--profile", action="store_true", help="launch each kernel once (for ncu)")
    ap.add_argument("--kernels", action="store_true", help="also print per-kernel durations of the backward")
    args = ap.parse_args()
    dev = "cuda"
    w = syn.WORKLOADS["base"]
    if os.environ.get("BENCH_BWD_MODE"):
        from bevformer_b200 import _lib
        assert _lib.load().bevf_msda_set_backward_mode(int(os.environ["BENCH_BWD_MODE"])) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = {
        "tsa_base": dict(bs=2, levels=[(200, 200)], nq=40000, pts=4),
        "sca_base": dict(bs=6, levels=list(w.levels), nq=9507, pts=8),
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    res = []
    for name, sh in shapes.items():
        if args.only and not any(o and o in name for o in args.only.split(",")):
            continue
        v, ss, lsi, loc, attn = syn.make_msda_inputs(sh["bs"], sh["levels"], sh["nq"], 8, 32,
                                                     sh["pts"], seed=0, device=dev)
        B, S, M, _ = v.shape
        Q, L, P = sh["nq"], len(sh["levels"]), sh["pts"]
        for dt in (torch.float32, torch.bfloat16):
            vd = v.to(dt)
            out = ops.msda_forward(vd, ss, lsi, loc, attn)
            g = torch.randn_like(out)
            gv = torch.zeros(v.shape, device=dev, dtype=torch.float32)
            sz = 4 if dt == torch.float32 else 2
            if args.profile:
                ops.msda_backward(vd, ss, lsi, loc, attn, g, gv)
                torch.cuda.synchronize()
                continue
            t_f, t_fmin = time_op(lambda: ops.msda_forward(vd, ss, lsi, loc, attn), args.iters, flush)
            t_b, t_bmin = time_op(lambda: ops.msda_backward(vd, ss, lsi, loc, attn, g, gv),
                                  args.iters, flush)
            bf, bb = alg_bytes(B, S, Q, M, L, P, sz, sz, False), alg_bytes(B, S, Q, M, L, P, sz, sz, True)
            r = dict(shape=name, dtype=str(dt).split(".")[-1], fwd_ms=round(t_f, 4), bwd_ms=round(t_b, 4),
                     fwd_min_ms=round(t_fmin, 4), bwd_min_ms=round(t_bmin, 4),
                     fwd_alg_MB=round(bf / 1e6, 1), bwd_alg_MB=round(bb / 1e6, 1),
                     fwd_GBs=round(bf / t_f / 1e6, 1), bwd_GBs=round(bb / t_b / 1e6, 1),
                     fwd_frac=round(bf / t_f / 1e6 / hbm, 4), bwd_frac=round(bb / t_b / 1e6 / hbm, 4))
            print(json.dumps(r), flush=True)
            res.append(r)
    # ---- the same kernels on the real geometry (what the encoder actually launches)
    for name in ("sca_rig", "tsa_rig", "tsa_rows"):
        if args.only and not any(o and o in name for o in args.only.split(",")):
            continue
        order = None
        if name == "sca_rig":
            v, ss, lsi, loc, attn, row_map = rig_sca_inputs(dev)
            R = loc.shape[0]
        elif name == "tsa_rows":
            v, ss, lsi, loc, attn, row_map, order = rig_tsa_rows_inputs(dev)
        else:
            v, ss, lsi, loc, attn = rig_tsa_inputs(dev)
            row_map = None
        B, S, M, _ = v.shape
        L, P = loc.shape[-3], loc.shape[-2]
        Q = loc.shape[0] if row_map is not None else loc.shape[1]
        nrows = Q if row_map is not None else B * Q
        for dt in (torch.float32, torch.bfloat16):
            vd = v.to(dt)
            if row_map is not None:
                fwd = lambda: ops.msda_rows_forward(vd, ss, lsi, loc, attn, row_map)
            else:
                fwd = lambda: ops.msda_forward(vd, ss, lsi, loc, attn)
            out = fwd()
            g = torch.randn_like(out)
            gv = torch.zeros(v.shape, device=dev, dtype=torch.float32)
            dense = None
            if name == "sca_rig" and dt == torch.bfloat16 and os.environ.get("BENCH_DENSE", "0") != "0":
                from bevformer_b200 import _lib
                assert _lib.load().bevf_msda_set_dense_backward(int(os.environ["BENCH_DENSE"])) == 0
                per_cam = torch.bincount(row_map[row_map >= 0].long(), minlength=B)
                ends = per_cam.cumsum(0)
                dense = (list(w.levels), torch.stack([ends - per_cam, ends], 1).to(torch.int32).contiguous())
            gvm = os.environ.get("BENCH_GV", "fp32")
            if row_map is not None and dt == torch.bfloat16 and gvm == "f16":
                bwd = lambda: ops.msda_rows_backward_f16acc(vd, ss, lsi, loc, attn, row_map, g, order)
            elif row_map is not None and dt == torch.bfloat16 and gvm.startswith("mixed") and L > 1:
                nf = int(gvm[5:] or 1)
                bwd = lambda: ops.msda_rows_backward_mixed(vd, ss, lsi, [tuple(x) for x in ss.tolist()], nf, loc, attn, row_map, g, order)
            elif row_map is not None:
                bwd = lambda: ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, g, gv, order, dense=dense)
            else:
                bwd = lambda: ops.msda_backward(vd, ss, lsi, loc, attn, g, gv)
            if args.profile:
                fwd(); bwd(); torch.cuda.synchronize(); continue
            sz = 4 if dt == torch.float32 else 2
            t_f, t_fmin = time_op(fwd, args.iters, flush)
            t_b, t_bmin = time_op(bwd, args.iters, flush)
            if args.kernels:                       # per-kernel durations of the backward (CUPTI records)
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(5):
                        flush.zero_(); bwd()
                    torch.cuda.synchronize()
                agg = {}
                for ev in prof.events():
                    if ev.device_type == torch.autograd.DeviceType.CUDA and "msda" in ev.name:
                        agg.setdefault(ev.name[:70], []).append(ev.device_time)
                for k, ts in agg.items():
                    print(f"    {name} {str(dt).split('.')[-1]:9s} {sum(ts) / len(ts):8.1f} us x{len(ts) // 5}  {k}", flush=True)
            C = M * 32
            bf = B * S * C * sz + nrows * M * L * P * 12 + nrows * C * sz
            bb = bf + B * S * C * 4 + nrows * M * L * P * 12
            r = dict(shape=name, rows=nrows, dtype=str(dt).split(".")[-1], gv=os.environ.get("BENCH_GV", "fp32"), bwd_mode=os.environ.get("BEVF_MSDA_BWD", "one"), dense=os.environ.get("BENCH_DENSE", "0") if dense is not None else "0",
                     splat_direct=os.environ.get("BEVF_SPLAT_DIRECT", "0"), fwd_ms=round(t_f, 4),
                     bwd_ms=round(t_b, 4), fwd_alg_MB=round(bf / 1e6, 1), bwd_alg_MB=round(bb / 1e6, 1),
                     fwd_GBs=round(bf / t_f / 1e6, 1), bwd_GBs=round(bb / t_b / 1e6, 1),
                     fwd_frac=round(bf / t_f / 1e6 / hbm, 4), bwd_frac=round(bb / t_b / 1e6 / hbm, 4))
            print(json.dumps(r), flush=True)
            res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = os.environ.get("BENCH_TAG", "")
    json.dump(res, open(f"gpurun_out/bench_msda{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
