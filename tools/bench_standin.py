"""GPU stand-in for "the reference CUDA op" (SURVEY.md §8d), development / measurement tool.

mmcv's ms_deform_attn CUDA kernel cannot be built here (third-party wheel, no network), so the number the
north-star's ">= 10x the reference CUDA op" is read against is the reference's OWN fallback arithmetic --
the grid_sample composition of multi_scale_deformable_attn_pytorch (restated in oracle/torch_ref.py,
validated against the reference modules) -- executed on the B200 by PyTorch's CUDA kernels, fp32 as the
reference runs it (custom_fwd(cast_inputs=float32)), TF32 matmuls allowed (the reference's default).

Reported, CUDA-event medians after warm-up:
  * op level, base shapes: SCA as the reference launches it (6 cameras x max_len = 9 507 zero-padded rows,
    4 levels, 8 points) and TSA (2 x 40 000 rows): stand-in fwd / fwd+bwd  vs  this repo's kernels on
    the work they actually do (44 511 in-view pairs, bf16 and fp32);
  * encoder level: the restated BEVFormerEncoder (6 layers, reference's padded re-batch and Python loops)
    fwd+bwd on the GPU vs nothing here (bench.py prints our number).
Baseline only; nothing of this file is on the product path.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevformer_b200 import ops, synthetic as syn  # noqa: E402
from oracle import torch_ref  # noqa: E402  (baseline leg only)


def med_ms(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(iters)]
    for s, e in ev:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2]


def op_level(dev, res):
    from tools.bench_msda import rig_sca_inputs, rig_tsa_inputs
    w = syn.WORKLOADS["base"]
    # ---- SCA: the reference re-batches to (bs*6, max_len, ...) zero-padded rows (spatial_cross_attention.py:141-153)
    v, ss, lsi, loc, attn, row_map = rig_sca_inputs(dev)
    R = loc.shape[0]
    rm = row_map.long()
    counts = torch.bincount(rm, minlength=6)
    max_len = int(counts.max())
    loc_p = torch.zeros(6, max_len, 8, 4, 8, 2, device=dev)
    att_p = torch.zeros(6, max_len, 8, 4, 8, device=dev)
    for cam in range(6):
        idx = (rm == cam).nonzero().flatten()
        loc_p[cam, : idx.numel()] = loc[idx]
        att_p[cam, : idx.numel()] = attn[idx]
    ssl = [(int(h), int(w_)) for h, w_ in ss.tolist()]

    def standin_fwd(vv, lp, ap, shapes):
        return torch_ref.msda_grid_sample(vv, shapes, lp, ap)

    def standin_fwdbwd(vv, lp, ap, shapes):
        vv = vv.detach().requires_grad_(True); lp = lp.detach().requires_grad_(True); ap = ap.detach().requires_grad_(True)
        out = torch_ref.msda_grid_sample(vv, shapes, lp, ap)
        out.backward(torch.ones_like(out))

    t_f = med_ms(lambda: standin_fwd(v, loc_p, att_p, ssl))
    t_fb = med_ms(lambda: standin_fwdbwd(v, loc_p, att_p, ssl))
    res["sca_standin_fp32_fwd_ms"], res["sca_standin_fp32_fwdbwd_ms"] = t_f, t_fb
    res["sca_standin_rows"] = 6 * max_len
    for dt in (torch.float32, torch.bfloat16):
        vd = v.to(dt)
        out = ops.msda_rows_forward(vd, ss, lsi, loc, attn, row_map)
        g = torch.ones_like(out)
        gv = torch.zeros(v.shape, device=dev, dtype=torch.float32)
        f = med_ms(lambda: ops.msda_rows_forward(vd, ss, lsi, loc, attn, row_map), 20, 3)
        b = med_ms(lambda: (gv.zero_(), ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, g, gv)), 20, 3)
        n = "bf16" if dt == torch.bfloat16 else "fp32"
        res[f"sca_ours_{n}_fwd_ms"], res[f"sca_ours_{n}_fwdbwd_ms"] = f, f + b
    res["sca_ours_rows"] = R
    # ---- TSA
    v, ss, lsi, loc, attn = rig_tsa_inputs(dev)
    ssl = [(200, 200)]
    res["tsa_standin_fp32_fwd_ms"] = med_ms(lambda: standin_fwd(v, loc, attn, ssl))
    res["tsa_standin_fp32_fwdbwd_ms"] = med_ms(lambda: standin_fwdbwd(v, loc, attn, ssl))
    for dt in (torch.float32, torch.bfloat16):
        vd = v.to(dt)
        out = ops.msda_forward(vd, ss, lsi, loc, attn)
        g = torch.ones_like(out)
        gv = torch.zeros(v.shape, device=dev, dtype=torch.float32)
        f = med_ms(lambda: ops.msda_forward(vd, ss, lsi, loc, attn), 20, 3)
        b = med_ms(lambda: (gv.zero_(), ops.msda_backward(vd, ss, lsi, loc, attn, g, gv)), 20, 3)
        n = "bf16" if dt == torch.bfloat16 else "fp32"
        res[f"tsa_ours_{n}_fwd_ms"], res[f"tsa_ours_{n}_fwdbwd_ms"] = f, f + b
    nq = w.num_query
    # BEV queries/s of the op pair (one SCA + one TSA call = one layer's sampler work), fwd+bwd
    res["msda_qps_standin_fp32"] = nq / ((res["sca_standin_fp32_fwdbwd_ms"] + res["tsa_standin_fp32_fwdbwd_ms"]) * 1e-3)
    res["msda_qps_ours_bf16"] = nq / ((res["sca_ours_bf16_fwdbwd_ms"] + res["tsa_ours_bf16_fwdbwd_ms"]) * 1e-3)
    res["msda_qps_ours_fp32"] = nq / ((res["sca_ours_fp32_fwdbwd_ms"] + res["tsa_ours_fp32_fwdbwd_ms"]) * 1e-3)
    res["msda_ratio_ours_bf16_over_standin"] = res["msda_qps_ours_bf16"] / res["msda_qps_standin_fp32"]
    res["msda_ratio_ours_fp32_over_standin"] = res["msda_qps_ours_fp32"] / res["msda_qps_standin_fp32"]


def encoder_level(dev, res, layers):
    """The restated reference encoder (padded re-batch, Python loops, grid_sample) on the GPU, fp32."""
    w = syn.WORKLOADS["base"]
    sd = {k: v.to(dev).requires_grad_(True) for k, v in syn.make_state_dict(w).items()}
    inp = syn.make_encoder_inputs(w, bs=1, seed=0)
    kw = inp.kwargs()
    for k in ("bev_pos", "prev_bev", "shift"):
        kw[k] = kw[k].to(dev)
    bq = inp.bev_query.to(dev).requires_grad_(True)
    ft = inp.feat.to(dev).requires_grad_(True)
    proj = torch.randn(1, w.num_query, w.embed_dims, device=dev)

    def step():
        for t in list(sd.values()) + [bq, ft]:
            t.grad = None
        with torch.device(dev):            # the restatement builds its small constant tensors on the default device
            out = torch_ref.encoder_forward(sd, layers, bq, ft, use_c_oracle=False, **kw)
        (out * proj).sum().backward()

    t = med_ms(step, 3, 1)
    res["encoder_standin_layers"] = layers
    res["encoder_standin_fp32_fwdbwd_ms"] = t * (w.num_layers / layers)
    res["encoder_qps_standin_fp32"] = w.num_query / (res["encoder_standin_fp32_fwdbwd_ms"] * 1e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--skip-encoder", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True       # the reference's default (tools/train.py:142-144)
    res = {"what": "grid_sample-on-B200 stand-in for the reference CUDA op vs this repo's kernels"}
    op_level(dev, res)
    if not args.skip_encoder:
        try:
            encoder_level(dev, res, args.layers)
        except Exception as exc:  # noqa: BLE001
            res["encoder_standin_error"] = repr(exc)[:300]
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_standin.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
