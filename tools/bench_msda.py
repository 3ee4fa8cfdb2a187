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
    args = ap.parse_args()
    dev = "cuda"
    w = syn.WORKLOADS["base"]
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
        if args.only and args.only not in name:
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
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_msda.json", "w"), indent=1)


if __name__ == "__main__":
    main()
