"""Which PyTorch (ATen) kernels are still on the encoder's hot path, and which line launches them.

Runs the base training step eagerly (gradient arena on, as bench.py's graph body) under torch.profiler with
Python stacks, then lists every kernel that is not one of this library's, grouped by the innermost
bevformer_b200 frame that launched it.  Development tool for DESIGN.md's "ATen glue" budget.
Usage: python tools/glue_report.py [--out gpurun_out/glue.txt]
"""
from __future__ import annotations

import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevformer_b200 import synthetic as syn  # noqa: E402
from bevformer_b200.plugin import build_transformer_layer_sequence  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = syn.WORKLOADS["base"]
    dtype = torch.bfloat16
    enc = build_transformer_layer_sequence(syn.encoder_cfg(w))
    enc.load_state_dict(syn.make_state_dict(w))
    enc = enc.to(dev, dtype).train()
    enc.enable_grad_arena(overlap=True)
    host = syn.make_encoder_inputs(w, bs=1, seed=0)
    inp = {k: getattr(host, k).to(dev, dtype) for k in ("bev_query", "feat", "bev_pos", "prev_bev")}
    l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in host.img_metas], dtype=np.float32)).to(dev)
    shift = host.shift.to(dev)
    ss, lsi = host.spatial_shapes.to(dev), host.level_start_index.to(dev)
    proj = torch.randn(1, w.num_query, w.embed_dims, device=dev, dtype=dtype)
    bq = inp["bev_query"].requires_grad_(True)
    ft = inp["feat"].requires_grad_(True)

    def step():
        for t in list(enc.parameters()) + [bq, ft]:
            t.grad = None
        out = enc(bq, ft, ft, bev_h=w.bev_h, bev_w=w.bev_w, bev_pos=inp["bev_pos"], spatial_shapes=ss,
                  level_start_index=lsi, prev_bev=inp["prev_bev"], shift=shift, img_metas=host.img_metas,
                  lidar2img=l2i)
        (out * proj).sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
                 record_shapes=True) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
            continue
        where = next((f for f in (ev.stack or []) if "bevformer_b200" in f or "tools/" in f), "?")
        where = where.split("bevformer_b200/")[-1][:70]
        shapes = str([s for s in (ev.input_shapes or []) if s])[:60]
        for k in ev.kernels:
            if "bevf::" in k.name or k.name.startswith("bevf") or "sum_n_kernel" in k.name:
                continue
            a = agg[(k.name.replace("void at::native::", "")[:58], ev.name, where, shapes)]
            a[0] += k.duration
            a[1] += 1
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    total = sum(v[0] for _, v in rows) / args.steps
    lines = [f"non-library kernels per step: {total / 1e3:.3f} ms over {sum(v[1] for _, v in rows) // args.steps} launches"]
    for (kname, op, where, shapes), (us, n) in rows:
        lines.append(f"{us / args.steps:8.1f} us {n // args.steps:4d}x  {kname:58s} {op:22s} {where}  {shapes}")
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
