"""Does a stream of weight-gradient GEMMs hide behind the sampler backward?  (development probe)

The sampler backward is bound by L2 reductions (issue slots ~40 % busy); the weight-gradient kernels
are short and latency-bound.  Runs one SCA sampler backward (real rig geometry) alone, eight wgrad
launches alone, and both on two streams, and prints the three times.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bevformer_b200 import ops  # noqa: E402
from bench_msda import rig_sca_inputs  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    value, ss, lsi, loc, attn, row_map = rig_sca_inputs(dev)
    value = value.to(torch.bfloat16)
    go = torch.randn(loc.shape[0], 256, device=dev, dtype=torch.bfloat16)
    gv = torch.zeros(value.shape, device=dev, dtype=torch.float32)
    dy = [torch.randn(40000, n, device=dev, dtype=torch.bfloat16) for n in (256, 512, 256, 768)]
    x = [torch.randn(40000, k, device=dev, dtype=torch.bfloat16) for k in (256, 256, 512, 256)]
    side = torch.cuda.Stream(dev, priority=-1)     # high priority: its CTAs are placed first when an SM frees up

    def sampler():
        ops.msda_rows_backward(value, ss, lsi, loc, attn, row_map, go, grad_value=gv)

    def wgrads():
        for _ in range(2):
            for a, b in zip(dy, x):
                ops.linear_wgrad_tc(a, b, with_bias=True)

    def both():
        main_s = torch.cuda.current_stream(dev)      # the capture stream while a graph is being recorded
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            wgrads()
        sampler()
        main_s.wait_stream(side)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    # capture each variant in a CUDA graph so host launch latency does not blur the comparison
    def graphed(fn):
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn()
        return g.replay

    def both_rev():
        main_s = torch.cuda.current_stream(dev)
        side.wait_stream(main_s)
        with torch.cuda.stream(side):
            sampler()
        wgrads()
        main_s.wait_stream(side)

    t_r = timed(graphed(both_rev))
    print(f"sampler on the high-priority stream, wgrads on main: {t_r:.3f} ms")
    t_s, t_w, t_b = timed(graphed(sampler)), timed(graphed(wgrads)), timed(graphed(both))
    print(f"sampler bwd alone {t_s:.3f} ms | 8 wgrads alone {t_w:.3f} ms | sum {t_s + t_w:.3f} ms | two streams {t_b:.3f} ms")


if __name__ == "__main__":
    main()
