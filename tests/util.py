"""Shared helpers of the test-suite: golden-file access and comparison metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def stats(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().cpu()
    return np.array([t.sum().item(), t.abs().sum().item(), t.square().sum().item(),
                     t.abs().max().item()])


def fixed_projection(shape, seed=11, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dtype)


def max_err(a, b) -> float:
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return (a - b).abs().max().item() if a.numel() else 0.0


def rel_err(a, b) -> float:
    """max |a-b| / max(1, max|b|): the form in which the 1e-3 (fp32) / 1e-2 (bf16) bars are read."""
    b_ = torch.as_tensor(b).detach().double().cpu()
    scale = max(1.0, b_.abs().max().item()) if b_.numel() else 1.0
    return max_err(a, b) / scale


def stats_close(got: np.ndarray, want: np.ndarray, rtol: float) -> bool:
    """sum |x|, sum x^2 and max |x| of the whole tensor agree (sum x itself cancels too much)."""
    return bool(np.all(np.abs(got[1:] - want[1:]) <= rtol * np.maximum(np.abs(want[1:]), 1e-12)))


def msda_case_inputs(g, dtype=torch.float32, device="cpu"):
    from bevformer_b200 import synthetic as syn
    bs, nq, heads, dim, pts, seed = [int(x) for x in g["meta"]]
    levels = [tuple(int(v) for v in row) for row in g["levels"]]
    return syn.make_msda_inputs(bs, levels, nq, heads, dim, pts, seed=seed, dtype=dtype,
                                device=device, loc_range=tuple(float(x) for x in g["loc_range"]),
                                value_scale=float(g["value_scale"]))
