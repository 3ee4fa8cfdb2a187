"""ctypes front-end of Oracle-S (``oracle/msda_ref.c``).  TEST INFRASTRUCTURE ONLY: importable
from tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / --impl reference legs; the
product (bevformer_b200/, projects/) never imports it.

Follows the call convention of the reference's op wrapper
(projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:94-128 forward,
:130-163 backward: zero-filled grads, accumulate in place, None for integer inputs).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmsda_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "msda_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _prep(value, spatial_shapes, level_start_index, loc, attn):
    dt = value.dtype
    if dt not in (torch.float32, torch.float64):
        dt = torch.float32
    v = value.detach().to("cpu", dt).contiguous().numpy()
    lo = loc.detach().to("cpu", dt).contiguous().numpy()
    at = attn.detach().to("cpu", dt).contiguous().numpy()
    hw = np.ascontiguousarray(torch.as_tensor(spatial_shapes).cpu().numpy().astype(np.int64))
    ls = np.ascontiguousarray(torch.as_tensor(level_start_index).cpu().numpy().astype(np.int64))
    B, S, M, D = v.shape
    _, Q, _, L, P, _ = lo.shape
    dims = [ctypes.c_int64(int(x)) for x in (B, S, M, D, Q, L, P)]
    return dt, v, lo, at, hw, ls, dims, (B, S, M, D, Q, L, P)


def msda_forward(value, spatial_shapes, level_start_index, loc, attn) -> torch.Tensor:
    lib = _load()
    dt, v, lo, at, hw, ls, dims, (B, S, M, D, Q, L, P) = _prep(
        value, spatial_shapes, level_start_index, loc, attn)
    out = np.empty((B, Q, M * D), dtype=v.dtype)
    fn = lib.msda_oracle_forward_f32 if dt == torch.float32 else lib.msda_oracle_forward_f64
    fn(_ptr(v), _ptr(hw), _ptr(ls), _ptr(lo), _ptr(at), _ptr(out), *dims)
    return torch.from_numpy(out)


def msda_backward(value, spatial_shapes, level_start_index, loc, attn, grad_out):
    lib = _load()
    dt, v, lo, at, hw, ls, dims, (B, S, M, D, Q, L, P) = _prep(
        value, spatial_shapes, level_start_index, loc, attn)
    go = grad_out.detach().to("cpu", dt).contiguous().numpy()
    gv, gl, ga = np.zeros_like(v), np.zeros_like(lo), np.zeros_like(at)
    fn = lib.msda_oracle_backward_f32 if dt == torch.float32 else lib.msda_oracle_backward_f64
    fn(_ptr(v), _ptr(hw), _ptr(ls), _ptr(lo), _ptr(at), _ptr(go), _ptr(gv), _ptr(gl), _ptr(ga),
       *dims)
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(ga)


class MSDAOracleFunction(torch.autograd.Function):
    """autograd wrapper so module-level restatements can train through Oracle-S on CPU."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, loc, attn, im2col_step=64):
        ctx.save_for_backward(value, spatial_shapes, level_start_index, loc, attn)
        return msda_forward(value, spatial_shapes, level_start_index, loc, attn).to(value.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        value, ss, lsi, loc, attn = ctx.saved_tensors
        gv, gl, ga = msda_backward(value, ss, lsi, loc, attn, grad_out)
        return gv.to(value.dtype), None, None, gl.to(loc.dtype), ga.to(attn.dtype), None
