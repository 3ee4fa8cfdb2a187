"""Python face of the C ABI: argument validation, stream plumbing, autograd glue.

Mirrors the reference's op wrapper
(projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py): same class
names, same ``apply`` signature, same 6-tuple of gradients.  PyTorch is used for device memory and
the current stream only; all arithmetic happens in ``libbevformer_b200.so``.
"""
from __future__ import annotations

import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (bevformer_b200 has no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")


def _level_tensors(value, spatial_shapes, level_start_index):
    ss = torch.as_tensor(spatial_shapes)
    ls = torch.as_tensor(level_start_index)
    if ss.dim() != 2 or ss.shape[1] != 2 or ls.dim() != 1 or ls.shape[0] != ss.shape[0]:
        raise RuntimeError("spatial_shapes must be (num_levels, 2) and level_start_index (num_levels,)")
    ss = ss.to(device=value.device, dtype=torch.int64).contiguous()
    ls = ls.to(device=value.device, dtype=torch.int64).contiguous()
    return ss, ls


def _dims(value, loc, attn):
    if value.dim() != 4 or loc.dim() != 6 or attn.dim() != 5 or loc.shape[-1] != 2:
        raise RuntimeError("expected value (B,S,M,D), sampling_locations (B,Q,M,L,P,2), "
                           "attention_weights (B,Q,M,L,P)")
    B, S, M, D = value.shape
    B2, Q, M2, L, P, _ = loc.shape
    if (B2, M2) != (B, M) or tuple(attn.shape) != (B, Q, M, L, P):
        raise RuntimeError("value / sampling_locations / attention_weights shapes disagree")
    return B, S, M, D, Q, L, P


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                 out_dtype=None):
    """ms_deform_attn_forward: returns (B, Q, M*D) in ``out_dtype`` (default: value's dtype)."""
    if value.dtype not in _DT:
        raise RuntimeError(f"value dtype {value.dtype} not supported (float32 or bfloat16)")
    for t, n in ((value, "value"), (sampling_locations, "sampling_loc"),
                 (attention_weights, "attn_weight")):
        _need_cuda(t, n)
    loc = sampling_locations if sampling_locations.dtype == torch.float32 else sampling_locations.float()
    attn = attention_weights if attention_weights.dtype == torch.float32 else attention_weights.float()
    B, S, M, D, Q, L, P = _dims(value, loc, attn)
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    if ss.shape[0] != L:
        raise RuntimeError("spatial_shapes and sampling_locations disagree on num_levels")
    out_dtype = out_dtype or value.dtype
    out = torch.empty((B, Q, M * D), device=value.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(value.device):
        st = lib.bevf_msda_forward(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                   loc.data_ptr(), attn.data_ptr(), out.data_ptr(), _DT[out_dtype],
                                   B, S, M, D, Q, L, P, _stream_ptr(value))
    _lib.check(st, lib)
    return out


def msda_backward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                  grad_output, grad_value=None):
    """ms_deform_attn_backward. ``grad_value`` (fp32, zero-filled) is accumulated into when given,
    otherwise allocated here. Returns (grad_value f32, grad_loc f32, grad_attn f32)."""
    for t, n in ((value, "value"), (sampling_locations, "sampling_loc"),
                 (attention_weights, "attn_weight"), (grad_output, "grad_output")):
        _need_cuda(t, n)
    if value.dtype not in _DT or grad_output.dtype not in _DT:
        raise RuntimeError("value / grad_output must be float32 or bfloat16")
    loc = sampling_locations if sampling_locations.dtype == torch.float32 else sampling_locations.float()
    attn = attention_weights if attention_weights.dtype == torch.float32 else attention_weights.float()
    B, S, M, D, Q, L, P = _dims(value, loc, attn)
    if tuple(grad_output.shape) != (B, Q, M * D):
        raise RuntimeError("grad_output must be (B, Q, M*D)")
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    if grad_value is None:
        grad_value = torch.zeros(value.shape, device=value.device, dtype=torch.float32)
    elif grad_value.dtype != torch.float32 or tuple(grad_value.shape) != tuple(value.shape):
        raise RuntimeError("grad_value must be float32 with value's shape")
    grad_loc = torch.empty(loc.shape, device=value.device, dtype=torch.float32)
    grad_attn = torch.empty(attn.shape, device=value.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(value.device):
        st = lib.bevf_msda_backward(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                    loc.data_ptr(), attn.data_ptr(), grad_output.data_ptr(),
                                    _DT[grad_output.dtype], grad_value.data_ptr(),
                                    grad_loc.data_ptr(), grad_attn.data_ptr(),
                                    B, S, M, D, Q, L, P, _stream_ptr(value))
    _lib.check(st, lib)
    return grad_value, grad_loc, grad_attn


def _check_im2col(batch: int, im2col_step) -> None:
    # mmcv asserts batch % min(batch, im2col_step) == 0; the step itself is not needed here
    step = min(batch, int(im2col_step)) if batch > 0 else 1
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")


class _MSDAFunction(Function):
    """Shared body of the two reference-named classes below."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        _check_im2col(value.shape[0], im2col_step)
        ctx.im2col_step = im2col_step
        out = msda_forward(value, value_spatial_shapes, value_level_start_index,
                           sampling_locations, attention_weights)
        ctx.save_for_backward(value, torch.as_tensor(value_spatial_shapes),
                              torch.as_tensor(value_level_start_index), sampling_locations,
                              attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, ss, ls, loc, attn = ctx.saved_tensors
        gv, gl, ga = msda_backward(value, ss, ls, loc, attn, grad_output.contiguous())
        return (gv.to(value.dtype), None, None, gl.to(loc.dtype), ga.to(attn.dtype), None)


def _cast_args(args, dtype):
    return tuple(a.to(dtype) if torch.is_tensor(a) and a.is_floating_point() else a for a in args)


class MultiScaleDeformableAttnFunction_fp32(_MSDAFunction):
    """Same contract as the reference class of this name
    (multi_scale_deformable_attn_function.py:90-163): under autocast every floating input is cast
    to fp32 first (``custom_fwd(cast_inputs=torch.float32)``, :93). Outside autocast a bf16 ``value``
    is consumed natively (bf16 storage, fp32 accumulation) -- an extension the reference lacks."""

    @staticmethod
    def forward(ctx, *args):
        if torch.is_autocast_enabled():
            with torch.autocast(device_type="cuda", enabled=False):
                return _MSDAFunction.forward(ctx, *_cast_args(args, torch.float32))
        return _MSDAFunction.forward(ctx, *args)


class MultiScaleDeformableAttnFunction_fp16(_MSDAFunction):
    """Name kept for import compatibility (spatial_cross_attention.py:24-25, decoder.py:27-28).
    The reference never selects it; half inputs are widened to fp32, the only half type the
    library stores being bf16."""

    @staticmethod
    def forward(ctx, *args):
        with torch.autocast(device_type="cuda", enabled=False):
            return _MSDAFunction.forward(ctx, *_cast_args(args, torch.float32))
