"""Python face of the C ABI: argument validation, stream plumbing, autograd glue.

Mirrors the reference's op wrapper
(projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py): same class
names, same ``apply`` signature, same 6-tuple of gradients.  PyTorch is used for device memory and
the current stream only; all arithmetic happens in ``libbevformer_b200.so``.
"""
from __future__ import annotations

import os

import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}

# bench.py's roofline needs the duration of individual launches inside the timed region: when a name
# is present in KERNEL_TIMERS, the wrapper brackets that launch with CUDA events on the launching
# stream and appends (start, end, tag) -- elapsed times are read after the region's final synchronize;
# tag = (rows, levels) of the launch, which tells the SCA launches (4 levels) from the TSA ones (1).
KERNEL_TIMERS: dict = {}


class _timed:
    def __init__(self, name, device, tag=None):
        self.rec = KERNEL_TIMERS.get(name)
        self.device = device
        self.tag = tag

    def __enter__(self):
        if self.rec is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *a):
        if self.rec is not None:
            self.e.record(torch.cuda.current_stream(self.device))
            self.rec.append((self.s, self.e, self.tag))


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
    with torch.cuda.device(value.device), _timed("msda_forward", value.device):
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
    with torch.cuda.device(value.device), _timed("msda_backward", value.device):
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


# =================================================================================================
# Row-list sampler + fused encoder-layer pieces (see include/bevformer_b200.h for what each replaces)
# =================================================================================================
def _i32(t, device):
    return torch.as_tensor(t).to(device=device, dtype=torch.int32).contiguous()


def msda_rows_forward(value, spatial_shapes, level_start_index, loc, attn, row_map, out_dtype=None):
    """value (NB,S,M,D); loc (R,M,L,P,2) f32; attn (R,M,L,P) f32; row_map (R,) int32 -> (R, M*D)."""
    for t, n in ((value, "value"), (loc, "sampling_loc"), (attn, "attn_weight"), (row_map, "row_map")):
        _need_cuda(t, n)
    if value.dtype not in _DT or loc.dtype != torch.float32 or attn.dtype != torch.float32:
        raise RuntimeError("value must be float32/bfloat16, sampling_loc and attn_weight float32")
    if row_map.dtype != torch.int32:
        raise RuntimeError("row_map must be int32")
    NB, S, M, D = value.shape
    R, M2, L, P, _ = loc.shape
    if M2 != M or tuple(attn.shape) != (R, M, L, P) or row_map.numel() != R:
        raise RuntimeError("value / sampling_loc / attn_weight / row_map shapes disagree")
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    out_dtype = out_dtype or value.dtype
    out = torch.empty((R, M * D), device=value.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(value.device), _timed("msda_rows_forward", value.device, (R, L)):
        st = lib.bevf_msda_rows_forward(value.data_ptr(), _DT[value.dtype], ss.data_ptr(),
                                        ls.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                        out.data_ptr(), _DT[out_dtype], row_map.data_ptr(),
                                        NB, S, M, D, R, L, P, _stream_ptr(value))
    _lib.check(st, lib)
    return out


def msda_rows_forward_staged(value, spatial_shapes, level_start_index, level_hw_host, loc, attn, map_range,
                             out_dtype=None):
    """msda_rows_forward for row lists grouped by value map, coarse levels TMA-staged in shared memory
    (bevf_msda_rows_forward_staged).  level_hw_host: [(h, w), ...] python ints; map_range (NB, 2) int32."""
    import ctypes
    for t, n in ((value, "value"), (loc, "sampling_loc"), (attn, "attn_weight"), (map_range, "map_range")):
        _need_cuda(t, n)
    if value.dtype not in _DT or loc.dtype != torch.float32 or attn.dtype != torch.float32:
        raise RuntimeError("value must be float32/bfloat16, sampling_loc and attn_weight float32")
    NB, S, M, D = value.shape
    R, M2, L, P, _ = loc.shape
    if M2 != M or tuple(attn.shape) != (R, M, L, P) or map_range.numel() != 2 * NB or len(level_hw_host) != L:
        raise RuntimeError("value / sampling_loc / attn_weight / map_range / level shapes disagree")
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    hw = (ctypes.c_int32 * (2 * L))(*[int(v) for hw_ in level_hw_host for v in hw_])
    out_dtype = out_dtype or value.dtype
    out = torch.empty((R, M * D), device=value.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(value.device), _timed("msda_rows_forward", value.device, (R, L)):
        st = lib.bevf_msda_rows_forward_staged(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                               ctypes.addressof(hw), loc.data_ptr(), attn.data_ptr(),
                                               out.data_ptr(), _DT[out_dtype], map_range.data_ptr(),
                                               NB, S, M, D, R, L, P, _stream_ptr(value))
    _lib.check(st, lib)
    return out


_DENSE_INIT = [False]


def _dense_mode_init(lib) -> None:
    """BEVF_MSDA_DENSE=2: the dense kernel on the library's second stream (created here, i.e. at the first eager
    backward -- never inside a stream capture); 0 / 1 are read by the library itself."""
    if not _DENSE_INIT[0]:
        _DENSE_INIT[0] = True
        if os.environ.get("BEVF_MSDA_DENSE", "") == "2" and not torch.cuda.is_current_stream_capturing():
            _lib.check(lib.bevf_msda_set_dense_backward(2), lib)


def msda_rows_backward(value, spatial_shapes, level_start_index, loc, attn, row_map, grad_output,
                       grad_value=None, group_order=None, dense=None):
    """``group_order`` (R,) int32: optional permutation of the rows in which runs of 64 entries are
    spatial neighbours on one value map (see bevf_msda_rows_backward_ordered).
    ``dense`` = (level_hw_host, map_range) for row lists grouped by value map: grad_value of the coarse levels
    through the tensor-core kernel (bevf_msda_rows_backward_dense)."""
    for t, n in ((value, "value"), (loc, "sampling_loc"), (attn, "attn_weight"),
                 (row_map, "row_map"), (grad_output, "grad_output")):
        _need_cuda(t, n)
    NB, S, M, D = value.shape
    R, _, L, P, _ = loc.shape
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    if grad_value is None:
        grad_value = torch.zeros(value.shape, device=value.device, dtype=torch.float32)
    grad_loc = torch.empty(loc.shape, device=value.device, dtype=torch.float32)
    grad_attn = torch.empty(attn.shape, device=value.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(value.device), _timed("msda_rows_backward", value.device, (R, L)):
        if group_order is not None and (group_order.dtype != torch.int32 or group_order.numel() != R
                                        or not group_order.is_cuda):
            raise RuntimeError("group_order must be a CUDA int32 tensor with one entry per row")
        if dense is not None and group_order is None:
            import ctypes
            _dense_mode_init(lib)
            level_hw_host, map_range = dense
            _need_cuda(map_range, "map_range")
            if len(level_hw_host) != L or map_range.numel() != 2 * NB or map_range.dtype != torch.int32:
                raise RuntimeError("dense backward: level_hw_host / map_range do not match value and sampling_loc")
            hw = (ctypes.c_int32 * (2 * L))(*[int(v) for hw_ in level_hw_host for v in hw_])
            st = lib.bevf_msda_rows_backward_dense(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                                   ctypes.addressof(hw), loc.data_ptr(), attn.data_ptr(),
                                                   grad_output.data_ptr(), _DT[grad_output.dtype],
                                                   grad_value.data_ptr(), grad_loc.data_ptr(),
                                                   grad_attn.data_ptr(), row_map.data_ptr(), map_range.data_ptr(),
                                                   NB, S, M, D, R, L, P, _stream_ptr(value))
            _lib.check(st, lib)
            return grad_value, grad_loc, grad_attn
        st = lib.bevf_msda_rows_backward_ordered(value.data_ptr(), _DT[value.dtype], ss.data_ptr(),
                                                 ls.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                                 grad_output.data_ptr(), _DT[grad_output.dtype],
                                                 grad_value.data_ptr(), grad_loc.data_ptr(),
                                                 grad_attn.data_ptr(), row_map.data_ptr(), _ptr(group_order),
                                                 NB, S, M, D, R, L, P, _stream_ptr(value))
    _lib.check(st, lib)
    return grad_value, grad_loc, grad_attn


def gv_mode_for(rows_per_map: float, num_points: int, level_hw_host):
    """How SamplerRows' backward accumulates grad_value for a bf16 value tensor (``gv_mode``):
    levels on which a (pixel, head) collects on average at most BEVF_GV_MAXCONTRIB (default 64) contributions
    -- rows_per_map * num_points * 4 / (H * W) -- accumulate in scaled fp16 (half the L2 reduction sectors), the
    others in fp32.  Measured on the base launches against Oracle-S (tests/test_msda_gpu.py): TSA (16 per pixel)
    3.0e-3, SCA level 0 (10) 5.2e-3, level 1 (41) 6.1e-3, level 2 (164) 9.6e-3, level 3 (630) 1.3e-2 of max|grad| --
    hence the cut at 64.  BEVF_GV_ACC=fp32 switches it off.  Returns None, "f16" or ("mixed", shapes, n_fine);
    only a PREFIX of the pyramid can be fp16 (the fine levels come first in every BEVFormer config)."""
    if os.environ.get("BEVF_GV_ACC", "f16") != "f16" or not level_hw_host:
        return None
    cap = float(os.environ.get("BEVF_GV_MAXCONTRIB", "64"))
    shapes = [(int(h), int(w)) for h, w in level_hw_host]
    nfine = 0
    for h, w in shapes:
        if rows_per_map * num_points * 4.0 / (h * w) > cap:
            break
        nfine += 1
    if nfine == 0:
        return None
    return "f16" if nfine == len(shapes) else ("mixed", shapes, nfine)


class LazyGradValue:
    """grad_value of a sampler backward still in accumulator form (scaled fp16 [+ fp32 side buffer]): ``materialize()``
    runs the one conversion pass (bevf_gv16_unscale / bevf_gv_merge) on the CURRENT stream and returns the bf16
    gradient.  Lets the consumer (plugin/linear.py's shared projections) do that pass off the critical path, exactly
    where it converts an fp32 grad_value."""

    def __init__(self, shape, fine, side, amax):
        self.shape, self.fine, self.side, self.amax = tuple(shape), fine, side, amax
        self.tensors = tuple(t for t in (fine, side, amax) if t is not None)
        self.device = fine.device

    def materialize(self) -> torch.Tensor:
        nb, s, m, d = self.shape
        gv = torch.empty(self.shape, device=self.device, dtype=torch.bfloat16)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if self.side is None:
                st = lib.bevf_gv16_unscale(self.fine.data_ptr(), self.amax.data_ptr(), gv.data_ptr(), gv.numel(),
                                           _stream_ptr(gv))
            else:
                st = lib.bevf_gv_merge(self.fine.data_ptr(), self.side.data_ptr(), self.amax.data_ptr(), gv.data_ptr(),
                                       nb, s, self.fine.shape[1], m * d, _stream_ptr(gv))
        _lib.check(st, lib)
        return gv


def abs_max_bits(x: torch.Tensor) -> torch.Tensor:
    """(1,) int32 device word holding the float bits of max|x| (bevf_abs_max): the scale source of the fp16-accumulated
    sampler backward."""
    _need_cuda(x, "x")
    x = x.contiguous()
    out = torch.empty(1, device=x.device, dtype=torch.int32)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_abs_max(x.data_ptr(), _DT[x.dtype], x.numel(), out.data_ptr(), _stream_ptr(x))
    _lib.check(st, lib)
    return out


def msda_rows_backward_f16acc(value, spatial_shapes, level_start_index, loc, attn, row_map, grad_output,
                              group_order=None, lazy=False):
    """Row-list backward with grad_value accumulated in scaled fp16 (bevf_msda_rows_backward_f16acc): half the L2
    reduction sectors of the fp32 path.  Returns (grad_value as bf16 -- or, with ``lazy``, a LazyGradValue --,
    grad_loc, grad_attn)."""
    for t, n in ((value, "value"), (loc, "sampling_loc"), (attn, "attn_weight"), (row_map, "row_map"),
                 (grad_output, "grad_output")):
        _need_cuda(t, n)
    if value.dtype != torch.bfloat16 or value.shape[-1] != 32:
        raise RuntimeError("fp16-accumulated backward: value must be bfloat16 with head_dim 32")
    NB, S, M, D = value.shape
    R, _, L, P, _ = loc.shape
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    grad_output = grad_output.contiguous()
    grad_loc = torch.empty(loc.shape, device=value.device, dtype=torch.float32)
    grad_attn = torch.empty(attn.shape, device=value.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(value.device), _timed("msda_rows_backward", value.device, (R, L)):
        # (the scale source and the zero-fill belong to the op: they are inside the bracket bench.py times)
        amax = abs_max_bits(grad_output)
        gv16 = torch.zeros(value.shape, device=value.device, dtype=torch.float16)
        st = lib.bevf_msda_rows_backward_f16acc(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                                loc.data_ptr(), attn.data_ptr(), grad_output.data_ptr(),
                                                _DT[grad_output.dtype], gv16.data_ptr(), amax.data_ptr(),
                                                grad_loc.data_ptr(), grad_attn.data_ptr(), row_map.data_ptr(),
                                                _ptr(group_order), NB, S, M, D, R, L, P, _stream_ptr(value))
        _lib.check(st, lib)
    gv = LazyGradValue(value.shape, gv16, None, amax)
    return (gv if lazy else gv.materialize()), grad_loc, grad_attn


def msda_rows_backward_mixed(value, spatial_shapes, level_start_index, level_hw_host, num_f16_levels, loc, attn,
                             row_map, grad_output, group_order=None, lazy=False):
    """Row-list backward with MIXED accumulation (bevf_msda_rows_backward_mixed): the first ``num_f16_levels`` levels
    in scaled fp16, the others in fp32 into a side buffer that only spans their pixels; one merge pass produces the
    bf16 gradient.  ``level_hw_host``: [(h, w), ...] python ints that MUST equal the device spatial_shapes."""
    import ctypes
    for t, n in ((value, "value"), (loc, "sampling_loc"), (attn, "attn_weight"), (row_map, "row_map"),
                 (grad_output, "grad_output")):
        _need_cuda(t, n)
    if value.dtype != torch.bfloat16 or value.shape[-1] != 32:
        raise RuntimeError("mixed-accumulation backward: value must be bfloat16 with head_dim 32")
    NB, S, M, D = value.shape
    R, _, L, P, _ = loc.shape
    if len(level_hw_host) != L or not (1 <= num_f16_levels < L):
        raise RuntimeError("mixed-accumulation backward: level_hw_host / num_f16_levels do not fit the pyramid")
    ss, ls = _level_tensors(value, spatial_shapes, level_start_index)
    s_fine = sum(int(h) * int(w) for h, w in level_hw_host[:num_f16_levels])
    grad_output = grad_output.contiguous()
    grad_loc = torch.empty(loc.shape, device=value.device, dtype=torch.float32)
    grad_attn = torch.empty(attn.shape, device=value.device, dtype=torch.float32)
    hw = (ctypes.c_int32 * (2 * L))(*[int(v) for hw_ in level_hw_host for v in hw_])
    lib = _lib.load()
    with torch.cuda.device(value.device), _timed("msda_rows_backward", value.device, (R, L)):
        amax = abs_max_bits(grad_output)
        fine = torch.zeros((NB, s_fine, M, D), device=value.device, dtype=torch.float16)
        side = torch.zeros((NB, S - s_fine, M, D), device=value.device, dtype=torch.float32)
        st = lib.bevf_msda_rows_backward_mixed(value.data_ptr(), _DT[value.dtype], ss.data_ptr(), ls.data_ptr(),
                                               ctypes.addressof(hw), loc.data_ptr(), attn.data_ptr(),
                                               grad_output.data_ptr(), _DT[grad_output.dtype], fine.data_ptr(),
                                               side.data_ptr(), amax.data_ptr(), int(num_f16_levels),
                                               grad_loc.data_ptr(), grad_attn.data_ptr(), row_map.data_ptr(),
                                               _ptr(group_order), NB, S, M, D, R, L, P, _stream_ptr(value))
        _lib.check(st, lib)
    gv = LazyGradValue(value.shape, fine, side, amax)
    return (gv if lazy else gv.materialize()), grad_loc, grad_attn


# Second stream for work that is off the critical path (weight gradients, zero-fills and projections that
# are needed later): set by BEVFormerEncoder.enable_grad_arena(overlap=True); None = everything in order.
AUX_STREAM: dict = {}


def aux_stream(device):
    return AUX_STREAM.get(torch.device(device))


class SamplerRows(Function):
    """Sampler over a compact list of query rows (SCA's in-view (camera, query) pairs)."""

    @staticmethod
    def forward(ctx, value, loc, attn, row_map, spatial_shapes, level_start_index, group_order=None,
                staged=None, gv_mode=None):
        """``staged`` = (level_hw_host, map_range): use the TMA-staged forward (rows grouped by value map)."""
        if value.dtype == torch.float16:      # the reference widens half inputs (…function.py:93)
            value = value.float()
        # measured slower than the plain kernel on B200 (profiles/README.md, r2g): opt-in for A/B runs
        if staged is not None and value.shape[-1] == 32 and os.environ.get("BEVF_MSDA_FWD", "plain") == "staged":
            out = msda_rows_forward_staged(value, spatial_shapes, level_start_index, staged[0], loc, attn,
                                           staged[1])
        else:
            out = msda_rows_forward(value, spatial_shapes, level_start_index, loc, attn, row_map)
        ctx.save_for_backward(value, loc, attn, row_map, spatial_shapes, level_start_index)
        ctx.group_order = group_order
        # grad_value accumulated in scaled fp16 -- "f16": every level, ("mixed", level_hw_host, n): the first n levels
        # -- (half the L2 reduction sectors of those levels): only where the caller asks for it
        ctx.gv_mode = gv_mode if (gv_mode is not None and value.dtype == torch.bfloat16 and value.shape[-1] == 32) else None
        ctx.dense = staged if (staged is not None and value.shape[-1] == 32) else None
        ctx.value_early = getattr(value, "_bevf_early", None)     # see plugin/linear.py::shared_input_projections
        ctx.gv_zero = None
        aux = aux_stream(value.device)
        if (aux is not None and value.requires_grad and torch.is_grad_enabled()
                and os.environ.get("BEVF_AUX_FILL", "0") == "1"):
            # the backward accumulates grad_value into a zero-filled fp32 buffer: fill it NOW on the second
            # stream (it overlaps the forward) instead of on the backward's critical path.  Measured neutral on
            # B200 (15.70 vs 15.73 ms) and it holds 1.6 GB from forward to backward: opt-in (BEVF_AUX_FILL=1)
            main = torch.cuda.current_stream(value.device)
            ev = torch.cuda.Event()
            ev.record(main)
            aux.wait_event(ev)
            with torch.cuda.stream(aux):
                gv0 = torch.zeros(value.shape, device=value.device, dtype=torch.float32)
                done = torch.cuda.Event()
                done.record(aux)
            gv0.record_stream(main)
            ctx.gv_zero = (gv0, done)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        value, loc, attn, row_map, ss, ls = ctx.saved_tensors
        gv0 = None
        if ctx.gv_zero is not None:
            gv0, done = ctx.gv_zero
            torch.cuda.current_stream(value.device).wait_event(done)
        dense_on = ctx.dense is not None and _lib.load().bevf_msda_get_dense_backward() != 0   # explicit opt-in wins
        if ctx.gv_mode is not None and gv0 is None and grad_out.dtype == torch.bfloat16 and not dense_on:
            if ctx.gv_mode == "f16":
                gv, gl, ga = msda_rows_backward_f16acc(value, ss, ls, loc, attn, row_map, grad_out, ctx.group_order,
                                                       lazy=True)
            else:
                _, hw_host, nfine = ctx.gv_mode
                if sum(h * w for h, w in hw_host) != value.shape[1]:
                    # a stale host copy of the pyramid (it no longer adds up to S): plain fp32 accumulation instead
                    gv, gl, ga = msda_rows_backward(value, ss, ls, loc, attn, row_map, grad_out.contiguous(), None,
                                                    group_order=ctx.group_order)
                    if ctx.value_early is not None and ctx.value_early(gv):
                        return None, gl, ga, None, None, None, None, None, None
                    return gv.to(value.dtype), gl, ga, None, None, None, None, None, None
                gv, gl, ga = msda_rows_backward_mixed(value, ss, ls, hw_host, nfine, loc, attn, row_map, grad_out,
                                                      ctx.group_order, lazy=True)
            if ctx.value_early is not None and ctx.value_early(gv):
                return None, gl, ga, None, None, None, None, None, None      # the producer converts it off the critical path
            return gv.materialize(), gl, ga, None, None, None, None, None, None
        gv, gl, ga = msda_rows_backward(value, ss, ls, loc, attn, row_map, grad_out.contiguous(), gv0,
                                        group_order=ctx.group_order, dense=ctx.dense)
        if ctx.value_early is not None and ctx.value_early(gv):
            # the producer of `value` took the gradient (conversion + its GEMMs run off the critical path)
            return None, gl, ga, None, None, None, None, None, None
        return gv.to(value.dtype), gl, ga, None, None, None, None, None, None


def sca_prep_forward(raw, ref_cam, pair_q, pair_cam, level_hw, B, Nq, M, L, P):
    _need_cuda(raw, "raw")
    R, D, ncam = pair_q.numel(), ref_cam.shape[3], ref_cam.shape[0]
    loc = torch.empty((B * R, M, L, P, 2), device=raw.device, dtype=torch.float32)
    attn = torch.empty((B * R, M, L, P), device=raw.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        st = lib.bevf_sca_prep_forward(raw.data_ptr(), ref_cam.data_ptr(), pair_q.data_ptr(),
                                       pair_cam.data_ptr(), level_hw.data_ptr(), loc.data_ptr(),
                                       attn.data_ptr(), B, Nq, R, M, L, P, D, ncam, _stream_ptr(raw))
    _lib.check(st, lib)
    return loc, attn


def sca_prep_backward(raw, grad_loc, grad_attn, pair_of, level_hw, B, Nq, R, M, L, P,
                      out_dtype=torch.float32):
    """d_raw of the SCA sampling-point prep; out_dtype=bfloat16 rounds in the kernel (the result then
    feeds the bf16 GEMMs of the head without a cast pass)."""
    ncam = pair_of.shape[0]
    d_raw = torch.empty(raw.shape, device=raw.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        st = lib.bevf_sca_prep_backward(raw.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
                                        pair_of.data_ptr(), level_hw.data_ptr(), d_raw.data_ptr(),
                                        _DT[out_dtype], B, Nq, R, M, L, P, ncam, _stream_ptr(raw))
    _lib.check(st, lib)
    return d_raw


class ScaPrep(Function):
    @staticmethod
    def forward(ctx, raw, ref_cam, pair_q, pair_cam, pair_of, level_hw, B, Nq, M, L, P):
        loc, attn = sca_prep_forward(raw, ref_cam, pair_q, pair_cam, level_hw, B, Nq, M, L, P)
        ctx.save_for_backward(raw, pair_of, level_hw)
        ctx.dims = (B, Nq, pair_q.numel(), M, L, P)
        return loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        raw, pair_of, level_hw = ctx.saved_tensors
        d_raw = sca_prep_backward(raw, grad_loc.contiguous(), grad_attn.contiguous(), pair_of,
                                  level_hw, *ctx.dims)
        return (d_raw,) + (None,) * 10


def tsa_prep_forward(raw, ref2d, level_hw, B, Nq, M, L, P, interleave=False):
    _need_cuda(raw, "raw")
    shape = (B * Nq * 2, M, L, P) if interleave else (B * 2, Nq, M, L, P)
    loc = torch.empty(shape + (2,), device=raw.device, dtype=torch.float32)
    attn = torch.empty(shape, device=raw.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        st = lib.bevf_tsa_prep_forward(raw.data_ptr(), ref2d.data_ptr(), level_hw.data_ptr(),
                                       loc.data_ptr(), attn.data_ptr(), B, Nq, M, L, P,
                                       int(interleave), _stream_ptr(raw))
    _lib.check(st, lib)
    return loc, attn


def tsa_prep_backward(raw, grad_loc, grad_attn, level_hw, B, Nq, M, L, P, interleave=0,
                      out_dtype=torch.float32):
    d_raw = torch.empty(raw.shape, device=raw.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        st = lib.bevf_tsa_prep_backward(raw.data_ptr(), grad_loc.contiguous().data_ptr(),
                                        grad_attn.contiguous().data_ptr(), level_hw.data_ptr(),
                                        d_raw.data_ptr(), _DT[out_dtype], B, Nq, M, L, P,
                                        int(interleave), _stream_ptr(raw))
    _lib.check(st, lib)
    return d_raw


class TsaPrep(Function):
    @staticmethod
    def forward(ctx, raw, ref2d, level_hw, B, Nq, M, L, P, interleave=False):
        loc, attn = tsa_prep_forward(raw, ref2d, level_hw, B, Nq, M, L, P, interleave)
        ctx.save_for_backward(raw, level_hw)
        ctx.dims = (B, Nq, M, L, P, int(interleave))
        return loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        raw, level_hw = ctx.saved_tensors
        d_raw = tsa_prep_backward(raw, grad_loc, grad_attn, level_hw, *ctx.dims)
        return (d_raw,) + (None,) * 8


def _ptr(t):
    return 0 if t is None else t.data_ptr()


_seed_counter = [0]
_seed_state: dict = {}        # device -> int64[1] step counter living on the device
_seed_snap: dict = {}         # device -> int64[1] copy of the counter taken by the last advance_seed()


def seed_state(device) -> torch.Tensor:
    """The dropout step key the kernels of the CURRENT forward read: the snapshot taken by the last
    ``advance_seed`` (the persistent counter itself before the first one).  Autograd nodes keep the
    tensor they were given, so a backward always regenerates the masks of its own forward even when
    other training-mode forwards ran in between."""
    key = torch.device(device)
    if key in _seed_snap:
        return _seed_snap[key]
    if key not in _seed_state:
        _seed_state[key] = torch.zeros(1, dtype=torch.int64, device=key)
    return _seed_state[key]


def advance_seed(device) -> None:
    """Bump the device-side dropout step counter and snapshot it (two tiny kernels; capturable in a
    CUDA graph: the in-place add lives on a persistent tensor, so every replay of a captured training
    step draws new masks, and the snapshot is recomputed by the replay)."""
    key = torch.device(device)
    if key not in _seed_state:
        _seed_state[key] = torch.zeros(1, dtype=torch.int64, device=key)
    _seed_state[key].add_(0x9E3779B97F4A7C1)
    _seed_snap[key] = _seed_state[key].clone()


def _next_seed() -> int:
    """A fresh 63-bit Philox key per dropout site, derived from torch's global seed (so
    torch.manual_seed makes runs repeatable) without touching the device."""
    _seed_counter[0] += 1
    return (torch.initial_seed() * 6364136223846793005 + _seed_counter[0] * 1442695040888963407) & ((1 << 63) - 1)


def _param_ptr(p, act_dtype):
    """Parameters are read in their own dtype when the kernel supports the combination."""
    if p.dtype == act_dtype or p.dtype == torch.float32:
        q = p.detach().contiguous()
    else:
        q = p.detach().float().contiguous()
    return q, _DT[q.dtype]


class LayerNormResidual(Function):
    """y = LayerNorm(dropout_p(x) + residual) * gamma + beta (fp32 statistics); residual may be None.
    The dropout keep-mask is regenerated from a Philox seed in backward (nothing stored).
    With ``pos`` the kernel also writes y + pos (the next block's positional-encoded query) and the
    call returns (y, y + pos); their two gradients are summed inside the backward kernel.
    With ``twin`` the call returns (y, y') where y' aliases y: the caller feeds y to the next block and y'
    to that block's residual connection, so their two gradients arrive separately and are summed inside
    the backward kernel too (autograd would otherwise add them with a three-pass elementwise kernel)."""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, eps, drop_p=0.0, pos=None, twin=False):
        _need_cuda(x, "x")
        if x.dtype not in _DT:
            raise RuntimeError("layernorm: float32 or bfloat16 only")
        C = x.shape[-1]
        rows = x.numel() // C
        rc = None if residual is None else residual.contiguous()
        g, pd = _param_ptr(gamma, x.dtype)
        b, pd2 = _param_ptr(beta, x.dtype)
        if pd != pd2:
            g, b, pd = g.float(), b.float(), F32
        y = torch.empty_like(x)
        y2 = None
        if pos is not None:
            if pos.dtype != x.dtype or pos.numel() != x.numel():
                raise RuntimeError("layernorm: pos must match x in dtype and size")
            pos = pos.contiguous()
            y2 = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        seed = _next_seed() if drop_p > 0.0 else 0
        sbase = seed_state(x.device) if drop_p > 0.0 else None
        lib = _lib.load()
        with torch.cuda.device(x.device):
            st = lib.bevf_layernorm_forward(x.data_ptr(), _ptr(rc), g.data_ptr(), b.data_ptr(), pd,
                                            _ptr(pos), y.data_ptr(), _ptr(y2), mean.data_ptr(),
                                            rstd.data_ptr(), rows, C, float(eps), float(drop_p), seed,
                                            _ptr(sbase), _DT[x.dtype], _stream_ptr(x))
        _lib.check(st, lib)
        ctx.save_for_backward(x, rc, g, mean, rstd)
        from .arena import arena_of
        ctx.arena, ctx.arena_accs = arena_of(gamma, beta)            # flat gradient arena, when enabled
        ctx.arena_params = (gamma, beta) if ctx.arena is not None else None
        ctx.sbase = sbase                 # the step key of THIS forward (see seed_state)
        ctx.pos_shape = None if pos is None else tuple(pos.shape)
        ctx.meta = (residual is not None, gamma.dtype, beta.dtype, pd, float(drop_p), seed, pos is not None)
        if pos is None and twin:
            return y, y.view(y.shape)
        return y if pos is None else (y, y2)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dy2=None):
        x, res, g, mean, rstd = ctx.saved_tensors
        has_res, gdt, bdt, pd, drop_p, seed, has_pos = ctx.meta
        C = x.shape[-1]
        rows = x.numel() // C
        # d(y + pos)/d pos = 1: pos (the learned BEV positional encoding) receives the gradient of y2
        d_pos = None
        if has_pos and ctx.needs_input_grad[6] and dy2 is not None:
            d_pos = dy2.reshape(ctx.pos_shape)
        if dy is None:                                   # only y + pos was used downstream
            dy, dy2 = dy2, None
        dy = dy.contiguous()
        ld2 = 0
        if dy2 is not None:
            # usually a column slice of the gradient of cat([prev_bev, query + pos]): rows at a
            # uniform stride are read in place
            r2 = dy2.reshape(-1, C) if dy2.is_contiguous() else dy2
            uniform = (r2.stride(-1) == 1 and r2.dim() >= 2 and r2.stride(-2) >= C
                       and all(r2.stride(i) == r2.stride(i + 1) * r2.shape[i + 1] for i in range(r2.dim() - 2))
                       and r2.stride(-2) % 8 == 0 and r2.data_ptr() % 16 == 0)
            if uniform:
                ld2 = int(r2.stride(-2))
                dy2 = r2
            else:
                dy2 = dy2.contiguous()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (has_res and drop_p > 0.0) else None
        if ctx.arena is not None:
            ctx.arena.touch(*ctx.arena_params)
            dgb = ctx.arena_accs                                    # accumulate straight into the arena
        else:
            dgb = torch.zeros(2, C, device=x.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            st = lib.bevf_layernorm_backward(x.data_ptr(), _ptr(res), g.data_ptr(), pd, mean.data_ptr(),
                                             rstd.data_ptr(), dy.data_ptr(), _ptr(dy2), ld2, dx.data_ptr(),
                                             _ptr(dres), dgb[0].data_ptr(), dgb[1].data_ptr(), rows, C,
                                             drop_p, seed,
                                             _ptr(ctx.sbase) if drop_p > 0.0 else 0,
                                             _DT[x.dtype], _stream_ptr(x))
        _lib.check(st, lib)
        d_res = None if not has_res else (dres if dres is not None else dx)
        if ctx.arena is not None:
            return dx, d_res, None, None, None, None, d_pos, None
        return dx, d_res, dgb[0].to(gdt), dgb[1].to(bdt), None, None, d_pos, None


class ScaCombine(Function):
    """Per-query mean over the cameras that see it (spatial_cross_attention.py:165-172)."""

    @staticmethod
    def forward(ctx, out, pair_of, pair_q, inv_count, B, Nq):
        _need_cuda(out, "out")
        C = out.shape[-1]
        R = pair_q.numel()
        ncam = pair_of.shape[0]
        slots = torch.empty((B, Nq, C), device=out.device, dtype=out.dtype)
        lib = _lib.load()
        with torch.cuda.device(out.device):
            st = lib.bevf_sca_combine_forward(out.data_ptr(), pair_of.data_ptr(),
                                              inv_count.data_ptr(), slots.data_ptr(), B, Nq, R, C,
                                              ncam, _DT[out.dtype], _stream_ptr(out))
        _lib.check(st, lib)
        ctx.save_for_backward(pair_q, inv_count)
        ctx.dims = (B, Nq, R, C)
        return slots

    @staticmethod
    @once_differentiable
    def backward(ctx, g_slots):
        pair_q, inv_count = ctx.saved_tensors
        B, Nq, R, C = ctx.dims
        g_slots = g_slots.contiguous()
        g_out = torch.empty((B * R, C), device=g_slots.device, dtype=g_slots.dtype)
        lib = _lib.load()
        with torch.cuda.device(g_slots.device):
            st = lib.bevf_sca_combine_backward(g_slots.data_ptr(), pair_q.data_ptr(),
                                               inv_count.data_ptr(), g_out.data_ptr(), B, Nq, R, C,
                                               _DT[g_slots.dtype], _stream_ptr(g_slots))
        _lib.check(st, lib)
        return g_out, None, None, None, None, None


def sca_plan_build(mask_u8, qorder, capacity):
    """Device-side pair list (bevf_sca_plan_build): mask_u8 (ncam, B, Nq, D) uint8, qorder (Nq,) int32 or
    None -> dict of the plan tensors; no host synchronisation."""
    _need_cuda(mask_u8, "bev_mask")
    if mask_u8.dtype != torch.uint8:
        raise RuntimeError("bev_mask must be uint8")
    ncam, B, Nq, D = mask_u8.shape
    dev = mask_u8.device
    lib = _lib.load()
    i32 = dict(device=dev, dtype=torch.int32)
    out = dict(pair_q=torch.empty(capacity, **i32), pair_cam=torch.empty(capacity, **i32),
               pair_of=torch.empty((ncam, Nq), **i32), row_map=torch.empty(B * capacity, **i32),
               inv_count=torch.empty((B, Nq), device=dev, dtype=torch.float32),
               map_range=torch.empty((B * ncam, 2), **i32), counters=torch.empty(2, **i32))
    ws = torch.empty(int(lib.bevf_sca_plan_workspace_ints(ncam, Nq)), **i32)
    with torch.cuda.device(dev):
        st = lib.bevf_sca_plan_build(mask_u8.data_ptr(), _ptr(qorder), out["pair_q"].data_ptr(),
                                     out["pair_cam"].data_ptr(), out["pair_of"].data_ptr(),
                                     out["row_map"].data_ptr(), out["inv_count"].data_ptr(),
                                     out["map_range"].data_ptr(), out["counters"].data_ptr(), ws.data_ptr(),
                                     B, ncam, Nq, D,
                                     int(capacity), _stream_ptr(mask_u8))
    _lib.check(st, lib)
    return out


def point_sampling(lidar2img, pc_range, z_norm, img_h, img_w, bev_h, bev_w, raw_mask=False):
    """lidar2img (B, ncam, 4, 4) f32 CUDA -> ref_cam (ncam,B,Nq,D,2) f32, bev_mask (ncam,B,Nq,D) bool
    (uint8 with ``raw_mask``: what the device-side plan builder reads)."""
    _need_cuda(lidar2img, "lidar2img")
    import ctypes
    B, ncam = lidar2img.shape[:2]
    D = len(z_norm)
    Nq = bev_h * bev_w
    ref_cam = torch.empty((ncam, B, Nq, D, 2), device=lidar2img.device, dtype=torch.float32)
    mask = torch.empty((ncam, B, Nq, D), device=lidar2img.device, dtype=torch.uint8)
    pc = (ctypes.c_float * 6)(*[float(v) for v in pc_range])
    zs = (ctypes.c_float * D)(*[float(v) for v in z_norm])
    lib = _lib.load()
    with torch.cuda.device(lidar2img.device):
        st = lib.bevf_point_sampling(lidar2img.data_ptr(), ctypes.addressof(pc), ctypes.addressof(zs),
                                     float(img_h), float(img_w), ref_cam.data_ptr(),
                                     mask.data_ptr(), B, ncam, bev_h, bev_w, D,
                                     _stream_ptr(lidar2img))
    _lib.check(st, lib)
    return ref_cam, (mask if raw_mask else mask.bool())


def linear_tc(x, weight, bias=None, residual=None, relu=False, out_dtype=None):
    """y = act(x @ weight.T + bias) (+ residual) on the tcgen05 GEMM.  x (..., K) bf16, weight (N, K)
    bf16, bias (N) any float dtype, residual (..., N) bf16; returns (..., N) bf16 or fp32."""
    _need_cuda(x, "x")
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise RuntimeError("linear_tc: x and weight must be bfloat16")
    K = x.shape[-1]
    N = weight.shape[0]
    M = x.numel() // K
    w = weight.contiguous()
    bq, bdt = (None, F32) if bias is None else _param_ptr(bias, torch.bfloat16)
    res = None if residual is None else residual.contiguous()
    out_dtype = out_dtype or torch.bfloat16
    y = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=out_dtype)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_linear_forward(x.data_ptr(), w.data_ptr(), _ptr(bq), bdt, _ptr(res), y.data_ptr(),
                                     _DT[out_dtype], M, N, K, int(bool(relu)), _stream_ptr(x))
    _lib.check(st, lib)
    return y


def linear_dgrad_tc(dy, weight, addend=None):
    """dx = dy @ weight (+ addend) on the tcgen05 GEMM, the weight (N, K) read in place (no transposed
    copy).  dy (M, N) bf16 -> (M, K) bf16; ``addend`` (M, K) bf16 is summed in the epilogue.  N or K not
    a multiple of 64: falls back to the transposed-copy form."""
    _need_cuda(dy, "dy")
    if dy.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16 or dy.shape[-1] != weight.shape[0]:
        raise RuntimeError("linear_dgrad_tc: dy (M,N) and weight (N,K) must be bfloat16")
    N, K = weight.shape
    if N % 64 or K % 64 or os.environ.get("BEVF_DGRAD", "mn") == "copy":
        dx = linear_tc(dy, weight.t().contiguous())
        return dx if addend is None else dx + addend.view(dx.shape)
    dy = dy.contiguous()
    w = weight.contiguous()
    M = dy.numel() // N
    dx = torch.empty(dy.shape[:-1] + (K,), device=dy.device, dtype=torch.bfloat16)
    lib = _lib.load()
    with torch.cuda.device(dy.device):
        if addend is None:
            st = lib.bevf_linear_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), M, N, K, _stream_ptr(dy))
        else:
            if addend.dtype != torch.bfloat16 or addend.numel() != M * K or not addend.is_contiguous():
                raise RuntimeError("linear_dgrad_tc: addend must be a contiguous bf16 (M, K) tensor")
            st = lib.bevf_linear_dgrad_acc(dy.data_ptr(), w.data_ptr(), addend.data_ptr(), dx.data_ptr(), M, N, K,
                                           _stream_ptr(dy))
    _lib.check(st, lib)
    return dx


def linear_wgrad_tc(dy, x, with_bias=False, out_dtype=None):
    """dW = dy^T @ x on the tcgen05 split-M kernel (and db = column sums of dy from the same pass).
    dy (M, N) bf16, x (M, K) bf16 -> (N, K) fp32 [, (N,) fp32].  The kernel accumulates in one fp32
    buffer holding [dW | db]; ``out_dtype`` converts that buffer once (dW and db are views of it)."""
    _need_cuda(dy, "dy")
    _need_cuda(x, "x")
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dy.shape[0] != x.shape[0]:
        raise RuntimeError("linear_wgrad_tc: dy (M,N) and x (M,K) must be bfloat16 with equal M")
    M, N = dy.shape
    K = x.shape[1]
    npad = (N + 3) // 4 * 4
    buf = torch.zeros(N * K + (npad if with_bias else 0), device=x.device, dtype=torch.float32)
    dw = buf[: N * K].view(N, K)
    db = buf[N * K: N * K + N] if with_bias else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), _ptr(db), M, N, K,
                                   _stream_ptr(x))
    _lib.check(st, lib)
    if out_dtype is not None and out_dtype != torch.float32:
        buf = buf.to(out_dtype)
        dw = buf[: N * K].view(N, K)
        db = buf[N * K: N * K + N] if with_bias else None
    return (dw, db) if with_bias else dw


def linear_wgrad_into(dy, x, dw_acc, db_acc=None):
    """dW += dy^T @ x (and db += column sums of dy) accumulated into caller-provided fp32 buffers (the
    gradient arena): no allocation, no zero-fill, no conversion here."""
    _need_cuda(dy, "dy")
    _need_cuda(x, "x")
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dy.shape[0] != x.shape[0]:
        raise RuntimeError("linear_wgrad_into: dy (M,N) and x (M,K) must be bfloat16 with equal M")
    M, N = dy.shape
    K = x.shape[1]
    if dw_acc.dtype != torch.float32 or dw_acc.numel() != N * K or not dw_acc.is_contiguous():
        raise RuntimeError("linear_wgrad_into: dw_acc must be a contiguous fp32 (N, K) buffer")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw_acc.data_ptr(), _ptr(db_acc), M, N, K,
                                   _stream_ptr(x))
    _lib.check(st, lib)


def linear_wgrad_out(dy, x, grad_dtype, with_bias):
    """Two-pass weight (+ bias) gradient written directly in ``grad_dtype``: returns (dW, db|None)."""
    _need_cuda(dy, "dy")
    _need_cuda(x, "x")
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dy.shape[0] != x.shape[0]:
        raise RuntimeError("linear_wgrad_out: dy (M,N) and x (M,K) must be bfloat16 with equal M")
    M, N = dy.shape
    K = x.shape[1]
    lib = _lib.load()
    with torch.cuda.device(x.device):
        need = int(lib.bevf_linear_wgrad_workspace_bytes(M, N, K))
        ws = torch.empty(need, device=x.device, dtype=torch.uint8)
        dw = torch.empty((N, K), device=x.device, dtype=grad_dtype)
        db = torch.empty((N,), device=x.device, dtype=grad_dtype) if with_bias else None
        st = lib.bevf_linear_wgrad_out(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), _ptr(db),
                                       _DT[grad_dtype], ws.data_ptr(), need, M, N, K, _stream_ptr(x))
    _lib.check(st, lib)
    return dw, db


def sum_tensors(ts):
    """Sum of up to 8 equally-shaped contiguous bf16 / f32 CUDA tensors in one pass (fp32 accumulation)."""
    import ctypes
    ts = [t.contiguous() for t in ts]
    if len(ts) == 1:
        return ts[0]
    t0 = ts[0]
    _need_cuda(t0, "tensors[0]")
    if len(ts) > 8 or any(t.shape != t0.shape or t.dtype != t0.dtype for t in ts) or t0.dtype not in _DT:
        raise RuntimeError("sum_tensors: 1..8 tensors of one shape, float32 or bfloat16")
    out = torch.empty_like(t0)
    arr = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    lib = _lib.load()
    with torch.cuda.device(t0.device):
        st = lib.bevf_sum_tensors(ctypes.addressof(arr), len(ts), out.data_ptr(), t0.numel(), _DT[t0.dtype],
                                  _stream_ptr(t0))
    _lib.check(st, lib)
    return out


def colsum(x):
    """fp32 column sums of a (rows, C) tensor (bias gradients)."""
    _need_cuda(x, "x")
    C = x.shape[-1]
    rows = x.numel() // C
    out = torch.zeros(C, device=x.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_colsum(x.data_ptr(), out.data_ptr(), rows, C, _DT[x.dtype], _stream_ptr(x))
    _lib.check(st, lib)
    return out


def dropout_inplace_(x, p):
    """x *= keep / (1 - p) in place with Philox bits keyed by the device-side step counter."""
    _need_cuda(x, "x")
    if p <= 0.0:
        return x
    lib = _lib.load()
    with torch.cuda.device(x.device):
        st = lib.bevf_dropout_inplace(x.data_ptr(), x.numel(), float(p), _next_seed(),
                                      seed_state(x.device).data_ptr(), _DT[x.dtype], _stream_ptr(x))
    _lib.check(st, lib)
    return x


def relu_dropout_backward(dy, h, p):
    """Gradient w.r.t. z of h = dropout_p(relu(z)), from the saved h alone."""
    _need_cuda(dy, "dy")
    dy = dy.contiguous()
    out = torch.empty_like(dy)
    lib = _lib.load()
    with torch.cuda.device(dy.device):
        st = lib.bevf_relu_dropout_backward(dy.data_ptr(), h.data_ptr(), out.data_ptr(), dy.numel(),
                                            1.0 / (1.0 - p), _DT[dy.dtype], _stream_ptr(dy))
    _lib.check(st, lib)
    return out


class FlattenFeats(Function):
    """Multi-level camera features [(bs, ncam, C, h, w)] -> (ncam, S, bs, C) with cams_embeds and
    level_embeds added (PerceptionTransformer.get_bev_features, transformer.py:161-181): one kernel
    per level instead of flatten / permute / add / add / cat / permute.  Call as
    ``FlattenFeats.apply(cams_embeds_or_None, level_embeds, *mlvl_feats)``."""

    @staticmethod
    def forward(ctx, cams_embeds, level_embeds, *feats):
        f0 = feats[0]
        _need_cuda(f0, "mlvl_feats[0]")
        if f0.dtype not in _DT:
            raise RuntimeError("flatten_feats: float32 or bfloat16 only")
        bs, ncam, C = f0.shape[:3]
        hws = [int(f.shape[3] * f.shape[4]) for f in feats]
        S = sum(hws)
        out = torch.empty((ncam, S, bs, C), device=f0.device, dtype=f0.dtype)
        ce = None if cams_embeds is None else cams_embeds.detach().float().contiguous()
        le = level_embeds.detach().float().contiguous()
        lib = _lib.load()
        start = 0
        with torch.cuda.device(f0.device):
            for lvl, f in enumerate(feats):
                if f.dtype != f0.dtype or tuple(f.shape[:3]) != (bs, ncam, C):
                    raise RuntimeError("flatten_feats: levels disagree in dtype / (bs, ncam, C)")
                fc = f.contiguous()
                st = lib.bevf_flatten_feats(fc.data_ptr(), _ptr(ce), le[lvl].data_ptr(), out.data_ptr(),
                                            bs, ncam, C, hws[lvl], S, start, _DT[f0.dtype],
                                            _stream_ptr(f0))
                _lib.check(st, lib)
                start += hws[lvl]
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.meta = (None if cams_embeds is None else cams_embeds.dtype, level_embeds.dtype)
        ctx.level_shape = tuple(level_embeds.shape)       # may hold more rows than there are levels
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        # data movement only (the transposes back to NCHW and three small reductions): torch ops
        cdt, ldt = ctx.meta
        need = ctx.needs_input_grad
        d_cams = dy.sum((1, 2), dtype=torch.float32).to(cdt) if (cdt is not None and need[0]) else None
        # the reference slices level_embeds[lvl:lvl+1] (transformer.py:172): rows beyond the pyramid's
        # levels (tiny / small: 1 level, num_feature_levels = 4) simply get a zero gradient
        d_level = torch.zeros(ctx.level_shape, device=dy.device, dtype=torch.float32) if need[1] else None
        d_feats, start = [], 0
        for i, (bs, ncam, C, h, w) in enumerate(ctx.shapes):
            sl = dy[:, start:start + h * w]                          # (ncam, hw, bs, C)
            if need[1]:
                d_level[i] = sl.sum((0, 1, 2), dtype=torch.float32)
            d_feats.append(sl.permute(2, 0, 3, 1).reshape(bs, ncam, C, h, w) if need[2 + i] else None)
            start += h * w
        if need[1]:
            d_level = d_level.to(ldt)
        return (d_cams, d_level, *d_feats)
