"""Flat gradient arena for the encoder's parameters (opt-in: ``BEVFormerEncoder.enable_grad_arena()``).

Without it every weight-gradient launch allocates and zero-fills its own fp32 accumulator and converts it
to the parameter dtype afterwards: 66 fills + 84 casts of 2-5 us per training step at base, a third of the
element-wise glue.  With it all parameter gradients of the encoder accumulate in ONE flat fp32 buffer that
is zeroed once per backward pass (one memset) and converted once at its end (one kernel); ``p.grad`` of
every parameter is a view of the flat result -- which is also exactly the bucket the data-parallel
all-reduce wants (``dist.average_gradients_flat`` then needs no packing).  The same idea as the contiguous
gradient buffers of large-scale trainers.

Semantics to know: gradients are OVERWRITTEN by every backward pass (as after ``zero_grad``), not added to
what ``p.grad`` held; the conversion runs in an end-of-backward callback of the autograd engine, so reading
``p.grad`` inside a backward hook of the same pass sees the previous values.
"""
from __future__ import annotations

from typing import Dict, List

import torch


class GradArena:
    def __init__(self, groups: List[List[torch.nn.Parameter]]):
        """groups: lists of parameters that must lie adjacently, in order (e.g. the two weights that the
        attention modules project with as one stacked matrix)."""
        params = [p for g in groups for p in g]
        if not params:
            raise ValueError("GradArena: no parameters")
        dev = params[0].device
        self.offset: Dict[int, int] = {}
        off = 0
        for g in groups:
            for i, p in enumerate(g):
                if p.device != dev:
                    raise ValueError("GradArena: parameters on several devices")
                if len(g) > 1 and p.numel() % 4 and i + 1 < len(g):
                    raise ValueError("GradArena: members of an adjacency group must be multiples of 4 elements")
                self.offset[id(p)] = off
                off += p.numel()
            off = (off + 3) // 4 * 4                     # 16-byte alignment of the next group
        self.numel = off
        self.params = params
        self.acc = torch.zeros(off, device=dev, dtype=torch.float32)
        dtypes = {p.dtype for p in params}
        self.out = {dt: (self.acc if dt == torch.float32 else torch.zeros(off, device=dev, dtype=dt))
                    for dt in dtypes}
        self.active = False
        self.touched: set = set()
        # Weight gradients feed nothing but this arena, so they are off the backward's critical path: with
        # ``side_stream`` set (enable_grad_arena(overlap=True)) every weight-gradient GEMM is issued on that
        # stream, behind an event of the launching stream, and the end-of-backward conversion waits for it.
        # The memory-bound dW GEMMs then run under the L2-reduction-bound sampler backward and the other
        # main-stream kernels instead of after them.  Capturable: the side stream forks from and joins the
        # capturing stream inside the pass.
        self.side_stream = None
        # data parallel: set to True when the caller averages the fp32 accumulator over the ranks itself
        # (all_reduce_mean) -- the conversion to the parameter dtype then happens after the all-reduce, so
        # gradients are averaged in fp32 like the reference's DDP does, not in bf16
        self.defer_conversion = False
        # The parameter VALUES move into flat buffers with the same offsets (p.data becomes a view): the
        # members of an adjacency group then already form the stacked matrix their module projects with, and
        # `stacked` hands out a view instead of launching a torch.cat per forward pass.
        self.pflat = {dt: torch.empty(off, device=dev, dtype=dt) for dt in dtypes}
        for p in params:
            o = self.offset[id(p)]
            home = self.pflat[p.dtype][o:o + p.numel()].view(p.shape)
            home.copy_(p.data)
            p.data = home
            p._bevf_acc = self.acc_view(p)
            p._bevf_arena = self

    # ---- views ------------------------------------------------------------------------------------
    def acc_view(self, p) -> torch.Tensor:
        o = self.offset[id(p)]
        return self.acc[o:o + p.numel()].view(p.shape)

    def span_view(self, ps, shape) -> torch.Tensor:
        """fp32 accumulator covering several adjacent parameters as one tensor of ``shape`` (None if they
        are not adjacent in this arena)."""
        o = self.offset.get(id(ps[0]))
        if o is None:
            return None
        end = o
        for p in ps:
            if self.offset.get(id(p)) != end:
                return None
            end += p.numel()
        return self.acc[o:end].view(shape)

    def grad_view(self, p) -> torch.Tensor:
        o = self.offset[id(p)]
        return self.out[p.dtype][o:o + p.numel()].view(p.shape)

    # ---- one backward pass ------------------------------------------------------------------------
    def touch(self, *params) -> None:
        """Called by a backward node before it accumulates into the arena: the first call of a pass zeroes
        the buffer and books the end-of-backward conversion."""
        if not self.active:
            self.acc.zero_()
            self.active = True
            self.touched = set()
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
        for t in params:
            for m in getattr(t, "_bevf_members", (t,)):      # a stacked weight stands for its member parameters
                self.touched.add(id(m))

    def run_off_critical_path(self, fn, *inputs) -> None:
        """Launch ``fn()`` (kernels that only accumulate into this arena) on the side stream when there is
        one, ordered after everything already queued on the current stream; otherwise inline."""
        if self.side_stream is None:
            fn()
            return
        main = torch.cuda.current_stream(self.acc.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side_stream.wait_event(ev)
        self.side_used = True
        with torch.cuda.stream(self.side_stream):
            fn()
        for t in inputs:                                   # the allocator must not recycle them early
            t.record_stream(self.side_stream)

    def _finalize(self) -> None:
        self.active = False
        if self.side_stream is not None and getattr(self, "side_used", False):
            # (only when this pass put work there: joining a stream that is not part of an ongoing capture is an error)
            torch.cuda.current_stream(self.acc.device).wait_stream(self.side_stream)
        self.side_used = False
        if not self.defer_conversion:
            self.convert()
        for p in self.params:
            if id(p) not in self.touched:
                continue
            view = self.grad_view(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view

    def convert(self) -> None:
        for dt, buf in self.out.items():
            if buf is not self.acc:
                buf.copy_(self.acc)                                   # the one conversion of the step

    def all_reduce_mean(self, world_size: int, group=None) -> None:
        """Average the fp32 accumulator over the process group with ONE all-reduce (19.8 MB for the base
        encoder), then convert once: the data-parallel exchange of a step (requires defer_conversion)."""
        import torch.distributed as dist
        dist.all_reduce(self.acc, group=group)
        self.acc.div_(world_size)
        self.convert()

    def flat_grad(self, dtype) -> torch.Tensor:
        return self.out[dtype]


def stacked(params, direct: bool = False) -> torch.Tensor:
    """torch.cat(params, 0) for parameters a module projects with as ONE matrix, tagged with the arena
    accumulator that spans them.  ``direct``: the caller guarantees that the consumer accumulates the
    gradient into the arena itself (plugin/linear.py's tcgen05 nodes) -- then, while the parameters still
    live adjacently in the arena's flat value buffer, the result is a VIEW of that buffer (no kernel): a leaf
    that requires grad and never receives one through autograd."""
    ar = getattr(params[0], "_bevf_arena", None)
    if ar is None or any(getattr(p, "_bevf_arena", None) is not ar for p in params):
        return torch.cat(params, 0)
    shape = (sum(p.shape[0] for p in params),) + tuple(params[0].shape[1:])
    span = ar.span_view(params, shape)
    if span is None:
        return torch.cat(params, 0)
    out = None
    if direct and shape[0] % 8 == 0:
        flat = ar.pflat[params[0].dtype]
        o, es = ar.offset[id(params[0])], flat.element_size()
        if all(p.dtype == flat.dtype and p.data_ptr() == flat.data_ptr() + ar.offset[id(p)] * es for p in params):
            out = flat[o:o + span.numel()].view(shape)
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                out.requires_grad_(True)
    if out is None:
        out = torch.cat(params, 0)
    out._bevf_arena, out._bevf_acc, out._bevf_members = ar, span, tuple(params)
    return out


def arena_of(*tensors):
    """(arena, [fp32 accumulators]) when every tensor carries an accumulator of one arena, else (None, None).
    A tensor is a parameter registered in an arena, or a stacked weight tagged by the module that built it."""
    arena, accs = None, []
    for t in tensors:
        if t is None:
            accs.append(None)
            continue
        a = getattr(t, "_bevf_arena", None)
        acc = getattr(t, "_bevf_acc", None)
        if a is None or acc is None or (arena is not None and a is not arena):
            return None, None
        arena = a
        accs.append(acc)
    return arena, accs
