"""CustomMSDeformableAttention on the B200 kernels: the decoder's cross-attention from object queries
into the BEV map, third call site of the sampler after TSA and SCA.

Drop-in for the reference class of the same name
(projects/mmdet3d_plugin/bevformer/modules/decoder.py:132-345): same registry name, constructor
arguments, parameter names / shapes / initialisers, forward signature and return convention.  The
surrounding ``DetectionTransformerDecoder`` (reference-point refinement loop around mmcv's
``DetrTransformerDecoderLayer``) is outside this library's scope.

The shapes here are small (900 queries, one 200x200 level, 4 points): the value projection over the
40 000 BEV cells is the only part with real work -- it and the other projections run on the tcgen05
GEMM, the gather on the sampler kernel; the few-KB softmax / offset arithmetic in between stays in
tensor ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .linear import linear, linear_fp32_out
from .registry import ATTENTION, _register
from .temporal_self_attention import _check_head_dim, ring_offsets_


class CustomMSDeformableAttention(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__()
        _check_head_dim(embed_dims, num_heads)
        self.init_cfg = init_cfg
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.fp16_enabled = False
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        """decoder.py:209-228: zero offset/attention weights, ring bias, xavier projections."""
        nn.init.zeros_(self.sampling_offsets.weight)
        ring_offsets_(self.sampling_offsets.bias, self.num_heads, self.num_levels, self.num_points)
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        for lin in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, flag="decoder", **kwargs):
        """query (num_query, bs, C) (or batch-first), value (num_key, bs, C), reference_points
        (bs, num_query, num_levels, 2 | 4).  Returns dropout(output_proj(sampled)) + identity in the
        caller's layout (decoder.py:233-345)."""
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, nq, _ = query.shape
        nv = value.shape[1]
        m, l, p = self.num_heads, self.num_levels, self.num_points
        ss = torch.as_tensor(spatial_shapes).to(device=query.device, dtype=torch.int64)
        lsi = torch.as_tensor(level_start_index).to(device=query.device, dtype=torch.int64)
        if ss.shape[0] != l:
            raise AssertionError("spatial_shapes does not have num_levels rows")
        # (the reference also asserts sum(h*w) == num_value, a host sync per call: the sampler's
        #  argument check covers it without one)

        v = linear(value, self.value_proj.weight, self.value_proj.bias)
        if key_padding_mask is not None:
            v = v.masked_fill(key_padding_mask[..., None], 0.0)
        v = v.reshape(bs, nv, m, -1)
        w = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
        b = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0)
        raw = linear_fp32_out(query, w, b)                                   # (bs, nq, m*l*p*3) fp32
        n_off = m * l * p * 2
        off = raw[..., :n_off].reshape(bs, nq, m, l, p, 2)
        att = raw[..., n_off:].reshape(bs, nq, m, l * p).softmax(-1).reshape(bs, nq, m, l, p)
        rp = reference_points.float()
        if rp.shape[-1] == 2:
            norm = torch.stack([ss[..., 1], ss[..., 0]], -1).to(torch.float32)
            loc = rp[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        elif rp.shape[-1] == 4:
            loc = rp[:, :, None, :, None, :2] + off / p * rp[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, "
                             f"but get {rp.shape[-1]} instead.")
        out = ops.MultiScaleDeformableAttnFunction_fp32.apply(v, ss, lsi, loc.contiguous(),
                                                              att.contiguous(), self.im2col_step)
        out = linear(out.to(query.dtype), self.output_proj.weight, self.output_proj.bias)
        if not self.batch_first:
            out = out.permute(1, 0, 2)
        return self.dropout(out) + identity


_register(ATTENTION, CustomMSDeformableAttention)
