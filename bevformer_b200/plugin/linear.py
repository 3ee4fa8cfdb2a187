"""The dense projections of the encoder layer (value_proj / output_proj / the combined
sampling_offsets|attention_weights head / FFN), as one function so that the backing kernel can be
swapped in one place.

Backing: cuBLASLt through torch for now (a plain library GEMM, bias fused by the library).  The
tcgen05 kernel with fused epilogues replaces it behind this same function.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def linear(x: torch.Tensor, weight: torch.Tensor, bias, relu: bool = False) -> torch.Tensor:
    y = F.linear(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))
    return F.relu(y, inplace=True) if relu else y


def linear_fp32_out(x: torch.Tensor, weight: torch.Tensor, bias) -> torch.Tensor:
    """Projection whose result is consumed in fp32 (sampling offsets / attention logits)."""
    y = F.linear(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))
    return y.float()
