"""Module-level CPU restatement of the BEV-encoder hot path, in functional form over a state_dict.

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  The product never imports this file.  It exists because
the GPU box has no ``/root/reference``: there this restatement (validated in the dev container
against the reference's own unmodified modules, tests/test_oracle.py::test_restatement_vs_reference)
is the checker for module-level parity and the timed arm of ``bench.py --impl reference``.

Each function cites the reference lines it follows (paths relative to
``projects/mmdet3d_plugin/bevformer/modules/``).  The third-party pieces (mmcv FFN / LayerNorm /
the grid_sample fallback) follow SURVEY.md Appendix A/B.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# the op (mmcv's multi_scale_deformable_attn_pytorch, called at
# spatial_cross_attention.py:394 and temporal_self_attention.py:252)
# ------------------------------------------------------------------------------------------------
def msda_grid_sample(value: Tensor, spatial_shapes, loc: Tensor, attn: Tensor) -> Tensor:
    b, _, m, d = value.shape
    _, q, _, nl, p, _ = loc.shape
    hw = [(int(h), int(w)) for h, w in spatial_shapes]
    chunks = value.split([h * w for h, w in hw], dim=1)
    cols = []
    for lvl, (h, w) in enumerate(hw):
        img = chunks[lvl].permute(0, 2, 3, 1).reshape(b * m, d, h, w)
        grid = (2.0 * loc[:, :, :, lvl] - 1.0).permute(0, 2, 1, 3, 4).reshape(b * m, q, p, 2)
        cols.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                  align_corners=False))                 # (b*m, d, q, p)
    samples = torch.cat(cols, dim=-1)                                   # (b*m, d, q, nl*p)
    wts = attn.permute(0, 2, 1, 3, 4).reshape(b * m, 1, q, nl * p)
    return (samples * wts).sum(-1).view(b, m * d, q).transpose(1, 2).contiguous()


def _sampler(use_c_oracle: bool):
    if not use_c_oracle:
        return lambda v, ss, lsi, loc, a: msda_grid_sample(v, ss, loc, a)
    from .msda_oracle import MSDAOracleFunction
    return lambda v, ss, lsi, loc, a: MSDAOracleFunction.apply(
        v, torch.as_tensor(ss, dtype=torch.int64), torch.as_tensor(lsi, dtype=torch.int64), loc, a, 64)


# ------------------------------------------------------------------------------------------------
# geometry (encoder.py:46-85 get_reference_points, :88-149 point_sampling)
# ------------------------------------------------------------------------------------------------
def reference_points_3d(h: int, w: int, z_extent: float, n_pillar: int, bs: int, dtype) -> Tensor:
    """(bs, D, H*W, 3) pillar anchors, row-major q = i*W + j (encoder.py:61-71)."""
    zs = torch.linspace(0.5, z_extent - 0.5, n_pillar, dtype=dtype) / z_extent
    xs = torch.linspace(0.5, w - 0.5, w, dtype=dtype) / w
    ys = torch.linspace(0.5, h - 0.5, h, dtype=dtype) / h
    grid = torch.stack([xs.view(1, 1, w).expand(n_pillar, h, w),
                        ys.view(1, h, 1).expand(n_pillar, h, w),
                        zs.view(n_pillar, 1, 1).expand(n_pillar, h, w)], -1)
    return grid.reshape(n_pillar, h * w, 3)[None].repeat(bs, 1, 1, 1)


def reference_points_2d(h: int, w: int, bs: int, dtype) -> Tensor:
    """(bs, H*W, 1, 2) as (x, y) (encoder.py:74-85)."""
    ys = (torch.linspace(0.5, h - 0.5, h, dtype=dtype) / h).view(h, 1).expand(h, w)
    xs = (torch.linspace(0.5, w - 0.5, w, dtype=dtype) / w).view(1, w).expand(h, w)
    return torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)[None, :, None].repeat(bs, 1, 1, 1)


def point_sampling(ref_3d: Tensor, pc_range: Sequence[float], img_metas) -> tuple:
    """lidar -> image projection in fp32 (encoder.py:95-144).
    Returns reference_points_cam (cam, B, Nq, D, 2) and bev_mask (cam, B, Nq, D) bool."""
    l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in img_metas]),
                          dtype=torch.float32)                          # (B, cam, 4, 4)
    ext = [pc_range[3] - pc_range[0], pc_range[4] - pc_range[1], pc_range[5] - pc_range[2]]
    pts = torch.stack([ref_3d[..., i] * ext[i] + pc_range[i] for i in range(3)], -1)   # :102-107,
    pts = torch.cat([pts, torch.ones_like(pts[..., :1])], -1)           # in the caller's dtype
    pts = pts.to(torch.float32)                                         # (B, D, Nq, 4)  :122-123
    b, d, nq = pts.shape[:3]
    cam = torch.matmul(l2i.view(1, b, -1, 1, 4, 4),                     # :116-123 (broadcast
                       pts.permute(1, 0, 2, 3).reshape(d, b, 1, nq, 4, 1))   # instead of repeat)
    cam = cam.squeeze(-1).permute(2, 1, 3, 0, 4)                        # (cam, B, Nq, D, 4)
    eps = 1e-5
    depth = cam[..., 2:3]
    mask = depth > eps                                                  # :126
    xy = cam[..., 0:2] / torch.maximum(depth, torch.full_like(depth, eps))
    img_h, img_w = img_metas[0]["img_shape"][0][0], img_metas[0]["img_shape"][0][1]
    xy = torch.stack([xy[..., 0] / img_w, xy[..., 1] / img_h], -1)      # :130-131
    mask = (mask & (xy[..., 1:2] > 0.0) & (xy[..., 1:2] < 1.0)
            & (xy[..., 0:1] < 1.0) & (xy[..., 0:1] > 0.0))              # :133-136
    return xy, mask.squeeze(-1)


# ------------------------------------------------------------------------------------------------
# attention blocks
# ------------------------------------------------------------------------------------------------
def _lin(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def temporal_self_attention(sd, pre, query: Tensor, prev_stack: Optional[Tensor], bev_pos: Tensor,
                            ref_2d: Tensor, bev_hw, sampler, num_heads=8, num_points=4) -> Tensor:
    """temporal_self_attention.py:128-272 with batch_first=True, num_levels=1, num_bev_queue=2.
    ``prev_stack`` is the (bs*2, Nq, C) [prev, cur] queue or None."""
    bs, nq, c = query.shape
    if prev_stack is None:                                              # :177-180
        prev_stack = torch.stack([query, query], 1).reshape(bs * 2, nq, c)
    identity = query                                                    # :184-185
    q = query + bev_pos                                                 # :186-187
    q = torch.cat([prev_stack[:bs], q], -1)                             # :197 (quirk 6)
    value = _lin(sd, pre + "value_proj", prev_stack).reshape(bs * 2, nq, num_heads, -1)
    off = _lin(sd, pre + "sampling_offsets", q).view(bs, nq, num_heads, 2, 1, num_points, 2)
    att = _lin(sd, pre + "attention_weights", q).view(bs, nq, num_heads, 2, num_points)
    att = att.softmax(-1).view(bs, nq, num_heads, 2, 1, num_points)     # :209-217 (quirk 7)
    att = att.permute(0, 3, 1, 2, 4, 5).reshape(bs * 2, nq, num_heads, 1, num_points).contiguous()
    off = off.permute(0, 3, 1, 2, 4, 5, 6).reshape(bs * 2, nq, num_heads, 1, num_points, 2)
    h, w = bev_hw
    norm = torch.tensor([[w, h]], dtype=off.dtype)                      # :225-226 (w, h)
    loc = ref_2d[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    ss = torch.tensor([[h, w]], dtype=torch.int64)
    out = sampler(value, ss, torch.zeros(1, dtype=torch.int64), loc.contiguous(), att)
    out = out.view(bs, 2, nq, c).mean(1)                                # :257-265
    return _lin(sd, pre + "output_proj", out) + identity                # :267-272 (eval dropout)


def msda3d(sd, pre, query: Tensor, value: Tensor, ref_cam: Tensor, spatial_shapes, lsi, sampler,
           num_heads=8, num_points=8) -> Tensor:
    """spatial_cross_attention.py:273-399 (batch_first=True, no identity, no output_proj)."""
    bs, nq, _ = query.shape
    nl = len(spatial_shapes)
    value = _lin(sd, pre + "value_proj", value).view(bs, value.shape[1], num_heads, -1)
    off = _lin(sd, pre + "sampling_offsets", query).view(bs, nq, num_heads, nl, num_points, 2)
    att = _lin(sd, pre + "attention_weights", query).view(bs, nq, num_heads, nl * num_points)
    att = att.softmax(-1).view(bs, nq, num_heads, nl, num_points)       # :343
    ss = torch.as_tensor(spatial_shapes)
    norm = torch.stack([ss[:, 1], ss[:, 0]], -1).to(off.dtype)          # :357-358
    d = ref_cam.shape[2]
    off = (off / norm[None, None, None, :, None, :]).view(
        bs, nq, num_heads, nl, num_points // d, d, 2)                   # :362-366 (quirk 3)
    loc = (ref_cam[:, :, None, None, None, :, :] + off).view(bs, nq, num_heads, nl, num_points, 2)
    return sampler(value, ss, torch.as_tensor(lsi), loc.contiguous(), att.contiguous())


def spatial_cross_attention(sd, pre, query: Tensor, feat: Tensor, ref_cam: Tensor,
                            bev_mask: Tensor, spatial_shapes, lsi, sampler,
                            num_heads=8, num_points=8, dense: bool = False) -> Tensor:
    """spatial_cross_attention.py:76-175.  ``feat`` (cam, S, bs, C); ``ref_cam`` (cam, bs, Nq, D, 2);
    ``bev_mask`` (cam, bs, Nq, D).  ``dense=False`` reproduces the per-camera re-batching with
    batch item 0's hit list (quirk 1); ``dense=True`` is the equivalent masked form."""
    bs, nq, c = query.shape
    ncam = feat.shape[0]
    d = ref_cam.shape[3]
    hit0 = [bev_mask[i, 0].sum(-1).nonzero().squeeze(-1) for i in range(ncam)]   # :138-140
    max_len = max(int(ix.numel()) for ix in hit0)
    q_re = query.new_zeros(bs, ncam, max_len, c)
    r_re = ref_cam.new_zeros(bs, ncam, max_len, d, 2)
    for j in range(bs):
        for i in range(ncam):
            n = hit0[i].numel()
            q_re[j, i, :n] = query[j, hit0[i]]
            r_re[j, i, :n] = ref_cam[i, j, hit0[i]]
    val = feat.permute(2, 0, 1, 3).reshape(bs * ncam, feat.shape[1], c)  # :157-160
    out = msda3d(sd, pre + "deformable_attention.", q_re.view(bs * ncam, max_len, c), val,
                 r_re.view(bs * ncam, max_len, d, 2), spatial_shapes, lsi, sampler,
                 num_heads, num_points).view(bs, ncam, max_len, c)
    slots = torch.zeros_like(query)
    for j in range(bs):
        for i in range(ncam):
            n = hit0[i].numel()
            slots[j, hit0[i]] = slots[j, hit0[i]] + out[j, i, :n]       # :165-167
    count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1).clamp(min=1.0)   # :169-171
    slots = slots / count[..., None]
    return _lin(sd, pre + "output_proj", slots) + query                 # :173-175 (eval dropout)


def ffn(sd, pre, x: Tensor) -> Tensor:
    """mmcv FFN: x + W2 relu(W1 x) in eval mode."""
    hidden = F.relu(_lin(sd, pre + "layers.0.0", x))
    return x + _lin(sd, pre + "layers.1", hidden)


def layer_norm(sd, pre, x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)


# ------------------------------------------------------------------------------------------------
# layer + encoder (encoder.py:287-406 and :151-239)
# ------------------------------------------------------------------------------------------------
def encoder_forward(sd: Dict[str, Tensor], num_layers: int, bev_query: Tensor, feat: Tensor, *,
                    bev_h: int, bev_w: int, bev_pos: Tensor, spatial_shapes, level_start_index,
                    prev_bev: Optional[Tensor], shift: Tensor, img_metas,
                    pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), num_points_in_pillar=4,
                    num_heads=8, tsa_points=4, sca_points=8, use_c_oracle=False,
                    return_intermediate=False, prefix: str = "") -> Tensor:
    """Functional BEVFormerEncoder.forward in eval mode. Inputs use the reference's layouts:
    bev_query/bev_pos/prev_bev (Nq, bs, C); feat (cam, S, bs, C). Returns (bs, Nq, C)."""
    sampler = _sampler(use_c_oracle)
    dtype = bev_query.dtype
    bs = bev_query.shape[1]
    ref_3d = reference_points_3d(bev_h, bev_w, pc_range[5] - pc_range[2], num_points_in_pillar,
                                 bs, dtype)
    ref_2d = reference_points_2d(bev_h, bev_w, bs, dtype)
    ref_cam, bev_mask = point_sampling(ref_3d, pc_range, img_metas)      # :193-194
    ref_cam = ref_cam.to(dtype)
    shift_ref = ref_2d + shift[:, None, None, :].to(dtype)               # :197-198 (quirk 9)
    q = bev_query.permute(1, 0, 2)
    pos = bev_pos.permute(1, 0, 2)
    nq = q.shape[1]
    if prev_bev is not None:                                             # :204-209 (quirk 8)
        prev_stack = torch.stack([prev_bev.permute(1, 0, 2), q], 1).reshape(bs * 2, nq, -1)
        hybrid = torch.stack([shift_ref, ref_2d], 1).reshape(bs * 2, nq, 1, 2)
    else:                                                                # :210-212
        prev_stack = None
        hybrid = torch.stack([ref_2d, ref_2d], 1).reshape(bs * 2, nq, 1, 2)
    ss = [(int(h), int(w)) for h, w in torch.as_tensor(spatial_shapes).tolist()]
    lsi = [int(v) for v in torch.as_tensor(level_start_index).tolist()]
    inter = []
    for i in range(num_layers):
        lp = f"{prefix}layers.{i}."
        q = temporal_self_attention(sd, lp + "attentions.0.", q, prev_stack, pos, hybrid,
                                    (bev_h, bev_w), sampler, num_heads, tsa_points)
        q = layer_norm(sd, lp + "norms.0.", q)
        q = spatial_cross_attention(sd, lp + "attentions.1.", q, feat, ref_cam, bev_mask, ss, lsi,
                                    sampler, num_heads, sca_points)
        q = layer_norm(sd, lp + "norms.1.", q)
        q = ffn(sd, lp + "ffns.0.", q)
        q = layer_norm(sd, lp + "norms.2.", q)
        if return_intermediate:
            inter.append(q)
    return torch.stack(inter) if return_intermediate else q


# ------------------------------------------------------------------------------------------------
# PerceptionTransformer.get_bev_features (modules/transformer.py:103-200): the encoder's caller
# ------------------------------------------------------------------------------------------------
def bev_shift(img_metas, bev_h: int, bev_w: int, grid_length, use_shift: bool = True):
    """Ego-motion shift of the BEV grid in normalised units, (bs, 2) float64 numpy (xy).
    transformer.py:122-140: translation length / heading from can_bus[0:2] and can_bus[-2]."""
    import numpy as np
    dx = np.array([m["can_bus"][0] for m in img_metas])
    dy = np.array([m["can_bus"][1] for m in img_metas])
    ego = np.array([m["can_bus"][-2] / np.pi * 180 for m in img_metas])
    length = np.sqrt(dx ** 2 + dy ** 2)
    ang = np.arctan2(dy, dx) / np.pi * 180
    bev_angle = ego - ang
    sy = length * np.cos(bev_angle / 180 * np.pi) / grid_length[0] / bev_h
    sx = length * np.sin(bev_angle / 180 * np.pi) / grid_length[1] / bev_w
    return np.stack([sx * use_shift, sy * use_shift], -1)


def rotate_prev_bev(prev_bev: Tensor, img_metas, bev_h: int, bev_w: int, rotate_center) -> Tensor:
    """prev_bev (Nq, bs, C) rotated per sample by can_bus[-1] degrees about rotate_center with
    torchvision's nearest-neighbour ``rotate`` (transformer.py:142-153).  Returns a NEW tensor (the
    reference writes into its argument)."""
    from torchvision.transforms.functional import rotate
    out = prev_bev.clone()
    for i in range(prev_bev.shape[1]):
        angle = img_metas[i]["can_bus"][-1]
        img = prev_bev[:, i].reshape(bev_h, bev_w, -1).permute(2, 0, 1)
        img = rotate(img, angle, center=rotate_center)
        out[:, i] = img.permute(1, 2, 0).reshape(bev_h * bev_w, -1)
    return out


def flatten_feats(mlvl_feats, cams_embeds: Optional[Tensor], level_embeds: Tensor):
    """(cam, S, bs, C) camera features + embeddings and the pyramid's shapes (transformer.py:161-181)."""
    flat, shapes = [], []
    for lvl, feat in enumerate(mlvl_feats):
        h, w = feat.shape[-2:]
        f = feat.flatten(3).permute(1, 0, 3, 2)                           # (cam, bs, hw, C)
        if cams_embeds is not None:
            f = f + cams_embeds[:, None, None, :].to(f.dtype)
        f = f + level_embeds[None, None, lvl:lvl + 1, :].to(f.dtype)
        flat.append(f)
        shapes.append((h, w))
    ss = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    return torch.cat(flat, 2).permute(0, 2, 1, 3), ss, lsi


def get_bev_features(sd: Dict[str, Tensor], num_layers: int, mlvl_feats, bev_queries: Tensor, bev_h: int,
                     bev_w: int, *, grid_length=(0.512, 0.512), bev_pos: Tensor, prev_bev: Optional[Tensor],
                     img_metas, rotate_center=(100, 100), rotate_prev=True, use_shift=True,
                     use_can_bus=True, can_bus_norm=True, use_cams_embeds=True, **encoder_kwargs) -> Tensor:
    """Functional PerceptionTransformer.get_bev_features (eval mode); ``sd`` holds the transformer's
    parameters with ``encoder.*`` for the encoder.  Returns (bs, Nq, C)."""
    dtype = bev_queries.dtype
    bs = mlvl_feats[0].shape[0]
    q = bev_queries.unsqueeze(1).repeat(1, bs, 1)
    pos = bev_pos.flatten(2).permute(2, 0, 1)
    shift = torch.as_tensor(bev_shift(img_metas, bev_h, bev_w, grid_length, use_shift)).to(dtype)
    if prev_bev is not None:
        if prev_bev.shape[1] == bev_h * bev_w:                            # (bs, Nq, C) -> (Nq, bs, C)
            prev_bev = prev_bev.permute(1, 0, 2)
        if rotate_prev:
            prev_bev = rotate_prev_bev(prev_bev, img_metas, bev_h, bev_w, list(rotate_center))
    cb = torch.tensor([m["can_bus"] for m in img_metas], dtype=dtype)       # new_tensor: caller dtype
    h = torch.relu(F.linear(cb, sd["can_bus_mlp.0.weight"], sd["can_bus_mlp.0.bias"]))
    h = torch.relu(F.linear(h, sd["can_bus_mlp.2.weight"], sd["can_bus_mlp.2.bias"]))
    if can_bus_norm:
        h = F.layer_norm(h, (h.shape[-1],), sd["can_bus_mlp.norm.weight"], sd["can_bus_mlp.norm.bias"])
    q = q + h[None] * use_can_bus
    feat, ss, lsi = flatten_feats(mlvl_feats, sd["cams_embeds"] if use_cams_embeds else None,
                                  sd["level_embeds"])
    return encoder_forward(sd, num_layers, q, feat, bev_h=bev_h, bev_w=bev_w, bev_pos=pos,
                           spatial_shapes=ss, level_start_index=lsi, prev_bev=prev_bev, shift=shift,
                           img_metas=img_metas, prefix="encoder.", **encoder_kwargs)


# ------------------------------------------------------------------------------------------------
# decoder cross-attention (modules/decoder.py:233-345), functional, batch-first inputs
# ------------------------------------------------------------------------------------------------
def custom_ms_deformable_attention(sd, pre, query: Tensor, value: Tensor, reference_points: Tensor,
                                   spatial_shapes, sampler, query_pos: Optional[Tensor] = None,
                                   key_padding_mask: Optional[Tensor] = None, num_heads=8,
                                   num_points=4) -> Tensor:
    """query (bs, Nq, C), value (bs, S, C), reference_points (bs, Nq, L, 2|4); eval mode.
    Returns output_proj(sampled) + query (identity = the query before query_pos is added)."""
    bs, nq, c = query.shape
    ss = [(int(h), int(w)) for h, w in torch.as_tensor(spatial_shapes).tolist()]
    nl = len(ss)
    q = query if query_pos is None else query + query_pos
    v = _lin(sd, pre + "value_proj", value)
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask[..., None], 0.0)
    v = v.view(bs, value.shape[1], num_heads, -1)
    off = _lin(sd, pre + "sampling_offsets", q).view(bs, nq, num_heads, nl, num_points, 2)
    att = _lin(sd, pre + "attention_weights", q).view(bs, nq, num_heads, nl * num_points).softmax(-1)
    att = att.view(bs, nq, num_heads, nl, num_points)
    if reference_points.shape[-1] == 2:
        norm = torch.tensor([[w, h] for h, w in ss], dtype=query.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / num_points * reference_points[:, :, None, :, None, 2:] * 0.5
    lsi = [0]
    for h, w in ss[:-1]:
        lsi.append(lsi[-1] + h * w)
    out = sampler(v, ss, lsi, loc, att)
    return _lin(sd, pre + "output_proj", out) + query


# ------------------------------------------------------------------------------------------------
# BEVFormerV2: PerceptionTransformerBEVEncoder.forward (modules/transformerV2.py:97-174)
# ------------------------------------------------------------------------------------------------
def bev_encoder_v2(sd: Dict[str, Tensor], num_layers: int, mlvl_feats, bev_queries: Tensor, bev_h: int,
                   bev_w: int, *, bev_pos: Tensor, img_metas, use_cams_embeds=True, **encoder_kwargs) -> Tensor:
    """Embeddings + flatten, encoder without temporal input (prev_bev=None, zero shift), then the
    optional resampling onto the augmented BEV grid.  Returns (bs, Nq, C)."""
    bs = mlvl_feats[0].shape[0]
    q = bev_queries.unsqueeze(1).repeat(1, bs, 1)
    pos = bev_pos.flatten(2).permute(2, 0, 1)
    feat, ss, lsi = flatten_feats(mlvl_feats, sd["cams_embeds"] if use_cams_embeds else None,
                                  sd["level_embeds"])
    bev = encoder_forward(sd, num_layers, q, feat, bev_h=bev_h, bev_w=bev_w, bev_pos=pos,
                          spatial_shapes=ss, level_start_index=lsi, prev_bev=None,
                          shift=q.new_tensor([0, 0]).unsqueeze(0), img_metas=img_metas,
                          prefix="encoder.", **encoder_kwargs)
    aug = img_metas[0].get("aug_param", {})
    if "GlobalRotScaleTransImage_param" not in aug:
        return bev
    _rot, _scale, _fx, _fy, bda_mat, only_gt = aug["GlobalRotScaleTransImage_param"]
    img = bev.reshape(bs, bev_h, bev_w, -1).permute(0, 3, 1, 2)
    if only_gt:
        ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, bev_h - 0.5, bev_h, dtype=q.dtype),
                                      torch.linspace(0.5, bev_w - 0.5, bev_w, dtype=q.dtype), indexing="ij")
        grid = torch.stack((ref_x / bev_w, ref_y / bev_h), -1) * 2.0 - 1.0
        grid = grid.unsqueeze(0).unsqueeze(-1)
        mat = torch.as_tensor(bda_mat)[:2, :2].to(grid).view(1, 1, 1, 2, 2)
        grid = torch.matmul(mat, grid).squeeze(-1)
        img = F.grid_sample(img, grid.expand(bs, -1, -1, -1), align_corners=False)
    return img.reshape(bs, -1, bev_h * bev_w).permute(0, 2, 1)
