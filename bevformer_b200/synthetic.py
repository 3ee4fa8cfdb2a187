"""Synthetic nuScenes-shaped workloads for the BEV-encoder hot path (SURVEY.md §8 shape table, §8d).

There is no dataset and no checkpoint on the box: every test, golden vector and bench line uses
tensors of the shapes the reference's shipped configs produce, filled from a seeded generator,
and a synthetic six-camera rig whose projection matrices play the role of
``img_metas[i]['lidar2img']`` (reference: projects/mmdet3d_plugin/datasets/nuscenes_dataset.py:105-167).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]


@dataclass(frozen=True)
class Workload:
    """One row of SURVEY.md §8's shape table."""
    name: str
    bev_h: int
    bev_w: int
    num_layers: int
    levels: Tuple[Tuple[int, int], ...]          # (H_l, W_l) per FPN level
    img_hw: Tuple[int, int]                      # padded image (H, W) == img_metas img_shape
    scale: float                                 # image scale applied to the 900x1600 rig
    num_cams: int = 6
    embed_dims: int = 256
    num_heads: int = 8
    ffn_dims: int = 512
    pillar_points: int = 4                       # num_points_in_pillar (D)
    sca_points: int = 8
    tsa_points: int = 4
    config_file: Optional[str] = None            # reference config this mirrors, if shipped

    @property
    def num_query(self) -> int:
        return self.bev_h * self.bev_w

    @property
    def num_value(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def level_start(self) -> List[int]:
        out, s = [], 0
        for h, w in self.levels:
            out.append(s)
            s += h * w
        return out


WORKLOADS = {
    # projects/configs/bevformer/bevformer_tiny.py:45-48,59,70,90,184-185
    "tiny": Workload("tiny", 50, 50, 3, ((15, 25),), (480, 800), 0.5,
                     config_file="bevformer/bevformer_tiny.py"),
    # projects/configs/bevformer/bevformer_small.py:41-44,54,68,88,182-183 (single-level C5)
    "small": Workload("small", 150, 150, 3, ((23, 40),), (736, 1280), 0.8,
                      config_file="bevformer/bevformer_small.py"),
    # BASELINE.json configs[2]: small with a synthetic 4-level pyramid
    "small4": Workload("small4", 150, 150, 3, ((92, 160), (46, 80), (23, 40), (12, 20)),
                       (736, 1280), 0.8),
    # projects/configs/bevformer/bevformer_base.py:34-37,47,60,80
    "base": Workload("base", 200, 200, 6, ((116, 200), (58, 100), (29, 50), (15, 25)),
                     (928, 1600), 1.0, config_file="bevformer/bevformer_base.py"),
    # a toy used by fast CPU tests (not a reference config)
    "toy": Workload("toy", 12, 10, 2, ((8, 14), (4, 7)), (64, 112), 0.07),
}


def encoder_cfg(w: Workload) -> dict:
    """The encoder dict the reference configs spell out (bevformer_base.py:78-105), for workloads
    that have no shipped config file (small4, toy). Shipped configs are read from their files."""
    return dict(
        type="BEVFormerEncoder", num_layers=w.num_layers, pc_range=list(PC_RANGE),
        num_points_in_pillar=w.pillar_points, return_intermediate=False,
        transformerlayers=dict(
            type="BEVFormerLayer",
            attn_cfgs=[
                dict(type="TemporalSelfAttention", embed_dims=w.embed_dims, num_levels=1),
                dict(type="SpatialCrossAttention", pc_range=list(PC_RANGE),
                     deformable_attention=dict(type="MSDeformableAttention3D",
                                               embed_dims=w.embed_dims,
                                               num_points=w.sca_points,
                                               num_levels=len(w.levels)),
                     embed_dims=w.embed_dims),
            ],
            feedforward_channels=w.ffn_dims, ffn_dropout=0.1,
            operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")))


# ------------------------------------------------------------------------------------------------
# camera rig: lidar frame x-right / y-forward / z-up
# ------------------------------------------------------------------------------------------------
_YAWS_DEG = (0.0, -55.0, 55.0, 180.0, 110.0, -110.0)
_FOCALS = (1266.4, 1260.8, 1272.6, 809.2, 1256.7, 1259.5)
_PRINCIPAL = (816.0, 491.5)
_CAM_AHEAD, _CAM_BELOW = 1.0, 0.3   # each camera sits 1 m along its own optical axis, 0.3 m below the lidar


def make_lidar2img(scale: float = 1.0, num_cams: int = 6) -> np.ndarray:
    """(num_cams, 4, 4) float64 projection matrices: pixel = K(scale) @ [R | -R c] @ (x, y, z, 1)."""
    mats = []
    for i in range(num_cams):
        yaw = math.radians(_YAWS_DEG[i % 6])
        fwd = np.array([-math.sin(yaw), math.cos(yaw), 0.0])
        right = np.array([math.cos(yaw), math.sin(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        rot = np.stack([right, down, fwd])                       # lidar -> camera axes
        ext = np.eye(4)
        ext[:3, :3] = rot
        pos = fwd * _CAM_AHEAD + np.array([0.0, 0.0, -_CAM_BELOW])
        ext[:3, 3] = -rot @ pos
        f = _FOCALS[i % 6] * scale
        k = np.eye(4)
        k[0, 0] = k[1, 1] = f
        k[0, 2] = _PRINCIPAL[0] * scale
        k[1, 2] = _PRINCIPAL[1] * scale
        mats.append(k @ ext)
    return np.stack(mats)


def make_can_bus(sample: int) -> list:
    """18 deterministic CAN-bus values in the layout the reference reads (transformer.py:122-150,
    datasets/nuscenes_dataset.py): [0:3] ego translation delta (m), [-2] ego yaw (rad), [-1] yaw
    delta (deg) -- the angle prev_bev is rotated by."""
    g = np.random.default_rng(1234 + sample)
    cb = g.standard_normal(18) * 0.1
    cb[0], cb[1], cb[2] = 0.8 + 0.3 * sample, -0.35 + 0.2 * sample, 0.0
    cb[-2] = 0.6 - 0.25 * sample
    cb[-1] = 3.5 - 5.0 * sample
    return [float(x) for x in cb]


def make_img_metas(w: Workload, bs: int = 1):
    l2i = make_lidar2img(w.scale, w.num_cams)
    h, wd = w.img_hw
    return [dict(lidar2img=[l2i[i].copy() for i in range(w.num_cams)],
                 img_shape=[(h, wd, 3)] * w.num_cams, can_bus=make_can_bus(i)) for i in range(bs)]


@dataclass
class EncoderInputs:
    bev_query: torch.Tensor          # (Nq, bs, C)
    feat: torch.Tensor               # (num_cams, S, bs, C)   key == value
    bev_pos: torch.Tensor            # (Nq, bs, C)
    prev_bev: Optional[torch.Tensor]  # (Nq, bs, C) or None
    shift: torch.Tensor              # (bs, 2)
    spatial_shapes: torch.Tensor     # (L, 2) int64 (h, w)
    level_start_index: torch.Tensor  # (L,) int64
    img_metas: list = field(default_factory=list)
    bev_h: int = 0
    bev_w: int = 0

    def kwargs(self):
        return dict(bev_h=self.bev_h, bev_w=self.bev_w, bev_pos=self.bev_pos,
                    spatial_shapes=self.spatial_shapes, level_start_index=self.level_start_index,
                    prev_bev=self.prev_bev, shift=self.shift, img_metas=self.img_metas)


def make_encoder_inputs(w: Workload, bs: int = 1, seed: int = 0, with_prev: bool = True,
                        dtype=torch.float32, device="cpu") -> EncoderInputs:
    """Seeded N(0,1) tensors of the encoder's input contract (modules/transformer.py:186-198)."""
    g = torch.Generator().manual_seed(seed)
    c = w.embed_dims

    def rn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    bev_query = rn(w.num_query, 1, c).repeat(1, bs, 1)           # one embedding table, per sample
    feat = rn(w.num_cams, w.num_value, bs, c)
    feat = feat + rn(w.num_cams, 1, 1, c)                        # cams_embeds
    lvl = rn(len(w.levels), c)
    for i, (s0, (h, ww)) in enumerate(zip(w.level_start, w.levels)):
        feat[:, s0:s0 + h * ww] += lvl[i]                        # level_embeds
    bev_pos = torch.rand(w.num_query, 1, c, generator=g).repeat(1, bs, 1)
    prev_bev = rn(w.num_query, bs, c) if with_prev else None
    shift = torch.tensor([[0.01, -0.02]], dtype=torch.float32).repeat(bs, 1)
    ss = torch.tensor(w.levels, dtype=torch.int64)
    lsi = torch.tensor(w.level_start, dtype=torch.int64)

    def cv(t):
        return None if t is None else t.to(device=device, dtype=dtype).contiguous()

    return EncoderInputs(cv(bev_query), cv(feat), cv(bev_pos), cv(prev_bev),
                         shift.to(device=device, dtype=dtype), ss.to(device), lsi.to(device),
                         make_img_metas(w, bs), w.bev_h, w.bev_w)


@dataclass
class PerceptionInputs:
    """Input contract of PerceptionTransformer.get_bev_features (modules/transformer.py:103-113)."""
    mlvl_feats: list                 # per level (bs, num_cams, C, h, w)
    bev_queries: torch.Tensor        # (Nq, C)   the BEV embedding table
    bev_pos: torch.Tensor            # (bs, C, bev_h, bev_w)
    prev_bev: Optional[torch.Tensor]  # (bs, Nq, C) or None
    img_metas: list = field(default_factory=list)
    bev_h: int = 0
    bev_w: int = 0


def make_perception_inputs(w: Workload, bs: int = 1, seed: int = 0, with_prev: bool = True,
                           dtype=torch.float32, device="cpu") -> PerceptionInputs:
    g = torch.Generator().manual_seed(4000 + seed)
    c = w.embed_dims
    feats = [torch.randn(bs, w.num_cams, c, h, ww, generator=g) for h, ww in w.levels]
    bev_queries = torch.randn(w.num_query, c, generator=g)
    bev_pos = torch.rand(bs, c, w.bev_h, w.bev_w, generator=g)
    prev = torch.randn(bs, w.num_query, c, generator=g) if with_prev else None

    def cv(t):
        return None if t is None else t.to(device=device, dtype=dtype).contiguous()

    return PerceptionInputs([cv(f) for f in feats], cv(bev_queries), cv(bev_pos), cv(prev),
                            make_img_metas(w, bs), w.bev_h, w.bev_w)


def make_perception_state_dict(w: Workload, seed: int = 0, trained_like: bool = True):
    """state_dict of a PerceptionTransformer without decoder: the transformer's own parameters
    (reference key names, modules/transformer.py:70-84) + ``encoder.*``."""
    g = torch.Generator().manual_seed(7000 + seed)
    c = w.embed_dims
    sd = {"level_embeds": torch.randn(len(w.levels), c, generator=g),
          "cams_embeds": torch.randn(w.num_cams, c, generator=g),
          "reference_points.weight": torch.randn(3, c, generator=g) * 0.05,
          "reference_points.bias": torch.zeros(3),
          "can_bus_mlp.0.weight": torch.randn(c // 2, 18, generator=g) * 0.2,
          "can_bus_mlp.0.bias": torch.randn(c // 2, generator=g) * 0.05,
          "can_bus_mlp.2.weight": torch.randn(c, c // 2, generator=g) * 0.1,
          "can_bus_mlp.2.bias": torch.randn(c, generator=g) * 0.05,
          "can_bus_mlp.norm.weight": 1.0 + 0.1 * torch.randn(c, generator=g),
          "can_bus_mlp.norm.bias": 0.1 * torch.randn(c, generator=g)}
    for k, v in make_state_dict(w, seed, trained_like).items():
        sd["encoder." + k] = v
    return sd


def randomize_trained_like(module: torch.nn.Module, seed: int = 1) -> None:
    """Second weight set of SURVEY.md §8d: the reference initialisers leave sampling offsets and
    attention logits query-independent (zero weights); perturb them so both depend on the query."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("sampling_offsets.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif name.endswith("attention_weights.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith("attention_weights.bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif ".norms." in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif ".norms." in name and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias") and "sampling_offsets" not in name:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))


# ------------------------------------------------------------------------------------------------
# op-level micro workloads (loc ~ U(-0.1, 1.1), attn = softmax(N(0,1)))
# ------------------------------------------------------------------------------------------------
def make_msda_inputs(bs, levels, num_query, num_heads=8, head_dim=32, num_points=4, seed=0,
                     dtype=torch.float32, device="cpu", value_scale=1.0, loc_range=(-0.1, 1.1)):
    g = torch.Generator().manual_seed(seed)
    s = sum(h * w for h, w in levels)
    nl = len(levels)
    value = torch.randn(bs, s, num_heads, head_dim, generator=g) * value_scale
    lo, hi = loc_range
    loc = torch.rand(bs, num_query, num_heads, nl, num_points, 2, generator=g) * (hi - lo) + lo
    attn = torch.randn(bs, num_query, num_heads, nl * num_points, generator=g).softmax(-1)
    attn = attn.view(bs, num_query, num_heads, nl, num_points)
    ss = torch.tensor(levels, dtype=torch.int64)
    lsi = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
    return (value.to(device=device, dtype=dtype).contiguous(), ss.to(device), lsi.to(device),
            loc.to(device=device, dtype=dtype).contiguous(),
            attn.to(device=device, dtype=dtype).contiguous())


# ------------------------------------------------------------------------------------------------
# deterministic encoder weights, independent of any module's construction order
# ------------------------------------------------------------------------------------------------
def _ring_bias(num_heads: int, groups: int, num_points: int) -> torch.Tensor:
    """The reference's sampling-offset bias: head m points along direction 2*pi*m/M (scaled to the
    unit square), point i at radius i+1 (spatial_cross_attention.py:255-267,
    temporal_self_attention.py:109-122). ``groups`` = num_levels (x num_bev_queue for TSA)."""
    th = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    d = torch.stack([th.cos(), th.sin()], -1)
    d = (d / d.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2).repeat(1, groups, num_points, 1)
    for i in range(num_points):
        d[:, :, i, :] *= i + 1
    return d.reshape(-1)


def make_state_dict(w: Workload, seed: int = 0, trained_like: bool = True, dtype=torch.float32):
    """Encoder ``state_dict`` with the reference's key names and shapes.
    ``trained_like=False`` reproduces the reference initialisers (zero offset/attention weights);
    ``True`` perturbs them so that offsets and attention logits depend on the query."""
    g = torch.Generator().manual_seed(1000 + seed)
    c, m, f = w.embed_dims, w.num_heads, w.ffn_dims
    nl = len(w.levels)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def rn(*s, std=1.0):
        return torch.randn(*s, generator=g) * std

    sd = {}
    for li in range(w.num_layers):
        p = f"layers.{li}."
        t, s = p + "attentions.0.", p + "attentions.1."
        n_off_t, n_att_t = 2 * m * 1 * w.tsa_points * 2, 2 * m * 1 * w.tsa_points
        sd[t + "sampling_offsets.weight"] = rn(n_off_t, 2 * c, std=0.02) if trained_like else torch.zeros(n_off_t, 2 * c)
        sd[t + "sampling_offsets.bias"] = _ring_bias(m, 1 * 2, w.tsa_points)
        sd[t + "attention_weights.weight"] = rn(n_att_t, 2 * c, std=0.1) if trained_like else torch.zeros(n_att_t, 2 * c)
        sd[t + "attention_weights.bias"] = rn(n_att_t, std=0.1) if trained_like else torch.zeros(n_att_t)
        sd[t + "value_proj.weight"] = xavier(c, c)
        sd[t + "value_proj.bias"] = rn(c, std=0.02) if trained_like else torch.zeros(c)
        sd[t + "output_proj.weight"] = xavier(c, c)
        sd[t + "output_proj.bias"] = rn(c, std=0.02) if trained_like else torch.zeros(c)
        d = s + "deformable_attention."
        n_off_s, n_att_s = m * nl * w.sca_points * 2, m * nl * w.sca_points
        sd[d + "sampling_offsets.weight"] = rn(n_off_s, c, std=0.02) if trained_like else torch.zeros(n_off_s, c)
        sd[d + "sampling_offsets.bias"] = _ring_bias(m, nl, w.sca_points)
        sd[d + "attention_weights.weight"] = rn(n_att_s, c, std=0.1) if trained_like else torch.zeros(n_att_s, c)
        sd[d + "attention_weights.bias"] = rn(n_att_s, std=0.1) if trained_like else torch.zeros(n_att_s)
        sd[d + "value_proj.weight"] = xavier(c, c)
        sd[d + "value_proj.bias"] = rn(c, std=0.02) if trained_like else torch.zeros(c)
        sd[s + "output_proj.weight"] = xavier(c, c)
        sd[s + "output_proj.bias"] = rn(c, std=0.02) if trained_like else torch.zeros(c)
        sd[p + "ffns.0.layers.0.0.weight"] = xavier(f, c)
        sd[p + "ffns.0.layers.0.0.bias"] = rn(f, std=0.02)
        sd[p + "ffns.0.layers.1.weight"] = xavier(c, f)
        sd[p + "ffns.0.layers.1.bias"] = rn(c, std=0.02)
        for k in range(3):
            sd[p + f"norms.{k}.weight"] = 1.0 + (rn(c, std=0.1) if trained_like else torch.zeros(c))
            sd[p + f"norms.{k}.bias"] = rn(c, std=0.1) if trained_like else torch.zeros(c)
    return {k: v.to(dtype) for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------
# generic seeded weights for modules whose state_dict layout is shared with the reference class
# ------------------------------------------------------------------------------------------------
def make_random_state_dict(module: torch.nn.Module, seed: int = 0, scale: float = 0.05) -> dict:
    """Deterministic "trained-like" weights for any module: iterates the state_dict in sorted key order
    with one seeded generator, so a reference module and its drop-in (same keys and shapes) receive
    identical values.  LayerNorm-style ``norms.*`` weights are 1 + noise; the ring-shaped
    ``sampling_offsets.bias`` keeps the module's own (deterministic) initialiser."""
    g = torch.Generator().manual_seed(31000 + seed)
    sd = module.state_dict()
    out = {}
    for k in sorted(sd):
        v = sd[k]
        if k.endswith("sampling_offsets.bias") or not v.is_floating_point():
            out[k] = v.clone()
        elif ".norms." in k and k.endswith("weight") or k.endswith("layer_norm.1.weight") or "bn" in k and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif k.endswith("sampling_offsets.weight"):
            out[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            out[k] = scale * torch.randn(v.shape, generator=g)
    return out


DECODER_CFG = dict(
    type="DetectionTransformerDecoder", num_layers=3, return_intermediate=True,
    transformerlayers=dict(
        type="DetrTransformerDecoderLayer",
        attn_cfgs=[dict(type="MultiheadAttention", embed_dims=256, num_heads=8, dropout=0.1),
                   dict(type="CustomMSDeformableAttention", embed_dims=256, num_levels=1)],
        feedforward_channels=512, ffn_dropout=0.1,
        operation_order=("self_attn", "norm", "cross_attn", "norm", "ffn", "norm")))


def make_decoder_inputs(w: Workload, bs: int = 2, num_query: int = 40, seed: int = 0):
    """Object queries, their positional part, the BEV memory (Nq_bev, bs, C), initial reference points in
    [0, 1]^3 and seeded regression branches (Linear(C, 10) per layer, the head's reg_branches)."""
    g = torch.Generator().manual_seed(12000 + seed)
    c = w.embed_dims
    query = torch.randn(num_query, bs, c, generator=g)
    query_pos = torch.randn(num_query, bs, c, generator=g)
    bev = torch.randn(w.num_query, bs, c, generator=g)
    ref = torch.rand(bs, num_query, 3, generator=g) * 0.9 + 0.05
    reg = torch.nn.ModuleList([torch.nn.Linear(c, 10) for _ in range(DECODER_CFG["num_layers"])])
    with torch.no_grad():
        for lin in reg:
            lin.weight.copy_(0.05 * torch.randn(lin.weight.shape, generator=g))
            lin.bias.copy_(0.05 * torch.randn(lin.bias.shape, generator=g))
    return query, query_pos, bev, ref, reg
