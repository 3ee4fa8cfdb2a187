"""GPU parity of PerceptionTransformer.get_bev_features (the encoder's caller) and of its
feature-flattening kernel, against the golden vectors made from the reference's own class."""
import pytest
import torch

from bevformer_b200 import _lib, ops, synthetic as syn
from bevformer_b200.plugin import PerceptionTransformer
from oracle import torch_ref
from tests.test_encoder_gpu import robust_close
from tests.util import fixed_projection, golden, max_err, rel_err, stats, stats_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _flatten_ref(feats, cams, lvl):
    out, _, _ = torch_ref.flatten_feats(feats, cams, lvl)
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("use_cams", [True, False])
@pytest.mark.parametrize("extra_level_rows", [0, 3])
def test_flatten_feats_kernel_exact(dtype, use_cams, extra_level_rows):
    """Same arithmetic as the reference's tensor ops (cast the embedding, add, round): bit-exact,
    including ragged tiles (hw and C not multiples of 32) and bs > 1.  ``extra_level_rows``: the
    shipped tiny / small configs feed ONE pyramid level but keep num_feature_levels = 4, so
    level_embeds has rows no level uses -- their gradient is zero (transformer.py:172)."""
    g = torch.Generator().manual_seed(3)
    bs, ncam, C = 2, 3, 72
    shapes = [(7, 9), (4, 5), (1, 3)]
    feats = [torch.randn(bs, ncam, C, h, w, generator=g).to(DEV, dtype).requires_grad_(True) for h, w in shapes]
    cams = torch.randn(ncam, C, generator=g).to(DEV).requires_grad_(True)
    lvl = torch.randn(len(shapes) + extra_level_rows, C, generator=g).to(DEV).requires_grad_(True)
    before = _lib.launch_count()
    out = ops.FlattenFeats.apply(cams if use_cams else None, lvl, *feats)
    assert _lib.launch_count() - before == len(shapes)
    ref = _flatten_ref([f.detach() for f in feats], cams.detach() if use_cams else None, lvl.detach())
    assert out.shape == ref.shape and torch.equal(out, ref)
    # backward = the transposes back + three reductions
    proj = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).to(DEV, dtype)
    (out * proj).sum().backward()
    f2 = [f.detach().clone().requires_grad_(True) for f in feats]
    c2, l2 = cams.detach().clone().requires_grad_(True), lvl.detach().clone().requires_grad_(True)
    (_flatten_ref(f2, c2 if use_cams else None, l2) * proj).sum().backward()
    for a, b in zip(feats, f2):
        assert torch.equal(a.grad, b.grad)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert lvl.grad.shape == lvl.shape and not lvl.grad[len(shapes):].any()
    assert max_err(lvl.grad, l2.grad) <= tol * max(1.0, l2.grad.abs().max().item())
    if use_cams:
        assert max_err(cams.grad, c2.grad) <= tol * max(1.0, c2.grad.abs().max().item())


def _grid_length(w):
    return (0.512 * 200 / w.bev_h, 0.512 * 200 / w.bev_w)


def _build(workload, dtype):
    w = syn.WORKLOADS[workload]
    m = PerceptionTransformer(num_feature_levels=len(w.levels), num_cams=w.num_cams,
                              encoder=syn.encoder_cfg(w), decoder=None, embed_dims=w.embed_dims,
                              rotate_center=[w.bev_h // 2, w.bev_w // 2])
    m.load_state_dict(syn.make_perception_state_dict(w))
    return w, m.to(DEV, dtype).eval()


@pytest.mark.parametrize("name,workload,bs,with_prev", [("toy", "toy", 2, True), ("toy_noprev", "toy", 1, False),
                                                        ("tiny", "tiny", 1, True)])
def test_get_bev_features_fp32_vs_golden(name, workload, bs, with_prev):
    g = golden("perception_" + name)
    w, m = _build(workload, torch.float32)
    inp = syn.make_perception_inputs(w, bs=bs, with_prev=with_prev, device=DEV)
    for f in inp.mlvl_feats:
        f.requires_grad_(True)
    inp.bev_queries.requires_grad_(True)
    prev0 = None if inp.prev_bev is None else inp.prev_bev.clone()
    before = _lib.launch_count()
    out = m.get_bev_features(inp.mlvl_feats, inp.bev_queries, w.bev_h, w.bev_w, grid_length=_grid_length(w),
                             bev_pos=inp.bev_pos, prev_bev=inp.prev_bev, img_metas=inp.img_metas)
    assert _lib.launch_count() > before                       # the CUDA library did the work
    assert out.shape == (bs, w.num_query, w.embed_dims)
    assert max_err(out[:, g["rows_q"]], g["out_rows"]) < 1e-3  # fp32 storage bound (BASELINE.json)
    assert stats_close(stats(out), g["out_stats"], 1e-3)
    if prev0 is not None:
        assert torch.equal(prev0, inp.prev_bev)               # input left untouched (the reference rotates in place)
    (out * fixed_projection(out.shape).to(DEV)).sum().backward()
    ok, info = robust_close(inp.bev_queries.grad[g["rows_q"]], g["grad_queries_rows"], 1e-3)
    assert ok, info
    for i, f in enumerate(inp.mlvl_feats):
        ok, info = robust_close(f.grad[:, :, :8, :2], g[f"grad_feat{i}_slice"], 1e-3)
        assert ok, (i, info)
    for k in ("level_embeds", "cams_embeds", "can_bus_mlp.0.weight", "can_bus_mlp.norm.bias"):
        p = dict(m.named_parameters())[k]
        ok, info = robust_close(p.grad, g["gfull:" + k], 2e-3)
        assert ok, (k, info)


@pytest.mark.parametrize("name,workload,bs,with_prev", [("toy", "toy", 2, True), ("tiny", "tiny", 1, True)])
def test_get_bev_features_bf16_vs_golden(name, workload, bs, with_prev):
    """bf16 storage against the fp32 golden vectors of the reference class; same bar as the bf16
    encoder test (six layers of bf16 GEMMs + LayerNorm sit between)."""
    g = golden("perception_" + name)
    w, m = _build(workload, torch.bfloat16)
    inp = syn.make_perception_inputs(w, bs=bs, with_prev=with_prev, device=DEV, dtype=torch.bfloat16)
    with torch.no_grad():
        out = m.get_bev_features(inp.mlvl_feats, inp.bev_queries, w.bev_h, w.bev_w, grid_length=_grid_length(w),
                                 bev_pos=inp.bev_pos, prev_bev=inp.prev_bev, img_metas=inp.img_metas)
    assert out.dtype == torch.bfloat16
    assert rel_err(out.float()[:, g["rows_q"]], g["out_rows"]) < 6e-2
    assert stats_close(stats(out.float()), g["out_stats"], 3e-2)


def test_forward_is_out_of_scope():
    _, m = _build("toy", torch.float32)
    with pytest.raises(NotImplementedError):
        m(None, None)
