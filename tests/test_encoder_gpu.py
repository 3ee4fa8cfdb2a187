"""GPU parity of the drop-in encoder (config -> registry -> modules -> libbevformer_b200.so) against
the golden vectors produced by the reference's own unmodified modules, forward and backward.
Tolerances: fp32 storage 1e-3, bf16 storage 1e-2 relative to max(1, max|ref|) for activations
(BASELINE.json north_star); gradients are compared through whole-tensor statistics and row samples
with looser bars because bf16 GEMMs feed them."""
import numpy as np
import pytest
import torch

from bevformer_b200 import _lib, synthetic as syn
from bevformer_b200.plugin import build_transformer_layer_sequence
from oracle import torch_ref
from tests.util import fixed_projection, golden, max_err, rel_err, stats, stats_close


def robust_close(got, want, tol, max_outlier_frac=0.02):
    """Gradients of bilinear sampling w.r.t. the sampling location are discontinuous where a sample
    sits on a cell boundary, so two correct implementations that differ by one ulp in a location
    disagree completely on the few samples that straddle a boundary.  Compare in relative L2 and by
    the fraction of outlying elements instead of the maximum error."""
    got = torch.as_tensor(got).detach().double().cpu().flatten()
    want = torch.as_tensor(want).detach().double().cpu().flatten()
    scale = max(1.0, want.abs().max().item())
    l2 = ((got - want).norm() / max(want.norm().item(), 1e-12)).item()
    frac = ((got - want).abs() > tol * scale).double().mean().item()
    return l2 < 5 * tol and frac < max_outlier_frac, (l2, frac)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _build(workload, dtype, seed=0):
    w = syn.WORKLOADS[workload]
    enc = build_transformer_layer_sequence(syn.encoder_cfg(w))
    enc.load_state_dict(syn.make_state_dict(w, seed=seed))
    return w, enc.to(DEV, dtype).eval()


def _inputs(w, bs, with_prev, dtype, seed=0):
    inp = syn.make_encoder_inputs(w, bs=bs, seed=seed, with_prev=with_prev)
    if bs > 1:
        g = torch.Generator().manual_seed(99)
        inp.feat = inp.feat + 0.5 * torch.randn(inp.feat.shape, generator=g)
        inp.bev_query = inp.bev_query + 0.1 * torch.randn(inp.bev_query.shape, generator=g)
    for k in ("bev_query", "feat", "bev_pos", "prev_bev", "shift"):
        t = getattr(inp, k)
        if t is not None:
            setattr(inp, k, t.to(DEV, dtype if k != "shift" else torch.float32))
    inp.spatial_shapes = inp.spatial_shapes.to(DEV)
    inp.level_start_index = inp.level_start_index.to(DEV)
    return inp


CASES = [("toy", "toy", 1, True), ("toy_bs2", "toy", 2, True), ("toy_noprev", "toy", 1, False),
         ("tiny", "tiny", 1, True), ("tiny_noprev", "tiny", 1, False), ("small", "small", 1, True),
         ("small4", "small4", 1, True), ("base", "base", 1, True)]


@pytest.mark.parametrize("name,workload,bs,with_prev", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_against_golden(name, workload, bs, with_prev, dtype):
    g = golden("encoder_" + name)
    w, enc = _build(workload, dtype)
    inp = _inputs(w, bs, with_prev, dtype)
    before = _lib.launch_count()
    with torch.no_grad():
        out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs())
    torch.cuda.synchronize()
    assert _lib.launch_count() - before >= 5 * w.num_layers, "the CUDA library did not run the layers"
    assert out.shape == (bs, w.num_query, 256) and out.dtype == dtype
    out = out.float().cpu()
    # the bf16 bar is wider than the op-level 1e-2: six layers of bf16 GEMMs + LayerNorm sit between
    tol = 1e-3 if dtype == torch.float32 else 6e-2
    assert rel_err(out[:, g["rows_q"]], g["out_rows"]) < tol
    if "out_full" in g:
        assert rel_err(out, g["out_full"]) < tol
    assert stats_close(stats(out), g["out_stats"], 2e-3 if dtype == torch.float32 else 3e-2)


@pytest.mark.parametrize("name,workload,bs,with_prev", [c for c in CASES if c[0] != "tiny_noprev"])
def test_backward_against_golden_fp32(name, workload, bs, with_prev):
    g = golden("encoder_" + name)
    w, enc = _build(workload, torch.float32)
    inp = _inputs(w, bs, with_prev, torch.float32)
    inp.bev_query.requires_grad_(True)
    inp.feat.requires_grad_(True)
    inp.bev_pos.requires_grad_(True)      # trains in the real model (LearnedPositionalEncoding)
    out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs())
    (out * fixed_projection(out.shape).to(DEV)).sum().backward()
    torch.cuda.synchronize()
    ok, m = robust_close(inp.bev_query.grad.cpu()[g["rows_q"]], g["grad_query_rows"], 2e-3)
    assert ok, ("grad_query", m)
    # every layer's temporal self-attention adds bev_pos to its query: all of them feed this gradient
    # (it is the sum of the TSA grad_loc paths of all layers: as boundary-sensitive as the
    # sampling_offsets rows below, hence their bar)
    ok, m = robust_close(inp.bev_pos.grad.cpu()[g["rows_q"]], g["grad_pos_rows"], 5e-3)
    assert ok, ("grad_pos", m)
    got, want = stats(inp.bev_pos.grad), g["grad_pos_stats"]
    assert np.all(np.abs(got[1:3] - want[1:3]) <= 5e-3 * np.abs(want[1:3])), ("grad_pos", got, want)
    ok, m = robust_close(inp.feat.grad.cpu()[:, g["rows_s"]], g["grad_feat_rows"], 2e-3)
    assert ok, ("grad_feat", m)
    # whole-tensor energy (sum |x|, sum x^2); max|x| is left out: one straddling sample can move it
    for got, want in ((stats(inp.bev_query.grad), g["grad_query_stats"]),
                      (stats(inp.feat.grad), g["grad_feat_stats"])):
        assert np.all(np.abs(got[1:3] - want[1:3]) <= 5e-3 * np.abs(want[1:3])), (got, want)
    for k, p in enc.named_parameters():
        assert p.grad is not None, k
        got, want = stats(p.grad), g["gstat:" + k]
        assert np.all(np.abs(got[1:3] - want[1:3]) <= 1e-2 * np.maximum(np.abs(want[1:3]), 1e-9)), (k, got, want)
        if "gfull:" + k in g.files:
            ok, m = robust_close(p.grad.cpu(), g["gfull:" + k], 5e-3, 0.15)
            assert ok, (k, m)
        elif "grows:" + k in g.files:
            rows = g["grows:" + k]
            # (the sampling_offsets weights collect grad_loc directly: the most boundary-sensitive rows)
            ok, m = robust_close(p.grad.cpu()[: rows.shape[0]], rows, 5e-3, 0.15)
            assert ok, (k, m)


def _cos(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return torch.nn.functional.cosine_similarity(a, b, dim=0).item()


@pytest.mark.parametrize("name,workload", [("toy", "toy"), ("tiny", "tiny"), ("small4", "small4"),
                                           ("base", "base")])
def test_backward_bf16_tracks_fp32_reference(name, workload):
    """bf16 storage (the headline configuration: base, bf16, fwd+bwd) against the fp32 golden gradients
    of the reference's own modules: rows of d bev_query / d feat / d bev_pos and every parameter's
    gradient.  bf16 GEMMs (rel. 2^-9 per product sum) sit between the sampler calls, so the bars are
    direction (cosine), energy and relative L2 of the sampled rows rather than the op-level 1e-2."""
    g = golden("encoder_" + name)
    w, enc = _build(workload, torch.bfloat16)
    inp = _inputs(w, 1, True, torch.bfloat16)
    inp.bev_query.requires_grad_(True)
    inp.feat.requires_grad_(True)
    inp.bev_pos.requires_grad_(True)
    out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs())
    (out.float() * fixed_projection(out.shape).to(DEV)).sum().backward()
    torch.cuda.synchronize()
    report = {}
    for key, got, want in (("grad_query", inp.bev_query.grad.float().cpu()[g["rows_q"]], g["grad_query_rows"]),
                           ("grad_feat", inp.feat.grad.float().cpu()[:, g["rows_s"]], g["grad_feat_rows"]),
                           ("grad_pos", inp.bev_pos.grad.float().cpu()[g["rows_q"]], g["grad_pos_rows"])):
        want = torch.from_numpy(want)
        l2 = ((got - want).norm() / want.norm()).item()
        report[key] = (round(_cos(got, want), 5), round(l2, 4))
    print(name, report)
    for key, (cos, l2) in report.items():
        assert cos > 0.99 and l2 < 0.12, (key, cos, l2)   # measured on B200: base 0.996 / 0.088, toy 0.999 / 0.05
    worst = (None, 0.0)
    for k, p in enc.named_parameters():
        got, ref = stats(p.grad.float()), g["gstat:" + k]
        dev = abs(got[2] - ref[2]) / max(ref[2], 1e-12)
        worst = max(worst, (k, dev), key=lambda t: t[1])
        assert dev <= 0.10, (k, got, ref)                                   # sum of squares (measured: <= 0.06)
        if "gfull:" + k in g.files:                                         # small tensors: element-wise direction
            assert _cos(p.grad.float().cpu(), g["gfull:" + k]) > 0.97, k
    print(name, "worst parameter-gradient energy deviation", worst)


def test_restatement_agrees_on_fresh_seed():
    """A case with no golden file: the repo-resident restatement (validated against the reference in
    the dev container) is the checker, on a seed the golden files do not cover."""
    w, enc = _build("toy", torch.float32, seed=5)
    inp = _inputs(w, 2, True, torch.float32, seed=5)
    with torch.no_grad():
        out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs()).cpu()
        cpu = syn.make_encoder_inputs(w, bs=2, seed=5)
        g = torch.Generator().manual_seed(99)
        cpu.feat = cpu.feat + 0.5 * torch.randn(cpu.feat.shape, generator=g)
        cpu.bev_query = cpu.bev_query + 0.1 * torch.randn(cpu.bev_query.shape, generator=g)
        ref = torch_ref.encoder_forward(syn.make_state_dict(w, seed=5), w.num_layers, cpu.bev_query,
                                        cpu.feat, use_c_oracle=True, **cpu.kwargs())
    assert rel_err(out, ref) < 1e-3


def test_point_sampling_kernel_matches_restatement():
    from bevformer_b200 import ops
    for name in ("tiny", "base"):
        w = syn.WORKLOADS[name]
        metas = syn.make_img_metas(w, 2)
        l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in metas], dtype=np.float32)).to(DEV)
        z = (torch.linspace(0.5, 7.5, 4) / 8.0).tolist()
        ref_cam, mask = ops.point_sampling(l2i, syn.PC_RANGE, z, w.img_hw[0], w.img_hw[1], w.bev_h, w.bev_w)
        r3 = torch_ref.reference_points_3d(w.bev_h, w.bev_w, 8.0, 4, 2, torch.float32)
        want_xy, want_mask = torch_ref.point_sampling(r3, syn.PC_RANGE, metas)
        flips = (mask.cpu() != want_mask).sum().item()
        assert flips <= 2, flips                       # only points within an ulp of the image border
        ok = want_mask & mask.cpu()
        assert max_err(ref_cam.cpu()[ok], want_xy[ok]) < 1e-5


def test_train_mode_dropout_runs_and_backpropagates():
    w, enc = _build("toy", torch.bfloat16)
    enc.train()
    inp = _inputs(w, 1, True, torch.bfloat16)
    inp.bev_query.requires_grad_(True)
    out = enc(inp.bev_query, inp.feat, inp.feat, **inp.kwargs())
    out.float().square().sum().backward()
    assert torch.isfinite(out.float()).all() and torch.isfinite(inp.bev_query.grad.float()).all()


def test_prep_and_reduction_kernels_against_torch():
    """The warp-cooperative SCA prep kernels (forward + backward) and the column-sum kernel against
    tensor-op restatements of spatial_cross_attention.py:338-372."""
    from bevformer_b200 import ops
    from bevformer_b200.plugin import ScaPlan
    for wl in ("toy", "tiny", "small4"):
        w = syn.WORKLOADS[wl]
        bs, m, l, p, d = 2, 8, len(w.levels), w.sca_points, 4
        metas = syn.make_img_metas(w, bs)
        l2i = torch.as_tensor(np.asarray([mm["lidar2img"] for mm in metas], dtype=np.float32)).to(DEV)
        z = (torch.linspace(0.5, 7.5, 4) / 8.0).tolist()
        ref_cam, mask = ops.point_sampling(l2i, syn.PC_RANGE, z, w.img_hw[0], w.img_hw[1], w.bev_h, w.bev_w)
        plan = ScaPlan.build(mask, ref_cam)
        nq, r = w.num_query, plan.num_pairs
        g = torch.Generator().manual_seed(3)
        raw = torch.randn(bs * nq, m * l * p * 3, generator=g).to(DEV).requires_grad_(True)
        ss = torch.tensor(w.levels, dtype=torch.int64, device=DEV)
        loc, attn = ops.ScaPrep.apply(raw, plan.ref_cam, plan.pair_q, plan.pair_cam, plan.pair_of, ss, bs,
                                      nq, m, l, p)
        # tensor-op restatement
        raw2 = raw.detach().clone().requires_grad_(True)
        off = raw2[:, : m * l * p * 2].view(bs, nq, m, l, p, 2)
        lg = raw2[:, m * l * p * 2:].view(bs, nq, m, l * p).softmax(-1).view(bs, nq, m, l, p)
        norm = torch.stack([ss[:, 1], ss[:, 0]], -1).float()
        off = off / norm[None, None, None, :, None, :]
        pq, pc = plan.pair_q.long(), plan.pair_cam.long()
        rc = plan.ref_cam[pc, :, pq].permute(1, 0, 2, 3)                      # (bs, R, D, 2)
        anchor = rc[:, :, None, None, :, :].expand(bs, r, 1, 1, d, 2)
        anchor = anchor[:, :, :, :, torch.arange(p) % d]                      # point p -> anchor p mod D
        want_loc = (off[:, pq] + anchor).reshape(bs * r, m, l, p, 2)
        want_attn = lg[:, pq].reshape(bs * r, m, l, p)
        assert max_err(loc, want_loc) < 1e-5 and max_err(attn, want_attn) < 1e-6
        gl, ga = torch.randn_like(loc), torch.randn_like(attn)
        (loc * gl).sum().backward(retain_graph=True)
        (attn * ga).sum().backward()
        ((want_loc * gl).sum() + (want_attn * ga).sum()).backward()
        assert rel_err(raw.grad, raw2.grad) < 1e-5
    x = torch.randn(40000, 512, device=DEV)
    assert rel_err(ops.colsum(x), x.double().sum(0)) < 1e-5
    xb = x.bfloat16()
    assert rel_err(ops.colsum(xb), xb.double().sum(0)) < 1e-5


def test_fused_dropout_layernorm_is_consistent():
    """The keep-mask is never stored: backward regenerates it from the Philox seed.  dx is exactly 0
    where an element was dropped, which reveals the mask; the forward output must equal a LayerNorm
    of the input scaled by that mask, and the drop rate must match p."""
    from bevformer_b200 import ops
    torch.manual_seed(123)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        x = torch.randn(5000, 256, device=DEV).to(dtype).requires_grad_(True)
        res = torch.randn(5000, 256, device=DEV).to(dtype).requires_grad_(True)
        gamma = (1 + 0.1 * torch.randn(256, device=DEV)).to(dtype).requires_grad_(True)
        beta = (0.1 * torch.randn(256, device=DEV)).to(dtype).requires_grad_(True)
        p = 0.3
        y = ops.LayerNormResidual.apply(x, res, gamma, beta, 1e-5, p)
        dy = torch.randn_like(y)
        y.backward(dy)
        keep = (x.grad != 0)
        rate = 1.0 - keep.float().mean().item()
        assert abs(rate - p) < 0.01, rate
        xs = x.detach().float() * keep / (1 - p) + res.detach().float()
        ref = torch.nn.functional.layer_norm(xs, (256,), gamma.detach().float(), beta.detach().float(), 1e-5)
        assert rel_err(y.float(), ref) < tol
        # gradient of the residual branch is the unmasked LayerNorm gradient
        xs2 = xs.clone().requires_grad_(True)
        torch.nn.functional.layer_norm(xs2, (256,), gamma.detach().float(), beta.detach().float(), 1e-5).backward(dy.float())
        assert rel_err(res.grad.float(), xs2.grad) < tol
        assert rel_err(x.grad.float(), xs2.grad * keep / (1 - p)) < tol
        # p = 0 reproduces plain residual LayerNorm and two calls draw different masks
        y0 = ops.LayerNormResidual.apply(x.detach(), res.detach(), gamma.detach(), beta.detach(), 1e-5, 0.0)
        ref0 = torch.nn.functional.layer_norm(x.detach().float() + res.detach().float(), (256,),
                                              gamma.detach().float(), beta.detach().float(), 1e-5)
        assert rel_err(y0.float(), ref0) < tol
        y_again = ops.LayerNormResidual.apply(x.detach(), res.detach(), gamma.detach(), beta.detach(), 1e-5, p)
        assert not torch.equal(y_again, y.detach())


@pytest.mark.parametrize("overlap", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_grad_arena_matches_per_parameter_gradients(dtype, overlap):
    """enable_grad_arena(): the same gradients as the default path (own buffer + conversion per parameter),
    delivered as views of one flat buffer; two passes in a row (the arena is zeroed per pass, not summed)."""
    w, enc = _build("tiny", dtype)
    inp = _inputs(w, 1, True, dtype)
    proj = fixed_projection((1, w.num_query, 256)).to(DEV, dtype)

    def run():
        for p in enc.parameters():
            p.grad = None
        q = inp.bev_query.detach().requires_grad_(True)
        out = enc(q, inp.feat, inp.feat, **inp.kwargs())
        (out * proj).sum().backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().float().clone() for k, p in enc.named_parameters()}, q.grad.float().clone()

    base, qg0 = run()
    arena = enc.enable_grad_arena(overlap=overlap)
    for _ in range(2):
        got, qg = run()
        flat = arena.flat_grad(dtype)
        for k, p in enc.named_parameters():
            if id(p) in arena.touched:      # (TSA's output_proj reaches its weight through autograd ops: not in the arena)
                assert flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + flat.numel() * flat.element_size(), k
            ref = base[k]
            tol = 2e-2 if dtype == torch.bfloat16 else 1e-4
            assert (got[k] - ref).abs().max().item() <= tol * max(1e-3, ref.abs().max().item()), k
        # (run-to-run: the sampler's reductions are unordered, one bf16 ulp may flip)
        assert (qg - qg0).abs().max().item() <= (1e-2 if dtype == torch.bfloat16 else 1e-3) * max(1.0, qg0.abs().max().item())
