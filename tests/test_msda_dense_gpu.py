"""GPU parity tests of the dense tensor-core backward (csrc/msda_dense.cu, bevf_msda_rows_backward_dense):
grad_value of the coarse pyramid levels as C[pixel, row] x grad_out[row, 32] on tcgen05, everything else on
the one-kernel backward.  Checked THROUGH THE C ABI against (a) the plain row-list backward, which the other
test files pin to Oracle-S and the golden vectors, on ragged row lists / split and merged pixel bins / 4 and 8
points, and (b) Oracle-S itself on the real SCA launch of the headline benchmark.
Bar: 1e-2 for bf16 storage (BASELINE.json north_star), max|err| / max(1, max|ref|); the coefficients of the
dense path are rounded to bf16 (2^-9 relative, independent per term), accumulation is fp32."""
import os

import pytest
import torch

from bevformer_b200 import _lib, ops, synthetic as syn
from oracle import msda_oracle
from tests.util import fixed_projection, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def dense_on():
    """The dense kernel is what these tests are about: switch it on explicitly (bevf_msda_set_dense_backward),
    whatever the library default / environment says, and restore the previous setting afterwards."""
    lib = _lib.load()
    assert lib.bevf_msda_set_dense_backward(1) == 0
    yield
    lib.bevf_msda_set_dense_backward(int(os.environ.get("BEVF_MSDA_DENSE", "0") or 0))


def _ragged_case(levels, rows_per_map, heads, pts, seed, gap=37):
    """Row list grouped by value map with unused (-1) rows between the maps, as the SCA plan lays it out."""
    g = torch.Generator().manual_seed(seed)
    nb = len(rows_per_map)
    s = sum(h * w for h, w in levels)
    nl = len(levels)
    total = sum(rows_per_map) + gap * nb
    value = torch.randn(nb, s, heads, 32, generator=g)
    loc = torch.rand(total, heads, nl, pts, 2, generator=g) * 1.3 - 0.15
    # neighbouring rows sample neighbouring places (collisions inside a step), a few rows far outside the map
    base = torch.rand(total // 16 + 1, 1, 1, 1, 2, generator=g).repeat_interleave(16, 0)[:total]
    loc = 0.6 * base + 0.4 * loc
    loc[::53] = loc[::53] * 40.0 - 20.0
    attn = torch.randn(total, heads, nl * pts, generator=g).softmax(-1).view(total, heads, nl, pts)
    attn[::29, :, :, 0] = 0.0
    row_map = torch.full((total,), -1, dtype=torch.int32)
    rng = torch.zeros(nb, 2, dtype=torch.int32)
    r = 0
    for b, n in enumerate(rows_per_map):
        row_map[r:r + n] = b
        rng[b, 0], rng[b, 1] = r, r + n
        r += n + gap
    ss = torch.tensor(levels, dtype=torch.int64)
    lsi = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
    return value, ss, lsi, loc.contiguous(), attn.contiguous(), row_map, rng


CASES = {
    # one bin holding four levels (all four scatter groups), three maps, one of them spanning 3 row chunks
    "merged4": dict(levels=[(20, 30), (10, 15), (5, 8), (3, 4)], rows=[700, 40, 1300], heads=8, pts=8),
    # level 0 split over 4 bins, levels 1 + 2 merged, level 3 its own bin; 4 points
    "split": dict(levels=[(64, 100), (32, 50), (16, 25), (13, 160)], rows=[530, 16, 1], heads=8, pts=4),
    # a level too large for the dense path in front (stays on the reductions), odd head count, empty map
    "mixed": dict(levels=[(100, 120), (29, 50), (15, 25)], rows=[300, 0, 513], heads=3, pts=8),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_dense_backward_equals_plain(case):
    c = CASES[case]
    v, ss, lsi, loc, attn, row_map, rng = _ragged_case(c["levels"], c["rows"], c["heads"], c["pts"], seed=len(case))
    vd = v.to(DEV, torch.bfloat16)
    ss, lsi, loc, attn, row_map, rng = (t.to(DEV) for t in (ss, lsi, loc, attn, row_map, rng))
    gout = fixed_projection((loc.shape[0], c["heads"] * 32)).to(DEV, torch.bfloat16)
    gv0, gl0, ga0 = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout)
    gv1, gl1, ga1 = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout, dense=(c["levels"], rng))
    torch.cuda.synchronize()
    used = row_map >= 0                                             # unused rows are never written (torch.empty)
    assert torch.equal(gl0[used], gl1[used]) and torch.equal(ga0[used], ga1[used])   # same kernel, same arithmetic
    err = rel_err(gv1.cpu(), gv0.cpu())
    lsl = lsi.tolist() + [int(v.shape[1])]
    per_level = [rel_err(gv1[:, lsl[i]:lsl[i + 1]].cpu(), gv0[:, lsl[i]:lsl[i + 1]].cpu()) for i in range(len(c["levels"]))]
    print(case, "grad_value dense vs plain:", err, per_level)
    assert err < 1e-2, per_level
    # the unused rows and the maps' own ranges only: nothing may leak into a neighbouring map
    assert torch.isfinite(gv1).all()
    # accumulation into a running sum, as the plain entry point does
    gv2, _, _ = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout, grad_value=gv1.clone(),
                                       dense=(c["levels"], rng))
    assert rel_err(gv2.cpu(), 2 * gv0.cpu()) < 1e-2


def test_dense_backward_stale_host_shapes_degrade_to_plain():
    """Host shapes that disagree with the device tensor: both kernels see it, every level stays on the reduction
    path, results are those of the plain backward (up to the order of the fp32 reductions)."""
    c = CASES["merged4"]
    v, ss, lsi, loc, attn, row_map, rng = _ragged_case(c["levels"], c["rows"], c["heads"], c["pts"], seed=3)
    vd = v.to(DEV, torch.bfloat16)
    ss, lsi, loc, attn, row_map, rng = (t.to(DEV) for t in (ss, lsi, loc, attn, row_map, rng))
    gout = fixed_projection((loc.shape[0], c["heads"] * 32)).to(DEV, torch.bfloat16)
    gv0, _, _ = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout)
    wrong = [(30, 20), (15, 10), (8, 5), (4, 3)]                     # same pixel counts, transposed
    gv1, _, _ = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout, dense=(wrong, rng))
    torch.cuda.synchronize()
    assert rel_err(gv1.cpu(), gv0.cpu()) < 1e-5


@pytest.mark.parametrize("mode", [1, 2], ids=["same_stream", "second_stream"])
def test_dense_backward_base_rig_against_oracle(mode):
    """The SCA launch of the headline benchmark (44 511 in-view pairs, 4 levels, 8 points, bf16) through the
    dense path, every output element against Oracle-S on the same storage-rounded inputs."""
    from tools.bench_msda import rig_sca_inputs
    lib = _lib.load()
    assert lib.bevf_msda_set_dense_backward(mode) == 0
    try:
        v, ss, lsi, loc, attn, row_map = rig_sca_inputs(DEV)
        w = syn.WORKLOADS["base"]
        per_cam = torch.bincount(row_map[row_map >= 0].long(), minlength=6)
        ends = per_cam.cumsum(0)
        rng = torch.stack([ends - per_cam, ends], 1).to(torch.int32).contiguous()
        vd = v.to(torch.bfloat16)
        gout = fixed_projection((loc.shape[0], 256)).to(DEV, torch.bfloat16)
        gv, gl, ga = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout, dense=(list(w.levels), rng))
        gvp, glp, gap = ops.msda_rows_backward(vd, ss, lsi, loc, attn, row_map, gout)
        torch.cuda.synchronize()
    finally:
        lib.bevf_msda_set_dense_backward(1)        # (the fixture restores the process default)
    used = row_map >= 0
    assert torch.equal(gl[used], glp[used]) and torch.equal(ga[used], gap[used])
    print("dense vs plain", rel_err(gv.cpu(), gvp.cpu()))
    if mode == 2:                   # the oracle comparison once is enough: same kernels, another stream
        assert rel_err(gv.cpu(), gvp.cpu()) < 1e-2
        return
    vr, gr = vd.float().cpu(), gout.float().cpu()
    loc_c, attn_c, rm = loc.cpu(), attn.cpu(), row_map.cpu().long()
    rgv = torch.zeros_like(gvp, device="cpu")
    for cam in range(vr.shape[0]):
        idx = (rm == cam).nonzero().flatten()
        a, _, _ = msda_oracle.msda_backward(vr[cam:cam + 1], ss.cpu(), lsi.cpu(), loc_c[idx][None].contiguous(),
                                            attn_c[idx][None].contiguous(), gr[idx][None].contiguous())
        rgv[cam] = a[0]
    lsl = lsi.tolist() + [int(v.shape[1])]
    per_level = [rel_err(gv[:, lsl[i]:lsl[i + 1]].cpu(), rgv[:, lsl[i]:lsl[i + 1]]) for i in range(4)]
    err = rel_err(gv.cpu(), rgv)
    print("grad_value dense vs Oracle-S:", err, per_level)
    assert err < 1e-2, per_level
