"""Host logic of the dense tensor-core backward (csrc/msda_dense.cu): the pixel-bin plan.  No GPU needed."""
import ctypes

import numpy as np
import pytest

from bevformer_b200 import _lib


def plan(levels, max_pix=8192, tiles=16):
    lib = _lib.load()
    hw = np.asarray(levels, dtype=np.int32).reshape(-1)
    out = np.full((12, 7), -7, dtype=np.int32)
    mask = ctypes.c_uint32(0)
    n = lib.bevf_msda_dense_plan(hw.ctypes.data, len(levels), max_pix, tiles, out.ctypes.data, 12,
                                 ctypes.addressof(mask))
    assert n >= 0, lib.bevf_last_error()
    return out[:n], mask.value


@pytest.mark.parametrize("tiles", [8, 16])
@pytest.mark.parametrize("levels", [
    [(116, 200), (58, 100), (29, 50), (15, 25)],          # bevformer_base
    [(92, 160), (46, 80), (23, 40), (12, 20)],            # bevformer_small (4-level variant)
    [(15, 25)],                                           # bevformer_tiny
    [(64, 100), (32, 50), (16, 25), (13, 160)],
    [(20, 30), (10, 15), (5, 8), (3, 4)],
    [(100, 120), (29, 50), (15, 25)],
    [(3, 4), (200, 200), (2, 2), (1, 1), (1, 1), (1, 1), (1, 1)],
])
def test_bins_partition_the_covered_levels(levels, tiles):
    bins, mask = plan(levels, tiles=tiles)
    cap = tiles * 128
    sizes = [h * w for h, w in levels]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    covered = np.zeros(starts[-1], dtype=np.int32)
    for s0, n, nlev, *lev in bins:
        assert 0 < n <= cap and 1 <= nlev <= 4
        covered[s0:s0 + n] += 1
        ls = [l for l in lev if l >= 0]
        assert len(ls) == nlev and ls == sorted(ls)
        # the listed levels are exactly those that intersect the bin
        hit = [l for l in range(len(levels)) if starts[l] < s0 + n and starts[l + 1] > s0]
        assert hit == ls
    for l, sz in enumerate(sizes):
        want = 1 if (mask >> l) & 1 else 0
        assert (covered[starts[l]:starts[l + 1]] == want).all()
        assert ((mask >> l) & 1) == (1 if sz <= 8192 else 0)


def test_base_plan_is_what_the_design_says():
    bins, mask = plan([(116, 200), (58, 100), (29, 50), (15, 25)])
    assert mask == 0b1110                                  # level 0 (23 200 pixels) stays on the reduction path
    assert [int(b[1]) for b in bins] == [2048, 2048, 1704, 1825]
    assert [int(b[2]) for b in bins] == [1, 1, 1, 2]
    bins2, mask2 = plan([(116, 200), (58, 100), (29, 50), (15, 25)], max_pix=2048)
    assert mask2 == 0b1100 and len(bins2) == 1


def test_gv_mode_policy(monkeypatch):
    """ops.gv_mode_for: which pyramid levels accumulate grad_value in scaled fp16 (host logic, no GPU)."""
    from bevformer_b200 import ops
    monkeypatch.delenv("BEVF_GV_ACC", raising=False)
    monkeypatch.delenv("BEVF_GV_MAXCONTRIB", raising=False)
    base = [(116, 200), (58, 100), (29, 50), (15, 25)]
    # SCA at base: 44 511 pairs over 6 cameras, 8 points -> 10 / 41 / 164 / 630 contributions per pixel
    mode = ops.gv_mode_for(44511 / 6, 8, base)
    assert mode[0] == "mixed" and mode[2] == 2 and mode[1] == base
    # TSA at base: 40 000 rows per BEV map, 4 points, one level of 40 000 pixels -> 16
    assert ops.gv_mode_for(40000.0, 4, [(1, 40000)]) == "f16"
    # bevformer_tiny: one 15 x 25 level, thousands of contributions per pixel -> fp32
    assert ops.gv_mode_for(2500.0, 8, [(15, 25)]) is None
    # only a prefix may be fp16: a coarse level in front keeps everything in fp32
    assert ops.gv_mode_for(44511 / 6, 8, [(15, 25), (116, 200)]) is None
    monkeypatch.setenv("BEVF_GV_MAXCONTRIB", "16")
    assert ops.gv_mode_for(44511 / 6, 8, base)[2] == 1
    monkeypatch.setenv("BEVF_GV_ACC", "fp32")
    assert ops.gv_mode_for(40000.0, 4, [(1, 40000)]) is None
