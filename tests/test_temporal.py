"""Temporal plumbing: the history-BEV recurrence of training and the streaming state of
video inference, against golden vectors produced by the reference's own PerceptionTransformer driven the
way detectors/bevformer.py:158-177 and :236-269 drive it (tests/golden/make_golden.py::temporal_case)."""
import copy

import numpy as np
import pytest
import torch

from bevformer_b200 import synthetic as syn
from bevformer_b200.plugin import BEVStream, obtain_history_bev
from tests.golden.make_golden import grid_length_of, sequence_inputs
from tests.util import golden, max_err, rel_err


class _Recorder(torch.nn.Module):
    """Stands in for the transformer: records what the temporal drivers feed the encoder (CPU logic test)."""

    def __init__(self):
        super().__init__()
        self.calls = []

    def get_bev_features(self, feats, bev_queries, bev_h, bev_w, grid_length=None, bev_pos=None, prev_bev=None,
                         img_metas=None):
        self.calls.append(dict(prev=None if prev_bev is None else float(prev_bev.flatten()[0]),
                               can_bus=np.array(img_metas[0]["can_bus"], dtype=np.float64).copy(),
                               training=self.training, grad=torch.is_grad_enabled(),
                               feat0=float(feats[0].flatten()[0])))
        return torch.full((1, bev_h * bev_w, 4), float(len(self.calls)))


def test_stream_state_machine_cpu():
    w = syn.WORKLOADS["toy"]
    feats, q, pos, metas = sequence_inputs(w, frames=4)
    rec = _Recorder().train()
    stream = BEVStream(rec)
    before = copy.deepcopy(metas)
    for i in range(4):
        stream.step([f[:, i] for f in feats], [metas[i]], q, w.bev_h, w.bev_w, pos)
    c = rec.calls
    assert [x["prev"] for x in c] == [None, 1.0, 2.0, None]           # history dropped at the scene change
    assert not any(x["training"] or x["grad"] for x in c) and rec.training    # eval + no_grad inside, mode restored
    assert np.all(c[0]["can_bus"][:3] == 0) and c[0]["can_bus"][-1] == 0      # first frame of a scene: zero ego motion
    for i in (1, 2):                                                          # later frames: deltas to the previous frame
        assert np.allclose(c[i]["can_bus"][:3], before[i]["can_bus"][:3] - before[i - 1]["can_bus"][:3])
        assert np.isclose(c[i]["can_bus"][-1], before[i]["can_bus"][-1] - before[i - 1]["can_bus"][-1])
    assert np.all(c[3]["can_bus"][:3] == 0) and c[3]["can_bus"][-1] == 0
    for a, b in zip(metas, before):                                           # caller's metas untouched
        assert np.array_equal(a["can_bus"], b["can_bus"])
    rec2 = _Recorder()
    s2 = BEVStream(rec2, video_test_mode=False)
    for i in range(3):
        s2.step([f[:, i] for f in feats], [metas[i]], q, w.bev_h, w.bev_w, pos)
    assert [x["prev"] for x in rec2.calls] == [None, None, None]              # bevformer.py:249-251


def test_history_recurrence_cpu():
    w = syn.WORKLOADS["toy"]
    feats, q, pos, metas = sequence_inputs(w, frames=4)
    rec = _Recorder().train()
    metas[2]["prev_bev_exists"] = False                                       # a cut inside the queue
    out = obtain_history_bev(rec, feats, [metas], q, w.bev_h, w.bev_w, pos)
    c = rec.calls
    assert len(c) == 4 and [x["prev"] for x in c] == [None, 1.0, None, 3.0]
    assert [x["feat0"] for x in c] == [float(feats[0][0, i].flatten()[0]) for i in range(4)]
    assert not any(x["training"] or x["grad"] for x in c) and rec.training
    assert float(out.flatten()[0]) == 4.0


def _transformer(w, dtype):
    from bevformer_b200.plugin import PerceptionTransformer
    m = PerceptionTransformer(num_feature_levels=len(w.levels), num_cams=w.num_cams, encoder=syn.encoder_cfg(w),
                              decoder=None, embed_dims=w.embed_dims, rotate_center=[w.bev_h // 2, w.bev_w // 2])
    m.load_state_dict(syn.make_perception_state_dict(w))
    return m.to("cuda", dtype).train()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_history_and_stream_against_reference_golden(dtype):
    g = golden("temporal_toy")
    w = syn.WORKLOADS["toy"]
    frames = int(g["meta"][0])
    feats, q, pos, metas = sequence_inputs(w, frames)
    m = _transformer(w, dtype)
    dfeats = [f.to("cuda", dtype) for f in feats]
    dq, dpos = q.to("cuda", dtype), pos.to("cuda", dtype)
    tol = 1e-3 if dtype == torch.float32 else 6e-2        # bf16: the encoder-level bar (tests/test_encoder_gpu.py)
    # training-time history: the dataset hands over can_bus deltas (nuscenes_dataset.py:86-103)
    tm = copy.deepcopy(metas[: frames - 1])
    for i in range(len(tm) - 1, 0, -1):
        tm[i]["can_bus"][:3] -= tm[i - 1]["can_bus"][:3]
        tm[i]["can_bus"][-1] -= tm[i - 1]["can_bus"][-1]
    tm[0]["can_bus"][:3] = 0
    tm[0]["can_bus"][-1] = 0
    hist = obtain_history_bev(m, [f[:, : frames - 1] for f in dfeats], [tm], dq, w.bev_h, w.bev_w, dpos,
                              grid_length_of(w))
    assert m.training and not hist.requires_grad
    assert rel_err(hist.float().cpu(), g["history"]) < tol
    stream = BEVStream(m)
    for i in range(frames):
        bev = stream.step([f[:, i] for f in dfeats], [metas[i]], dq, w.bev_h, w.bev_w, dpos, grid_length_of(w))
        assert rel_err(bev.float().cpu(), g[f"stream{i}"]) < tol, i
    # the last frame opened a new scene: its BEV must not depend on the frames before it
    fresh = BEVStream(m).step([f[:, frames - 1] for f in dfeats], [metas[frames - 1]], dq, w.bev_h, w.bev_w, dpos,
                              grid_length_of(w))
    assert max_err(fresh, bev) == 0.0
