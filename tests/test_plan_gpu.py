"""Device-side SCA pair list (bevf_sca_plan_build) and the frame-valid CUDA graph built on it.

The reference sizes its per-camera index lists with nonzero() + max() on the host in every layer
(spatial_cross_attention.py:138-141).  Here the list is compacted on the device into fixed-capacity
buffers, so one captured graph serves every frame: these tests replay ONE graph with two different camera
rigs and hold each result to the restated reference encoder."""
import copy

import numpy as np
import pytest
import torch

from bevformer_b200 import ops, synthetic as syn
from bevformer_b200.plugin import ScaPlan, build_transformer_layer_sequence
from oracle import torch_ref
from tests.util import max_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _geometry(w, metas):
    l2i = torch.as_tensor(np.asarray([m["lidar2img"] for m in metas], dtype=np.float32)).to(DEV)
    z = (torch.linspace(0.5, 7.5, 4) / 8.0).tolist()
    return ops.point_sampling(l2i, syn.PC_RANGE, z, w.img_hw[0], w.img_hw[1], w.bev_h, w.bev_w, raw_mask=True)


@pytest.mark.parametrize("workload,bs", [("toy", 1), ("toy", 2), ("tiny", 1), ("base", 1)])
@pytest.mark.parametrize("ordered", [True, False])
def test_device_plan_equals_host_plan(workload, bs, ordered):
    w = syn.WORKLOADS[workload]
    metas = syn.make_img_metas(w, bs)
    if bs > 1:   # a different rig for the second sample: the lists still come from sample 0 (quirk 1)
        metas[1]["lidar2img"] = [m.copy() for m in syn.make_lidar2img(w.scale * 0.9, w.num_cams)]
    ref_cam, mask = _geometry(w, metas)
    host = ScaPlan.build(mask.bool(), ref_cam, (w.bev_h, w.bev_w) if ordered else None)
    cap = host.num_pairs + 37
    qorder = ScaPlan.tile_order(w.bev_h, w.bev_w, DEV) if ordered else None
    plan = ScaPlan.build_device(mask, ref_cam, qorder, cap)
    found, over = plan.counters.tolist()
    n = host.num_pairs
    assert (found, over) == (n, 0)
    assert torch.equal(plan.pair_q[:n], host.pair_q) and torch.equal(plan.pair_cam[:n], host.pair_cam)
    assert bool((plan.pair_q[n:] == -1).all()) and bool((plan.pair_cam[n:] == -1).all())
    assert torch.equal(plan.pair_of, host.pair_of)
    assert torch.equal(plan.inv_count, host.inv_count)
    # row ranges per value map: host plan rows are b*n + r, device plan rows b*cap + r
    hr = host.map_range.view(bs, -1, 2) - (torch.arange(bs, device=DEV) * n).view(bs, 1, 1).int()
    dr = plan.map_range.view(bs, -1, 2) - (torch.arange(bs, device=DEV) * cap).view(bs, 1, 1).int()
    assert torch.equal(hr, dr)
    rm = plan.row_map.view(bs, cap)
    assert torch.equal(rm[:, :n].reshape(-1), host.row_map) and bool((rm[:, n:] == -1).all())


def test_device_plan_reports_overflow():
    w = syn.WORKLOADS["tiny"]
    ref_cam, mask = _geometry(w, syn.make_img_metas(w, 1))
    n = ScaPlan.build(mask.bool(), ref_cam).num_pairs
    plan = ScaPlan.build_device(mask, ref_cam, None, n - 5)
    assert plan.counters.tolist() == [n, 1]
    assert int((plan.pair_of >= 0).sum()) == n - 5          # the excess is dropped, never written out of bounds


def _other_rig(metas, w):
    """A second frame: the rig yawed by 9 degrees and moved, so other queries are in view."""
    out = copy.deepcopy(metas)
    ang = np.deg2rad(9.0)
    rot = np.eye(4)
    rot[:2, :2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    rot[:3, 3] = [1.5, -0.7, 0.1]
    for m in out:
        m["lidar2img"] = [np.asarray(a) @ rot for a in m["lidar2img"]]
    return out


@pytest.mark.parametrize("workload", ["toy", "tiny"])
def test_one_graph_two_camera_rigs(workload):
    """forward + backward captured once (point sampling, plan construction and all layers inside the
    graph), replayed with two lidar2img sets: each replay matches the restated reference encoder run
    with that rig."""
    w = syn.WORKLOADS[workload]
    sd = syn.make_state_dict(w)
    enc = build_transformer_layer_sequence(syn.encoder_cfg(w))
    enc.load_state_dict(sd)
    enc = enc.to(DEV).eval()
    inp = syn.make_encoder_inputs(w)
    metas_a = inp.img_metas
    metas_b = _other_rig(metas_a, w)
    dev_in = {k: getattr(inp, k).to(DEV) for k in ("bev_query", "feat", "bev_pos", "prev_bev", "shift")}
    dev_in["bev_query"].requires_grad_(True)
    ss, lsi = inp.spatial_shapes.to(DEV), inp.level_start_index.to(DEV)
    l2i = torch.zeros(1, w.num_cams, 4, 4, device=DEV)
    proj = torch.randn(1, w.num_query, 256, generator=torch.Generator().manual_seed(3)).to(DEV)

    def set_rig(metas):
        l2i.copy_(torch.as_tensor(np.asarray([m["lidar2img"] for m in metas], dtype=np.float32)))

    def body():
        dev_in["bev_query"].grad = None
        out = enc(dev_in["bev_query"], dev_in["feat"], dev_in["feat"], bev_h=w.bev_h, bev_w=w.bev_w,
                  bev_pos=dev_in["bev_pos"], spatial_shapes=ss, level_start_index=lsi,
                  prev_bev=dev_in["prev_bev"], shift=dev_in["shift"], img_metas=metas_a, lidar2img=l2i)
        (out * proj).sum().backward()
        return out

    set_rig(metas_a)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()                                              # eager warm-up: sizes the pair list (one sync)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    dev_in["bev_query"].grad = None
    with torch.cuda.graph(g):
        out = body()
    results = {}
    for name, metas in (("a", metas_a), ("b", metas_b), ("a2", metas_a)):
        set_rig(metas)
        g.replay()
        torch.cuda.synchronize()
        results[name] = (out.detach().cpu().clone(), dev_in["bev_query"].grad.detach().cpu().clone())
    enc.check_plan()
    assert torch.equal(results["a"][0], results["a2"][0])   # the graph holds no per-frame state
    assert max_err(results["a"][0], results["b"][0]) > 1e-2   # and the rigs really differ
    for name, metas in (("a", metas_a), ("b", metas_b)):
        q = inp.bev_query.clone().requires_grad_(True)
        kw = dict(inp.kwargs(), img_metas=metas)
        ref = torch_ref.encoder_forward(sd, w.num_layers, q, inp.feat, use_c_oracle=True, **kw)
        (ref * proj.cpu()).sum().backward()
        assert max_err(results[name][0], ref) < 1e-3, name
        gerr = (results[name][1] - q.grad).norm() / q.grad.norm()
        assert gerr < 5e-3, (name, float(gerr))
