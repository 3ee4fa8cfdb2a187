"""BEVFormerV2's BEV-encoder wrapper: restatement against the reference's own
class on CPU, CUDA drop-in against the restatement on GPU -- with and without the BEV augmentation
resampling branch."""
import numpy as np
import pytest
import torch

from bevformer_b200 import synthetic as syn
from oracle import mmcv_stub, torch_ref
from tests.util import max_err, rel_err

W = syn.WORKLOADS["toy"]


def _metas(bs, aug):
    metas = syn.make_img_metas(W, bs)
    if aug is not None:
        a = np.deg2rad(12.0)
        mat = torch.tensor([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0],
                            [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float32) * 1.05
        for m in metas:
            m["aug_param"] = {"GlobalRotScaleTransImage_param": (12.0, 1.05, False, False, mat, aug == "only_gt")}
    return metas


def _sd(dtype=torch.float32):
    sd = syn.make_perception_state_dict(W)
    return {k: v.to(dtype) for k, v in sd.items()
            if k.startswith("encoder.") or k in ("level_embeds", "cams_embeds")}


def _restatement(inp, metas, sd):
    return torch_ref.bev_encoder_v2(sd, W.num_layers, inp.mlvl_feats, inp.bev_queries, W.bev_h, W.bev_w,
                                    bev_pos=inp.bev_pos, img_metas=metas, tsa_points=W.tsa_points,
                                    sca_points=W.sca_points)


@pytest.mark.skipif(not mmcv_stub.reference_available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("aug", [None, "only_gt", "images_too"])
@pytest.mark.parametrize("bs", [1, 2])
def test_restatement_vs_reference_class_fp64(aug, bs):
    if aug == "only_gt" and bs > 1:
        pytest.skip("the reference's resampling branch builds a batch-1 grid: it only runs with 1 sample per GPU")
    cls = mmcv_stub.load_reference_transformer_v2()
    m = cls(num_feature_levels=len(W.levels), num_cams=W.num_cams, encoder=syn.encoder_cfg(W),
            embed_dims=W.embed_dims).double().eval()
    sd = _sd(torch.float64)
    m.load_state_dict(sd)
    inp = syn.make_perception_inputs(W, bs=bs, dtype=torch.float64)
    metas = _metas(bs, aug)
    with torch.no_grad():
        want = m(inp.mlvl_feats, inp.bev_queries, W.bev_h, W.bev_w, bev_pos=inp.bev_pos, prev_bev=inp.prev_bev,
                 img_metas=metas)
        got = _restatement(inp, metas, sd)
    assert max_err(got, want) < 1e-9


@pytest.mark.skipif(not mmcv_stub.reference_available(), reason="/root/reference not mounted")
def test_dropin_parameters_match_reference():
    from bevformer_b200.plugin import PerceptionTransformerBEVEncoder
    kw = dict(num_feature_levels=len(W.levels), num_cams=W.num_cams, encoder=syn.encoder_cfg(W), embed_dims=W.embed_dims)
    for extra in (dict(), dict(use_cams_embeds=False)):
        a = PerceptionTransformerBEVEncoder(**kw, **extra)
        b = mmcv_stub.load_reference_transformer_v2()(**kw, **extra)
        assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in b.state_dict().items()}


def test_no_cpu_path():
    from bevformer_b200.plugin import PerceptionTransformerBEVEncoder
    m = PerceptionTransformerBEVEncoder(num_feature_levels=len(W.levels), num_cams=W.num_cams,
                                        encoder=syn.encoder_cfg(W), embed_dims=W.embed_dims)
    inp = syn.make_perception_inputs(W, bs=1)
    with pytest.raises(RuntimeError):
        m(inp.mlvl_feats, inp.bev_queries, W.bev_h, W.bev_w, bev_pos=inp.bev_pos, img_metas=inp.img_metas)


@pytest.mark.gpu
@pytest.mark.parametrize("aug", [None, "only_gt", "images_too"])
def test_dropin_on_gpu_vs_restatement(aug):
    from bevformer_b200.plugin import PerceptionTransformerBEVEncoder
    bs = 2
    m = PerceptionTransformerBEVEncoder(num_feature_levels=len(W.levels), num_cams=W.num_cams,
                                        encoder=syn.encoder_cfg(W), embed_dims=W.embed_dims)
    sd = _sd()
    m.load_state_dict(sd)
    m = m.cuda().eval()
    metas = _metas(bs, aug)
    cpu = syn.make_perception_inputs(W, bs=bs)
    with torch.no_grad():
        want = _restatement(cpu, metas, sd)
    dev = syn.make_perception_inputs(W, bs=bs, device="cuda")
    with torch.no_grad():
        got = m(dev.mlvl_feats, dev.bev_queries, W.bev_h, W.bev_w, bev_pos=dev.bev_pos, prev_bev=dev.prev_bev,
                img_metas=metas)
    assert got.shape == want.shape
    assert rel_err(got.cpu(), want) < 1e-3
