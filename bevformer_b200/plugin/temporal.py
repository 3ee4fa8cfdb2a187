"""Temporal plumbing around the BEV encoder: the history-BEV recurrence of training and the streaming
state of video inference.

Reference: ``BEVFormer.obtain_history_bev`` (projects/mmdet3d_plugin/bevformer/detectors/bevformer.py:158-177)
runs the encoder over the ``queue_length - 1`` earlier frames without gradients (3 extra encoder passes per
training step at base) and ``BEVFormer.forward_test`` (:236-269, state initialised at :59-64) carries
``prev_frame_info`` from frame to frame.  Both reach the encoder through
``pts_bbox_head(..., only_bev=True)`` -> ``transformer.get_bev_features`` (dense_heads/bevformer_head.py);
here they call ``PerceptionTransformer.get_bev_features`` of this package directly: the detector and head
classes themselves (backbone, FPN, query embeddings, losses) are outside the hot path.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import numpy as np
import torch


def obtain_history_bev(transformer, feats_queue: Sequence[torch.Tensor], img_metas_list, bev_queries,
                       bev_h: int, bev_w: int, bev_pos, grid_length=(0.512, 0.512)) -> Optional[torch.Tensor]:
    """The no-grad recurrence over the history frames (bevformer.py:158-177).

    feats_queue   per pyramid level (bs, len_queue, num_cams, C, h, w) -- what ``extract_feat(img,
                  len_queue=len_queue)`` returns (:150-156)
    img_metas_list  per sample, indexable by frame: ``img_metas_list[b][i]`` is frame i's meta dict with
                  ``prev_bev_exists`` / ``can_bus`` / ``lidar2img`` / ``img_shape`` (:169-171)
    Returns the BEV of the last history frame (bs, Nq, C), or None for an empty queue.  The transformer
    is put in eval mode for the loop and returned to the mode it was in (the reference calls
    ``self.train()`` unconditionally, :176)."""
    was_training = transformer.training
    transformer.eval()
    try:
        with torch.no_grad():
            prev_bev = None
            len_queue = feats_queue[0].shape[1]
            for i in range(len_queue):
                img_metas = [each[i] for each in img_metas_list]
                if not img_metas[0]["prev_bev_exists"]:               # :170-171, sample 0 decides for the batch
                    prev_bev = None
                img_feats = [lvl[:, i] for lvl in feats_queue]
                prev_bev = transformer.get_bev_features(img_feats, bev_queries, bev_h, bev_w,
                                                        grid_length=list(grid_length), bev_pos=bev_pos,
                                                        prev_bev=prev_bev, img_metas=img_metas)
            return prev_bev
    finally:
        transformer.train(was_training)


class BEVStream:
    """Streaming inference state: the previous frame's BEV, ego position and heading
    (``prev_frame_info``, bevformer.py:59-64) and the per-frame update of forward_test (:236-269):
    a new scene (or ``video_test_mode=False``) drops the history; CAN-bus position / angle are turned into
    deltas against the previous frame before the encoder sees them, zeros on a scene's first frame."""

    def __init__(self, transformer, video_test_mode: bool = True):
        self.transformer = transformer
        self.video_test_mode = video_test_mode
        self.reset()

    def reset(self) -> None:
        self.prev_frame_info = {"prev_bev": None, "scene_token": None, "prev_pos": 0, "prev_angle": 0}

    @torch.no_grad()
    def step(self, mlvl_feats: List[torch.Tensor], img_metas: List[dict], bev_queries, bev_h: int, bev_w: int,
             bev_pos, grid_length=(0.512, 0.512)) -> torch.Tensor:
        """One frame: mlvl_feats per level (bs, num_cams, C, h, w), img_metas one dict per sample with
        ABSOLUTE ``can_bus`` and a ``scene_token``.  Returns this frame's BEV (bs, Nq, C), which also becomes
        the next frame's ``prev_bev``.  The caller's metas are not modified (the reference edits them in
        place, :254-261; the values the encoder sees are the same)."""
        info = self.prev_frame_info
        if img_metas[0].get("scene_token") != info["scene_token"]:
            info["prev_bev"] = None                                  # :243-245
        info["scene_token"] = img_metas[0].get("scene_token")
        if not self.video_test_mode:
            info["prev_bev"] = None                                  # :249-251
        metas = [dict(m) for m in img_metas]
        can_bus = np.array(metas[0]["can_bus"], dtype=np.float64, copy=True)
        tmp_pos, tmp_angle = can_bus[:3].copy(), copy.deepcopy(can_bus[-1])      # :254-255
        if info["prev_bev"] is not None:
            can_bus[:3] -= info["prev_pos"]                          # :257-258
            can_bus[-1] -= info["prev_angle"]
        else:
            can_bus[-1] = 0                                          # :260-261
            can_bus[:3] = 0
        metas[0]["can_bus"] = can_bus
        was_training = self.transformer.training
        self.transformer.eval()
        try:
            bev = self.transformer.get_bev_features(mlvl_feats, bev_queries, bev_h, bev_w,
                                                    grid_length=list(grid_length), bev_pos=bev_pos,
                                                    prev_bev=info["prev_bev"], img_metas=metas)
        finally:
            self.transformer.train(was_training)
        info["prev_pos"], info["prev_angle"], info["prev_bev"] = tmp_pos, tmp_angle, bev    # :266-268
        return bev
