/*
 * bevformer_b200 -- C ABI of the B200-native BEV-encoder hot path.
 *
 * This is the drop-in boundary.  The reference reaches its native code through the
 * pybind module `mmcv._ext` (third-party wheel mmcv-full==1.4.0, not in the reference tree):
 *
 *   ext_module.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc,
 *                                     attn_weight, im2col_step) -> Tensor
 *       projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:118-124
 *   ext_module.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc,
 *                                      attn_weight, grad_output, grad_value, grad_sampling_loc,
 *                                      grad_attn_weight, im2col_step) -> None   (in place)
 *       .../multi_scale_deformable_attn_function.py:150-160
 *
 * Those two calls are replaced 1:1 by bevf_msda_forward / bevf_msda_backward below.  The remaining
 * entry points are the fused pieces of the encoder layer that the reference spells as ATen /
 * cuBLAS launches inside TemporalSelfAttention / SpatialCrossAttention / BEVFormerLayer /
 * BEVFormerEncoder (each cites the Python lines it replaces).
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, an opaque cudaStream_t passed as void*; no torch types.
 *   - the caller owns every buffer; the library never allocates, frees or synchronises, and keeps no
 *     mutable global state besides a thread-local error string, so every call is re-entrant and
 *     safe under CUDA-graph capture.
 *   - every function returns 0 on success, non-zero on error; bevf_last_error() then holds a
 *     message for the calling thread (the Python wrapper raises RuntimeError with it, which is
 *     what mmcv's TORCH_CHECK failures surface as).
 *   - tensors are dense row-major in the layouts named per function; "dtype" arguments take the
 *     BEVF_DTYPE_* codes.  Device pointers must be 16-byte aligned.
 *   - there is NO CPU implementation behind this ABI.
 */
#ifndef BEVFORMER_B200_H_
#define BEVFORMER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEVF_ABI_VERSION 1

#if defined(__GNUC__)
#define BEVF_API __attribute__((visibility("default")))
#else
#define BEVF_API
#endif

enum bevf_dtype { BEVF_DTYPE_F32 = 0, BEVF_DTYPE_BF16 = 1 };

/* ABI version of the loaded library (== BEVF_ABI_VERSION it was built with). */
BEVF_API int bevf_version(void);

/* Message of the last failing call on this thread ("" if none). Never NULL. */
BEVF_API const char *bevf_last_error(void);

/* Number of kernel launches issued through this library by the calling process so far
 * (bench.py reports it as gpu_launches). */
BEVF_API int64_t bevf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention sampler.
 *
 * replaces: mmcv._ext.ms_deform_attn_forward, as called at
 *   multi_scale_deformable_attn_function.py:118-124 (from temporal_self_attention.py:247,
 *   spatial_cross_attention.py:390, decoder.py:332).
 *
 *   value        (B, S, M, D)        value_dtype (f32 | bf16)
 *   level_hw     (L, 2) int64 DEVICE (h, w) per level          -- the reference's spatial_shapes
 *   level_start  (L,)   int64 DEVICE first row of each level   -- the reference's level_start_index
 *   loc          (B, Q, M, L, P, 2) f32, (x, y) normalised to [0,1] over each level
 *   attn         (B, Q, M, L, P)    f32
 *   out          (B, Q, M*D)        out_dtype (f32 | bf16), fully overwritten
 *
 * out[b,q,m,:] = sum_l sum_p attn * bilinear(value_l, x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5),
 * zero padding, a sample contributes only if -1 < x < W_l and -1 < y < H_l.
 * Accumulation is fp32 for both value dtypes.  `im2col_step` of the reference has no equivalent:
 * the whole batch is one launch.  L <= 16.
 */
BEVF_API int bevf_msda_forward(const void *value, int value_dtype, const int64_t *level_hw,
                      const int64_t *level_start, const float *loc, const float *attn, void *out,
                      int out_dtype, int B, int S, int M, int D, int Q, int L, int P, void *stream);

/*
 * replaces: mmcv._ext.ms_deform_attn_backward, as called at
 *   multi_scale_deformable_attn_function.py:150-160.
 *
 *   grad_out    (B, Q, M*D)  grad_out_dtype (f32 | bf16)
 *   grad_value  (B, S, M, D) f32 -- ACCUMULATED INTO (caller zero-fills, as the reference does at
 *                                   multi_scale_deformable_attn_function.py:146)
 *   grad_loc    (B, Q, M, L, P, 2) f32 -- fully overwritten (zeros for skipped samples)
 *   grad_attn   (B, Q, M, L, P)    f32 -- fully overwritten
 * grad_value uses vector fp32 reductions in L2; its summation order is not deterministic.
 */
BEVF_API int bevf_msda_backward(const void *value, int value_dtype, const int64_t *level_hw,
                       const int64_t *level_start, const float *loc, const float *attn,
                       const void *grad_out, int grad_out_dtype, float *grad_value,
                       float *grad_loc, float *grad_attn, int B, int S, int M, int D, int Q, int L,
                       int P, void *stream);

/*
 * Row-list form of the sampler: R query rows, each sampling the value map named by row_map[r]
 * (0 <= row_map[r] < B).  bevf_msda_forward is the special case row_map[r] = r / Q.
 *
 * replaces: the zero-padded per-camera re-batching around the op in
 *   spatial_cross_attention.py:138-167 (nonzero -> max_len -> queries_rebatch -> op -> slots +=):
 *   SpatialCrossAttention hands the op only the (camera, query) pairs that are actually in view,
 *   in one launch, instead of num_cams x max_len padded rows.
 *
 *   loc (R, M, L, P, 2) f32; attn (R, M, L, P) f32; out / grad_out (R, M*D); row_map (R,) int32 DEVICE.
 * Other arguments and semantics as for bevf_msda_forward / bevf_msda_backward.
 */
BEVF_API int bevf_msda_rows_forward(const void *value, int value_dtype, const int64_t *level_hw,
                                    const int64_t *level_start, const float *loc, const float *attn,
                                    void *out, int out_dtype, const int32_t *row_map, int B, int S,
                                    int M, int D, int R, int L, int P, void *stream);

BEVF_API int bevf_msda_rows_backward(const void *value, int value_dtype, const int64_t *level_hw,
                                     const int64_t *level_start, const float *loc,
                                     const float *attn, const void *grad_out, int grad_out_dtype,
                                     float *grad_value, float *grad_loc, float *grad_attn,
                                     const int32_t *row_map, int B, int S, int M, int D, int R,
                                     int L, int P, void *stream);

/*
 * Same as bevf_msda_rows_backward, plus `group_order` (R,) int32 (or NULL = identity): a permutation of
 * the rows that lists them so that runs of 64 consecutive entries are spatial neighbours on ONE value map
 * (for the interleaved TSA rows: 8x8 BEV tiles of one frame).  The grad_value half of the backward
 * merges the contributions of such a run in registers before they reach L2 (csrc/msda_splat.cuh); the
 * order changes which additions are merged, never the result beyond fp32 summation order.
 * Level maps must satisfy H_l, W_l < 32768.
 */
BEVF_API int bevf_msda_rows_backward_ordered(const void *value, int value_dtype, const int64_t *level_hw,
                                             const int64_t *level_start, const float *loc,
                                             const float *attn, const void *grad_out,
                                             int grad_out_dtype, float *grad_value, float *grad_loc,
                                             float *grad_attn, const int32_t *row_map,
                                             const int32_t *group_order, int B, int S, int M, int D,
                                             int R, int L, int P, void *stream);

/*
 * grad_value accumulated in SCALED fp16 instead of fp32.  The sampler backward is bound by the L2 reduction-sector
 * rate; an fp16 running sum needs half the sectors (one 16-byte f16x2 vector reduction per lane and corner).
 * fp16 needs a scale: bevf_abs_max puts max|grad_out| (as float bits) into one device word, every kernel of the
 * path derives the same power of two from it (the maximum lands in [8, 16): sums of thousands of contributions stay
 * below 65504, values down to 2^-17 of the maximum stay normal numbers).  The running sum is rounded to 11 bits at
 * every addition, so this is for maps / levels where a (pixel, head) collects few contributions; measured errors
 * in tests/test_msda_gpu.py (bf16 accumulation -- what the reference's fp16 op class does,
 * multi_scale_deformable_attn_function.py:146-160 -- was measured at 1.4e-2 on the TSA launch, above the bar).
 *
 *   bevf_abs_max(x, dtype, n, amax_bits)      *amax_bits = float bits of max|x| (n a multiple of 8 / 4 elements)
 *   bevf_msda_rows_backward_f16acc            bevf_msda_rows_backward_ordered with grad_value_f16 (B, S, M, D) fp16,
 *                                             zero-filled or holding a running sum in the same scale
 *   bevf_gv16_unscale(gv16, amax, out, n)     out (bf16) = gv16 / scale
 *   bevf_msda_rows_backward_mixed             the first num_f16_levels pyramid levels (pixels [0, S_fine) of every
 *                                             map) in scaled fp16 into grad_value_fine_f16 (B, S_fine, M, D), the
 *                                             others in fp32 into grad_value_side (B, S - S_fine, M, D); S_fine is
 *                                             taken from level_hw_host (L, 2) int32 HOST, the kernel re-derives which
 *                                             levels lie before / after it from the DEVICE pyramid (a device level
 *                                             that straddles S_fine is a caller bug and traps)
 *   bevf_gv_merge(fine, side, amax, out, B, S, S_fine, row_elems)   out (B, S, row_elems) bf16 from both
 * All need a bf16 value tensor and head_dim 32; grad_loc / grad_attn are those of the fp32 path bit for bit.
 */
BEVF_API int bevf_abs_max(const void *x, int dtype, int64_t n, uint32_t *amax_bits, void *stream);
BEVF_API int bevf_msda_rows_backward_f16acc(const void *value, int value_dtype, const int64_t *level_hw,
                                            const int64_t *level_start, const float *loc, const float *attn,
                                            const void *grad_out, int grad_out_dtype, void *grad_value_f16,
                                            const uint32_t *amax_bits, float *grad_loc, float *grad_attn,
                                            const int32_t *row_map, const int32_t *group_order, int B, int S,
                                            int M, int D, int R, int L, int P, void *stream);
BEVF_API int bevf_gv16_unscale(const void *gv16, const uint32_t *amax_bits, void *out_bf16, int64_t n, void *stream);
BEVF_API int bevf_msda_rows_backward_mixed(const void *value, int value_dtype, const int64_t *level_hw,
                                           const int64_t *level_start, const int32_t *level_hw_host,
                                           const float *loc, const float *attn, const void *grad_out,
                                           int grad_out_dtype, void *grad_value_fine_f16, float *grad_value_side,
                                           const uint32_t *amax_bits, int num_f16_levels, float *grad_loc,
                                           float *grad_attn, const int32_t *row_map, const int32_t *group_order,
                                           int B, int S, int M, int D, int R, int L, int P, void *stream);
BEVF_API int bevf_gv_merge(const void *fine_f16, const float *side_f32, const uint32_t *amax_bits, void *out_bf16,
                           int B, int S, int S_fine, int row_elems, void *stream);

/*
 * bevf_msda_rows_forward with the coarse pyramid levels staged in shared memory by TMA.
 * For row lists whose rows are grouped by value map (the SCA pair list: camera-major): a CTA takes one
 * (value map, head) and a share of that map's rows, loads every level that fits -- coarsest first, whole
 * level of that head, one cp.async.bulk.tensor per level -- and gathers those levels from shared memory;
 * finer levels keep the global path.  Same results as bevf_msda_rows_forward.
 *   level_hw_host  (L, 2) int32 HOST copy of level_hw (the tensor maps are built on the host); a level
 *                  whose host shape disagrees with the device-side level_hw / level_start is not staged
 *   map_range      (B, 2) int32 DEVICE: [first, end) rows of every value map (bevf_sca_plan_build)
 *   B = number of value maps; levels must be stored back to back (level_start[l] = sum of H*W before l,
 *   the reference's own definition, transformer.py:178-180); head_dim 32.
 */
BEVF_API int bevf_msda_rows_forward_staged(const void *value, int value_dtype, const int64_t *level_hw,
                                           const int64_t *level_start, const int32_t *level_hw_host,
                                           const float *loc, const float *attn, void *out, int out_dtype,
                                           const int32_t *map_range, int B, int S, int M, int D, int R,
                                           int L, int P, void *stream);

/*
 * bevf_msda_rows_backward for row lists grouped by value map (the SCA pair list), with grad_value of the
 * COARSE pyramid levels computed as a dense product on the tensor cores (csrc/msda_dense.cu) instead of one
 * L2 reduction per (sample, corner): for a (value map, head) the col2im scatter of
 * multi_scale_deformable_attn_function.py:150-160 is  C[pixel, row] x grad_out[row, 32]  with a sparse
 * coefficient matrix C (attention weight x bilinear weight, rounded to bf16); bins of <= 2048 pixels keep
 * their fp32 accumulators in TMEM, the coefficients of 16 rows at a time are written into shared memory by
 * one thread per (row, level) and multiplied with tcgen05.mma.  Levels with more than BEVF_DENSE_MAXPIX
 * pixels (default 8192), grad_loc and grad_attn come from the one-kernel backward as before.
 * Used when grad_out is bf16 and head_dim is 32 with 4 or 8 points per level; any other configuration
 * runs exactly bevf_msda_rows_backward.  grad_value must be zero-filled or hold a running sum, as there.
 *   level_hw_host  (L, 2) int32 HOST copy of level_hw (the bins are planned on the host; a device-side
 *                  mismatch is a caller bug and traps)
 *   map_range      (B, 2) int32 DEVICE: [first, end) rows of every value map (bevf_sca_plan_build)
 * bevf_msda_set_dense_backward: 0 = never use the dense kernel, 1 = on the caller's stream (default, or the
 * environment variable BEVF_MSDA_DENSE), 2 = on a library-owned second stream next to the reduction kernel
 * (fork / join with events, capturable).
 */
BEVF_API int bevf_msda_rows_backward_dense(const void *value, int value_dtype, const int64_t *level_hw,
                                           const int64_t *level_start, const int32_t *level_hw_host,
                                           const float *loc, const float *attn, const void *grad_out,
                                           int grad_out_dtype, float *grad_value, float *grad_loc,
                                           float *grad_attn, const int32_t *row_map, const int32_t *map_range,
                                           int B, int S, int M, int D, int R, int L, int P, void *stream);
BEVF_API int bevf_msda_set_dense_backward(int mode);
BEVF_API int bevf_msda_get_dense_backward(void);
/*
 * Host-only helper: the pixel bins bevf_msda_rows_backward_dense plans for a pyramid.  bins_out receives 7 int32
 * per bin {first pixel, pixel count, number of levels, level ids (4, -1 padded)}; *level_mask the levels covered;
 * returns the number of bins (0: nothing applies), negative on a bad argument.  tiles = 8 or 16 accumulator
 * tiles of 128 pixels per bin.  No device work.
 */
BEVF_API int bevf_msda_dense_plan(const int32_t *level_hw_host, int L, int max_pix, int tiles, int32_t *bins_out,
                                  int bins_cap, uint32_t *level_mask);

/*
 * Selects how bevf_msda_*backward* computes grad_value (process-wide; default 0, or the environment
 * variable BEVF_MSDA_BWD=split).  0: one kernel, one 16 B-vector L2 reduction per corner contribution.
 * 1: gather kernel + a splat kernel that merges the contributions of 64 neighbouring rows in registers
 * before they reach L2 (fewer reductions, more instructions; kept for A/B measurements).
 * 2: hybrid -- the coarse half of the pyramid through the splat kernel on a library-owned second stream
 * (forked from / joined to the caller's stream with events: capturable), the rest in the one kernel.
 * Results agree up to fp32 summation order.
 */
BEVF_API int bevf_msda_set_backward_mode(int mode);

/* ------------------------------------------------------------------------------------------------
 * Fused memory-bound pieces of one encoder layer.  "raw" is the fp32 output of the layer's combined
 * sampling_offsets|attention_weights GEMM, one row per BEV query.
 * ---------------------------------------------------------------------------------------------- */

/*
 * In-view (camera, query) pair list of SpatialCrossAttention, built on the device without a host
 * synchronisation (capturable in a CUDA graph; a new lidar2img every frame just changes the contents).
 * replaces spatial_cross_attention.py:138-141 (per-camera nonzero() + max_len: two host syncs per layer)
 * and the zero-padded re-batch at :144-153.
 *   bev_mask   (ncam, B, Nq, D) uint8    the in-view mask bevf_point_sampling writes
 *   qorder     (Nq,) int32 or NULL       order of the queries inside a camera's list (e.g. 8x8 BEV tiles)
 *   pair_q, pair_cam (capacity,) int32   out; camera-major; entries >= num_pairs are -1
 *   pair_of    (ncam, Nq) int32          out; row of (cam, q) or -1
 *   row_map    (B*capacity,) int32       out; value map b*ncam+cam of sampler row b*capacity+r, -1 = unused
 *   inv_count  (B, Nq) f32               out; 1 / max(1, #cameras seeing q in batch item b)  (:169-171)
 *   map_range  (B*ncam, 2) int32         out; [first, end) sampler rows of value map b*ncam+cam (its rows are
 *                                        contiguous) -- what bevf_msda_rows_forward_staged partitions by
 *   counters   (2,) int32                out; [0] = number of pairs found, [1] = 1 if it exceeded capacity
 *                                        (the pairs beyond capacity are dropped: the caller must check)
 *   workspace  bevf_sca_plan_workspace_ints(ncam, Nq) int32
 * The lists come from batch item 0's mask for every batch item, as in the reference (:139).  Every
 * row-list entry point of this library skips rows whose pair_q / row_map entry is -1.
 */
BEVF_API int64_t bevf_sca_plan_workspace_ints(int ncam, int Nq);
BEVF_API int bevf_sca_plan_build(const unsigned char *bev_mask, const int32_t *qorder, int32_t *pair_q,
                                 int32_t *pair_cam, int32_t *pair_of, int32_t *row_map, float *inv_count,
                                 int32_t *map_range, int32_t *counters, int32_t *workspace, int B, int ncam,
                                 int Nq, int D, int capacity, void *stream);

/*
 * SCA sampling points.  replaces spatial_cross_attention.py:338-372 (view, softmax over L*P,
 * offset / (W_l, H_l), Z-anchor broadcast "point p uses anchor p mod D", add) for the in-view
 * (camera, query) pairs only.
 *   raw      (B*Nq, M*L*P*3) f32: [offsets (M,L,P,2) | logits (M,L*P)]
 *   ref_cam  (ncam, B, Nq, D, 2) f32       pair_q, pair_cam (R,) int32
 *   loc      (B*R, M, L, P, 2) f32 out      attn (B*R, M, L, P) f32 out   (row = b*R + r)
 */
BEVF_API int bevf_sca_prep_forward(const float *raw, const float *ref_cam, const int32_t *pair_q,
                                   const int32_t *pair_cam, const int64_t *level_hw, float *loc,
                                   float *attn, int B, int Nq, int R, int M, int L, int P, int D,
                                   int ncam, void *stream);

/* Backward of the above into d_raw (B*Nq, M*L*P*3), fully overwritten, stored as out_dtype (f32, or
 * bf16 when it feeds the bf16 dX / dW GEMMs of the head directly); pair_of (ncam, Nq) int32 holds the
 * pair row of (camera, query) or -1. */
BEVF_API int bevf_sca_prep_backward(const float *raw, const float *grad_loc, const float *grad_attn,
                                    const int32_t *pair_of, const int64_t *level_hw, void *d_raw,
                                    int out_dtype, int B, int Nq, int R, int M, int L, int P, int ncam,
                                    void *stream);

/*
 * TSA sampling points.  replaces temporal_self_attention.py:206-229 (view, softmax over L*P per
 * queue entry, the two permute+reshape copies, offset / (W, H) + reference point).
 *   raw    (B*Nq, M*2*L*P*3) f32: [offsets (M,2,L,P,2) | logits (M,2,L*P)]
 *   ref2d  (B*2, Nq, L, 2) f32 (the encoder's hybird_ref_2d)
 *   loc    (B*2, Nq, M, L, P, 2) f32 out     attn (B*2, Nq, M, L, P) f32 out
 *   interleave != 0 orders the output rows (b, q, frame) instead of (b, frame, q): the two frames of a
 *   query become adjacent rows, so the row-list sampler's output is (B*Nq, 2*C) and the average over
 *   the frames (temporal_self_attention.py:257-265) folds into the output projection.
 */
BEVF_API int bevf_tsa_prep_forward(const float *raw, const float *ref2d, const int64_t *level_hw,
                                   float *loc, float *attn, int B, int Nq, int M, int L, int P,
                                   int interleave, void *stream);

/* d_raw (B*Nq, M*2*L*P*3) stored as out_dtype (f32 or bf16), fully overwritten. */
BEVF_API int bevf_tsa_prep_backward(const float *raw, const float *grad_loc, const float *grad_attn,
                                    const int64_t *level_hw, void *d_raw, int out_dtype, int B, int Nq,
                                    int M, int L, int P, int interleave, void *stream);

/*
 * y = LayerNorm(dropout(x) + residual) * gamma + beta, optionally also y_plus_pos = y + pos.
 * replaces the `norm` steps of BEVFormerLayer.forward (encoder.py:377-379) together with the
 * preceding "self.dropout(output) + identity" of the attention / FFN
 * (temporal_self_attention.py:272, spatial_cross_attention.py:175, mmcv FFN) and TSA's
 * `query + query_pos` (temporal_self_attention.py:186-187).
 *   x, residual, pos, y, y_plus_pos: (rows, C) in `dtype`; gamma, beta: (C) in `param_dtype`;
 *   mean, rstd: (rows) f32.  residual / pos / y_plus_pos / mean / rstd may be NULL.  C in {256, 512}.
 *   drop_p in [0,1): inverted dropout on x with keep-mask bits from Philox4x32-10(seed, row*32+lane);
 *   the backward regenerates the same bits from `seed`, no mask tensor exists.  drop_p = 0: no dropout.
 *   seed_base (DEVICE, may be NULL) is added to seed inside the kernel: a step counter kept on the
 *   device lets a captured CUDA graph draw new masks on every replay.
 */
BEVF_API int bevf_layernorm_forward(const void *x, const void *residual, const void *gamma,
                                    const void *beta, int param_dtype, const void *pos, void *y,
                                    void *y_plus_pos, float *mean, float *rstd, int64_t rows, int C,
                                    float eps, float drop_p, uint64_t seed, const uint64_t *seed_base,
                                    int dtype, void *stream);

/* dx (rows, C) = gradient of x, fully overwritten; dres = gradient of residual (may be NULL when
 * drop_p == 0: it then equals dx); dgamma / dbeta (C,) f32 are ACCUMULATED INTO.  dy_plus_pos may be
 * NULL; its rows may be strided (dy_plus_pos_ld elements between rows, 0 = C): it usually arrives as a
 * column slice of the gradient of cat([prev_bev, query + pos]).  drop_p / seed must be the forward call's. */
BEVF_API int bevf_layernorm_backward(const void *x, const void *residual, const void *gamma,
                                     int param_dtype, const float *mean, const float *rstd,
                                     const void *dy, const void *dy_plus_pos, int64_t dy_plus_pos_ld,
                                     void *dx, void *dres, float *dgamma, float *dbeta, int64_t rows,
                                     int C, float drop_p, uint64_t seed, const uint64_t *seed_base,
                                     int dtype, void *stream);

/*
 * slots[b,q,:] = inv_count[b,q] * sum_{cameras seeing q} out[b*R + pair_of[cam][q], :]
 * replaces spatial_cross_attention.py:165-172 (python scatter-add loops, count, divide).
 */
BEVF_API int bevf_sca_combine_forward(const void *out, const int32_t *pair_of,
                                      const float *inv_count, void *slots, int B, int Nq, int R,
                                      int C, int ncam, int dtype, void *stream);

BEVF_API int bevf_sca_combine_backward(const void *g_slots, const int32_t *pair_q,
                                       const float *inv_count, void *g_out, int B, int Nq, int R,
                                       int C, int dtype, void *stream);

/*
 * One pyramid level of camera features into the encoder's key/value layout, embeddings added.
 * replaces PerceptionTransformer.get_bev_features' flatten / permute / +cams_embeds / +level_embeds /
 * cat / permute (transformer.py:161-181).
 *   feat (bs, ncam, C, hw) T in;  cams_embeds (ncam, C) f32 or null;  level_embed (C) f32;
 *   out (ncam, S, bs, C) T: rows [level_start, level_start + hw) of every camera are written.
 */
BEVF_API int bevf_flatten_feats(const void *feat, const float *cams_embeds, const float *level_embed,
                                void *out, int bs, int ncam, int C, int hw, int S, int level_start,
                                int dtype, void *stream);

/*
 * Projection of the pillar anchors into every camera + in-view mask, fp32 without FMA contraction.
 * replaces BEVFormerEncoder.get_reference_points(dim='3d') + point_sampling (encoder.py:46-71,
 * 88-149), including the 61 MB repeated-matrix materialisation.
 *   lidar2img (B, ncam, 4, 4) f32 DEVICE; pc_range (6) and z_norm (D) HOST arrays;
 *   ref_cam (ncam, B, Nq, D, 2) f32 out; bev_mask (ncam, B, Nq, D) uint8 out; Nq = bev_h*bev_w.
 */
BEVF_API int bevf_point_sampling(const float *lidar2img, const float *pc_range, const float *z_norm,
                                 float img_h, float img_w, float *ref_cam, uint8_t *bev_mask, int B,
                                 int ncam, int bev_h, int bev_w, int D, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Dense projection on the tcgen05 tensor cores:
 *     y[M,N] = act( x[M,K] . w[N,K]^T + bias[N] ) (+ residual[M,N])
 * replaces nn.Linear (cuBLAS GEMM + bias) and the ReLU / "+ identity" launches that follow it in
 * TemporalSelfAttention (temporal_self_attention.py:198,206-209,267), MSDeformableAttention3D /
 * SpatialCrossAttention (spatial_cross_attention.py:334,338-341,173) and mmcv's FFN.
 *   x (M,K) bf16, w (N,K) bf16 (nn.Linear layout), bias (N) in bias_dtype (f32 | bf16) or NULL,
 *   residual (M,N) bf16 or NULL,
 *   y (M,N) in y_dtype (bf16 | f32, straight from the fp32 accumulator).  relu != 0 applies
 *   max(.,0) before the residual add.  K % 64 == 0, N % 16 == 0.
 */
BEVF_API int bevf_linear_forward(const void *x, const void *w, const void *bias, int bias_dtype,
                                 const void *residual, void *y, int y_dtype, int64_t M, int N, int K,
                                 int relu, void *stream);

/*
 * Input gradient of a projection: dx (M, K) bf16 = dy (M, N) bf16 . w (N, K) bf16 -- the
 * `grad_input = grad_output.mm(weight)` half of F.linear's backward.  Same weight-stationary tcgen05
 * kernel as bevf_linear_forward, but the weight is read in place as an MN-major operand (no W^T copy).
 * N and K must be multiples of 64.
 */
BEVF_API int bevf_linear_dgrad(const void *dy, const void *w, void *dx, int64_t M, int N, int K,
                               void *stream);

/*
 * dX = addend + dY . W : the same kernel with a (M, K) bf16 addend read in the epilogue (may alias dx).
 * Layers that share an input (the camera features feed every layer's value_proj; the BEV queue feeds every
 * layer's temporal value_proj) chain their input gradients through it instead of materialising one
 * gradient per layer and summing them with separate element-wise kernels.
 */
BEVF_API int bevf_linear_dgrad_acc(const void *dy, const void *w, const void *addend, void *dx, int64_t M,
                                   int N, int K, void *stream);

/*
 * Weight gradient of the projection above:  dw[N,K] += dy[M,N]^T . x[M,K]   (fp32, ACCUMULATED
 * INTO: the caller zero-fills).  dy, x bf16 row-major; split over the M rows across the SMs, partial
 * tiles combined with 16 B fp32 reductions.  replaces the cuBLAS call autograd makes for
 * nn.Linear.weight.grad.  K % 64 == 0, N % 8 == 0.  db (N) f32, optional: the bias gradient
 * db[n] += sum_m dy[m, n], summed from the dY tiles while they sit in shared memory (no second pass
 * over dy; replaces the at::reduce_kernel behind nn.Linear.bias.grad).
 */
BEVF_API int bevf_linear_wgrad(const void *dy, const void *x, float *dw, float *db, int64_t M, int N,
                               int K, void *stream);

/* out[c] += sum over rows of x[r, c]  (fp32, ACCUMULATED INTO).  The bias gradient of the projections:
 * replaces the at::reduce_kernel autograd launches for nn.Linear.bias.grad.  x (rows, C) f32 | bf16. */
/* out = srcs[0] + ... + srcs[n-1] (n <= 8 device tensors of `numel` elements, bf16 or f32, fp32 accumulation):
 * the one-pass sum of the per-layer input gradients of a shared input (replaces autograd's chain of
 * pairwise add kernels).  `srcs` is a HOST array of device pointers. */
BEVF_API int bevf_sum_tensors(const void *const *srcs, int n, void *out, int64_t numel, int dtype, void *stream);

BEVF_API int bevf_colsum(const void *x, float *out, int64_t rows, int C, int dtype, void *stream);

/*
 * The FFN's hidden dropout (mmcv FFN: Linear -> ReLU -> Dropout; the ReLU itself is fused into
 * bevf_linear_forward's epilogue) and the joint backward of dropout(relu(z)).
 *   bevf_dropout_inplace: x *= keep / (1 - p) with Philox4x32-10(seed [+ *seed_base], element / vec) bits,
 *     no mask tensor.  replaces at::native::fused_dropout.
 *   bevf_relu_dropout_backward: out = dy * scale where h != 0, else 0, with h the SAVED forward
 *     activation dropout(relu(z)) (h != 0 <=> z > 0 and kept) and scale = 1 / (1 - p).  replaces
 *     masked_scale + threshold_backward.  dy, h, out: n elements in `dtype`.
 */
BEVF_API int bevf_dropout_inplace(void *x, int64_t n, float p, uint64_t seed, const uint64_t *seed_base,
                                  int dtype, void *stream);
BEVF_API int bevf_relu_dropout_backward(const void *dy, const void *h, void *out, int64_t n, float scale,
                                        int dtype, void *stream);

/*
 * Two-pass form of the weight (+ bias) gradient: every split of the M rows stores its partial 128-row
 * tiles into its own slab of `workspace` with plain stores (no 148-way contended reductions), then a
 * small kernel sums the slabs and writes dw (N, K) and db (N) -- fully OVERWRITTEN, in grad_dtype
 * (f32 | bf16, i.e. the parameter's dtype: no zero-fill and no cast launch around the call).
 * bevf_linear_wgrad_workspace_bytes gives the scratch size for a problem; db may be NULL.
 */
BEVF_API int64_t bevf_linear_wgrad_workspace_bytes(int64_t M, int N, int K);
BEVF_API int bevf_linear_wgrad_out(const void *dy, const void *x, void *dw, void *db, int grad_dtype,
                                   void *workspace, int64_t workspace_bytes, int64_t M, int N, int K,
                                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BEVFORMER_B200_H_ */
