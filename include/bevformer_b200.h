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

#ifdef __cplusplus
}
#endif
#endif /* BEVFORMER_B200_H_ */
