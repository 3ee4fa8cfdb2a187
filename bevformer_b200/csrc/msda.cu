// Multi-scale deformable attention sampler for sm_100a: the irregular multi-camera / multi-level
// bilinear gather + attention-weighted reduce, and its backward.
//
// Replaces the reference's native op (mmcv._ext.ms_deform_attn_{forward,backward}; call sites
// projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:118-124 and
// :150-160).  Arithmetic: SURVEY.md Appendix A.  This is HBM/L2-bound gather work: no tensor cores.
//
// Mapping (head_dim == 32 fast path).  One (pixel, head) row of `value` is 32 contiguous channels:
// 128 B in fp32, 64 B in bf16.  A lane owns 16 B of it (4 fp32 / 8 bf16 channels), so a row is
// covered by 8 (fp32) or 4 (bf16) adjacent lanes and one warp works on 4 / 8 consecutive
// (query, head) pairs at once.  Every corner fetch is therefore one fully-used 16 B vector load per
// lane, a (query, head) pair's sampling locations / weights are read once per lane group, and the
// output row is written with 16 B stores that are contiguous across the warp.
#include "common.cuh"

namespace bevf {

constexpr int kMaxLevels = 16;
constexpr int kThreads = 256;

struct Corner {
    int off00, off01, off10, off11;   // element offsets of the four corners (clamped in range)
    float w00, w01, w10, w11;         // bilinear weights, zero for corners outside the map
    float f00, f01, f10, f11;         // 1 if the corner lies inside the map and the sample counts
    float lx, ly;
    bool valid;
};

// Range test in float BEFORE any int conversion: projected anchors behind a camera reach |x| ~ 1e9.
__device__ __forceinline__ Corner make_corner(float locx, float locy, int H, int W, int pix_stride) {
    Corner c;
    float x = locx * (float)W - 0.5f, y = locy * (float)H - 0.5f;
    c.valid = (x > -1.f) && (y > -1.f) && (x < (float)W) && (y < (float)H);
    if (!c.valid) { x = 0.f; y = 0.f; }
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf, x1 = x0 + 1, y1 = y0 + 1;
    c.lx = x - xf; c.ly = y - yf;
    const float hx = 1.f - c.lx, hy = 1.f - c.ly;
    const bool x0ok = x0 >= 0, x1ok = x1 <= W - 1, y0ok = y0 >= 0, y1ok = y1 <= H - 1;
    c.f00 = (c.valid && x0ok && y0ok) ? 1.f : 0.f;
    c.f01 = (c.valid && x1ok && y0ok) ? 1.f : 0.f;
    c.f10 = (c.valid && x0ok && y1ok) ? 1.f : 0.f;
    c.f11 = (c.valid && x1ok && y1ok) ? 1.f : 0.f;
    c.w00 = c.f00 * hy * hx;
    c.w01 = c.f01 * hy * c.lx;
    c.w10 = c.f10 * c.ly * hx;
    c.w11 = c.f11 * c.ly * c.lx;
    const int x0c = max(x0, 0), x1c = min(x1, W - 1), y0c = max(y0, 0), y1c = min(y1, H - 1);
    c.off00 = (y0c * W + x0c) * pix_stride;
    c.off01 = (y0c * W + x1c) * pix_stride;
    c.off10 = (y1c * W + x0c) * pix_stride;
    c.off11 = (y1c * W + x1c) * pix_stride;
    return c;
}

__device__ __forceinline__ void load_levels(const int64_t *level_hw, const int64_t *level_start,
                                            int L, int *s_h, int *s_w, int *s_start) {
    if ((int)threadIdx.x < L) {
        s_h[threadIdx.x] = (int)level_hw[2 * threadIdx.x];
        s_w[threadIdx.x] = (int)level_hw[2 * threadIdx.x + 1];
        s_start[threadIdx.x] = (int)level_start[threadIdx.x];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward, head_dim == 32
// ------------------------------------------------------------------------------------------------
template <typename T, typename TO>
__global__ void __launch_bounds__(kThreads)
msda_fwd_d32(const T *__restrict__ value, const int64_t *__restrict__ level_hw,
             const int64_t *__restrict__ level_start, const float *__restrict__ loc,
             const float *__restrict__ attn, TO *__restrict__ out, int S, int M, int Q, int L, int P,
             long long rows) {
    constexpr int VEC = Row<T>::kVec, LANES = 32 / VEC, G = 32 / LANES;
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    load_levels(level_hw, level_start, L, s_h, s_w, s_start);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES, grp = lane / LANES;
    const long long row = ((long long)blockIdx.x * (kThreads / 32) + warp) * G + grp;
    if (row >= rows) return;
    const int m = (int)(row % M);
    const int b = (int)(row / ((long long)M * Q));
    const int pix = M * 32;
    const T *vbase = value + ((long long)b * S * M + m) * 32 + sub * VEC;
    const float2 *locp = reinterpret_cast<const float2 *>(loc) + row * L * P;
    const float *attp = attn + row * L * P;

    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

    for (int l = 0; l < L; ++l) {
        const int H = s_h[l], W = s_w[l];
        const T *vl = vbase + (long long)s_start[l] * pix;
#pragma unroll 4
        for (int p = 0; p < P; ++p) {
            const float2 xy = __ldg(locp + l * P + p);
            const float a = __ldg(attp + l * P + p);
            const Corner c = make_corner(xy.x, xy.y, H, W, pix);
            float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
            Row<T>::load(vl + c.off00, v00);
            Row<T>::load(vl + c.off01, v01);
            Row<T>::load(vl + c.off10, v10);
            Row<T>::load(vl + c.off11, v11);
            const float w00 = c.w00 * a, w01 = c.w01 * a, w10 = c.w10 * a, w11 = c.w11 * a;
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                acc[i] += w00 * v00[i] + w01 * v01[i] + w10 * v10[i] + w11 * v11[i];
        }
    }
    store_vec<TO, VEC>(out + row * 32 + sub * VEC, acc);
}

// ------------------------------------------------------------------------------------------------
// backward, head_dim == 32
// ------------------------------------------------------------------------------------------------
template <typename T, typename TG>
__global__ void __launch_bounds__(kThreads)
msda_bwd_d32(const T *__restrict__ value, const int64_t *__restrict__ level_hw,
             const int64_t *__restrict__ level_start, const float *__restrict__ loc,
             const float *__restrict__ attn, const TG *__restrict__ grad_out,
             float *__restrict__ grad_value, float *__restrict__ grad_loc,
             float *__restrict__ grad_attn, int S, int M, int Q, int L, int P, long long rows) {
    constexpr int VEC = Row<T>::kVec, LANES = 32 / VEC, G = 32 / LANES;
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    load_levels(level_hw, level_start, L, s_h, s_w, s_start);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LANES, grp = lane / LANES;
    long long row = ((long long)blockIdx.x * (kThreads / 32) + warp) * G + grp;
    const bool live = row < rows;          // dead groups still take part in the shuffles
    if (!live) row = rows - 1;
    const int m = (int)(row % M);
    const int b = (int)(row / ((long long)M * Q));
    const int pix = M * 32;
    const long long voff = ((long long)b * S * M + m) * 32 + sub * VEC;
    const float2 *locp = reinterpret_cast<const float2 *>(loc) + row * L * P;
    const float *attp = attn + row * L * P;

    float g[VEC];   // grad_out row, in the gradient's own storage type
    load_vec<TG, VEC>(grad_out + row * 32 + sub * VEC, g);

    for (int l = 0; l < L; ++l) {
        const int H = s_h[l], W = s_w[l];
        const long long lbase = voff + (long long)s_start[l] * pix;
        const T *vl = value + lbase;
        float *gvl = grad_value + lbase;
#pragma unroll 2
        for (int p = 0; p < P; ++p) {
            const long long si = row * L * P + l * P + p;
            const float2 xy = __ldg(locp + l * P + p);
            const float a = __ldg(attp + l * P + p);
            const Corner c = make_corner(xy.x, xy.y, H, W, pix);
            float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
            Row<T>::load(vl + c.off00, v00);
            Row<T>::load(vl + c.off01, v01);
            Row<T>::load(vl + c.off10, v10);
            Row<T>::load(vl + c.off11, v11);
            const float hx = 1.f - c.lx, hy = 1.f - c.ly;
            float ga = 0.f, gx = 0.f, gy = 0.f;
            float t00[VEC], t01[VEC], t10[VEC], t11[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float t = g[i] * a;
                t00[i] = c.w00 * t; t01[i] = c.w01 * t; t10[i] = c.w10 * t; t11[i] = c.w11 * t;
                // corners outside the map read as zero, whatever their (clamped) address holds
                const float a00 = c.f00 * v00[i], a01 = c.f01 * v01[i], a10 = c.f10 * v10[i],
                            a11 = c.f11 * v11[i];
                ga += g[i] * (hy * (hx * a00 + c.lx * a01) + c.ly * (hx * a10 + c.lx * a11));
                gx += t * (hy * (a01 - a00) + c.ly * (a11 - a10));
                gy += t * (hx * (a10 - a00) + c.lx * (a11 - a01));
            }
            // scatter: 16 B vector reductions, skipped entirely for zero-weight corners
            if (live) {
#pragma unroll
                for (int i = 0; i < VEC; i += 4) {
                    if (c.w00 != 0.f) red_add_v4(gvl + c.off00 + i, t00[i], t00[i + 1], t00[i + 2], t00[i + 3]);
                    if (c.w01 != 0.f) red_add_v4(gvl + c.off01 + i, t01[i], t01[i + 1], t01[i + 2], t01[i + 3]);
                    if (c.w10 != 0.f) red_add_v4(gvl + c.off10 + i, t10[i], t10[i + 1], t10[i + 2], t10[i + 3]);
                    if (c.w11 != 0.f) red_add_v4(gvl + c.off11 + i, t11[i], t11[i + 1], t11[i + 2], t11[i + 3]);
                }
            }
            // reduce the three scalars over the lanes that share this row
#pragma unroll
            for (int s = LANES / 2; s > 0; s >>= 1) {
                ga += __shfl_xor_sync(0xffffffffu, ga, s);
                gx += __shfl_xor_sync(0xffffffffu, gx, s);
                gy += __shfl_xor_sync(0xffffffffu, gy, s);
            }
            if (live && sub == 0) {
                grad_attn[si] = ga;
                reinterpret_cast<float2 *>(grad_loc)[si] = make_float2((float)W * gx, (float)H * gy);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// any head_dim: one warp per (query, head) row, lanes stride over channels
// ------------------------------------------------------------------------------------------------
template <typename T, typename TO>
__global__ void __launch_bounds__(kThreads)
msda_fwd_generic(const T *__restrict__ value, const int64_t *__restrict__ level_hw,
                 const int64_t *__restrict__ level_start, const float *__restrict__ loc,
                 const float *__restrict__ attn, TO *__restrict__ out, int S, int M, int D, int Q,
                 int L, int P, long long rows) {
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    load_levels(level_hw, level_start, L, s_h, s_w, s_start);
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int m = (int)(row % M);
    const int b = (int)(row / ((long long)M * Q));
    const int pix = M * D;
    for (int c0 = lane; c0 < D; c0 += 32) {
        float acc = 0.f;
        for (int l = 0; l < L; ++l) {
            const int H = s_h[l], W = s_w[l];
            const T *vl = value + ((long long)b * S + s_start[l]) * pix + (long long)m * D + c0;
            for (int p = 0; p < P; ++p) {
                const long long si = row * L * P + l * P + p;
                const Corner c = make_corner(loc[2 * si], loc[2 * si + 1], H, W, pix);
                const float a = attn[si];
                acc += a * (c.w00 * Row<T>::load1(vl + c.off00) + c.w01 * Row<T>::load1(vl + c.off01) +
                            c.w10 * Row<T>::load1(vl + c.off10) + c.w11 * Row<T>::load1(vl + c.off11));
            }
        }
        Row<TO>::store1(out + row * D + c0, acc);
    }
}

template <typename T, typename TG>
__global__ void __launch_bounds__(kThreads)
msda_bwd_generic(const T *__restrict__ value, const int64_t *__restrict__ level_hw,
                 const int64_t *__restrict__ level_start, const float *__restrict__ loc,
                 const float *__restrict__ attn, const TG *__restrict__ grad_out,
                 float *__restrict__ grad_value, float *__restrict__ grad_loc,
                 float *__restrict__ grad_attn, int S, int M, int D, int Q, int L, int P,
                 long long rows) {
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    load_levels(level_hw, level_start, L, s_h, s_w, s_start);
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= rows) return;   // whole warp leaves together
    const int m = (int)(row % M);
    const int b = (int)(row / ((long long)M * Q));
    const int pix = M * D;
    for (int l = 0; l < L; ++l) {
        const int H = s_h[l], W = s_w[l];
        const long long lbase = ((long long)b * S + s_start[l]) * pix + (long long)m * D;
        for (int p = 0; p < P; ++p) {
            const long long si = row * L * P + l * P + p;
            const float lxn = loc[2 * si], lyn = loc[2 * si + 1];
            const Corner c = make_corner(lxn, lyn, H, W, pix);
            const float a = attn[si];
            const float hx = 1.f - c.lx, hy = 1.f - c.ly;
            const float f00 = c.f00, f01 = c.f01, f10 = c.f10, f11 = c.f11;
            float ga = 0.f, gx = 0.f, gy = 0.f;
            for (int c0 = lane; c0 < D; c0 += 32) {
                const float gc = Row<TG>::load1(grad_out + row * D + c0), t = gc * a;
                const float v00 = f00 * Row<T>::load1(value + lbase + c.off00 + c0);
                const float v01 = f01 * Row<T>::load1(value + lbase + c.off01 + c0);
                const float v10 = f10 * Row<T>::load1(value + lbase + c.off10 + c0);
                const float v11 = f11 * Row<T>::load1(value + lbase + c.off11 + c0);
                if (c.w00 != 0.f) atomicAdd(grad_value + lbase + c.off00 + c0, c.w00 * t);
                if (c.w01 != 0.f) atomicAdd(grad_value + lbase + c.off01 + c0, c.w01 * t);
                if (c.w10 != 0.f) atomicAdd(grad_value + lbase + c.off10 + c0, c.w10 * t);
                if (c.w11 != 0.f) atomicAdd(grad_value + lbase + c.off11 + c0, c.w11 * t);
                ga += gc * (hy * (hx * v00 + c.lx * v01) + c.ly * (hx * v10 + c.lx * v11));
                gx += t * (hy * (v01 - v00) + c.ly * (v11 - v10));
                gy += t * (hx * (v10 - v00) + c.lx * (v11 - v01));
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                ga += __shfl_xor_sync(0xffffffffu, ga, s);
                gx += __shfl_xor_sync(0xffffffffu, gx, s);
                gy += __shfl_xor_sync(0xffffffffu, gy, s);
            }
            if (lane == 0) {
                grad_attn[si] = ga;
                grad_loc[2 * si] = (float)W * gx;
                grad_loc[2 * si + 1] = (float)H * gy;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int check_dims(const char *who, int B, int S, int M, int D, int Q, int L, int P) {
    if (B < 0 || S < 0 || M <= 0 || D <= 0 || Q < 0 || L <= 0 || P <= 0)
        return fail("%s: negative or zero dimension", who);
    if (L > kMaxLevels) return fail("%s: at most 16 levels are supported (got %lld)", who, L);
    if ((long long)S * M * D >= (1ll << 31))
        return fail("%s: one batch item of value exceeds 2^31 elements", who);
    return 0;
}

template <typename T, typename TO>
static int launch_fwd(const void *value, const int64_t *hw, const int64_t *ls, const float *loc,
                      const float *attn, void *out, int S, int M, int D, int Q, int L, int P,
                      long long rows, cudaStream_t st) {
    if (D == 32) {
        constexpr int G = 32 / (32 / Row<T>::kVec);
        const long long per_block = (long long)(kThreads / 32) * G;
        const unsigned grid = (unsigned)((rows + per_block - 1) / per_block);
        msda_fwd_d32<T, TO><<<grid, kThreads, 0, st>>>((const T *)value, hw, ls, loc, attn, (TO *)out,
                                                       S, M, Q, L, P, rows);
    } else {
        const unsigned grid = (unsigned)((rows + kThreads / 32 - 1) / (kThreads / 32));
        msda_fwd_generic<T, TO><<<grid, kThreads, 0, st>>>((const T *)value, hw, ls, loc, attn,
                                                           (TO *)out, S, M, D, Q, L, P, rows);
    }
    return check_launch("bevf_msda_forward");
}

template <typename T, typename TG>
static int launch_bwd(const void *value, const int64_t *hw, const int64_t *ls, const float *loc,
                      const float *attn, const void *go, float *gv, float *gl, float *ga, int S,
                      int M, int D, int Q, int L, int P, long long rows, cudaStream_t st) {
    if (D == 32) {
        constexpr int G = 32 / (32 / Row<T>::kVec);
        const long long per_block = (long long)(kThreads / 32) * G;
        const unsigned grid = (unsigned)((rows + per_block - 1) / per_block);
        msda_bwd_d32<T, TG><<<grid, kThreads, 0, st>>>((const T *)value, hw, ls, loc, attn,
                                                       (const TG *)go, gv, gl, ga, S, M, Q, L, P, rows);
    } else {
        const unsigned grid = (unsigned)((rows + kThreads / 32 - 1) / (kThreads / 32));
        msda_bwd_generic<T, TG><<<grid, kThreads, 0, st>>>((const T *)value, hw, ls, loc, attn,
                                                           (const TG *)go, gv, gl, ga, S, M, D, Q, L,
                                                           P, rows);
    }
    return check_launch("bevf_msda_backward");
}

}  // namespace bevf

using namespace bevf;

extern "C" int bevf_msda_forward(const void *value, int value_dtype, const int64_t *level_hw,
                                 const int64_t *level_start, const float *loc, const float *attn,
                                 void *out, int out_dtype, int B, int S, int M, int D, int Q, int L,
                                 int P, void *stream) {
    if (int e = check_dims("bevf_msda_forward", B, S, M, D, Q, L, P)) return e;
    if (!value || !level_hw || !level_start || !loc || !attn || !out)
        return fail("%s: null pointer argument", "bevf_msda_forward");
    if (!aligned16(value) || !aligned16(loc) || !aligned16(attn) || !aligned16(out))
        return fail("%s: device pointers must be 16-byte aligned", "bevf_msda_forward");
    const long long rows = (long long)B * Q * M;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vb = value_dtype == BEVF_DTYPE_BF16, ob = out_dtype == BEVF_DTYPE_BF16;
    if ((value_dtype != BEVF_DTYPE_F32 && !vb) || (out_dtype != BEVF_DTYPE_F32 && !ob))
        return fail("%s: unsupported dtype code", "bevf_msda_forward");
    if (!vb && !ob) return launch_fwd<float, float>(value, level_hw, level_start, loc, attn, out, S, M, D, Q, L, P, rows, st);
    if (vb && ob) return launch_fwd<bf16, bf16>(value, level_hw, level_start, loc, attn, out, S, M, D, Q, L, P, rows, st);
    if (vb && !ob) return launch_fwd<bf16, float>(value, level_hw, level_start, loc, attn, out, S, M, D, Q, L, P, rows, st);
    return fail("%s: fp32 value with bf16 output is not supported", "bevf_msda_forward");
}

extern "C" int bevf_msda_backward(const void *value, int value_dtype, const int64_t *level_hw,
                                  const int64_t *level_start, const float *loc, const float *attn,
                                  const void *grad_out, int grad_out_dtype, float *grad_value,
                                  float *grad_loc, float *grad_attn, int B, int S, int M, int D,
                                  int Q, int L, int P, void *stream) {
    if (int e = check_dims("bevf_msda_backward", B, S, M, D, Q, L, P)) return e;
    if (!value || !level_hw || !level_start || !loc || !attn || !grad_out || !grad_value ||
        !grad_loc || !grad_attn)
        return fail("%s: null pointer argument", "bevf_msda_backward");
    if (!aligned16(value) || !aligned16(loc) || !aligned16(attn) || !aligned16(grad_out) ||
        !aligned16(grad_value) || !aligned16(grad_loc) || !aligned16(grad_attn))
        return fail("%s: device pointers must be 16-byte aligned", "bevf_msda_backward");
    const long long rows = (long long)B * Q * M;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vb = value_dtype == BEVF_DTYPE_BF16, gb = grad_out_dtype == BEVF_DTYPE_BF16;
    if ((value_dtype != BEVF_DTYPE_F32 && !vb) || (grad_out_dtype != BEVF_DTYPE_F32 && !gb))
        return fail("%s: unsupported dtype code", "bevf_msda_backward");
    if (!vb && !gb) return launch_bwd<float, float>(value, level_hw, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, S, M, D, Q, L, P, rows, st);
    if (vb && gb) return launch_bwd<bf16, bf16>(value, level_hw, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, S, M, D, Q, L, P, rows, st);
    if (vb && !gb) return launch_bwd<bf16, float>(value, level_hw, level_start, loc, attn, grad_out, grad_value, grad_loc, grad_attn, S, M, D, Q, L, P, rows, st);
    return fail("%s: fp32 value with bf16 grad_out is not supported", "bevf_msda_backward");
}
