// Device helpers shared by the sampler kernels (msda.cu, msda_splat.cuh, msda_staged.cu): the bilinear
// corner arithmetic of SURVEY.md Appendix A, the 16-byte row slices and their fp32 / bf16 math.
#pragma once

#include "common.cuh"

namespace bevf {

constexpr int kMaxLevels = 16;
constexpr int kThreads = 256;

struct Corner {
    int x0, y0;                       // top-left cell, unclamped: x0 in [-1, W-1], y0 in [-1, H-1]
    int pidx;                         // pixel index y0c*W + x0c of the (clamped) top-left corner
    int dx, dy;                       // 1 if the right / bottom neighbour is a distinct in-map pixel
    float w00, w01, w10, w11;         // bilinear weights (zero for corners outside the map / skipped)
    float lx, ly;
    float f00, f01, f10, f11;         // 1 if the corner lies inside the map and the sample counts
    bool valid;
};

// Range test in float BEFORE any int conversion: projected anchors behind a camera reach |x| ~ 1e9.
__device__ __forceinline__ Corner make_corner(float locx, float locy, int H, int W) {
    Corner c;
    // unfused multiply / add: the sampling cell is floor(x), so x must round exactly like the
    // reference expression loc * W - 0.5 (an FMA would flip floor() for samples on a cell boundary)
    float x = __fadd_rn(__fmul_rn(locx, (float)W), -0.5f), y = __fadd_rn(__fmul_rn(locy, (float)H), -0.5f);
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
    // Clamp into the map.  When x0 == -1 the only in-map column is x1 == 0: corner "00" then aliases
    // pixel 0 with weight 0 and corner "01" must also address pixel 0, hence dx = 0 (same for y).
    c.x0 = x0; c.y0 = y0;
    c.pidx = max(y0, 0) * W + max(x0, 0);
    c.dx = (x0ok && x1ok) ? 1 : 0;
    c.dy = (y0ok && y1ok) ? 1 : 0;
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

// Per-level tables for the fast kernels: element offset of the level's first pixel row (times the
// pixel stride) and element stride between image rows, so the inner loop does no multiplies.
struct LevelTab {
    int h[kMaxLevels], w[kMaxLevels], rs[kMaxLevels];
    long long lofs[kMaxLevels];
};
__device__ __forceinline__ void load_level_tab(const int64_t *level_hw, const int64_t *level_start, int L,
                                               int pix, LevelTab &t) {
    if ((int)threadIdx.x < L) {
        const int l = threadIdx.x;
        t.h[l] = (int)level_hw[2 * l];
        t.w[l] = (int)level_hw[2 * l + 1];
        t.rs[l] = t.w[l] * pix;
        t.lofs[l] = (long long)level_start[l] * pix;
    }
    __syncthreads();
}

// The pyramid as the HOST planned with it (dense tensor-core backward, msda_dense.cu).  Both kernels of that path
// evaluate the same predicate on the device -- host shapes == device spatial_shapes / level_start -- so a stale
// host copy degrades to the plain reduction path instead of producing a wrong gradient.
struct HostLevels {
    int h[kMaxLevels], w[kMaxLevels], start[kMaxLevels];
};
__device__ __forceinline__ bool host_levels_match(const HostLevels &hl, const int64_t *level_hw,
                                                  const int64_t *level_start, int L) {
    bool ok = true;
    for (int l = 0; l < L; ++l)
        ok = ok && (int)level_hw[2 * l] == hl.h[l] && (int)level_hw[2 * l + 1] == hl.w[l] &&
             (int)level_start[l] == hl.start[l];
    return ok;
}

// ---- per-storage-type math ---------------------------------------------------------------------
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    float d;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void split16(uint32_t u, unsigned short &lo, unsigned short &hi) {
    asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(u));
}

template <typename T> struct Vec;          // 16 B of a row as loaded
template <> struct Vec<float> {
    float4 v;
    static constexpr int N = 4;
    __device__ __forceinline__ void load(const float *p) { v = __ldg(reinterpret_cast<const float4 *>(p)); }
    __device__ __forceinline__ void axpy(float w, float (&acc)[4]) const {
        acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
        acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
    }
    __device__ __forceinline__ float dot(const float (&g)[4]) const {
        return fmaf(g[0], v.x, fmaf(g[1], v.y, fmaf(g[2], v.z, g[3] * v.w)));
    }
};
template <> struct Vec<bf16> {
    uint4 v;
    static constexpr int N = 8;
    __device__ __forceinline__ void load(const bf16 *p) { v = __ldg(reinterpret_cast<const uint4 *>(p)); }
    // acc += w * v with w already rounded to bf16 (products exact, fp32 accumulation)
    __device__ __forceinline__ void axpy_h(unsigned short w, float (&acc)[8]) const {
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi;
            split16(u[i], lo, hi);
            acc[2 * i] = fhfma(lo, w, acc[2 * i]);
            acc[2 * i + 1] = fhfma(hi, w, acc[2 * i + 1]);
        }
    }
    // <g, v> with g given as packed bf16 (exact products)
    __device__ __forceinline__ float dot_h(const uint4 &g) const {
        const uint32_t u[4] = {v.x, v.y, v.z, v.w}, q[4] = {g.x, g.y, g.z, g.w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi, glo, ghi;
            split16(u[i], lo, hi);
            split16(q[i], glo, ghi);
            d = fhfma(lo, glo, d);
            d = fhfma(hi, ghi, d);
        }
        return d;
    }
    // <g, v> with g in fp32
    __device__ __forceinline__ float dot(const float (&g)[8]) const {
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            d = fmaf(g[2 * i], bf16_lo(u[i]), d);
            d = fmaf(g[2 * i + 1], bf16_hi(u[i]), d);
        }
        return d;
    }
};

template <int LANES> struct GroupMask;
template <> struct GroupMask<4> { static constexpr unsigned kBits = 0x11111111u; };
template <> struct GroupMask<8> { static constexpr unsigned kBits = 0x01010101u; };
template <> struct GroupMask<16> { static constexpr unsigned kBits = 0x00010001u; };

// level of flat sample index s (= s / P) without an integer division: magic = ceil(2^16 / P),
// exact while s * P < 2^16 (checked on the host: L * P * P < 65536).
__device__ __forceinline__ int level_of(int s, int magic) { return (s * magic) >> 16; }

__device__ __forceinline__ int value_map_of(const int *row_map, long long row, int M, int Q) {
    return row_map ? __ldg(row_map + row / M) : (int)(row / ((long long)M * Q));
}


template <typename T> __device__ __forceinline__ decltype(Vec<T>::v) vec_bits(const uint4 &u);
template <> __device__ __forceinline__ uint4 vec_bits<bf16>(const uint4 &u) { return u; }
template <> __device__ __forceinline__ float4 vec_bits<float>(const uint4 &u) {
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
}

}  // namespace bevf
