// grad_value of the multi-scale deformable attention sampler ("splat" half of the backward), sm_100a.
//
// Reference arithmetic: projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py
// :130-163 (-> mmcv ms_deform_attn_backward): grad_value[pixel] += w_corner * attn * grad_out[row].
//
// Why a separate kernel.  The one-kernel backward issues a 128 B row reduction per (sample, corner):
// 163 M L2 reduction sectors per SCA launch at base, 27x the compulsory grad_value bytes, and the L2
// reduction rate (~200 G sectors/s) is what bounded it (profiles/r1p_*).  Neighbouring BEV queries hit
// neighbouring pixels, so most of those reductions collide (tools/analysis/sca_window_stats.py: for 64
// consecutive tile-ordered pairs, one head, one (level, point) the 256 corner contributions touch 7 % /
// 12 % / 20 % / 35 % as many distinct pixels at levels 3 / 2 / 1 / 0).  Shared-memory atomics are no way
// out (fp32 ATOMS is a CAS loop, integer ATOMS runs at 2 clk per lane), so the merge happens in
// REGISTERS:
//   * a CTA owns kSplatG consecutive rows of the (tile-ordered) row list, one warp per head -- rows of
//     different heads never meet, so warps need no synchronisation;
//   * lane = channel (32 lanes = the 32 channels of a head).  For one (level, point) "slice" the warp
//     first lays out every sample's top-left cell and its four attention-scaled bilinear weights in a
//     small per-warp staging buffer (one sample per lane), reduces the cells' bounding box with REDUX,
//     then walks the samples one by one: cell index and weights are warp-uniform broadcasts, the lane
//     reads its channel of grad_out from the CTA's staged tile and a switch on the cell index adds the
//     four products into a (kWX+1) x (kWY+1) window of per-lane accumulators -- statically indexed
//     registers.  Boxes larger than the window are swept in passes (ballot-selected members);
//   * each touched window cell is flushed with ONE full-line reduction (32 lanes x 4 B = 128 B);
//   * slices that do not cluster (uniform random locations, more than kSplatMaxPasses passes) take a
//     direct path: lanes = 4 corners x 8 x 16 B, one red.global.add.v4.f32 instruction per sample.
#pragma once

namespace bevf {

constexpr int kSplatG = 64;                       // rows per CTA (2 samples per lane and slice)
constexpr int kWX = 10, kWY = 5;                  // window, in top-left cells
constexpr int kWXP = kWX + 1, kWYP = kWY + 1, kCells = kWXP * kWYP;
constexpr int kChunk = 4;                         // slices laid out per staging round
constexpr int kSplatMaxPasses = 8;
constexpr int kSplatMaxHeads = 8;

struct __align__(16) SplatStage {                 // per warp
    float4 w[kChunk][kSplatG];                    // attention-scaled corner weights {00, 01, 10, 11}
    int cell[kChunk][kSplatG];                    // (y0 + 1) << 16 | (x0 + 1), -1 = sample out of the map
    float4 lw[kSplatG];                           // the members of the current window pass, compacted:
    unsigned lm[kSplatG];                         //   weights, (byte offset of the row's grad_out << 8) | window position
};

__device__ __forceinline__ void red_add_f32(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

template <typename TG> __device__ __forceinline__ float splat_g1(const TG *p);
template <> __device__ __forceinline__ float splat_g1<float>(const float *p) { return *p; }
template <> __device__ __forceinline__ float splat_g1<bf16>(const bf16 *p) {
    return __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short *>(p)) << 16);
}
template <typename TG> __device__ __forceinline__ void splat_g4(const TG *p, float (&g)[4]);
template <> __device__ __forceinline__ void splat_g4<float>(const float *p, float (&g)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
}
template <> __device__ __forceinline__ void splat_g4<bf16>(const bf16 *p, float (&g)[4]) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    g[0] = bf16_lo(t.x); g[1] = bf16_hi(t.x); g[2] = bf16_lo(t.y); g[3] = bf16_hi(t.y);
}

#include "msda_splat_switch.inc"   // BEVF_SPLAT_DISPATCH: one brx.idx jump table (tools/gen_splat_switch.py)
static_assert(kWX == 10 && kWY == 5, "msda_splat_switch.inc is generated for a 10 x 5 window");

// kM: the head count when known at compile time (row pitch C = 32 kM becomes an immediate), 0 = run time.
template <typename TG, int kM>
__global__ void __launch_bounds__(32 * kSplatMaxHeads, 2)
msda_bwd_splat_d32(const float *__restrict__ loc, const float *__restrict__ attn,
                   const TG *__restrict__ grad_out, float *__restrict__ grad_value,
                   const int *__restrict__ row_map, const int *__restrict__ order,
                   const int64_t *__restrict__ level_hw, const int64_t *__restrict__ level_start,
                   int S, int M, int Q, int L, int P, long long pairs, unsigned direct_mask,
                   unsigned level_mask) {
    extern __shared__ __align__(16) unsigned char splat_smem[];
    const int C = kM ? kM * 32 : M * 32;
    TG *gs = reinterpret_cast<TG *>(splat_smem);                                   // [kSplatG][C]
    SplatStage *stages = reinterpret_cast<SplatStage *>(splat_smem + (size_t)kSplatG * C * sizeof(TG));
    __shared__ int row_s[kSplatG], map_s[kSplatG];
    __shared__ LevelTab tab;

    const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
    const long long gb = (long long)blockIdx.x * kSplatG;
    for (int k = threadIdx.x; k < kSplatG; k += blockDim.x) {
        const long long idx = gb + k;
        long long r = -1;
        if (idx < pairs) r = order ? (long long)__ldg(order + idx) : idx;
        const int mp = r < 0 ? -1 : (row_map ? __ldg(row_map + r) : (int)(r / Q));
        row_s[k] = mp < 0 ? -1 : (int)r;                  // row_map -1: unused row of a fixed-capacity list
        map_s[k] = mp;
    }
    load_level_tab(level_hw, level_start, L, C, tab);                              // ends with __syncthreads()
    {   // the CTA's grad_out tile, in its storage type
        constexpr int VEC = 16 / sizeof(TG);
        const int per_row = C / VEC;
        for (int i = threadIdx.x; i < kSplatG * per_row; i += blockDim.x) {
            const int k = i / per_row, c = (i - k * per_row) * VEC;
            const int r = row_s[k];
            uint4 v = make_uint4(0, 0, 0, 0);
            if (r >= 0) v = __ldg(reinterpret_cast<const uint4 *>(grad_out + (long long)r * C + c));
            *reinterpret_cast<uint4 *>(gs + k * C + c) = v;
        }
    }
    __syncthreads();

    SplatStage &st = stages[m];
    const int LP = L * P;
    const TG *gcol = gs + m * 32;                       // + k * C + channel
    const int r0 = row_s[lane], r1 = row_s[lane + 32];
    const int map0 = map_s[lane], map1 = map_s[lane + 32];
    const float2 *loc0 = reinterpret_cast<const float2 *>(loc) + ((long long)max(r0, 0) * M + m) * LP;
    const float2 *loc1 = reinterpret_cast<const float2 *>(loc) + ((long long)max(r1, 0) * M + m) * LP;
    const float *att0 = attn + ((long long)max(r0, 0) * M + m) * LP;
    const float *att1 = attn + ((long long)max(r1, 0) * M + m) * LP;
    const unsigned FULL = 0xffffffffu;
    const char *gbytes = reinterpret_cast<const char *>(gcol + lane);      // + row byte offset = this lane's channel

    for (int l = 0; l < L; ++l) {
        if (!((level_mask >> l) & 1u)) continue;        // hybrid mode: the other levels are scattered by msda_bwd_d32
        const int H = tab.h[l], W = tab.w[l];
        const long long lofs = tab.lofs[l];
        const bool direct_level = (direct_mask >> l) & 1u;
        for (int p0 = 0; p0 < P; p0 += kChunk) {
            const int nsl = min(kChunk, P - p0);
            // ---- lay out the samples of up to kChunk slices: lane k <-> rows k and k + 32
            // lay out the samples of up to kChunk slices; lane k <-> rows k and k + 32, one row at a time (the
            // register window stays live across this code: keep the transient state small).  16 B loads when
            // the chunk is a whole aligned group of four points.
            const int sbase = l * P + p0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int rr = half ? r1 : r0;
                const float2 *lp = (half ? loc1 : loc0) + sbase;
                const float *ap = (half ? att1 : att0) + sbase;
                float2 xy[kChunk];
                float at[kChunk];
                if (rr >= 0) {
                    if (nsl == kChunk && (P & 3) == 0) {
                        const float4 a = __ldg(reinterpret_cast<const float4 *>(lp));
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(lp) + 1);
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(ap));
                        xy[0] = make_float2(a.x, a.y); xy[1] = make_float2(a.z, a.w);
                        xy[2] = make_float2(b.x, b.y); xy[3] = make_float2(b.z, b.w);
                        at[0] = t.x; at[1] = t.y; at[2] = t.z; at[3] = t.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < kChunk; ++j)
                            if (j < nsl) { xy[j] = __ldg(lp + j); at[j] = __ldg(ap + j); }
                    }
                }
#pragma unroll
                for (int j = 0; j < kChunk; ++j) {
                    if (j < nsl) {
                        int cell = -1;
                        float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (rr >= 0) {
                            const Corner c = make_corner(xy[j].x, xy[j].y, H, W);
                            if (c.valid) {
                                cell = ((c.y0 + 1) << 16) | (c.x0 + 1);
                                wv = make_float4(c.w00 * at[j], c.w01 * at[j], c.w10 * at[j], c.w11 * at[j]);
                            }
                        }
                        st.w[j][lane + 32 * half] = wv;
                        st.cell[j][lane + 32 * half] = cell;
                    }
                }
            }
            __syncwarp();
            // ---- one slice at a time
#pragma unroll 1
            for (int j = 0; j < nsl; ++j) {
                float acc[kCells];                       // the register window; every flush leaves it zero
#pragma unroll
                for (int i = 0; i < kCells; ++i) acc[i] = 0.f;
                const int c0 = st.cell[j][lane], c1 = st.cell[j][lane + 32];
                const int x0a = (c0 & 0xffff) - 1, y0a = (c0 >> 16) - 1;       // garbage when c0 < 0 (masked)
                const int x0b = (c1 & 0xffff) - 1, y0b = (c1 >> 16) - 1;
                unsigned rem0 = __ballot_sync(FULL, c0 >= 0), rem1 = __ballot_sync(FULL, c1 >= 0);
                while (rem0 | rem1) {
                    // the samples of ONE value map (a group normally lies inside one camera / frame)
                    const int kf = rem0 ? (__ffs(rem0) - 1) : (31 + __ffs(rem1));
                    const int bmap = map_s[kf];
                    const bool in0 = ((rem0 >> lane) & 1u) && map0 == bmap;
                    const bool in1 = ((rem1 >> lane) & 1u) && map1 == bmap;
                    const unsigned same0 = __ballot_sync(FULL, in0), same1 = __ballot_sync(FULL, in1);
                    rem0 &= ~same0; rem1 &= ~same1;
                    float *gbase = grad_value + (long long)bmap * S * C + lofs + m * 32;
                    const int BIG = 1 << 30;
                    const int minx = __reduce_min_sync(FULL, min(in0 ? x0a : BIG, in1 ? x0b : BIG));
                    const int miny = __reduce_min_sync(FULL, min(in0 ? y0a : BIG, in1 ? y0b : BIG));
                    const int maxx = __reduce_max_sync(FULL, max(in0 ? x0a : -BIG, in1 ? x0b : -BIG));
                    const int maxy = __reduce_max_sync(FULL, max(in0 ? y0a : -BIG, in1 ? y0b : -BIG));
                    const int npx = (maxx - minx) / kWX + 1, npy = (maxy - miny) / kWY + 1;
                    if (direct_level || npx * npy > kSplatMaxPasses) {
                        // ---- direct: lanes = 4 corners x 8 x (4 channels), one v4 reduction per sample
                        const int corner = lane >> 3, chunk = lane & 7;
                        for (int half = 0; half < 2; ++half) {
                            for (unsigned mm = half ? same1 : same0; mm; mm &= mm - 1) {
                                const int k = __ffs(mm) - 1 + 32 * half;
                                const int cv = st.cell[j][k];
                                const float wv = reinterpret_cast<const float *>(&st.w[j][k])[corner];
                                const int x = (cv & 0xffff) - 1 + (corner & 1), y = (cv >> 16) - 1 + (corner >> 1);
                                if (wv != 0.f && (unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
                                    float g[4];
                                    splat_g4<TG>(gcol + k * C + chunk * 4, g);
                                    red_add_v4(gbase + (long long)(y * W + x) * C + chunk * 4,
                                               wv * g[0], wv * g[1], wv * g[2], wv * g[3]);
                                }
                            }
                        }
                        continue;
                    }
                    // ---- merged: sweep the bounding box with the register window
                    for (int oy = miny; oy <= maxy; oy += kWY) {
                        for (int ox = minx; ox <= maxx; ox += kWX) {
                            const bool me0 = in0 && (unsigned)(x0a - ox) < (unsigned)kWX &&
                                             (unsigned)(y0a - oy) < (unsigned)kWY;
                            const bool me1 = in1 && (unsigned)(x0b - ox) < (unsigned)kWX &&
                                             (unsigned)(y0b - oy) < (unsigned)kWY;
                            const unsigned mem0 = __ballot_sync(FULL, me0), mem1 = __ballot_sync(FULL, me1);
                            if (!(mem0 | mem1)) continue;
                            // compact the members into the pass list: position = rank among the members
                            const unsigned lt = (1u << lane) - 1u;
                            const int n0 = __popc(mem0), n = n0 + __popc(mem1);
                            if (me0) {
                                const int pos = __popc(mem0 & lt);
                                st.lw[pos] = st.w[j][lane];
                                st.lm[pos] = ((unsigned)(lane * C * (int)sizeof(TG)) << 8) |
                                             (unsigned)((y0a - oy) * kWX + (x0a - ox));
                            }
                            if (me1) {
                                const int pos = n0 + __popc(mem1 & lt);
                                st.lw[pos] = st.w[j][lane + 32];
                                st.lm[pos] = ((unsigned)((lane + 32) * C * (int)sizeof(TG)) << 8) |
                                             (unsigned)((y0b - oy) * kWX + (x0b - ox));
                            }
                            __syncwarp();
                            // apply: sequential entries, two per trip (independent loads ahead of each dispatch)
                            int e = 0;
                            for (; e + 2 <= n; e += 2) {
                                const float4 w0 = st.lw[e], w1 = st.lw[e + 1];
                                const unsigned m0 = st.lm[e], m1 = st.lm[e + 1];
                                const float g0 = splat_g1<TG>(reinterpret_cast<const TG *>(gbytes + (m0 >> 8)));
                                const float g1 = splat_g1<TG>(reinterpret_cast<const TG *>(gbytes + (m1 >> 8)));
                                { const int c = (int)(m0 & 0xffu); BEVF_SPLAT_DISPATCH(acc, w0, g0, c); }
                                { const int c = (int)(m1 & 0xffu); BEVF_SPLAT_DISPATCH(acc, w1, g1, c); }
                            }
                            if (e < n) {
                                const float4 w0 = st.lw[e];
                                const unsigned m0 = st.lm[e];
                                const float g0 = splat_g1<TG>(reinterpret_cast<const TG *>(gbytes + (m0 >> 8)));
                                const int c = (int)(m0 & 0xffu);
                                BEVF_SPLAT_DISPATCH(acc, w0, g0, c);
                            }
                            __syncwarp();                                   // the list is rewritten by the next pass
                            // ---- flush: one full-line reduction per touched cell, which is zeroed on the way
#pragma unroll
                            for (int cy = 0; cy < kWYP; ++cy) {
                                const int y = oy + cy;
                                const bool yok = (unsigned)y < (unsigned)H;
                                // one address per window row; the cells of the row sit C floats apart
                                float *rowp = gbase + ((long long)y * W + ox) * C + lane;
#pragma unroll
                                for (int cx = 0; cx < kWXP; ++cx) {
                                    const float v = acc[cy * kWXP + cx];
                                    if (__any_sync(FULL, v != 0.f)) {
                                        if (yok && (unsigned)(ox + cx) < (unsigned)W) red_add_f32(rowp + cx * C, v);
                                        acc[cy * kWXP + cx] = 0.f;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

#undef BEVF_SPLAT_DISPATCH

inline size_t splat_smem_bytes(int M, size_t sizeof_tg) {
    return (size_t)kSplatG * M * 32 * sizeof_tg + (size_t)M * sizeof(SplatStage);
}

}  // namespace bevf
