// Sampler backward: grad_value of the COARSE pyramid levels as a dense product on the tensor cores.
//
//   grad_value[b, pixel, m, :] = sum over the samples that touch `pixel` of  (attn * bilinear weight) * grad_out[row, m, :]
//
// (the col2im half of mmcv's ms_deform_attn_backward; call site
// projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:150-160, arithmetic
// SURVEY.md Appendix A).  The one-kernel backward (msda.cu) issues one 128 B L2 reduction per (sample, corner):
// 45.6 M of them per SpatialCrossAttention launch at base, and L2's reduction-sector rate bounds the kernel
// (profiles/README.md).  On the coarse levels a pixel receives hundreds of contributions, so for one
// (camera, head) the scatter is a sparse matrix  C[pixel, row]  (<= 4 P non-zeros per column and level) times
// the dense  grad_out[row, 32]  -- and with <= 2048 pixels per bin it is cheap to treat C as DENSE:
//
//   * a CTA owns (value map b, head m, pixel bin, chunk of rows) and keeps  D[bin pixels, 32 channels]  in
//     TMEM (kTiles accumulator tiles of 128 lanes x 32 fp32 columns, kTiles * 32 <= 512 columns);
//   * per step of 16 rows, ONE THREAD per (row, level) writes that row's coefficients into a bf16
//     16 x (bin pixels) slab of shared memory in the UMMA "MN-major, SWIZZLE_128B" layout (thread-level
//     16-bit read-modify-writes: a column belongs to one thread, so no atomics), the 16 grad_out rows go
//     into a K-major tile, and one elected thread issues  tcgen05.mma (M = 128, N = 32, K = 16)  for every
//     accumulator tile the step touched;
//   * slabs are double buffered (the scatter of step i+1 runs under the MMAs of step i, completion through
//     tcgen05.commit -> mbarrier) and are cleaned by re-zeroing exactly the entries that were written;
//   * at the end of the unit the touched tiles are read back (tcgen05.ld), transposed through shared memory
//     and added to grad_value with full-line 16 B vector reductions (rows that stayed zero are skipped).
//
// Per sample this costs a few THREAD-level instructions instead of four warp-wide reductions; the L2
// reduction traffic of these levels drops from 16 sectors per sample to one flush per (unit, touched pixel).
// Coefficients are rounded to bf16 (relative 2^-9, independent per term); accumulation is fp32 in TMEM.
// Levels too large for this treatment (pixels > the launcher's limit) stay on msda_bwd_d32's reduction path
// (its red_skip mask excludes the levels handled here); grad_loc / grad_attn always come from that kernel.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "msda_common.cuh"

namespace bevf {

constexpr int kDnK = 16;                 // sampler rows (reduction index) per step == one MMA K
constexpr int kDnMaxBins = 12;
constexpr int kDnBinLevels = 4;          // levels that may share one bin == scatter thread groups
constexpr int kDnTileBytes = 128 * kDnK * 2;           // one accumulator tile's slab: 128 pixels x 16 rows bf16
constexpr int kDnTransposeBytes = 32 * 144;            // per warp: 32 rows x (32 + 4 pad) floats

struct DenseBins {
    int nbins, L;
    int s0[kDnMaxBins], n[kDnMaxBins];                  // first pixel (flattened over the pyramid), pixel count
    int nlev[kDnMaxBins];
    int lev[kDnMaxBins][kDnBinLevels];                  // levels that intersect the bin
    HostLevels hl;                                      // the host's view of the pyramid (checked on the device)
};

// ---- PTX wrappers (the same instructions gemm.cu uses) -------------------------------------------
namespace dn {
__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DN_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DN_WAIT_DONE;\n"
        "bra DN_WAIT_LOOP;\n"
        "DN_WAIT_DONE:\n"
        "}\n" ::"r"(s32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the tensor core's (async proxy) operand reads
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A operand (the coefficient slab): MN-major, SWIZZLE_128B.  64 consecutive pixels (128 B) per reduction row,
// 8-row groups 1024 B apart (SBO), the next 64-pixel chunk 16 rows x 128 B = 2048 B further (LBO).  Same
// convention as gemm.cu's weight-gradient operands (smem_desc_mn_sw128), with a 16-row chunk.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t chunk_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// B operand (grad_out^T: 32 channel rows x 16 reduction columns): K-major, SWIZZLE_128B (gemm.cu smem_desc_k_sw128)
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16, bf16 x bf16 -> fp32, M = 128, N = 32, A MN-major (bit 15), B K-major
__host__ __device__ constexpr uint32_t idesc_bf16_m128_n32_amn() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (static_cast<uint32_t>(32 >> 3) << 17) |
           (static_cast<uint32_t>(128 >> 4) << 24);
}

// byte offset of coefficient (pixel row `row` of the bin, reduction column k) inside a slab
__device__ __forceinline__ uint32_t slab_off(int row, int k) {
    return ((uint32_t)(row >> 6) << 11) + ((uint32_t)k << 7) + ((((uint32_t)(row >> 3) & 7u) ^ ((uint32_t)k & 7u)) << 4) +
           (((uint32_t)row & 7u) << 1);
}
__device__ __forceinline__ unsigned short lds16(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr) : "memory");
    return h;
}
__device__ __forceinline__ void sts16(uint32_t addr, unsigned short h) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(h) : "memory");
}
__device__ __forceinline__ unsigned short bf16_add(unsigned short old, float x) {
    const float v = __uint_as_float((uint32_t)old << 16) + x;
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
}  // namespace dn

template <int P> struct DnSamples {       // one (row, head, level): P sampling points
    float4 l[P / 2];                      // (x, y) pairs
    float4 a[P / 4];                      // attention weights
};

// Thread organisation: NT scatter teams of 64 threads (team t owns slab t and handles steps t, t + NT, ... of the
// unit) and one issuer warp.  A team thread is (reduction column k = 0..15, level slot 0..3 of the bin); the team's
// second warp also transposes the 16 grad_out rows of the step into the K-major tile.  Hand-over per step:
//   team:   wait empty[t] -> zero what it wrote two uses ago -> scatter -> fence.proxy.async -> arrive full[t]
//   issuer: wait full[t] -> tcgen05.mma for every touched accumulator tile -> tcgen05.commit -> empty[t]
// so the scatter of up to NT steps runs under the MMAs of the previous ones.
template <int kTiles, int P, int NT>
__global__ void __launch_bounds__(64 * NT + 32, 1)
msda_bwd_dense_tc(const __grid_constant__ DenseBins bins, const int64_t *__restrict__ level_hw,
                  const int64_t *__restrict__ level_start, const float *__restrict__ loc,
                  const float *__restrict__ attn, const bf16 *__restrict__ grad_out,
                  float *__restrict__ grad_value, const int *__restrict__ map_range, int NB, int S, int M, int L,
                  int chunk_rows) {
    static_assert(P == 4 || P == 8, "points per level: 4 or 8");
    static_assert(kTiles == 8 || kTiles == 16, "accumulator tiles per bin");
    static_assert(NT >= 2 && NT <= 6, "scatter teams");
    constexpr int kThreadsAll = 64 * NT + 32;
    constexpr int kSlabBytes = kTiles * kDnTileBytes;
    constexpr uint32_t kTmemCols = kTiles * 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *slab = smem;                                   // NT x kSlabBytes
    uint8_t *gtile = smem + NT * kSlabBytes;                // NT x 4096 (32 channel rows x 128 B)
    float *tr = reinterpret_cast<float *>(gtile + NT * 4096);                       // 4 warps x 32 x 36 floats
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(tr) + 4 * kDnTransposeBytes);
    uint64_t *bar_empty = bar_full + NT;
    uint64_t *bar_unit = bar_empty + NT;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_unit + 1);
    uint32_t *s_dirty = tmem_slot + 1;                      // [t]: tiles touched by the step in slab t; [NT]: by the unit
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    __shared__ int s_bad;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int team = tid >> 6, tt = tid & 63;                // team NT = the issuer warp
    const bool is_team = team < NT;

    // ---- set-up: pyramid table (device copy, checked against the host's), barriers, TMEM, clean slabs
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (tid < L) {
        s_h[tid] = (int)level_hw[2 * tid];
        s_w[tid] = (int)level_hw[2 * tid + 1];
        s_start[tid] = (int)level_start[tid];
    }
    if (tid == 64 && !host_levels_match(bins.hl, level_hw, level_start, L)) s_bad = 1;
    if (tid == 32) {
        for (int i = 0; i < NT; ++i) { dn::mbar_init(&bar_full[i], 64); dn::mbar_init(&bar_empty[i], 1); }
        dn::mbar_init(bar_unit, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i <= NT; ++i) s_dirty[i] = 0;
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const int n16 = (NT * kSlabBytes + NT * 4096) / 16;
        for (int i = tid; i < n16; i += kThreadsAll) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    // spatial_shapes on the device differ from the shapes the launcher planned with: msda_bwd_d32 evaluates the
    // same predicate and then keeps every level on its reduction path
    if (s_bad) return;
    if (warp == 0) dn::tmem_alloc(tmem_slot, kTmemCols);
    dn::fence_async_smem();
    dn::tc_fence_before();
    __syncthreads();
    dn::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- unit list: (map, chunk, head, bin), bin fastest
    const int per_chunk = M * bins.nbins;
    int total = 0;
    for (int b = 0; b < NB; ++b) {
        const int n = __ldg(map_range + 2 * b + 1) - __ldg(map_range + 2 * b);
        total += ((n + chunk_rows - 1) / chunk_rows) * per_chunk;
    }

    const int k = tt & 15, slot = tt >> 4;                   // scatter role: reduction column, level slot of the bin
    const bool g_role = is_team && tt >= 32;                 // grad_out role: row k, channel groups 2 * gh, 2 * gh + 1
    const int gh = (tt >> 4) & 1;
    const uint32_t my_slab = dn::s32(slab + (is_team ? team : 0) * kSlabBytes) + ((uint32_t)k << 7);   // + column term
    const uint32_t kx = (uint32_t)k & 7u;
    const uint32_t my_gt = dn::s32(gtile + (is_team ? team : 0) * 4096);
    const uint32_t idesc = dn::idesc_bf16_m128_n32_amn();
    uint32_t nuse = 0;                                       // team: arrivals on full[team] so far == uses of its slab
    uint32_t nfull[NT];                                      // issuer: completed waits per slab
#pragma unroll
    for (int i = 0; i < NT; ++i) nfull[i] = 0;
    uint32_t unit_phase = 0;

    // byte offset of pixel row `row` in this thread's column (without the slab base)
    auto row_off = [&](int row) -> uint32_t {
        return ((uint32_t)(row >> 6) << 11) + ((((uint32_t)(row >> 3) & 7u) ^ kx) << 4) + (((uint32_t)row & 7u) << 1);
    };

    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        // ---- decode
        int b = 0, rem = u, ps = 0, pe = 0;
        for (; b < NB; ++b) {
            ps = __ldg(map_range + 2 * b); pe = __ldg(map_range + 2 * b + 1);
            const int cnt = ((pe - ps + chunk_rows - 1) / chunk_rows) * per_chunk;
            if (rem < cnt) break;
            rem -= cnt;
        }
        const int chunk = rem / per_chunk, rem2 = rem - chunk * per_chunk;
        const int m = rem2 / bins.nbins, bin = rem2 - m * bins.nbins;
        const int r_begin = ps + chunk * chunk_rows, r_end = min(pe, r_begin + chunk_rows);
        const int nsteps = (r_end - r_begin + kDnK - 1) / kDnK;
        const int s0 = bins.s0[bin], nb = bins.n[bin];

        if (is_team) {
            // ================================ scatter teams ================================
            const bool s_role = slot < bins.nlev[bin];
            const int lvl = s_role ? bins.lev[bin][slot] : 0;
            const int H = s_h[lvl], W = s_w[lvl], lbase = s_start[lvl] - s0;
            // rec[p]: (byte offset / 2) | 0x8000 of the four entries written for point p, 0 = none
            uint32_t rec_a[P], rec_b[P];
#pragma unroll
            for (int p = 0; p < P; ++p) rec_a[p] = rec_b[p] = 0;

            auto load_samples = [&](int step, DnSamples<P> &q) {
                const int r = r_begin + step * kDnK + k;
                if (s_role && step < nsteps && r < r_end) {
                    const long long e = (((long long)r * M + m) * L + lvl) * P;
                    const float4 *lp = reinterpret_cast<const float4 *>(loc + 2 * e);
                    const float4 *ap = reinterpret_cast<const float4 *>(attn + e);
#pragma unroll
                    for (int i = 0; i < P / 2; ++i) q.l[i] = __ldg(lp + i);
#pragma unroll
                    for (int i = 0; i < P / 4; ++i) q.a[i] = __ldg(ap + i);
                } else {
#pragma unroll
                    for (int i = 0; i < P / 2; ++i) q.l[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < P / 4; ++i) q.a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            auto load_gout = [&](int step, uint4 &g0, uint4 &g1) {
                const int r = r_begin + step * kDnK + k;
                if (g_role && step < nsteps && r < r_end) {
                    const uint4 *gp = reinterpret_cast<const uint4 *>(grad_out + ((long long)r * M + m) * 32 + gh * 16);
                    g0 = __ldg(gp); g1 = __ldg(gp + 1);
                } else {
                    g0 = g1 = make_uint4(0, 0, 0, 0);
                }
            };
            auto unscatter = [&]() {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint32_t ra = rec_a[p], rb = rec_b[p];
                    if ((ra | rb) != 0) {
                        if (ra & 0x8000u) dn::sts16(my_slab + ((ra & 0x7fffu) << 1), 0);
                        if (ra & 0x80000000u) dn::sts16(my_slab + (((ra >> 16) & 0x7fffu) << 1), 0);
                        if (rb & 0x8000u) dn::sts16(my_slab + ((rb & 0x7fffu) << 1), 0);
                        if (rb & 0x80000000u) dn::sts16(my_slab + (((rb >> 16) & 0x7fffu) << 1), 0);
                        rec_a[p] = rec_b[p] = 0;
                    }
                }
            };
            auto scatter = [&](const DnSamples<P> &q) -> uint32_t {
                uint32_t dirty = 0;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float4 lq = q.l[p >> 1];
                    const float4 aq = q.a[p >> 2];
                    const float x = (p & 1) ? lq.z : lq.x, y = (p & 1) ? lq.w : lq.y;
                    const float a = (p & 3) == 0 ? aq.x : (p & 3) == 1 ? aq.y : (p & 3) == 2 ? aq.z : aq.w;
                    if (a == 0.f) continue;
                    {   // cheap rejection first: a level cut into several bins is scanned once per bin, and most of
                        // its samples fall into another one (same rounding as make_corner: unfused multiply / add)
                        const float yy = __fadd_rn(__fmul_rn(y, (float)H), -0.5f);
                        if (!(yy > -1.f) || !(yy < (float)H)) continue;
                        const int y0 = (int)floorf(yy);
                        if (lbase + (min(y0 + 1, H - 1) + 1) * W <= 0 || lbase + max(y0, 0) * W >= nb) continue;
                    }
                    const Corner c = make_corner(x, y, H, W);
                    if (!c.valid) continue;
                    const int r00 = lbase + c.pidx, r01 = r00 + c.dx, r10 = r00 + c.dy * W, r11 = r10 + c.dx;
                    const float c00 = c.w00 * a, c01 = c.w01 * a, c10 = c.w10 * a, c11 = c.w11 * a;
                    const bool p00 = c00 != 0.f && (unsigned)r00 < (unsigned)nb, p01 = c01 != 0.f && (unsigned)r01 < (unsigned)nb;
                    const bool p10 = c10 != 0.f && (unsigned)r10 < (unsigned)nb, p11 = c11 != 0.f && (unsigned)r11 < (unsigned)nb;
                    if (!(p00 || p01 || p10 || p11)) continue;
                    const uint32_t f00 = row_off(r00), f01 = row_off(r01), f10 = row_off(r10), f11 = row_off(r11);
                    // the four corners are distinct rows (a coinciding pair has one zero weight): loads first
                    unsigned short o00 = 0, o01 = 0, o10 = 0, o11 = 0;
                    if (p00) o00 = dn::lds16(my_slab + f00);
                    if (p01) o01 = dn::lds16(my_slab + f01);
                    if (p10) o10 = dn::lds16(my_slab + f10);
                    if (p11) o11 = dn::lds16(my_slab + f11);
                    uint32_t ra = 0, rb = 0;
                    if (p00) { dn::sts16(my_slab + f00, dn::bf16_add(o00, c00)); dirty |= 1u << (r00 >> 7); ra |= (f00 >> 1) | 0x8000u; }
                    if (p01) { dn::sts16(my_slab + f01, dn::bf16_add(o01, c01)); dirty |= 1u << (r01 >> 7); ra |= ((f01 >> 1) | 0x8000u) << 16; }
                    if (p10) { dn::sts16(my_slab + f10, dn::bf16_add(o10, c10)); dirty |= 1u << (r10 >> 7); rb |= (f10 >> 1) | 0x8000u; }
                    if (p11) { dn::sts16(my_slab + f11, dn::bf16_add(o11, c11)); dirty |= 1u << (r11 >> 7); rb |= ((f11 >> 1) | 0x8000u) << 16; }
                    rec_a[p] = ra; rec_b[p] = rb;
                }
                return dirty;
            };
            // grad_out rows of the step, transposed: channel n = row of the K-major tile, reduction column k
            auto fill_gout = [&](const uint4 &g0, const uint4 &g1) {
                const uint32_t w8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int n = gh * 16 + i;
                    const unsigned short h = (unsigned short)((i & 1) ? (w8[i >> 1] >> 16) : (w8[i >> 1] & 0xffffu));
                    dn::sts16(my_gt + (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + (((uint32_t)k & 7u) << 1), h);
                }
            };

            DnSamples<P> cur;
            uint4 g0, g1;
            load_samples(team, cur);
            load_gout(team, g0, g1);
            for (int step = team; step < nsteps; step += NT) {
                DnSamples<P> nxt;
                uint4 n0, n1;
                load_samples(step + NT, nxt);                    // next step's inputs are in flight during this one
                load_gout(step + NT, n0, n1);
                if (nuse > 0) dn::mbar_wait(&bar_empty[team], (nuse - 1) & 1);   // MMAs that read this slab retired
                if (s_role) {
                    unscatter();
                    const uint32_t dirty = scatter(cur);
                    if (dirty) atomicOr(&s_dirty[team], dirty);
                }
                if (g_role) fill_gout(g0, g1);
                dn::fence_async_smem();
                dn::mbar_arrive(&bar_full[team]);
                nuse++;
                cur = nxt; g0 = n0; g1 = n1;
            }
            // unit end: every MMA retired -> clean the slab
            dn::mbar_wait(bar_unit, unit_phase);
            dn::tc_fence_after();
            if (s_role) unscatter();
        } else {
            // ================================ issuer warp ================================
            if (lane == 0) {
                uint32_t udirty = 0;                          // tiles that hold data of this unit
                int t = 0;
                for (int step = 0; step < nsteps; ++step) {
                    uint32_t par = 0;
#pragma unroll
                    for (int i = 0; i < NT; ++i) if (i == t) { par = nfull[i] & 1u; nfull[i]++; }
                    dn::mbar_wait(&bar_full[t], par);
                    dn::tc_fence_after();
                    const uint32_t mask = atomicExch(&s_dirty[t], 0u);
                    const uint32_t sa = dn::s32(slab + t * kSlabBytes);
                    const uint64_t db = dn::desc_k_sw128(dn::s32(gtile + t * 4096));
#pragma unroll 1
                    for (int i = 0; i < kTiles; ++i)
                        if ((mask >> i) & 1u)
                            dn::umma_bf16(tmem_base + (uint32_t)(i * 32), dn::desc_mn_sw128(sa + (uint32_t)i * kDnTileBytes, 2048),
                                          db, idesc, (udirty >> i) & 1u);
                    udirty |= mask;
                    dn::umma_commit(&bar_empty[t]);
                    if (++t == NT) t = 0;
                }
                s_dirty[NT] = udirty;
                dn::umma_commit(bar_unit);
            }
            __syncwarp();
            dn::mbar_wait(bar_unit, unit_phase);
            dn::tc_fence_after();
        }
        unit_phase ^= 1u;
        __syncthreads();
        // ---- flush: touched tiles -> grad_value (warps 0..3 = TMEM lane quarters)
        const uint32_t touched = s_dirty[NT];
        if (warp < 4) {
            float *trw = tr + warp * (kDnTransposeBytes / 4);
#pragma unroll 1
            for (int t = 0; t < kTiles; ++t) {
                if (!((touched >> t) & 1u)) continue;            // CTA-uniform
                uint32_t r0[16], r1[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 32);
                dn::tmem_ld16(taddr, r0);
                dn::tmem_ld16(taddr + 16, r1);
                dn::tmem_ld_wait();
                uint32_t any = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) any |= (r0[i] | r1[i]) & 0x7fffffffu;
                const int row = t * 128 + warp * 32 + lane;
                const unsigned live = __ballot_sync(0xffffffffu, any != 0 && row < nb);
                if (live == 0) continue;                          // warp-uniform
                float4 *mine = reinterpret_cast<float4 *>(trw + lane * 36);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    mine[i] = make_float4(__uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1]),
                                          __uint_as_float(r0[4 * i + 2]), __uint_as_float(r0[4 * i + 3]));
                    mine[4 + i] = make_float4(__uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1]),
                                              __uint_as_float(r1[4 * i + 2]), __uint_as_float(r1[4 * i + 3]));
                }
                __syncwarp();
                // 8 lanes x 16 B cover one 128 B (pixel, head) row of grad_value; 4 rows per instruction
                const int sub = lane & 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rr = 4 * j + (lane >> 3);
                    if ((live >> rr) & 1u) {
                        const float4 v = *reinterpret_cast<const float4 *>(trw + rr * 36 + sub * 4);
                        const long long pix = (long long)s0 + t * 128 + warp * 32 + rr;
                        float *gp = grad_value + (((long long)b * S + pix) * M + m) * 32 + sub * 4;
                        red_add_v4(gp, v.x, v.y, v.z, v.w);
                    }
                }
                __syncwarp();
            }
        }
        dn::tc_fence_before();
        __syncthreads();                                      // accumulators and slabs are free for the next unit
        dn::tc_fence_after();
    }
    dn::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        dn::tc_fence_after();
        dn::tmem_dealloc(tmem_base, kTmemCols);
    }
}


// ------------------------------------------------------------------------------------------------
// sequential variant: the whole CTA walks the steps in lock step (two slabs, __syncthreads per step)
// ------------------------------------------------------------------------------------------------
constexpr int kDnThreads = 128;          // sequential variant (BEVF_DENSE_TEAMS=1): one CTA-wide pipeline, kept for A/B runs
// rec layout: bits [0,18) = row of corner 00 in the bin + 65536, 18 = dx, 19 = dy, [20,24) = corners written
constexpr int kDnRecBias = 65536;

template <int kTiles, int P>
__global__ void __launch_bounds__(kDnThreads, 1)
msda_bwd_dense_seq(const __grid_constant__ DenseBins bins, const int64_t *__restrict__ level_hw,
                  const int64_t *__restrict__ level_start, const float *__restrict__ loc,
                  const float *__restrict__ attn, const bf16 *__restrict__ grad_out,
                  float *__restrict__ grad_value, const int *__restrict__ map_range, int NB, int S, int M, int L,
                  int chunk_rows) {
    static_assert(P == 4 || P == 8, "points per level: 4 or 8");
    static_assert(kTiles == 8 || kTiles == 16, "accumulator tiles per bin");
    constexpr int kSlabBytes = kTiles * kDnTileBytes;
    constexpr uint32_t kTmemCols = kTiles * 32;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *slab = smem;                                   // 2 x kSlabBytes
    uint8_t *gtile = smem + 2 * kSlabBytes;                 // 2 x 4096 (32 channel rows x 128 B)
    float *tr = reinterpret_cast<float *>(gtile + 2 * 4096);                        // 4 warps x 32 x 36 floats
    uint64_t *bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(tr) + 4 * kDnTransposeBytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 3);
    uint32_t *s_dirty = tmem_slot + 1;                      // [0], [1]: tiles touched by the step in that slab; [2]: by the unit
    __shared__ int s_h[kMaxLevels], s_w[kMaxLevels], s_start[kMaxLevels];
    __shared__ int s_bad;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- set-up: pyramid table (device copy, checked against the host's), barriers, TMEM, clean slabs
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (tid < L) {
        s_h[tid] = (int)level_hw[2 * tid];
        s_w[tid] = (int)level_hw[2 * tid + 1];
        s_start[tid] = (int)level_start[tid];
    }
    if (tid == 64 && !host_levels_match(bins.hl, level_hw, level_start, L)) s_bad = 1;
    if (tid == 32) {
        dn::mbar_init(&bar[0], 1); dn::mbar_init(&bar[1], 1); dn::mbar_init(&bar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_dirty[0] = s_dirty[1] = s_dirty[2] = 0;
    }
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const int n16 = (2 * kSlabBytes + 2 * 4096) / 16;
        for (int i = tid; i < n16; i += kDnThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    // spatial_shapes on the device differ from the shapes the launcher planned with: msda_bwd_d32 evaluates the
    // same predicate and then keeps every level on its reduction path
    if (s_bad) return;
    if (warp == 0) dn::tmem_alloc(tmem_slot, kTmemCols);
    dn::fence_async_smem();
    dn::tc_fence_before();
    __syncthreads();
    dn::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- unit list: (map, chunk, head, bin), bin fastest
    const int per_chunk = M * bins.nbins;
    int total = 0;
    for (int b = 0; b < NB; ++b) {
        const int n = __ldg(map_range + 2 * b + 1) - __ldg(map_range + 2 * b);
        total += ((n + chunk_rows - 1) / chunk_rows) * per_chunk;
    }

    const int k = tid & 15, slot = tid >> 4;                 // scatter role: reduction column, level slot of the bin
    const int gk = tid & 15, gc = (tid >> 4) & 3;            // grad_out role (threads 64..127): row, 8-channel group
    const bool g_role = tid >= 64;
    const uint32_t slab_a[2] = {dn::s32(slab), dn::s32(slab + kSlabBytes)};
    const uint32_t gt_a[2] = {dn::s32(gtile), dn::s32(gtile + 4096)};
    const uint32_t idesc = dn::idesc_bf16_m128_n32_amn();
    uint32_t nuse[2] = {0, 0};                               // commits issued so far on bar[0], bar[1] (CTA lifetime)
    uint32_t unit_phase = 0;

    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        // ---- decode
        int b = 0, rem = u, ps = 0, pe = 0;
        for (; b < NB; ++b) {
            ps = __ldg(map_range + 2 * b); pe = __ldg(map_range + 2 * b + 1);
            const int cnt = ((pe - ps + chunk_rows - 1) / chunk_rows) * per_chunk;
            if (rem < cnt) break;
            rem -= cnt;
        }
        const int chunk = rem / per_chunk, rem2 = rem - chunk * per_chunk;
        const int m = rem2 / bins.nbins, bin = rem2 - m * bins.nbins;
        const int r_begin = ps + chunk * chunk_rows, r_end = min(pe, r_begin + chunk_rows);
        const int nsteps = (r_end - r_begin + kDnK - 1) / kDnK;
        const int s0 = bins.s0[bin], nb = bins.n[bin];
        const bool s_role = slot < bins.nlev[bin];
        const int lvl = s_role ? bins.lev[bin][slot] : 0;
        const int H = s_h[lvl], W = s_w[lvl], lbase = s_start[lvl] - s0;

        uint32_t rec0[P], rec1[P];
#pragma unroll
        for (int p = 0; p < P; ++p) rec0[p] = rec1[p] = 0;
        uint32_t udirty = 0;                                 // issuer only: tiles that hold data of this unit

        auto load_samples = [&](int step, DnSamples<P> &q) {
            const int r = r_begin + step * kDnK + k;
            if (s_role && step < nsteps && r < r_end) {
                const long long e = (((long long)r * M + m) * L + lvl) * P;
                const float4 *lp = reinterpret_cast<const float4 *>(loc + 2 * e);
                const float4 *ap = reinterpret_cast<const float4 *>(attn + e);
#pragma unroll
                for (int i = 0; i < P / 2; ++i) q.l[i] = __ldg(lp + i);
#pragma unroll
                for (int i = 0; i < P / 4; ++i) q.a[i] = __ldg(ap + i);
            } else {
#pragma unroll
                for (int i = 0; i < P / 2; ++i) q.l[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < P / 4; ++i) q.a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto load_gout = [&](int step) -> uint4 {
            const int r = r_begin + step * kDnK + gk;
            if (g_role && step < nsteps && r < r_end)
                return __ldg(reinterpret_cast<const uint4 *>(grad_out + ((long long)r * M + m) * 32 + gc * 8));
            return make_uint4(0, 0, 0, 0);
        };
        // zero the entries a previous step wrote into this slab
        auto unscatter = [&](uint32_t base, uint32_t (&rec)[P]) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const uint32_t rc = rec[p];
                if (rc != 0) {
                    const int r00 = (int)(rc & 0x3ffffu) - kDnRecBias;
                    const int dx = (rc >> 18) & 1, dy = (rc >> 19) & 1;
                    const int r01 = r00 + dx, r10 = r00 + dy * W, r11 = r10 + dx;
                    if (rc & (1u << 20)) dn::sts16(base + dn::slab_off(r00, k), 0);
                    if (rc & (1u << 21)) dn::sts16(base + dn::slab_off(r01, k), 0);
                    if (rc & (1u << 22)) dn::sts16(base + dn::slab_off(r10, k), 0);
                    if (rc & (1u << 23)) dn::sts16(base + dn::slab_off(r11, k), 0);
                    rec[p] = 0;
                }
            }
        };
        auto scatter = [&](uint32_t base, const DnSamples<P> &q, uint32_t (&rec)[P]) -> uint32_t {
            uint32_t dirty = 0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float4 lq = q.l[p >> 1];
                const float4 aq = q.a[p >> 2];
                const float x = (p & 1) ? lq.z : lq.x, y = (p & 1) ? lq.w : lq.y;
                const float a = (p & 3) == 0 ? aq.x : (p & 3) == 1 ? aq.y : (p & 3) == 2 ? aq.z : aq.w;
                if (a == 0.f) continue;
                const Corner c = make_corner(x, y, H, W);
                if (!c.valid) continue;
                const int r00 = lbase + c.pidx, r01 = r00 + c.dx, r10 = r00 + c.dy * W, r11 = r10 + c.dx;
                const float c00 = c.w00 * a, c01 = c.w01 * a, c10 = c.w10 * a, c11 = c.w11 * a;
                const bool p00 = c00 != 0.f && (unsigned)r00 < (unsigned)nb, p01 = c01 != 0.f && (unsigned)r01 < (unsigned)nb;
                const bool p10 = c10 != 0.f && (unsigned)r10 < (unsigned)nb, p11 = c11 != 0.f && (unsigned)r11 < (unsigned)nb;
                if (!(p00 || p01 || p10 || p11)) continue;
                const uint32_t a00 = base + dn::slab_off(r00, k), a01 = base + dn::slab_off(r01, k);
                const uint32_t a10 = base + dn::slab_off(r10, k), a11 = base + dn::slab_off(r11, k);
                // the four corners are distinct rows (a coinciding pair has one zero weight): loads first
                unsigned short o00 = 0, o01 = 0, o10 = 0, o11 = 0;
                if (p00) o00 = dn::lds16(a00);
                if (p01) o01 = dn::lds16(a01);
                if (p10) o10 = dn::lds16(a10);
                if (p11) o11 = dn::lds16(a11);
                if (p00) { dn::sts16(a00, dn::bf16_add(o00, c00)); dirty |= 1u << (r00 >> 7); }
                if (p01) { dn::sts16(a01, dn::bf16_add(o01, c01)); dirty |= 1u << (r01 >> 7); }
                if (p10) { dn::sts16(a10, dn::bf16_add(o10, c10)); dirty |= 1u << (r10 >> 7); }
                if (p11) { dn::sts16(a11, dn::bf16_add(o11, c11)); dirty |= 1u << (r11 >> 7); }
                rec[p] = (uint32_t)(r00 + kDnRecBias) | ((uint32_t)c.dx << 18) | ((uint32_t)c.dy << 19) |
                         ((uint32_t)p00 << 20) | ((uint32_t)p01 << 21) | ((uint32_t)p10 << 22) | ((uint32_t)p11 << 23);
            }
            return dirty;
        };
        // grad_out rows of the step, transposed: channel n = row of the K-major tile, reduction column gk
        auto fill_gout = [&](uint32_t base, const uint4 &v) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = gc * 8 + i;
                const unsigned short h = (unsigned short)((i & 1) ? (w4[i >> 1] >> 16) : (w4[i >> 1] & 0xffffu));
                dn::sts16(base + (uint32_t)n * 128u + ((((uint32_t)gk >> 3) ^ ((uint32_t)n & 7u)) << 4) + (((uint32_t)gk & 7u) << 1), h);
            }
        };
        auto do_step = [&](int step, int buf, DnSamples<P> &cur, uint4 &gcur, uint32_t (&rec)[P]) {
            DnSamples<P> nxt;
            load_samples(step + 1, nxt);                     // next step's inputs are in flight during this one
            const uint4 gnxt = load_gout(step + 1);
            if (nuse[buf] > 0) dn::mbar_wait(&bar[buf], (nuse[buf] - 1) & 1);     // MMAs that read this slab retired
            if (s_role) {
                unscatter(slab_a[buf], rec);
                const uint32_t dirty = scatter(slab_a[buf], cur, rec);
                if (dirty) atomicOr(&s_dirty[buf], dirty);
            }
            if (g_role) fill_gout(gt_a[buf], gcur);
            dn::fence_async_smem();
            __syncthreads();
            if (tid == 96) {                                 // issuer
                dn::tc_fence_after();
                const uint32_t mask = s_dirty[buf];
                s_dirty[buf] = 0;
                const uint64_t db = dn::desc_k_sw128(gt_a[buf]);
#pragma unroll 1
                for (int t = 0; t < kTiles; ++t)
                    if ((mask >> t) & 1u)
                        dn::umma_bf16(tmem_base + (uint32_t)(t * 32), dn::desc_mn_sw128(slab_a[buf] + (uint32_t)t * kDnTileBytes, 2048),
                                      db, idesc, (udirty >> t) & 1u);
                udirty |= mask;
                dn::umma_commit(&bar[buf]);
            }
            nuse[buf]++;
            cur = nxt;
            gcur = gnxt;
        };

        DnSamples<P> cur;
        load_samples(0, cur);
        uint4 gcur = load_gout(0);
        for (int step = 0; step < nsteps; step += 2) {
            do_step(step, 0, cur, gcur, rec0);
            if (step + 1 < nsteps) do_step(step + 1, 1, cur, gcur, rec1);
        }
        // ---- unit end: every MMA retired, slabs cleaned, accumulators flushed
        if (tid == 96) {
            s_dirty[2] = udirty;
            dn::umma_commit(&bar[2]);
        }
        dn::mbar_wait(&bar[2], unit_phase);
        unit_phase ^= 1u;
        dn::tc_fence_after();
        if (s_role) { unscatter(slab_a[0], rec0); unscatter(slab_a[1], rec1); }
        __syncthreads();
        const uint32_t touched = s_dirty[2];
        float *trw = tr + warp * (kDnTransposeBytes / 4);
#pragma unroll 1
        for (int t = 0; t < kTiles; ++t) {
            if (!((touched >> t) & 1u)) continue;            // CTA-uniform
            uint32_t r0[16], r1[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 32);
            dn::tmem_ld16(taddr, r0);
            dn::tmem_ld16(taddr + 16, r1);
            dn::tmem_ld_wait();
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) any |= (r0[i] | r1[i]) & 0x7fffffffu;
            const int row = t * 128 + warp * 32 + lane;
            const unsigned live = __ballot_sync(0xffffffffu, any != 0 && row < nb);
            if (live == 0) continue;                          // warp-uniform
            float4 *mine = reinterpret_cast<float4 *>(trw + lane * 36);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mine[i] = make_float4(__uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1]),
                                      __uint_as_float(r0[4 * i + 2]), __uint_as_float(r0[4 * i + 3]));
                mine[4 + i] = make_float4(__uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1]),
                                          __uint_as_float(r1[4 * i + 2]), __uint_as_float(r1[4 * i + 3]));
            }
            __syncwarp();
            // 8 lanes x 16 B cover one 128 B (pixel, head) row of grad_value; 4 rows per instruction
            const int sub = lane & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rr = 4 * j + (lane >> 3);
                if ((live >> rr) & 1u) {
                    const float4 v = *reinterpret_cast<const float4 *>(trw + rr * 36 + sub * 4);
                    const long long pix = (long long)s0 + t * 128 + warp * 32 + rr;
                    float *gp = grad_value + (((long long)b * S + pix) * M + m) * 32 + sub * 4;
                    red_add_v4(gp, v.x, v.y, v.z, v.w);
                }
            }
            __syncwarp();
        }
        dn::tc_fence_before();
        __syncthreads();                                      // accumulators and slabs are free for the next unit
        dn::tc_fence_after();
    }
    dn::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        dn::tc_fence_after();
        dn::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace bevf

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace bevf {

static long dense_env(const char *name, long dflt) {
    const char *e = getenv(name);
    return e ? strtol(e, nullptr, 0) : dflt;
}

// Bins of at most `cap` consecutive pixels over the levels with at most `max_pix` pixels: a level larger than a bin
// is cut into ceil(n / cap) bins, consecutive small levels share one (the kernel gives each level of a bin its own
// group of scatter threads).  Returns the mask of the covered levels, -1 for a bad shape, -2 if the table overflows.
static int plan_dense_bins(const int32_t *hw_host, int L, long max_pix, int cap, DenseBins &bins, long long *total) {
    memset(&bins, 0, sizeof(bins));
    bins.L = L;
    long long start = 0;
    unsigned mask = 0;
    int open = -1;                                       // bin that may still take the next level
    for (int l = 0; l < L; ++l) {
        const int h = hw_host[2 * l], w = hw_host[2 * l + 1];
        if (h <= 0 || w <= 0 || h >= 32768 || w >= 32768) return -1;
        const long long n = (long long)h * w;
        bins.hl.h[l] = h; bins.hl.w[l] = w; bins.hl.start[l] = (int)start;
        if (n <= max_pix) {
            if (n <= cap && open >= 0 && bins.n[open] + n <= cap && bins.nlev[open] < kDnBinLevels) {
                bins.lev[open][bins.nlev[open]++] = l;   // contiguous with the previous coarse level
                bins.n[open] += (int)n;
            } else {
                const int parts = (int)((n + cap - 1) / cap);
                if (bins.nbins + parts > kDnMaxBins) return -2;
                long long off = 0;
                for (int i = 0; i < parts; ++i) {
                    const int bi = bins.nbins++;
                    bins.s0[bi] = (int)(start + off);
                    bins.n[bi] = (int)((n - off) < cap ? (n - off) : cap);
                    bins.nlev[bi] = 1;
                    bins.lev[bi][0] = l;
                    off += bins.n[bi];
                }
                open = bins.nbins - 1;
            }
            mask |= 1u << l;
        } else {
            open = -1;
        }
        start += n;
    }
    *total = start;
    return (int)mask;
}

// Plans the bins from the HOST copy of the pyramid and launches the dense kernel for every level with at most
// `BEVF_DENSE_MAXPIX` pixels.  *handled = mask of the levels whose grad_value it produced (0: not applicable --
// the caller then leaves every level to the reduction path).
int dense_coarse_backward(const char *who, const int64_t *hw_dev, const int64_t *ls_dev, const int32_t *hw_host,
                          const float *loc, const float *attn, const void *grad_out, float *grad_value,
                          const int32_t *map_range, int NB, int S, int M, int L, int P, cudaStream_t st,
                          unsigned *handled, HostLevels *host_levels) {
    *handled = 0;
    static const long max_pix = dense_env("BEVF_DENSE_MAXPIX", 8192);
    static const long tiles = dense_env("BEVF_DENSE_TILES", 16);
    static const long chunk = dense_env("BEVF_DENSE_CHUNK", 512);
    if ((P != 4 && P != 8) || L > kMaxLevels || NB <= 0 || max_pix <= 0) return 0;
    if (tiles != 8 && tiles != 16) return fail("%s: BEVF_DENSE_TILES must be 8 or 16", who);
    if (chunk < kDnK || chunk % kDnK) return fail("%s: BEVF_DENSE_CHUNK must be a positive multiple of 16", who);
    const int cap = (int)tiles * 128;
    DenseBins bins;
    long long start = 0;
    const int pm = plan_dense_bins(hw_host, L, max_pix, cap, bins, &start);
    if (pm == -1) return fail("%s: bad host level shape", who);
    if (pm == -2) return 0;                              // pyramid too large for the bin table
    const unsigned mask = (unsigned)pm;
    if (start != S) return fail("%s: host level shapes do not add up to S (%lld vs %lld)", who, start, S);
    if (mask == 0) return 0;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    static const long teams_env = dense_env("BEVF_DENSE_TEAMS", 0);
    const int teams = teams_env > 0 ? (int)teams_env : (tiles == 16 ? 3 : 2);
    const size_t smem = 1024 + (size_t)teams * ((size_t)tiles * kDnTileBytes + 4096) + 4 * kDnTransposeBytes + 256;
    auto launch = [&](auto kern, int ctas_per_sm) -> int {
        static bool attr_done = false;                   // per instantiation of this generic lambda
        if (!attr_done) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                cudaGetLastError();
                return fail("%s: cannot reserve shared memory for the dense backward", who);
            }
            attr_done = true;
        }
        kern<<<(unsigned)(sms * ctas_per_sm), 64 * teams + 32, smem, st>>>(bins, hw_dev, ls_dev, loc, attn,
                                                                          (const bf16 *)grad_out, grad_value, map_range,
                                                                          NB, S, M, L, (int)chunk);
        return check_launch(who);
    };
    int e;
    if (teams == 1) {
        const size_t smem1 = 1024 + 2 * (size_t)tiles * kDnTileBytes + 2 * 4096 + 4 * kDnTransposeBytes + 64;
        auto launch1 = [&](auto kern, int ctas_per_sm) -> int {
            static bool attr_done = false;
            if (!attr_done) {
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1) != cudaSuccess) {
                    cudaGetLastError();
                    return fail("%s: cannot reserve shared memory for the dense backward", who);
                }
                attr_done = true;
            }
            kern<<<(unsigned)(sms * ctas_per_sm), kDnThreads, smem1, st>>>(bins, hw_dev, ls_dev, loc, attn, (const bf16 *)grad_out,
                                                                          grad_value, map_range, NB, S, M, L, (int)chunk);
            return check_launch(who);
        };
        if (tiles == 16) e = (P == 8) ? launch1(msda_bwd_dense_seq<16, 8>, 1) : launch1(msda_bwd_dense_seq<16, 4>, 1);
        else e = (P == 8) ? launch1(msda_bwd_dense_seq<8, 8>, 2) : launch1(msda_bwd_dense_seq<8, 4>, 2);
    } else if (tiles == 16 && teams == 3) e = (P == 8) ? launch(msda_bwd_dense_tc<16, 8, 3>, 1) : launch(msda_bwd_dense_tc<16, 4, 3>, 1);
    else if (tiles == 16 && teams == 2) e = (P == 8) ? launch(msda_bwd_dense_tc<16, 8, 2>, 1) : launch(msda_bwd_dense_tc<16, 4, 2>, 1);
    else if (tiles == 8 && teams == 2) e = (P == 8) ? launch(msda_bwd_dense_tc<8, 8, 2>, 2) : launch(msda_bwd_dense_tc<8, 4, 2>, 2);
    else if (tiles == 8 && teams == 5) e = (P == 8) ? launch(msda_bwd_dense_tc<8, 8, 5>, 1) : launch(msda_bwd_dense_tc<8, 4, 5>, 1);
    else return fail("%s: unsupported BEVF_DENSE_TILES / BEVF_DENSE_TEAMS combination (x:1, 16:3, 16:2, 8:2, 8:5)", who);
    if (e) return e;
    *handled = mask;
    *host_levels = bins.hl;
    return 0;
}

}  // namespace bevf

// Host-only: the bin plan the dense backward would use for this pyramid (tests / diagnostics; no device work).
extern "C" int bevf_msda_dense_plan(const int32_t *level_hw_host, int L, int max_pix, int tiles, int32_t *bins_out,
                                    int bins_cap, uint32_t *level_mask) {
    using namespace bevf;
    const char *who = "bevf_msda_dense_plan";
    if (!level_hw_host || !bins_out || !level_mask || L <= 0 || L > kMaxLevels || (tiles != 8 && tiles != 16))
        return -fail("%s: bad argument", who);
    DenseBins bins;
    long long total = 0;
    const int pm = plan_dense_bins(level_hw_host, L, max_pix, tiles * 128, bins, &total);
    if (pm == -1) return -fail("%s: bad host level shape", who);
    if (pm == -2) { *level_mask = 0; return 0; }
    if (bins.nbins > bins_cap) return -fail("%s: output too small", who);
    for (int i = 0; i < bins.nbins; ++i) {
        int32_t *o = bins_out + i * (3 + kDnBinLevels);
        o[0] = bins.s0[i]; o[1] = bins.n[i]; o[2] = bins.nlev[i];
        for (int j = 0; j < kDnBinLevels; ++j) o[3 + j] = j < bins.nlev[i] ? bins.lev[i][j] : -1;
    }
    *level_mask = (unsigned)pm;
    return bins.nbins;
}
