// Dense projections of the encoder layer on the 5th-generation tensor cores (sm_100a only).
//
//   Y[M,N] = act( X[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])        bf16 in, fp32 accumulate
//
// replaces the cuBLAS GEMM + separate bias/ReLU/residual kernels behind every nn.Linear of
// TemporalSelfAttention / MSDeformableAttention3D / SpatialCrossAttention / mmcv FFN
// (temporal_self_attention.py:198,206-209,267; spatial_cross_attention.py:173,334,338-341;
// custom_base_transformer_layer.py:157-158).
//
// Structure (one CTA per SM, persistent over 128 x BN output tiles):
//   warp 0     TMA producer: cp.async.bulk.tensor loads of the A (128 x 64) and B (BN x 64) k-blocks
//              into a ring of SWIZZLE_128B shared-memory stages, completion on mbarriers
//   warp 1     MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//              K=16) x 4 per k-block, accumulating in TMEM; tcgen05.commit releases the smem stage
//              and, at the end of a tile, publishes the accumulator
//   warp 2     TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-7  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (one TMEM lane = one output row per
//              thread), add bias / residual, ReLU, convert, 16 B stores; the double-buffered TMEM
//              lets the MMA of tile i+1 run under the epilogue of tile i.
// These GEMMs have K = 256 or 512 only: they are bound by streaming X / Y through HBM, so the tile
// is sized for full-width rows (BN up to 256) rather than for tensor-pipe peak.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace bevf {

constexpr int kBM = 128;           // rows per tile == TMEM lanes
constexpr int kBK = 64;            // bf16 elements per k-block row == 128 B == one swizzle atom
constexpr int kGemmThreads = 256;
constexpr int kAccStages = 2;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols));
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
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

// Shared-memory matrix descriptor of a K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row
// groups 1024 B apart (SBO), version 1 (Blackwell), layout type 2.  (Bit layout: CuTe
// cute/arch/mma_sm100_desc.hpp, union SmemDescriptor.)
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);            // start address  [0,14)
    d |= static_cast<uint64_t>(1) << 16;                           // LBO (unused)   [16,30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO            [32,46)
    d |= static_cast<uint64_t>(1) << 46;                           // version        [46,48)
    d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B   [61,64)
    return d;
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

struct GemmParams {
    int M, N, K;
    int BN;                // output-tile width (multiple of 16, <= 256, divides N)
    int stages;            // smem ring depth
    int relu;
    const void *bias;      // (N) f32 or bf16, or null
    int bias_bf16;
    const bf16 *residual;  // (M, N) or null
    void *y;               // (M, N) bf16 or fp32
};

template <typename TO>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve-up: [stages x (A 16 KB | B BN*128 B)] [barriers] [tmem ptr]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_bytes = kBM * 128, b_bytes = p.BN * 128, stage_bytes = a_bytes + b_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *acc_full = empty + p.stages;
    uint64_t *acc_empty = acc_full + kAccStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + kBM - 1) / kBM, tiles_n = p.N / p.BN;
    const int num_tiles = tiles_m * tiles_n;
    const int kblocks = p.K / kBK;
    const uint32_t tmem_cols = (kAccStages * p.BN <= 32) ? 32 : (kAccStages * p.BN <= 64) ? 64
                             : (kAccStages * p.BN <= 128) ? 128 : (kAccStages * p.BN <= 256) ? 256 : 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int tm = t / tiles_n, tn = t % tiles_n;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t *sa = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    tma_load_2d(sa, &map_a, &full[s], kb * kBK, tm * kBM);
                    tma_load_2d(sa + a_bytes, &map_b, &full[s], kb * kBK, tn * p.BN);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(kBM, p.BN);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t da = smem_desc_k_sw128(a_addr), db = smem_desc_k_sw128(a_addr + a_bytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16 B units
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty[s]);                       // frees the smem stage when the MMAs retire
                    if (kb == kblocks - 1) umma_commit(&acc_full[as]);   // accumulator complete
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                if (++as == kAccStages) { as = 0; aph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                                    // TMEM lane quarter of this warp
        int as = 0; uint32_t aph = 0;
        TO *y = reinterpret_cast<TO *>(p.y);
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int tm = t / tiles_n, tn = t % tiles_n;
            const int row = tm * kBM + q * 32 + lane;
            const bool row_ok = row < p.M;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                tmem_ld_wait();
                const int col = tn * p.BN + c0;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                if (p.bias) {
                    if (p.bias_bf16) {
                        const uint4 *bp = reinterpret_cast<const uint4 *>(reinterpret_cast<const bf16 *>(p.bias) + col);
                        const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
                        const uint32_t u[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo(u[i]); v[2 * i + 1] += bf16_hi(u[i]); }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.bias) + col + i));
                            v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (row_ok) {
                    if (p.residual) {
                        const uint4 *rp = reinterpret_cast<const uint4 *>(p.residual + (size_t)row * p.N + col);
                        const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                        const uint32_t u[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo(u[i]); v[2 * i + 1] += bf16_hi(u[i]); }
                    }
                    TO *dst = y + (size_t)row * p.N + col;
                    if constexpr (sizeof(TO) == 2) {
                        uint4 o0, o1;
                        o0.x = pack_bf16x2(v[0], v[1]); o0.y = pack_bf16x2(v[2], v[3]);
                        o0.z = pack_bf16x2(v[4], v[5]); o0.w = pack_bf16x2(v[6], v[7]);
                        o1.x = pack_bf16x2(v[8], v[9]); o1.y = pack_bf16x2(v[10], v[11]);
                        o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
                        reinterpret_cast<uint4 *>(dst)[0] = o0;
                        reinterpret_cast<uint4 *>(dst)[1] = o1;
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            reinterpret_cast<float4 *>(dst)[i / 4] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (++as == kAccStages) { as = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}



// ================================================================================================
// v2 of the NT projection: weight-stationary.
//
// ncu on v1 (profiles/README.md): tensor pipe 15 % active, DRAM 13 %, 8 resident warps idle -- every
// 128 x BN output tile re-fetched its BN x K weight tile (128 KB) from L2, and one SM's share of L2
// bandwidth is only ~42 B/clk.  Here each CTA loads the weight tile of the current column block ONCE
// into shared memory (<= 128 KB, SWIZZLE_128B k-blocks) and streams only the activation rows past it:
//   warp 0     TMA producer: B once per column block, then A k-blocks (128 x 64) through a ring
//   warp 1     MMA issuer: tcgen05.mma with the A descriptor of the ring stage and the B descriptor of
//              the resident k-block; accumulators double-buffered in TMEM
//   warps 4-7  epilogue: tcgen05.ld 64 (bf16 out) / 32 (fp32 out) columns at a time, bias from shared
//              memory, ReLU, convert, write a 32-row x 128 B slab into a SWIZZLE_128B staging buffer
//              and hand it to the TMA store engine (cp.async.bulk.tensor ... global <- shared): the
//              threads issue no global stores, rows beyond M are clipped by the tensor map.
// ================================================================================================
struct GemmWsParams {
    int M, N, K;
    int BN;                // column block (multiple of 16, <= 256, BN * K * 2 <= 128 KB, divides N)
    int stages;            // A ring depth
    int relu;
    const void *bias;      // (N) f32 or bf16, or null
    int bias_bf16;
    int out_f32;
    int b_mn;              // 0: B = weight (N, K) row-major, K-major tiles (Y = X W^T)
                           // 1: B = (K, N) row-major, MN-major tiles (Y = X B; dX = dY W without W^T)
    const bf16 *addend;    // optional (M, N) bf16 added in the epilogue (bf16 output only): the running sum
                           // of the input gradients of layers that share an input (dX_total = dX_prev + dY W)
};

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t chunk_bytes);

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool kAddend>                    // compile-time: the plain projections pay nothing for the addend path
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_ws_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_y, const GemmWsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kblocks = p.K / kBK;
    const int b_kb_bytes = p.BN * 128;                    // one resident k-block of the weight tile
    const int b_bytes = kblocks * b_kb_bytes;             // <= 128 KB, multiple of 2 KB
    const int a_bytes = kBM * 128;
    uint8_t *smem_b = smem;
    uint8_t *smem_a = smem_b + ((b_bytes + 1023) & ~1023);
    uint8_t *smem_c = smem_a + (size_t)p.stages * a_bytes;           // 4 warps x 2 buffers x 4 KB
    float *s_bias = reinterpret_cast<float *>(smem_c + 4 * 2 * 4096);  // BN floats (<= 1 KB)
    uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_bias) + 1024);
    uint64_t *empty = full + p.stages;
    uint64_t *acc_full = empty + p.stages;
    uint64_t *acc_empty = acc_full + kAccStages;
    uint64_t *b_full = acc_empty + kAccStages;
    uint64_t *b_empty = b_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = (p.M + kBM - 1) / kBM, tiles_n = p.N / p.BN;
    const uint32_t tmem_cols = (kAccStages * p.BN <= 32) ? 32 : (kAccStages * p.BN <= 64) ? 64
                             : (kAccStages * p.BN <= 128) ? 128 : (kAccStages * p.BN <= 256) ? 256 : 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            // One flat list of (column block, row tile) units, column-block-major, dealt round-robin to the
            // CTAs: ceil(tiles_m * tiles_n / grid) rounds instead of tiles_n * ceil(tiles_m / grid) (313 row
            // tiles on 148 SMs: 5 rounds instead of 6 for two column blocks, 7 instead of 9 for three).  A
            // CTA's units have non-decreasing column block, so it still loads each weight block at most once.
            int cur_tn = -1, wloads = 0;
            for (int u = blockIdx.x; u < tiles_m * tiles_n; u += gridDim.x) {
                const int tn = u / tiles_m, tm = u - tn * tiles_m;
                if (tn != cur_tn) {
                    mbar_wait(b_empty, (uint32_t)((wloads & 1) ^ 1));   // previous column block fully consumed
                    mbar_expect_tx(b_full, (uint32_t)b_bytes);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        if (!p.b_mn) {
                            tma_load_2d(smem_b + (size_t)kb * b_kb_bytes, &map_b, b_full, kb * kBK, tn * p.BN);
                        } else {                   // 64 reduction rows x BN columns as BN/64 chunks of 8 KB
                            for (int c = 0; c < p.BN / 64; ++c)
                                tma_load_2d(smem_b + (size_t)kb * b_kb_bytes + c * 8192, &map_b, b_full,
                                            tn * p.BN + c * 64, kb * kBK);
                        }
                    }
                    cur_tn = tn; ++wloads;
                }
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)a_bytes);
                    tma_load_2d(smem_a + (size_t)s * a_bytes, &map_a, &full[s], kb * kBK, tm * kBM);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(kBM, p.BN) | (p.b_mn ? (1u << 16) : 0u);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            int cur_tn = -1, wloads = 0;
            const uint32_t b_addr = smem_u32(smem_b);
            for (int u = blockIdx.x; u < tiles_m * tiles_n; u += gridDim.x) {
                const int tn = u / tiles_m;
                if (tn != cur_tn) {
                    if (cur_tn >= 0) umma_commit(b_empty);       // all MMAs reading the previous weight block have retired
                    mbar_wait(b_full, (uint32_t)(wloads & 1));
                    tc_fence_after();
                    cur_tn = tn; ++wloads;
                }
                mbar_wait(&acc_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t da = smem_desc_k_sw128(smem_u32(smem_a + (size_t)s * a_bytes));
                    // K-major B: 16 k-columns further = 32 B; MN-major B: 16 k-rows = 2048 B
                    const uint64_t db = p.b_mn ? smem_desc_mn_sw128(b_addr + (uint32_t)(kb * b_kb_bytes), 8192)
                                               : smem_desc_k_sw128(b_addr + (uint32_t)(kb * b_kb_bytes));
                    const uint64_t db_step = p.b_mn ? 128 : 2;
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)
                        umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + db_step * (uint64_t)k, idesc,
                                  (uint32_t)((kb | k) != 0));
                    umma_commit(&empty[s]);
                    if (kb == kblocks - 1) umma_commit(&acc_full[as]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                if (++as == kAccStages) { as = 0; aph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int et = threadIdx.x - 128;                             // 0..127 among the epilogue threads
        uint8_t *my_c = smem_c + (size_t)q * 2 * 4096;
        const int cols_per_chunk = p.out_f32 ? 32 : 64;               // 128 B of output per row
        int as = 0; uint32_t aph = 0;
        int cbuf = 0;
        int cur_tn = -1;
        for (int u = blockIdx.x; u < tiles_m * tiles_n; u += gridDim.x) {
            const int tn = u / tiles_m, tm = u - tn * tiles_m;
            if (tn != cur_tn) {
                // bias of this column block -> shared memory (all 128 epilogue threads)
                named_bar_sync(1, 128);                                    // previous block's readers are done
                for (int i = et; i < p.BN; i += 128) {
                    float bv = 0.f;
                    if (p.bias) {
                        const int col = tn * p.BN + i;
                        bv = p.bias_bf16 ? __bfloat162float(reinterpret_cast<const bf16 *>(p.bias)[col])
                                         : reinterpret_cast<const float *>(p.bias)[col];
                    }
                    s_bias[i] = bv;
                }
                named_bar_sync(1, 128);
                cur_tn = tn;
            }
            {
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
                const int row0 = tm * kBM + q * 32;
                for (int c0 = 0; c0 < p.BN; c0 += cols_per_chunk) {
                    uint8_t *buf = my_c + (size_t)cbuf * 4096;
                    // the TMA store that last read this buffer (two chunks ago) must have drained it
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    if (!p.out_f32) {
                        uint32_t r[4][16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) tmem_ld16(taddr + (uint32_t)(c0 + 16 * i), r[i]);
                        tmem_ld_wait();
                        // addend: this thread's output row, 128 contiguous bytes fetched up front
                        uint4 av[8];
                        if constexpr (kAddend) {
                            const int grow = row0 + lane;
#pragma unroll
                            for (int i = 0; i < 8; ++i) av[i] = make_uint4(0, 0, 0, 0);
                            if (grow < p.M) {
                                const uint4 *arow = reinterpret_cast<const uint4 *>(
                                    p.addend + (size_t)grow * p.N + tn * p.BN + c0);
#pragma unroll
                                for (int i = 0; i < 8; ++i) av[i] = __ldg(arow + i);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {                 // eight 16 B chunks = 8 columns each
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int c = 8 * i + j;
                                float t = __uint_as_float(r[c >> 4][c & 15]) + s_bias[c0 + c];
                                if constexpr (kAddend) {
                                    const uint32_t aw[4] = {av[i].x, av[i].y, av[i].z, av[i].w};
                                    t += (j & 1) ? bf16_hi(aw[j >> 1]) : bf16_lo(aw[j >> 1]);
                                }
                                v[j] = p.relu ? fmaxf(t, 0.f) : t;
                            }
                            uint4 o;
                            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                            *reinterpret_cast<uint4 *>(buf + lane * 128 + ((i ^ (lane & 7)) << 4)) = o;
                        }
                    } else {
                        uint32_t r[2][16];
#pragma unroll
                        for (int i = 0; i < 2; ++i) tmem_ld16(taddr + (uint32_t)(c0 + 16 * i), r[i]);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {                 // eight 16 B chunks = 4 columns each
                            float v[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = 4 * i + j;
                                float t = __uint_as_float(r[c >> 4][c & 15]) + s_bias[c0 + c];
                                v[j] = p.relu ? fmaxf(t, 0.f) : t;
                            }
                            *reinterpret_cast<float4 *>(buf + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                                make_float4(v[0], v[1], v[2], v[3]);
                        }
                    }
                    fence_async_smem();                               // generic-proxy writes -> async proxy
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_y, buf, tn * p.BN + c0, row0);
                        tma_store_commit();
                    }
                    cbuf ^= 1;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[as]);
                if (++as == kAccStages) { as = 0; aph ^= 1; }
            }
        }
        if (lane == 0) tma_store_wait_read<0>();                     // staging memory must outlive the reads
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// ================================================================================================
// Weight gradient:  dW[N, K] += dY[M, N]^T . X[M, K]      (reduction over the M rows)
//
// Both operands are read in their natural row-major layout, which makes them MN-major for this
// product (the non-reduced index is the contiguous one): TMA boxes of 64 columns x 64 rows land in
// shared memory exactly as the UMMA "MN-major, SWIZZLE_128B" canonical layout wants them
// (64 contiguous MN elements = 128 B per K row, 8-row groups 1024 B apart = SBO, successive
// 64-column chunks one box (8 KB) apart = LBO).  The reduction is split over CTAs (split-M); every
// CTA adds its 128 x BN partial tile into dW with 16 B fp32 reductions.
// ================================================================================================
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t addr, uint32_t chunk_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;            // LBO: next 64-element MN chunk
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO: next group of 8 K rows
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

struct WgradParams {
    int M, N, K;           // dY (M, N), X (M, K), dW (N, K)
    int BN;                // tile width along K (multiple of 64, <= 256, divides K)
    int stages;
    int rows_per_split;    // multiple of 64
    int splits;
    float *dw;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad_bf16(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kChunk = 64 * 128;                               // one 64 x 64 bf16 box
    const int a_bytes = 2 * kChunk, b_bytes = (p.BN / 64) * kChunk, stage_bytes = a_bytes + b_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *acc_full = empty + p.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (p.N + kBM - 1) / kBM, tiles_k = p.K / p.BN;
    const int units = tiles_n * tiles_k * p.splits;
    const uint32_t tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int s = 0; uint32_t ph = 0;          // smem ring state (producer and MMA walk it identically)
    uint32_t aph = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = u % p.splits, tile = u / p.splits;
        const int tn = tile / tiles_k, tk = tile % tiles_k;
        const int m_begin = split * p.rows_per_split;
        const int m_end = min(p.M, m_begin + p.rows_per_split);
        const int kblocks = (m_end - m_begin + 63) / 64;
        if (warp == 0 && lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t *sa = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                const int m0 = m_begin + kb * 64;
                tma_load_2d(sa, &map_dy, &full[s], tn * kBM, m0);
                tma_load_2d(sa + kChunk, &map_dy, &full[s], tn * kBM + 64, m0);
                for (int c = 0; c < p.BN / 64; ++c)
                    tma_load_2d(sa + a_bytes + c * kChunk, &map_x, &full[s], tk * p.BN + c * 64, m0);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1 && lane == 0) {
            // MN-major A and B: instruction-descriptor bits 15 and 16
            const uint32_t idesc = instr_desc_bf16(kBM, p.BN) | (1u << 15) | (1u << 16);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t da = smem_desc_mn_sw128(a_addr, kChunk), db = smem_desc_mn_sw128(a_addr + a_bytes, kChunk);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // 16 reduction rows further = 2048 B = 128 in 16 B units
                    umma_bf16(tmem_base, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), idesc,
                              (uint32_t)((kb | k) != 0));
                }
                umma_commit(&empty[s]);
                if (kb == kblocks - 1) umma_commit(acc_full);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        // every warp waits for the accumulator of this unit; warps 4-7 drain it
        if (kblocks > 0) {
            if (warp >= 4) {
                mbar_wait(acc_full, aph);
                tc_fence_after();
                const int q = warp & 3;
                const int row = tn * kBM + q * 32 + lane;          // output row = N index
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
                for (int c0 = 0; c0 < p.BN; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + (uint32_t)c0, r);
                    tmem_ld_wait();
                    if (row < p.N) {
                        float *dst = p.dw + (size_t)row * p.K + tk * p.BN + c0;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                       __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                    }
                }
                tc_fence_before();
            }
            aph ^= 1;
        }
        __syncthreads();   // the accumulator is reused by the next unit: drain before the next MMA
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}


// ------------------------------------------------------------------------------------------------
// v2 of the weight gradient: a unit now covers up to TWO 128-row blocks of dW (two TMEM accumulators
// of BN columns) against one shared X tile, so X streams through L2 once instead of once per block
// (ncu on v1: 123 MB through L2 for 82 MB of operands, tensor pipe 14 %), and the three roles run
// free behind mbarriers instead of meeting at a __syncthreads per unit.
// ------------------------------------------------------------------------------------------------
constexpr int kWgradEpiBytes = 4 * 32 * 33 * 4;            // four epilogue warps x (32 x 33) fp32 transpose tile
constexpr int kWgradPipeBytes = 208 * 1024;                 // + 1 KB alignment + barriers + tiles <= 227 KB
struct Wgrad2Params {
    int M, N, K;
    int BN;                // tile width along K (multiple of 64, <= 256, divides K)
    int NB;                // 128-row blocks of dW per unit (1 or 2)
    int stages;
    int rows_per_split;    // multiple of 64
    int splits;
    float *dw;
    float *db;             // (N) f32 bias gradient = column sums of dY, accumulated into; may be null
    float *ws;             // two-pass mode: (splits, n_pad, K) partial tiles, plain stores; else null
    float *ws_b;           // two-pass mode: (splits, n_pad) partial column sums
    int n_pad;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad2_bf16(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                 const Wgrad2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kChunk = 64 * 128;
    const int a_bytes = p.NB * 2 * kChunk, b_bytes = (p.BN / 64) * kChunk, stage_bytes = a_bytes + b_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *acc_full = empty + p.stages;
    uint64_t *acc_empty = acc_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);
    // per-epilogue-warp 32 x 33 fp32 transpose tile (see the epilogue)
    float *epi_tiles = reinterpret_cast<float *>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blocks_n = (p.N + kBM - 1) / kBM;
    const int groups_n = (blocks_n + p.NB - 1) / p.NB, tiles_k = p.K / p.BN;
    const int units = groups_n * tiles_k * p.splits;
    const uint32_t need = (uint32_t)(p.NB * p.BN);
    const uint32_t tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        // a stage is released by the MMA commit and, when the bias gradient is wanted, by warp 3 too
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], p.db ? 2 : 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit -> (group of dW row blocks, K tile, split of the reduction rows)
    auto decode = [&](int u, int &gn, int &tk, int &m_begin, int &kblocks) {
        const int split = u % p.splits, tile = u / p.splits;
        gn = tile / tiles_k; tk = tile % tiles_k;
        m_begin = split * p.rows_per_split;
        const int m_end = min(p.M, m_begin + p.rows_per_split);
        kblocks = (m_end - m_begin + 63) / 64;
    };

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int gn, tk, m_begin, kblocks;
                decode(u, gn, tk, m_begin, kblocks);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t *sa = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    const int m0 = m_begin + kb * 64;
                    for (int c = 0; c < p.NB * 2; ++c)
                        tma_load_2d(sa + c * kChunk, &map_dy, &full[s], gn * p.NB * kBM + c * 64, m0);
                    for (int c = 0; c < p.BN / 64; ++c)
                        tma_load_2d(sa + a_bytes + c * kChunk, &map_x, &full[s], tk * p.BN + c * 64, m0);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = instr_desc_bf16(kBM, p.BN) | (1u << 15) | (1u << 16);
            int s = 0; uint32_t ph = 0;
            uint32_t eph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int gn, tk, m_begin, kblocks;
                decode(u, gn, tk, m_begin, kblocks);
                if (kblocks == 0) continue;
                mbar_wait(acc_empty, eph ^ 1);               // epilogue has drained the accumulators
                eph ^= 1;
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t db = smem_desc_mn_sw128(a_addr + a_bytes, kChunk);
                    for (int nb = 0; nb < p.NB; ++nb) {
                        const uint64_t da = smem_desc_mn_sw128(a_addr + nb * 2 * kChunk, kChunk);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem_base + (uint32_t)(nb * p.BN), da + (uint64_t)(128 * k),
                                      db + (uint64_t)(128 * k), idesc, (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&empty[s]);
                    if (kb == kblocks - 1) umma_commit(acc_full);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== bias gradient: column sums of the dY tiles, read from shared memory ====
        // lane = (16 B unit j = lane % 8 of a 64-column chunk, row group lane / 8); the tile is stored
        // MN-major with the 128 B swizzle: unit j of K row r sits at position j ^ (r & 7).
        if (p.db) {
            int s = 0; uint32_t ph = 0;
            const int j = lane & 7, rg = lane >> 3;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int gn, tk, m_begin, kblocks;
                decode(u, gn, tk, m_begin, kblocks);
                float acc[4][8];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    if (tk == 0) {                                    // count every dY element once
                        const uint8_t *sa = smem + (size_t)s * stage_bytes;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < p.NB * 2) {
#pragma unroll 4
                                for (int r = rg; r < 64; r += 4) {
                                    const uint4 v = *reinterpret_cast<const uint4 *>(sa + c * kChunk + r * 128 + ((j ^ (r & 7)) << 4));
                                    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) { acc[c][2 * e] += bf16_lo(w4[e]); acc[c][2 * e + 1] += bf16_hi(w4[e]); }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                if (tk == 0 && kblocks > 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float t = acc[c][e];
                            t += __shfl_xor_sync(0xffffffffu, t, 8);
                            t += __shfl_xor_sync(0xffffffffu, t, 16);
                            acc[c][e] = t;
                        }
                        const int col = gn * p.NB * kBM + c * 64 + j * 8;
                        if (rg == 0 && c < p.NB * 2) {
                            if (p.ws_b) {
                                float *dst = p.ws_b + (size_t)(u % p.splits) * p.n_pad + col;
                                *reinterpret_cast<float4 *>(dst) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                                *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
                            } else if (col < p.N) {
                                red_add_v4(p.db + col, acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
                                red_add_v4(p.db + col + 4, acc[c][4], acc[c][5], acc[c][6], acc[c][7]);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t aph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int gn, tk, m_begin, kblocks;
            decode(u, gn, tk, m_begin, kblocks);
            if (kblocks == 0) continue;
            mbar_wait(acc_full, aph);
            aph ^= 1;
            tc_fence_after();
            for (int nb = 0; nb < p.NB; ++nb) {
                const int row = (gn * p.NB + nb) * kBM + q * 32 + lane;       // dW row = N index
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(nb * p.BN);
                for (int c0 = 0; c0 < p.BN; c0 += 32) {
                    uint32_t r[2][16];
                    tmem_ld16(taddr + (uint32_t)c0, r[0]);
                    tmem_ld16(taddr + (uint32_t)(c0 + 16), r[1]);
                    tmem_ld_wait();
                    if (p.ws) {
                        // two-pass: this split's partial tile goes to its own slab with plain stores
                        const int split = u % p.splits;
                        float *dst = p.ws + ((size_t)split * p.n_pad + row) * p.K + tk * p.BN + c0;
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4 *>(dst + 16 * h + i) =
                                    make_float4(__uint_as_float(r[h][i]), __uint_as_float(r[h][i + 1]),
                                                __uint_as_float(r[h][i + 2]), __uint_as_float(r[h][i + 3]));
                    } else {
                        // A lane holds 32 consecutive columns of ONE dW row, so reducing straight from
                        // the registers would touch 32 rows x 16 B per instruction (half-filled sectors,
                        // ncu: 32 sectors / request).  Transpose the 32 x 32 block through shared memory
                        // (row stride 33 words: conflict-free both ways) so that every reduction
                        // instruction covers 4 rows x 128 contiguous bytes.
                        float *tile = epi_tiles + q * (32 * 33);
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 16; ++i) tile[lane * 33 + 16 * h + i] = __uint_as_float(r[h][i]);
                        __syncwarp();
                        const int rr = lane >> 3, cc = (lane & 7) * 4;
                        const int row0 = (gn * p.NB + nb) * kBM + q * 32;
#pragma unroll
                        for (int ps = 0; ps < 8; ++ps) {
                            const int rl = ps * 4 + rr;
                            const float *src = tile + rl * 33 + cc;
                            if (row0 + rl < p.N)
                                red_add_v4(p.dw + (size_t)(row0 + rl) * p.K + tk * p.BN + c0 + cc, src[0], src[1], src[2], src[3]);
                        }
                        __syncwarp();
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// second pass of the two-pass weight gradient: dW[n, k] = sum over splits of the partial slabs, written
// in the parameter's own dtype (no zero-fill, no cast kernel); db likewise.
template <typename TO>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ ws, const float *__restrict__ ws_b, TO *__restrict__ dw,
                    TO *__restrict__ db, int N, int K, int n_pad, int splits) {
    const int kv = K / 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)N * kv;
    if (t < total) {
        const int n = (int)(t / kv), c = (int)(t % kv) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *src = ws + (size_t)n * K + c;
        const size_t slab = (size_t)n_pad * K;
#pragma unroll 4
        for (int s2 = 0; s2 < splits; ++s2) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(src + s2 * slab));
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if constexpr (sizeof(TO) == 2) {
            *reinterpret_cast<uint2 *>(dw + (size_t)n * K + c) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
        } else {
            *reinterpret_cast<float4 *>(dw + (size_t)n * K + c) = a;
        }
    } else if (db && t < total + N) {
        const int n = (int)(t - total);
        float a = 0.f;
        for (int s2 = 0; s2 < splits; ++s2) a += ws_b[(size_t)s2 * n_pad + n];
        if constexpr (sizeof(TO) == 2) db[n] = __float2bfloat16_rn(a); else db[n] = a;
    }
}

// ---- host side ----------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver-API symbol; resolve it through the runtime at first use so that
// the library carries no link-time dependency on libcuda.so (it must load on GPU-less build hosts).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static int make_map_2d(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return -1;
    // row-major (rows, cols) bf16; box = box_rows x 64 elements (128 B inner), SWIZZLE_128B
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(uint16_t)};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims,
                                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

static int make_map_out(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, bool f32) {
    // output (rows, cols): box = 32 rows x 128 B (64 bf16 / 32 fp32 columns), SWIZZLE_128B
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return -1;
    const uint32_t esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * esz};
    cuuint32_t box[2] = {128 / esz, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

// weight-stationary tile width: the largest multiple of 16 that divides N, is <= 256, whose weight
// tile fits 128 KB of shared memory, and (for the 128 B output slabs) is a multiple of `gran` columns
static int pick_bn_ws(int N, int K, int gran) {
    for (int bn = 256; bn >= 16; bn -= 16)
        if (N % bn == 0 && bn % gran == 0 && (long long)bn * K * 2 <= 128 * 1024) return bn;
    return 0;
}

static int pick_bn(int N) {
    for (int bn : {256, 192, 128, 64}) if (N % bn == 0) return bn;
    if (N <= 256 && N % 16 == 0) return N;
    for (int bn = 240; bn >= 16; bn -= 16) if (N % bn == 0) return bn;
    return 0;
}

// weight-stationary launch shared by the forward projection (b_mn = 0, B = W (N, K)) and the input
// gradient (b_mn = 1, B = W (K_red, N_out) read in place)
static int launch_ws(const char *who, const void *x, const void *b, const void *bias, int bias_dtype, void *y,
                     bool f32, int64_t M, int N, int K, int bn_ws, int relu, int b_mn, void *stream,
                     const void *addend = nullptr) {
    CUtensorMap map_a, map_b, map_y;
    if (int e = make_map_2d(&map_a, x, (uint64_t)M, (uint64_t)K, kBM))
        return fail("%s: cuTensorMapEncodeTiled(A) failed (%lld)", who, e);
    if (int e = b_mn ? make_map_2d(&map_b, b, (uint64_t)K, (uint64_t)N, 64)
                     : make_map_2d(&map_b, b, (uint64_t)N, (uint64_t)K, (uint32_t)bn_ws))
        return fail("%s: cuTensorMapEncodeTiled(B) failed (%lld)", who, e);
    if (int e = make_map_out(&map_y, y, (uint64_t)M, (uint64_t)N, f32))
        return fail("%s: cuTensorMapEncodeTiled(Y) failed (%lld)", who, e);
    GemmWsParams q;
    q.M = (int)M; q.N = N; q.K = K; q.BN = bn_ws; q.relu = relu; q.bias = bias;
    q.bias_bf16 = bias_dtype == BEVF_DTYPE_BF16; q.out_f32 = f32 ? 1 : 0; q.b_mn = b_mn;
    q.addend = f32 ? nullptr : reinterpret_cast<const bf16 *>(addend);
    const int b_bytes = ((bn_ws * K * 2) + 1023) & ~1023;
    int stages = (227 * 1024 - 1024 - b_bytes - 32 * 1024 - 1024 - 256) / (kBM * 128);
    if (stages > 6) stages = 6;
    if (stages < 2) return fail("%s: not enough shared memory for the A ring", who);
    q.stages = stages;
    const size_t smem = 1024 + (size_t)b_bytes + (size_t)stages * kBM * 128 + 32 * 1024 + 1024 + 256;
    static int sms_ws = 0;
    if (sms_ws == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms_ws, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_nt_ws_bf16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(gemm_nt_ws_bf16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    const int tiles_m = (int)((M + kBM - 1) / kBM);
    const int grid = tiles_m < sms_ws ? tiles_m : sms_ws;
    if (q.addend)
        gemm_nt_ws_bf16<true><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(map_a, map_b, map_y, q);
    else
        gemm_nt_ws_bf16<false><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(map_a, map_b, map_y, q);
    return check_launch(who);
}

}  // namespace bevf

using namespace bevf;

static int linear_dgrad_impl(const char *who, const void *dy, const void *w, const void *addend, void *dx, int64_t M,
                             int N, int K, void *stream);

extern "C" int bevf_linear_dgrad(const void *dy, const void *w, void *dx, int64_t M, int N, int K, void *stream) {
    return linear_dgrad_impl("bevf_linear_dgrad", dy, w, nullptr, dx, M, N, K, stream);
}

extern "C" int bevf_linear_dgrad_acc(const void *dy, const void *w, const void *addend, void *dx, int64_t M, int N,
                                     int K, void *stream) {
    if (addend && !aligned16(addend)) return fail("%s: addend must be 16-byte aligned", "bevf_linear_dgrad_acc");
    return linear_dgrad_impl("bevf_linear_dgrad_acc", dy, w, addend, dx, M, N, K, stream);
}

static int linear_dgrad_impl(const char *who, const void *dy, const void *w, const void *addend, void *dx, int64_t M,
                             int N, int K, void *stream) {
    if (M < 0 || N <= 0 || K <= 0) return fail("%s: bad dimension", who);
    if (M == 0) return 0;
    if (!dy || !w || !dx) return fail("%s: null pointer argument", who);
    if (N % kBK != 0 || K % 64 != 0) return fail("%s: N and K must be multiples of 64 (got %lld, %lld)", who, N, K);
    if (M >= (1ll << 31)) return fail("%s: M too large", who);
    if (!aligned16(dy) || !aligned16(w) || !aligned16(dx)) return fail("%s: pointers must be 16-byte aligned", who);
    // output columns = K of the weight, reduction = its N rows
    const int bn = pick_bn_ws(K, N, 64);
    if (bn == 0) return fail("%s: no 64-column tile of the weight fits shared memory", who);
    return launch_ws(who, dy, w, nullptr, BEVF_DTYPE_F32, dx, false, M, K, N, bn, 0, 1, stream, addend);
}

extern "C" int bevf_linear_forward(const void *x, const void *w, const void *bias, int bias_dtype,
                                   const void *residual, void *y, int y_dtype, int64_t M, int N, int K,
                                   int relu, void *stream) {
    const char *who = "bevf_linear_forward";
    if (M < 0 || N <= 0 || K <= 0) return fail("%s: bad dimension", who);
    if (M == 0) return 0;
    if (!x || !w || !y) return fail("%s: null pointer argument", who);
    if (K % kBK != 0) return fail("%s: K must be a multiple of 64 (got %lld)", who, K);
    if (N % 16 != 0) return fail("%s: N must be a multiple of 16 (got %lld)", who, N);
    if (M >= (1ll << 31)) return fail("%s: M too large", who);
    if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (bias && !aligned16(bias)) ||
        (residual && !aligned16(residual)))
        return fail("%s: pointers must be 16-byte aligned", who);
    if (y_dtype != BEVF_DTYPE_BF16 && y_dtype != BEVF_DTYPE_F32) return fail("%s: unsupported dtype code", who);
    if (bias && bias_dtype != BEVF_DTYPE_BF16 && bias_dtype != BEVF_DTYPE_F32)
        return fail("%s: unsupported bias dtype code", who);
    static int use_ws = -1;
    if (use_ws < 0) {
        const char *e = getenv("BEVF_GEMM_WS");
        use_ws = e ? atoi(e) : 1;
    }
    const bool f32 = y_dtype == BEVF_DTYPE_F32;
    const int bn_ws = pick_bn_ws(N, K, f32 ? 32 : 64);
    if (use_ws && !residual && bn_ws > 0)
        return launch_ws(who, x, w, bias, bias_dtype, y, f32, M, N, K, bn_ws, relu, 0, stream);
    const int bn = pick_bn(N);
    if (bn == 0) return fail("%s: no tile width divides N", who);

    CUtensorMap map_a, map_b;
    if (int e = make_map_2d(&map_a, x, (uint64_t)M, (uint64_t)K, kBM))
        return fail("%s: cuTensorMapEncodeTiled(A) failed (%lld)", who, e);
    if (int e = make_map_2d(&map_b, w, (uint64_t)N, (uint64_t)K, (uint32_t)bn))
        return fail("%s: cuTensorMapEncodeTiled(B) failed (%lld)", who, e);

    GemmParams p;
    p.M = (int)M; p.N = N; p.K = K; p.BN = bn; p.relu = relu;
    p.bias = bias; p.bias_bf16 = bias_dtype == BEVF_DTYPE_BF16;
    p.residual = reinterpret_cast<const bf16 *>(residual); p.y = y;
    const int stage_bytes = (kBM + bn) * 128;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages > K / kBK * 2) stages = K / kBK * 2;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 /*align*/ + (2 * stages + 2 * kAccStages) * 8 + 16;

    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_nt_bf16<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(gemm_nt_bf16<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    const int tiles = (int)((M + kBM - 1) / kBM) * (N / bn);
    const int grid = tiles < num_sms ? tiles : num_sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (y_dtype == BEVF_DTYPE_BF16)
        gemm_nt_bf16<bf16><<<grid, kGemmThreads, smem, st>>>(map_a, map_b, p);
    else
        gemm_nt_bf16<float><<<grid, kGemmThreads, smem, st>>>(map_a, map_b, p);
    return check_launch(who);
}


extern "C" int bevf_linear_wgrad(const void *dy, const void *x, float *dw, float *db, int64_t M, int N,
                                 int K, void *stream) {
    const char *who = "bevf_linear_wgrad";
    if (M < 0 || N <= 0 || K <= 0) return fail("%s: bad dimension", who);
    if (M == 0) return 0;
    if (!dy || !x || !dw) return fail("%s: null pointer argument", who);
    if (K % 64 != 0 || N % 8 != 0) return fail("%s: K must be a multiple of 64 and N of 8", who);
    if (db && !aligned16(db)) return fail("%s: pointers must be 16-byte aligned", who);
    if (M >= (1ll << 31)) return fail("%s: M too large", who);
    if (!aligned16(dy) || !aligned16(x) || !aligned16(dw)) return fail("%s: pointers must be 16-byte aligned", who);
    int bn = 0;
    for (int c : {256, 192, 128, 64}) if (K % c == 0) { bn = c; break; }
    CUtensorMap map_dy, map_x;
    if (int e = make_map_2d(&map_dy, dy, (uint64_t)M, (uint64_t)N, 64))
        return fail("%s: cuTensorMapEncodeTiled(dY) failed (%lld)", who, e);
    if (int e = make_map_2d(&map_x, x, (uint64_t)M, (uint64_t)K, 64))
        return fail("%s: cuTensorMapEncodeTiled(X) failed (%lld)", who, e);
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_wgrad_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    static int use_v2 = -1;
    if (use_v2 < 0) {
        const char *e = getenv("BEVF_WGRAD_V2");
        use_v2 = e ? atoi(e) : 1;
        cudaFuncSetAttribute(gemm_wgrad2_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    if (use_v2) {
        Wgrad2Params q;
        q.M = (int)M; q.N = N; q.K = K; q.BN = bn; q.dw = dw; q.db = db;
        q.ws = nullptr; q.ws_b = nullptr; q.n_pad = 0;
        const int blocks_n = (N + kBM - 1) / kBM;
        q.NB = (blocks_n >= 2 && 2 * bn <= 512) ? 2 : 1;
        const int groups = (blocks_n + q.NB - 1) / q.NB;
        const int tiles2 = groups * (K / bn);
        int splits2 = (num_sms + tiles2 - 1) / tiles2;                 // ~1 unit per SM
        int rows2 = (int)((M + splits2 - 1) / splits2);
        rows2 = ((rows2 + 63) / 64) * 64;
        splits2 = (int)((M + rows2 - 1) / rows2);
        q.rows_per_split = rows2; q.splits = splits2;
        const int stage2 = q.NB * 2 * 8192 + (bn / 64) * 8192;
        int st2 = (kWgradPipeBytes) / stage2;
        if (st2 > 6) st2 = 6;
        if (st2 < 2) st2 = 2;
        q.stages = st2;
        const size_t smem2 = (size_t)st2 * stage2 + 1024 + (2 * st2 + 2) * 8 + 16 + kWgradEpiBytes;
        const int units2 = tiles2 * splits2;
        const int grid2 = units2 < num_sms ? units2 : num_sms;
        gemm_wgrad2_bf16<<<grid2, kGemmThreads, smem2, (cudaStream_t)stream>>>(map_dy, map_x, q);
        return check_launch(who);
    }
    if (db) return fail("%s: the bias gradient needs the v2 kernel (unset BEVF_WGRAD_V2=0)", who);
    WgradParams p;
    p.M = (int)M; p.N = N; p.K = K; p.BN = bn; p.dw = dw;
    const int tiles = ((N + kBM - 1) / kBM) * (K / bn);
    int splits = (2 * num_sms + tiles - 1) / tiles;                // ~2 units per SM
    int rows = (int)((M + splits - 1) / splits);
    rows = ((rows + 63) / 64) * 64;
    splits = (int)((M + rows - 1) / rows);
    p.rows_per_split = rows; p.splits = splits;
    const int stage_bytes = 2 * 8192 + (bn / 64) * 8192;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 1) * 8 + 16;
    const int units = tiles * splits;
    const int grid = units < num_sms ? units : num_sms;
    gemm_wgrad_bf16<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(map_dy, map_x, p);
    return check_launch(who);
}


// plan shared by the workspace query and the launch
static void wgrad_plan(int64_t M, int N, int K, int num_sms, int &bn, int &NB, int &splits, int &rows, int &n_pad) {
    bn = 0;
    for (int c : {256, 192, 128, 64}) if (K % c == 0) { bn = c; break; }
    const int blocks_n = (N + kBM - 1) / kBM;
    NB = (blocks_n >= 2 && 2 * bn <= 512) ? 2 : 1;
    const int groups = (blocks_n + NB - 1) / NB;
    const int tiles = groups * (K / bn);
    splits = (num_sms + tiles - 1) / tiles;
    rows = (int)((M + splits - 1) / splits);
    rows = ((rows + 63) / 64) * 64;
    splits = (int)((M + rows - 1) / rows);
    n_pad = groups * NB * kBM;
}

extern "C" int64_t bevf_linear_wgrad_workspace_bytes(int64_t M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int bn, NB, splits, rows, n_pad;
    wgrad_plan(M, N, K, sms, bn, NB, splits, rows, n_pad);
    return (int64_t)splits * n_pad * ((int64_t)K + 1) * 4 + 256;
}

extern "C" int bevf_linear_wgrad_out(const void *dy, const void *x, void *dw, void *db, int grad_dtype,
                                     void *workspace, int64_t workspace_bytes, int64_t M, int N, int K,
                                     void *stream) {
    const char *who = "bevf_linear_wgrad_out";
    if (M <= 0 || N <= 0 || K <= 0) return fail("%s: bad dimension", who);
    if (!dy || !x || !dw || !workspace) return fail("%s: null pointer argument", who);
    if (K % 64 != 0 || N % 8 != 0) return fail("%s: K must be a multiple of 64 and N of 8", who);
    if (M >= (1ll << 31)) return fail("%s: M too large", who);
    if (!aligned16(dy) || !aligned16(x) || !aligned16(dw) || !aligned16(workspace))
        return fail("%s: pointers must be 16-byte aligned", who);
    if (grad_dtype != BEVF_DTYPE_BF16 && grad_dtype != BEVF_DTYPE_F32) return fail("%s: unsupported dtype code", who);
    if (workspace_bytes < bevf_linear_wgrad_workspace_bytes(M, N, K))
        return fail("%s: workspace too small (need %lld bytes)", who, bevf_linear_wgrad_workspace_bytes(M, N, K));
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_wgrad2_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    Wgrad2Params q;
    int rows;
    wgrad_plan(M, N, K, sms, q.BN, q.NB, q.splits, rows, q.n_pad);
    q.M = (int)M; q.N = N; q.K = K; q.rows_per_split = rows;
    q.dw = nullptr; q.db = db ? reinterpret_cast<float *>(1) : nullptr;   // non-null => warp 3 sums columns
    q.ws = reinterpret_cast<float *>(workspace);
    q.ws_b = q.ws + (size_t)q.splits * q.n_pad * K;
    CUtensorMap map_dy, map_x;
    if (int e = make_map_2d(&map_dy, dy, (uint64_t)M, (uint64_t)N, 64))
        return fail("%s: cuTensorMapEncodeTiled(dY) failed (%lld)", who, e);
    if (int e = make_map_2d(&map_x, x, (uint64_t)M, (uint64_t)K, 64))
        return fail("%s: cuTensorMapEncodeTiled(X) failed (%lld)", who, e);
    const int stage = q.NB * 2 * 8192 + (q.BN / 64) * 8192;
    int st = (kWgradPipeBytes) / stage;
    if (st > 6) st = 6;
    if (st < 2) st = 2;
    q.stages = st;
    const size_t smem = (size_t)st * stage + 1024 + (2 * st + 2) * 8 + 16 + kWgradEpiBytes;
    const int groups = q.n_pad / (q.NB * kBM);
    const int units = groups * (K / q.BN) * q.splits;
    const int grid = units < sms ? units : sms;
    cudaStream_t cs = (cudaStream_t)stream;
    gemm_wgrad2_bf16<<<grid, kGemmThreads, smem, cs>>>(map_dy, map_x, q);
    if (int e = check_launch(who)) return e;
    const long long total = (long long)N * (K / 4) + (db ? N : 0);
    const unsigned rgrid = (unsigned)((total + 255) / 256);
    if (grad_dtype == BEVF_DTYPE_BF16)
        wgrad_reduce_kernel<bf16><<<rgrid, 256, 0, cs>>>(q.ws, q.ws_b, (bf16 *)dw, (bf16 *)db, N, K, q.n_pad, q.splits);
    else
        wgrad_reduce_kernel<float><<<rgrid, 256, 0, cs>>>(q.ws, q.ws_b, (float *)dw, (float *)db, N, K, q.n_pad, q.splits);
    return check_launch(who);
}
