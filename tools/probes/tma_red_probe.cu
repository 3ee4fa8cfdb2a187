// Probe (development only): rate of 128-byte fp32 row reductions into global memory through
//   (a) REDG.E.ADD.F32x4 issued by the lanes (the sampler backward's path),
//   (b) cp.reduce.async.bulk .add.f32 from shared memory (TMA engine),
//   (c) both at once (half of the rows each).
// Rows are pseudo-random 128 B-aligned addresses in a buffer of 16..512 MB, 8 rows per warp-instruction
// like the bf16 sampler.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_red_probe tma_red_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ void red4(float* p, float a) { asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(a) : "memory"); }

__device__ __forceinline__ void red4_bf16x2(void* p, uint32_t a) { asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(a) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* buf, uint32_t rows_mask, int iters) {
    __shared__ __align__(128) float stage[8][2][8][32];            // warp, buffer, row, 32 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 3, grp = lane >> 2;                     // 4 lanes per row, 8 rows per warp
    uint32_t seed = (blockIdx.x * 8 + warp) * 977u + 12345u;
    for (int it = 0; it < iters; ++it) {
        const uint32_t r = hash32(seed + it * 8 + grp) & rows_mask;  // row index (128 B rows)
        float* dst = buf + (size_t)r * 32;
        const bool use_tma = (MODE == 1) || (MODE == 2 && (it & 1));
        if (MODE == 4) {                                             // full 128 B line per row per instruction
            const int sub8 = lane & 7, g4 = lane >> 3;                 // 8 lanes per row, 4 rows per instruction
            const uint32_t ra = hash32(seed + it * 8 + g4) & rows_mask, rb = hash32(seed + it * 8 + 4 + g4) & rows_mask;
            red4(buf + (size_t)ra * 32 + 4 * sub8, 1.0f);
            red4(buf + (size_t)rb * 32 + 4 * sub8, 1.0f);
        } else if (MODE == 3) {                                             // same row, bf16 accumulation: 64 B per row
            red4_bf16x2(reinterpret_cast<char*>(dst) + 16 * sub, 0x3c003c00u);
        } else if (!use_tma) {
            red4(dst + 4 * sub, 1.0f);
            red4(dst + 16 + 4 * sub, 1.0f);
        } else {
            const int b = (it >> (MODE == 2 ? 1 : 0)) & 1;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
            *reinterpret_cast<float4*>(&stage[warp][b][grp][4 * sub]) = v;
            *reinterpret_cast<float4*>(&stage[warp][b][grp][16 + 4 * sub]) = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (sub == 0) {
                uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[warp][b][grp][0]);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" :: "l"(dst), "r"(s) : "memory");
            }
            __syncwarp();
            if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int run(size_t mb) {
    const size_t bytes = mb << 20;
    printf("-- target buffer %zu MB\n", mb);
    float* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    const uint32_t mask = (uint32_t)(bytes / 128) - 1;
    const int iters = 512, grid = 148 * 8;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[5] = {"REDG.v4 (lanes)", "cp.reduce.async.bulk (TMA)", "half / half", "REDG.v4.bf16x2 (64 B rows)", "REDG.v4 full line / instr"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (mode == 0) probe<0><<<grid, 256>>>(buf, mask, iters);
            if (mode == 1) probe<1><<<grid, 256>>>(buf, mask, iters);
            if (mode == 2) probe<2><<<grid, 256>>>(buf, mask, iters);
            if (mode == 3) probe<3><<<grid, 256>>>(buf, mask, iters);
            if (mode == 4) probe<4><<<grid, 256>>>(buf, mask, iters);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double rows = (double)grid * 8 * iters * 8;
        printf("%-28s %8.3f ms  %7.2f G rows/s  %6.2f TB/s payload  err=%s\n", names[mode], ms, rows / ms / 1e6,
               rows * 128 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(buf);
    return 0;
}

int main(int argc, char** argv) {
    // buffer sizes either side of the 126 MB L2: hit-dominated vs miss-dominated reductions
    const size_t sizes[] = {32, 256};
    for (size_t mb : sizes) run(mb);
    return 0;
}
