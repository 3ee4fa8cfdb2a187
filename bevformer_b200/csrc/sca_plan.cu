// Device-side construction of SpatialCrossAttention's in-view (camera, query) pair list -- no host
// synchronisation, capturable in a CUDA graph, valid for a new lidar2img every frame.
//
// Replaces the reference's per-layer host round trip
//     indexes = [mask_per_img[0].sum(-1).nonzero().squeeze(-1) ...];  max_len = max(len(each))
// (projects/mmdet3d_plugin/bevformer/modules/spatial_cross_attention.py:138-141) and the zero-padded
// re-batch that follows (:144-153).  Quirk kept: the lists come from batch item 0's mask for every batch
// item (:139); the divisor `count` uses each item's own mask (:169-171).
//
// Output contract (what every row-list kernel of this library understands):
//   pair_q / pair_cam (capacity,)    the in-view pairs, camera-major; inside a camera the queries follow
//                                    `qorder` (8x8 BEV tiles) or ascend; entries >= num_pairs hold -1
//   row_map (B * capacity,)          value map b * ncam + cam of sampler row b * capacity + r, -1 for unused rows
//   pair_of (ncam, Nq)               row of (cam, q) or -1
//   inv_count (B, Nq)                1 / max(1, #cameras that see q in batch item b)
//   map_range (B * ncam, 2)          [first, end) sampler rows of every value map (rows of a map are contiguous)
//   counters[0] = num_pairs (may exceed capacity), counters[1] = 1 if it did (pairs beyond capacity are dropped)
// Stable stream compaction in three small launches: per-block hit counts, block offsets + local scan,
// tail marking + inv_count.
#include "common.cuh"

namespace bevf {

constexpr int kPlanThreads = 256;

__device__ __forceinline__ bool plan_hit(const unsigned char *mask, int cam, int q, int B, int Nq, int D) {
    const unsigned char *p = mask + (((long long)cam * B + 0) * Nq + q) * D;     // batch item 0 (quirk 1)
    bool h = false;
    for (int d = 0; d < D; ++d) h = h || (p[d] != 0);
    return h;
}

__global__ void __launch_bounds__(kPlanThreads)
sca_plan_count(const unsigned char *__restrict__ mask, const int *__restrict__ qorder,
               int *__restrict__ block_counts, int *__restrict__ cam_first, int B, int ncam, int Nq, int D) {
    const long long i = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
    if (blockIdx.x == 0 && (int)threadIdx.x < ncam) cam_first[threadIdx.x] = 0x7fffffff;
    bool hit = false;
    if (i < (long long)ncam * Nq) {
        const int cam = (int)(i / Nq), k = (int)(i % Nq);
        hit = plan_hit(mask, cam, qorder ? __ldg(qorder + k) : k, B, Nq, D);
    }
    const int n = __syncthreads_count(hit);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(kPlanThreads)
sca_plan_fill(const unsigned char *__restrict__ mask, const int *__restrict__ qorder,
              const int *__restrict__ block_counts, int *__restrict__ cam_first, int *__restrict__ pair_q,
              int *__restrict__ pair_cam, int *__restrict__ pair_of, int *__restrict__ row_map,
              int *__restrict__ counters, int B, int ncam, int Nq, int D, int capacity) {
    __shared__ int s_red[kPlanThreads / 32];
    __shared__ int s_base;
    // ---- offset of this block = sum of the counts of the blocks before it
    int part = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += kPlanThreads) part += block_counts[j];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kPlanThreads / 32; ++w) t += s_red[w];
        s_base = t;
    }
    __syncthreads();
    const int base = s_base;
    __syncthreads();
    // ---- stable local scan
    const long long i = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
    const bool in = i < (long long)ncam * Nq;
    int cam = 0, q = 0;
    bool hit = false;
    if (in) {
        cam = (int)(i / Nq);
        const int k = (int)(i % Nq);
        q = qorder ? __ldg(qorder + k) : k;
        hit = plan_hit(mask, cam, q, B, Nq, D);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_red[w];
    const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
    if (in) {
        int slot = -1;
        if (hit && pos < capacity) {
            slot = pos;
            pair_q[pos] = q;
            pair_cam[pos] = cam;
            for (int b = 0; b < B; ++b) row_map[(long long)b * capacity + pos] = b * ncam + cam;
        }
        pair_of[(long long)cam * Nq + q] = slot;
        if (slot >= 0) atomicMin(cam_first + cam, slot);         // first row of this camera's list
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        int total = base;
        for (int w = 0; w < kPlanThreads / 32; ++w) total += s_red[w];
        counters[0] = total;
        counters[1] = total > capacity ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kPlanThreads)
sca_plan_finish(const unsigned char *__restrict__ mask, const int *__restrict__ counters,
                int *__restrict__ pair_q, int *__restrict__ pair_cam, int *__restrict__ row_map,
                float *__restrict__ inv_count, const int *__restrict__ cam_first,
                int *__restrict__ map_range, int B, int ncam, int Nq, int D, int capacity) {
    const long long t = (long long)blockIdx.x * kPlanThreads + threadIdx.x;
    const int n = min(counters[0], capacity);
    if (t < (long long)B * ncam) {                              // row range of value map b * ncam + cam
        const int cam = (int)(t % ncam), b = (int)(t / ncam);
        int first = n, end = n;                                 // a camera that sees nothing gets an empty range
        for (int c = ncam - 1; c >= cam; --c) {
            const int f = min(cam_first[c], n);
            if (c > cam) end = min(end, f); else first = min(f, end);
        }
        map_range[2 * t] = b * capacity + first;
        map_range[2 * t + 1] = b * capacity + end;
    }
    if (t < capacity && t >= n) {                               // unused tail rows
        pair_q[t] = -1;
        pair_cam[t] = -1;
        for (int b = 0; b < B; ++b) row_map[(long long)b * capacity + t] = -1;
    }
    if (t < (long long)B * Nq) {                                // spatial_cross_attention.py:169-171
        const int q = (int)(t % Nq), b = (int)(t / Nq);
        int cnt = 0;
        for (int cam = 0; cam < ncam; ++cam) {
            const unsigned char *p = mask + (((long long)cam * B + b) * Nq + q) * D;
            bool h = false;
            for (int d = 0; d < D; ++d) h = h || (p[d] != 0);
            cnt += h ? 1 : 0;
        }
        inv_count[t] = 1.f / (float)max(cnt, 1);
    }
}

}  // namespace bevf

using namespace bevf;

extern "C" int64_t bevf_sca_plan_workspace_ints(int ncam, int Nq) {
    return ((int64_t)ncam * Nq + kPlanThreads - 1) / kPlanThreads + ncam;      // block counts + first row per camera
}

extern "C" int bevf_sca_plan_build(const unsigned char *bev_mask, const int32_t *qorder, int32_t *pair_q,
                                   int32_t *pair_cam, int32_t *pair_of, int32_t *row_map, float *inv_count,
                                   int32_t *map_range, int32_t *counters, int32_t *workspace, int B, int ncam,
                                   int Nq, int D, int capacity, void *stream) {
    const char *who = "bevf_sca_plan_build";
    if (B <= 0 || ncam <= 0 || Nq <= 0 || D <= 0 || capacity <= 0) return fail("%s: non-positive dimension", who);
    if (!bev_mask || !pair_q || !pair_cam || !pair_of || !row_map || !inv_count || !map_range || !counters ||
        !workspace)
        return fail("%s: null pointer argument", who);
    cudaStream_t st = (cudaStream_t)stream;
    const long long seq = (long long)ncam * Nq;
    const unsigned blocks = (unsigned)((seq + kPlanThreads - 1) / kPlanThreads);
    int32_t *cam_first = workspace + blocks;
    if (ncam > kPlanThreads) return fail("%s: at most 256 cameras", who);
    sca_plan_count<<<blocks, kPlanThreads, 0, st>>>(bev_mask, qorder, workspace, cam_first, B, ncam, Nq, D);
    if (int e = check_launch(who)) return e;
    sca_plan_fill<<<blocks, kPlanThreads, 0, st>>>(bev_mask, qorder, workspace, cam_first, pair_q, pair_cam,
                                                  pair_of, row_map, counters, B, ncam, Nq, D, capacity);
    if (int e = check_launch(who)) return e;
    long long fin = (long long)B * Nq > capacity ? (long long)B * Nq : capacity;
    if (fin < (long long)B * ncam) fin = (long long)B * ncam;
    sca_plan_finish<<<(unsigned)((fin + kPlanThreads - 1) / kPlanThreads), kPlanThreads, 0, st>>>(
        bev_mask, counters, pair_q, pair_cam, row_map, inv_count, cam_first, map_range, B, ncam, Nq, D, capacity);
    return check_launch(who);
}
