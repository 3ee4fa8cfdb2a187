/*
 * Oracle-S: plain-C restatement of the multi-scale deformable attention sampler.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product
 * (bevformer_b200/, projects/).  Allowed callers: tests/, __graft_entry__.smoke(), and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * What it restates.  The reference reaches this arithmetic through
 *   projects/mmdet3d_plugin/bevformer/modules/multi_scale_deformable_attn_function.py:118-124
 *     (ext_module.ms_deform_attn_forward)  and  :150-160 (ms_deform_attn_backward),
 * whose implementation lives in the third-party wheel mmcv-full==1.4.0 (docs/install.md:27), which
 * is NOT under /root/reference and not installable here.  The arithmetic below is the published
 * algorithm (Deformable-DETR's im2col/col2im sampler) as specified in SURVEY.md Appendix A:
 *   x = loc_x*W - 0.5, y = loc_y*H - 0.5; sample skipped unless -1 < x < W and -1 < y < H (float
 *   compare before any int cast); 4-corner bilinear with zero padding; output = sum_l sum_p A*S.
 * Backward accumulates INTO caller-zeroed buffers, like the reference's call site
 *   (multi_scale_deformable_attn_function.py:146-160).
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path, so
 * this file is pinned against outputs of the reference's own Python modules run in the dev
 * container (oracle/mmcv_stub.py -> tests/golden/ (.npz files), made by tests/golden/make_golden.py) and
 * against autograd of the grid_sample form in fp64 (tests/test_oracle.py).
 *
 * Layouts (row-major): value (B,S,M,D); loc (B,Q,M,L,P,2) as (x,y); attn (B,Q,M,L,P);
 * out / grad_out (B,Q,M*D); level_hw (L,2) as (h,w); level_start (L).
 * Compiled twice through REAL = float / double (see oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#ifndef REAL
#define REAL float
#endif
#ifndef SUFFIX
#define SUFFIX f32
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* ---- minimal pthread parallel-for (libgomp is not installed in this image) ---------------------
 * Threads = $MSDA_ORACLE_THREADS or the online core count; iterations are handed out one at a
 * time from an atomic counter, and every iteration writes disjoint outputs, so results do not
 * depend on the thread count. */
typedef void (*FN(job_fn))(int64_t i, void *ctx);
typedef struct { FN(job_fn) fn; void *ctx; int64_t n; int64_t next; } FN(pool_t);
static void *FN(worker)(void *arg) {
    FN(pool_t) *p = (FN(pool_t) *)arg;
    for (;;) {
        int64_t i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (i >= p->n) break;
        p->fn(i, p->ctx);
    }
    return 0;
}
static void FN(parallel_for)(int64_t n, FN(job_fn) fn, void *ctx) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    const char *e = getenv("MSDA_ORACLE_THREADS");
    if (e && atoi(e) > 0) nt = atoi(e);
    if (nt > 256) nt = 256;
    if (nt > n) nt = (long)n;
    FN(pool_t) pool = {fn, ctx, n, 0};
    if (nt <= 1) { FN(worker)(&pool); return; }
    pthread_t th[256];
    long started = 0;
    for (long t = 0; t < nt - 1; ++t)
        if (pthread_create(&th[started], 0, FN(worker), &pool) == 0) ++started;
    FN(worker)(&pool);
    for (long t = 0; t < started; ++t) pthread_join(th[t], 0);
}

typedef struct {
    const REAL *value; const int64_t *level_hw; const int64_t *level_start; const REAL *loc;
    const REAL *attn; const REAL *grad_out; REAL *out; REAL *grad_value; REAL *grad_loc;
    REAL *grad_attn; int64_t B, S, M, D, Q, L, P, chunk;
} FN(args_t);

static inline int in_range(REAL x, REAL y, int H, int W) {
    return x > (REAL)-1 && y > (REAL)-1 && x < (REAL)W && y < (REAL)H;
}

/* forward: parallel over chunks of (b, q) rows; every output element written exactly once. */
static void FN(forward_job)(int64_t job, void *ctx) {
    const FN(args_t) *A = (const FN(args_t) *)ctx;
    const REAL *value = A->value, *loc = A->loc, *attn = A->attn;
    const int64_t *level_hw = A->level_hw, *level_start = A->level_start;
    REAL *out = A->out;
    const int64_t S = A->S, M = A->M, D = A->D, Q = A->Q, L = A->L, P = A->P;
    const int64_t rows = A->B * Q;
    const int64_t r_end = (job + 1) * A->chunk < rows ? (job + 1) * A->chunk : rows;
    for (int64_t r = job * A->chunk; r < r_end; ++r) {
        const int64_t b = r / Q;
        for (int64_t m = 0; m < M; ++m) {
            REAL *o = out + (r * M + m) * D;
            for (int64_t c = 0; c < D; ++c) o[c] = 0;
            for (int64_t l = 0; l < L; ++l) {
                const int H = (int)level_hw[2 * l], W = (int)level_hw[2 * l + 1];
                const REAL *vl = value + ((b * S + level_start[l]) * M + m) * D;
                const int64_t pix = M * D; /* stride between consecutive pixels */
                for (int64_t p = 0; p < P; ++p) {
                    const int64_t si = ((r * M + m) * L + l) * P + p;
                    const REAL x = loc[2 * si] * (REAL)W - (REAL)0.5;
                    const REAL y = loc[2 * si + 1] * (REAL)H - (REAL)0.5;
                    if (!in_range(x, y, H, W)) continue;
                    const REAL a = attn[si];
                    const int x0 = (int)floor((double)x), y0 = (int)floor((double)y);
                    const int x1 = x0 + 1, y1 = y0 + 1;
                    const REAL lx = x - (REAL)x0, ly = y - (REAL)y0, hx = 1 - lx, hy = 1 - ly;
                    const REAL w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
                    const REAL *v00 = (y0 >= 0 && x0 >= 0) ? vl + ((int64_t)y0 * W + x0) * pix : 0;
                    const REAL *v01 = (y0 >= 0 && x1 <= W - 1) ? vl + ((int64_t)y0 * W + x1) * pix : 0;
                    const REAL *v10 = (y1 <= H - 1 && x0 >= 0) ? vl + ((int64_t)y1 * W + x0) * pix : 0;
                    const REAL *v11 = (y1 <= H - 1 && x1 <= W - 1) ? vl + ((int64_t)y1 * W + x1) * pix : 0;
                    for (int64_t c = 0; c < D; ++c) {
                        REAL s = 0;
                        if (v00) s += w00 * v00[c];
                        if (v01) s += w01 * v01[c];
                        if (v10) s += w10 * v10[c];
                        if (v11) s += w11 * v11[c];
                        o[c] += a * s;
                    }
                }
            }
        }
    }
}

void FN(msda_oracle_forward)(const REAL *value, const int64_t *level_hw, const int64_t *level_start,
                             const REAL *loc, const REAL *attn, REAL *out,
                             int64_t B, int64_t S, int64_t M, int64_t D, int64_t Q, int64_t L,
                             int64_t P) {
    FN(args_t) a = {value, level_hw, level_start, loc, attn, 0, out, 0, 0, 0,
                    B, S, M, D, Q, L, P, 64};
    FN(parallel_for)((B * Q + a.chunk - 1) / a.chunk, FN(forward_job), &a);
}

/* backward: parallel over (b, m); grad_value[b,:,m,:] slices are disjoint between iterations, so
 * the scatter needs no atomics and the summation order is deterministic (q, l, p ascending). */
static void FN(backward_job)(int64_t j, void *ctx) {
    const FN(args_t) *A = (const FN(args_t) *)ctx;
    const REAL *value = A->value, *loc = A->loc, *attn = A->attn, *grad_out = A->grad_out;
    const int64_t *level_hw = A->level_hw, *level_start = A->level_start;
    REAL *grad_value = A->grad_value, *grad_loc = A->grad_loc, *grad_attn = A->grad_attn;
    const int64_t S = A->S, M = A->M, D = A->D, Q = A->Q, L = A->L, P = A->P;
    {
        const int64_t b = j / M, m = j % M;
        const int64_t pix = M * D;
        for (int64_t q = 0; q < Q; ++q) {
            const int64_t r = b * Q + q;
            const REAL *g = grad_out + (r * M + m) * D;
            for (int64_t l = 0; l < L; ++l) {
                const int H = (int)level_hw[2 * l], W = (int)level_hw[2 * l + 1];
                const int64_t base = ((b * S + level_start[l]) * M + m) * D;
                for (int64_t p = 0; p < P; ++p) {
                    const int64_t si = ((r * M + m) * L + l) * P + p;
                    const REAL x = loc[2 * si] * (REAL)W - (REAL)0.5;
                    const REAL y = loc[2 * si + 1] * (REAL)H - (REAL)0.5;
                    if (!in_range(x, y, H, W)) continue;
                    const REAL a = attn[si];
                    const int x0 = (int)floor((double)x), y0 = (int)floor((double)y);
                    const int x1 = x0 + 1, y1 = y0 + 1;
                    const REAL lx = x - (REAL)x0, ly = y - (REAL)y0, hx = 1 - lx, hy = 1 - ly;
                    const int ok00 = (y0 >= 0 && x0 >= 0), ok01 = (y0 >= 0 && x1 <= W - 1);
                    const int ok10 = (y1 <= H - 1 && x0 >= 0), ok11 = (y1 <= H - 1 && x1 <= W - 1);
                    const int64_t i00 = base + ((int64_t)y0 * W + x0) * pix;
                    const int64_t i01 = base + ((int64_t)y0 * W + x1) * pix;
                    const int64_t i10 = base + ((int64_t)y1 * W + x0) * pix;
                    const int64_t i11 = base + ((int64_t)y1 * W + x1) * pix;
                    REAL ga = 0, gx = 0, gy = 0;
                    for (int64_t c = 0; c < D; ++c) {
                        const REAL gc = g[c], t = gc * a;
                        const REAL v00 = ok00 ? value[i00 + c] : 0, v01 = ok01 ? value[i01 + c] : 0;
                        const REAL v10 = ok10 ? value[i10 + c] : 0, v11 = ok11 ? value[i11 + c] : 0;
                        if (ok00) grad_value[i00 + c] += hy * hx * t;
                        if (ok01) grad_value[i01 + c] += hy * lx * t;
                        if (ok10) grad_value[i10 + c] += ly * hx * t;
                        if (ok11) grad_value[i11 + c] += ly * lx * t;
                        ga += gc * (hy * hx * v00 + hy * lx * v01 + ly * hx * v10 + ly * lx * v11);
                        gx += t * (-hy * v00 + hy * v01 - ly * v10 + ly * v11);
                        gy += t * (-hx * v00 - lx * v01 + hx * v10 + lx * v11);
                    }
                    grad_attn[si] += ga;
                    grad_loc[2 * si] += (REAL)W * gx;
                    grad_loc[2 * si + 1] += (REAL)H * gy;
                }
            }
        }
    }
}

void FN(msda_oracle_backward)(const REAL *value, const int64_t *level_hw,
                              const int64_t *level_start, const REAL *loc, const REAL *attn,
                              const REAL *grad_out, REAL *grad_value, REAL *grad_loc,
                              REAL *grad_attn, int64_t B, int64_t S, int64_t M, int64_t D,
                              int64_t Q, int64_t L, int64_t P) {
    FN(args_t) a = {value, level_hw, level_start, loc, attn, grad_out, 0, grad_value, grad_loc,
                    grad_attn, B, S, M, D, Q, L, P, 1};
    FN(parallel_for)(B * M, FN(backward_job), &a);
}
