/*
 * oracle/chamfer_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the reference's Chamfer nearest-neighbour search, used as the checker for
 * the sm_100a kernels in geometrics_b200/csrc/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this.
 *
 * Parity status: the reference ships NO golden vectors for this path.  This oracle
 * is pinned against (1) oracle/_ref/libref_cpu.so = the reference's own my_lib.c compiled
 * unmodified (tests/test_oracle_vs_ref.py, CPU) and (2) oracle/_ref/libref_cuda.so = the
 * reference's own chamfer_distance.cu compiled unmodified, run on the GPU box
 * (tests/test_gpu_chamfer.py), and against tests/golden/ fixtures generated from (2).
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: every rounding below is spelled out with fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "parfor.h"

#define ORACLE_TILE 512 /* chamfer_distance.cu:15  const int batch=512 */

/* Squared distance exactly as nvcc 12.9 compiles chamfer_distance.cu:35-38 for sm_100a:
 *   x2 = cand.x - query.x ...; d = x2*x2 + y2*y2 + z2*z2
 * becomes FMUL(y2,y2) -> FFMA(x2,x2,.) -> FFMA(z2,z2,.)   (SASS inspected). */
static inline float sqdist_ref(float qx, float qy, float qz, float cx, float cy, float cz)
{
    float x2 = cx - qx;
    float y2 = cy - qy;
    float z2 = cz - qz;
    return fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
}

/*
 * One launch of ChamferDistanceKernel (chamfer_distance.cu:6-55), device semantics restated
 * serially: 512-candidate tiles (:19), per-tile scan truncated to a multiple of 4 (:31,:33 -- the
 * "tail skip"), strict '<' inside a tile (:39), strict '>' across tiles (:47).  `result` /
 * `result_i` need no initialisation (first tile always writes).
 */
typedef struct { int b, n, m; const float *xyz, *xyz2; float *result; int *result_i; } dir_ctx;

static void chamfer_dir_rows(long begin, long end, void *vctx)
{
    dir_ctx *c = (dir_ctx *)vctx;
    const int n = c->n, m = c->m;
    const float *xyz = c->xyz, *xyz2 = c->xyz2;
    float *result = c->result;
    int *result_i = c->result_i;
    for (long row = begin; row < end; row++) {
        {
            const int i = (int)(row / n), j = (int)(row % n);
            const float x1 = xyz[((size_t)i * n + j) * 3 + 0];
            const float y1 = xyz[((size_t)i * n + j) * 3 + 1];
            const float z1 = xyz[((size_t)i * n + j) * 3 + 2];
            float res = 0.0f;
            int res_i = 0;
            for (int k2 = 0; k2 < m; k2 += ORACLE_TILE) {
                int end_k = (m < k2 + ORACLE_TILE ? m : k2 + ORACLE_TILE) - k2;
                int end_ka = end_k - (end_k & 3);
                const float *buf = xyz2 + ((size_t)i * m + k2) * 3;
                float best = 0.0f;
                int best_i = 0;
                for (int k = 0; k < end_ka; k++) {
                    float d = sqdist_ref(x1, y1, z1, buf[k * 3 + 0], buf[k * 3 + 1], buf[k * 3 + 2]);
                    if (k == 0 || d < best) {
                        best = d;
                        best_i = k + k2;
                    }
                }
                if (k2 == 0 || res > best) {
                    res = best;
                    res_i = best_i;
                }
            }
            result[(size_t)i * n + j] = res;
            result_i[(size_t)i * n + j] = res_i;
        }
    }
}

void oracle_chamfer_dir(int b, int n, const float *xyz, int m, const float *xyz2,
                        float *result, int *result_i)
{
    dir_ctx c = {b, n, m, xyz, xyz2, result, result_i};
    parfor((long)b * n, chamfer_dir_rows, &c);
}

/* ChamferDistanceKernelLauncher (chamfer_distance.cu:57-73): two launches, roles swapped. */
void oracle_chamfer_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                        float *dist1, int *idx1, float *dist2, int *idx2)
{
    oracle_chamfer_dir(b, n, xyz1, m, xyz2, dist1, idx1);
    oracle_chamfer_dir(b, m, xyz2, n, xyz1, dist2, idx2);
}

/*
 * "Exact" variant (no tail skip): what the kernel would compute had the legacy tail loop
 * (old_GEOMetrics/chamfer_distance/src/nnd_cuda.cu:110-119) not been deleted.  Checker for the
 * product's opt-in GEOM_FLAG_EXACT_TAIL mode.  Lowest index among exact fp32 minima.
 */
static void chamfer_dir_exact_rows(long begin, long end, void *vctx)
{
    dir_ctx *c = (dir_ctx *)vctx;
    const int n = c->n, m = c->m;
    const float *xyz = c->xyz, *xyz2 = c->xyz2;
    float *result = c->result;
    int *result_i = c->result_i;
    for (long row = begin; row < end; row++) {
        {
            const int i = (int)(row / n), j = (int)(row % n);
            const float x1 = xyz[((size_t)i * n + j) * 3 + 0];
            const float y1 = xyz[((size_t)i * n + j) * 3 + 1];
            const float z1 = xyz[((size_t)i * n + j) * 3 + 2];
            const float *buf = xyz2 + (size_t)i * m * 3;
            float best = 0.0f;
            int best_i = 0;
            for (int k = 0; k < m; k++) {
                float d = sqdist_ref(x1, y1, z1, buf[k * 3 + 0], buf[k * 3 + 1], buf[k * 3 + 2]);
                if (k == 0 || d < best) {
                    best = d;
                    best_i = k;
                }
            }
            result[(size_t)i * n + j] = best;
            result_i[(size_t)i * n + j] = best_i;
        }
    }
}

void oracle_chamfer_dir_exact(int b, int n, const float *xyz, int m, const float *xyz2,
                              float *result, int *result_i)
{
    dir_ctx c = {b, n, m, xyz, xyz2, result, result_i};
    parfor((long)b * n, chamfer_dir_exact_rows, &c);
}

/*
 * Restatement of the reference CPU path nnsearch (old_GEOMetrics/chamfer_distance/src/my_lib.c:4-26):
 * single thread, fp32 products summed in fp32 WITHOUT fma, running best widened to double.
 * This is the "port" CPU baseline; the unmodified original lives in oracle/_ref/libref_cpu.so.
 */
void oracle_nnsearch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx)
{
    for (int i = 0; i < b; i++) {
        for (int j = 0; j < n; j++) {
            float x1 = xyz1[((size_t)i * n + j) * 3 + 0];
            float y1 = xyz1[((size_t)i * n + j) * 3 + 1];
            float z1 = xyz1[((size_t)i * n + j) * 3 + 2];
            double best = 0;
            int besti = 0;
            const float *q = xyz2 + (size_t)i * m * 3;
            for (int k = 0; k < m; k++) {
                float x2 = q[k * 3 + 0] - x1;
                float y2 = q[k * 3 + 1] - y1;
                float z2 = q[k * 3 + 2] - z1;
                double d = x2 * x2 + y2 * y2 + z2 * z2;
                if (k == 0 || d < best) {
                    best = d;
                    besti = k;
                }
            }
            dist[(size_t)i * n + j] = (float)best;
            idx[(size_t)i * n + j] = besti;
        }
    }
}

/*
 * Chamfer loss forward as the caller computes it from the indices
 * (utils.py:402-417 batch_point_to_point; utils.py:451-462):
 *   sum1 = sum_j ||xyz2[idx2->...]...   -- here restated on raw (xyz1, xyz2, idx1, idx2):
 *   s1 = sum over xyz1 rows j of ||xyz2[idx1[j]] - xyz1[j]||^2     (mean -> dist_2 in utils.py:417)
 *   s2 = sum over xyz2 rows k of ||xyz1[idx2[k]] - xyz2[k]||^2     (mean -> dist_1 in utils.py:416)
 * Accumulated in double so the checker is order-independent; the GPU result is compared to
 * <=1e-5 relative.
 */
void oracle_chamfer_loss_sums(int b, int n, const float *xyz1, int m, const float *xyz2,
                              const int *idx1, const int *idx2, double *s1, double *s2)
{
    double a1 = 0, a2 = 0;
    for (int i = 0; i < b; i++) {
        for (int j = 0; j < n; j++) {
            const float *p = xyz1 + ((size_t)i * n + j) * 3;
            const float *q = xyz2 + ((size_t)i * m + idx1[(size_t)i * n + j]) * 3;
            for (int c = 0; c < 3; c++) {
                float df = q[c] - p[c];
                a1 += (double)df * df;
            }
        }
        for (int k = 0; k < m; k++) {
            const float *p = xyz2 + ((size_t)i * m + k) * 3;
            const float *q = xyz1 + ((size_t)i * n + idx2[(size_t)i * m + k]) * 3;
            for (int c = 0; c < 3; c++) {
                float df = q[c] - p[c];
                a2 += (double)df * df;
            }
        }
    }
    *s1 = a1;
    *s2 = a2;
}

/*
 * Gradient of  L = w1 * s1 + w2 * s2  (s1, s2 as above, indices held constant), restating the
 * scatter in nnd_backward (old_GEOMetrics/chamfer_distance/src/my_lib.c:67-108) with
 * graddist == w1 / w2 everywhere.  This equals what autograd produces for utils.py:413-417
 * (index_select -> index_add).  Accumulates in double, writes fp32.
 */
void oracle_chamfer_loss_bwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                             const int *idx1, const int *idx2, float w1, float w2,
                             float *grad1, float *grad2)
{
    size_t n1 = (size_t)b * n * 3, n2 = (size_t)b * m * 3;
    double *acc1 = (double *)__builtin_malloc(sizeof(double) * (n1 + n2));
    double *acc2 = acc1 + n1;
    memset(acc1, 0, sizeof(double) * (n1 + n2));
    for (int i = 0; i < b; i++) {
        for (int j = 0; j < n; j++) {
            size_t pj = ((size_t)i * n + j) * 3;
            size_t qj = ((size_t)i * m + idx1[(size_t)i * n + j]) * 3;
            for (int c = 0; c < 3; c++) {
                double g = 2.0 * w1 * ((double)xyz1[pj + c] - (double)xyz2[qj + c]);
                acc1[pj + c] += g;
                acc2[qj + c] -= g;
            }
        }
        for (int k = 0; k < m; k++) {
            size_t pk = ((size_t)i * m + k) * 3;
            size_t qk = ((size_t)i * n + idx2[(size_t)i * m + k]) * 3;
            for (int c = 0; c < 3; c++) {
                double g = 2.0 * w2 * ((double)xyz2[pk + c] - (double)xyz1[qk + c]);
                acc2[pk + c] += g;
                acc1[qk + c] -= g;
            }
        }
    }
    for (size_t t = 0; t < n1; t++) grad1[t] = (float)acc1[t];
    for (size_t t = 0; t < n2; t++) grad2[t] = (float)acc2[t];
    __builtin_free(acc1);
}
