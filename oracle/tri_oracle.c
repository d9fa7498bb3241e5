/*
 * oracle/tri_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement of the reference's point-to-triangle search TriDistanceKernel
 * (tri_distance/tri_distance.cu:94-211 + device helpers :19-91).  The reference has NO CPU path
 * and no golden vectors for this kernel: parity is pinned on the GPU box against
 * oracle/_ref/libref_cuda.so (the reference .cu compiled unmodified) and against tests/golden/
 * fixtures produced by it.
 *
 * The arithmetic below is the kernel AS COMPILED by nvcc 12.9.86 for sm_100a (default flags,
 * -fmad=true), decoded from `cuobjdump -sass` (see DESIGN.md "TriDistance arithmetic"):
 *   S1  dot(a,b)      = fma(a.z,b.z, fma(a.x,b.x, rn(a.y*b.y)))          (tri_distance.cu:38-40)
 *   S2  cross(X,Y).x  = fma(X.y,Y.z, -rn(X.z*Y.y)), .y/.z cyclic          (tri_distance.cu:34-37)
 *   S3  point_at      = fma(D.c, t, A.c)                                   (tri_distance.cu:73-76)
 *   S4  Project_Edge  = IEEE div( S1(p-A, D), S1(D,D) )                    (tri_distance.cu:46-52)
 *   S5  Project_Plane : n.c = rn(T.c * rcp(sqrt(S1(T,T)))); dist = S1(p-A, n);
 *                       proj.c = fma(-n.c, dist, p.c)                      (tri_distance.cu:77-91)
 *   S6  d = S1(p-p2, p-p2); select  k==0 || d<best                         (tri_distance.cu:191-198)
 * Build with -ffp-contract=off so that gcc adds no contraction of its own.
 */
#include <math.h>
#include <stddef.h>
#include "parfor.h"

#define ORACLE_TILE 512 /* tri_distance.cu:106 */

typedef struct { float x, y, z; } V3;

static inline V3 v3(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static inline V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float dot_s1(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }
static inline V3 cross_s2(V3 X, V3 Y)
{
    return v3(fmaf(X.y, Y.z, -(X.z * Y.y)),
              fmaf(X.z, Y.x, -(X.x * Y.z)),
              fmaf(X.x, Y.y, -(X.y * Y.x)));
}
static inline V3 point_at_s3(V3 A, V3 D, float t)
{
    return v3(fmaf(D.x, t, A.x), fmaf(D.y, t, A.y), fmaf(D.z, t, A.z));
}
static inline float project_edge_s4(V3 A, V3 D, V3 p)
{
    return dot_s1(vsub(p, A), D) / dot_s1(D, D);
}
/* in_range (tri_distance.cu:59-66): NaN is out of range. */
static inline int in_range(float a) { return (a <= 1.0f) && (a >= 0.0f); }
/* is_above (tri_distance.cu:67-71) */
static inline int is_above(V3 A, V3 D, V3 T, V3 p)
{
    return dot_s1(cross_s2(T, D), vsub(p, A)) > 0.0f;
}
static inline V3 project_plane_s5(V3 orig, V3 T, V3 p)
{
    float inv = 1.0f / sqrtf(dot_s1(T, T));
    V3 n = v3(T.x * inv, T.y * inv, T.z * inv);
    float dist = dot_s1(vsub(p, orig), n);
    return v3(fmaf(-n.x, dist, p.x), fmaf(-n.y, dist, p.y), fmaf(-n.z, dist, p.z));
}

/*
 * One (point, triangle) evaluation: returns squared distance, writes the region option
 * (tri_distance.cu:136-191).  Note the option-6 quirk (:180): the CA-edge closest point is built
 * with AB's direction vector; reproduced because it decides the ranking of triangles.
 */
float oracle_tri_pair(const float *p_, const float *a_, const float *b_, const float *c_, int *option)
{
    V3 p = v3(p_[0], p_[1], p_[2]);
    V3 A = v3(a_[0], a_[1], a_[2]), B = v3(b_[0], b_[1], b_[2]), C = v3(c_[0], c_[1], c_[2]);
    V3 AB = vsub(B, A), BC = vsub(C, B), CA = vsub(A, C);
    V3 T = cross_s2(vsub(A, B), vsub(A, C));
    float uab = project_edge_s4(A, AB, p);
    float uca = project_edge_s4(C, CA, p);
    V3 p2;
    int opt;
    if (uca > 1.0f && uab < 0.0f) {
        p2 = A; opt = 1;
    } else {
        float ubc = project_edge_s4(B, BC, p);
        if (uab > 1.0f && ubc < 0.0f) {
            p2 = B; opt = 2;
        } else if (ubc > 1.0f && uca < 0.0f) {
            p2 = C; opt = 3;
        } else if (in_range(uab) && !is_above(A, AB, T, p)) {
            p2 = point_at_s3(A, AB, uab); opt = 4;
        } else if (in_range(ubc) && !is_above(B, BC, T, p)) {
            p2 = point_at_s3(B, BC, ubc); opt = 5;
        } else if (in_range(uca) && !is_above(C, CA, T, p)) {
            p2 = point_at_s3(C, AB, uca); opt = 6; /* sic: AB_Delta, tri_distance.cu:180 */
        } else {
            p2 = project_plane_s5(A, T, p); opt = 0;
        }
    }
    *option = opt;
    V3 df = vsub(p, p2);
    return dot_s1(df, df);
}

/*
 * TriDistanceKernel + launcher (tri_distance.cu:94-228), device semantics restated serially:
 * 512-triangle tiles (:106-118), scan truncated to a multiple of 4 per tile (:129,:134),
 * best initialised to 10000 (:127), strict '<' in tile (:194), strict '>' across tiles (:202),
 * option carried as float and truncated on store (:128,:204).
 */
typedef struct { int b, n, m; const float *xyz, *tri1, *tri2, *tri3; float *dist; int *point, *index; } tri_ctx;
#define TRI_CTX_UNPACK \
    tri_ctx *c = (tri_ctx *)vctx; const int n = c->n, m = c->m; const float *xyz = c->xyz; \
    const float *tri1 = c->tri1, *tri2 = c->tri2, *tri3 = c->tri3; float *dist = c->dist; \
    int *point = c->point, *index = c->index;

static void tri_rows(long begin, long end, void *vctx)
{
    TRI_CTX_UNPACK
    for (long row = begin; row < end; row++) {
        {
            const int i = (int)(row / n), j = (int)(row % n);
            const float *p = xyz + ((size_t)i * n + j) * 3;
            float res = 0.0f;
            int res_opt = 0, res_i = 0;
            for (int k2 = 0; k2 < m; k2 += ORACLE_TILE) {
                int end_k = (m < k2 + ORACLE_TILE ? m : k2 + ORACLE_TILE) - k2;
                int end_ka = end_k - (end_k & 3);
                float best = 10000.0f;
                float best_option = 0.0f;
                int best_i = 0;
                for (int k = 0; k < end_ka; k++) {
                    size_t t = ((size_t)i * m + k2 + k) * 3;
                    int option;
                    float d = oracle_tri_pair(p, tri1 + t, tri2 + t, tri3 + t, &option);
                    if (k == 0 || d < best) {
                        best = d;
                        best_option = (float)option;
                        best_i = k + k2;
                    }
                }
                if (k2 == 0 || res > best) {
                    res = best;
                    res_opt = (int)best_option;
                    res_i = best_i;
                }
            }
            dist[(size_t)i * n + j] = res;
            point[(size_t)i * n + j] = res_opt;
            index[(size_t)i * n + j] = res_i;
        }
    }
}

void oracle_tri_fwd(int b, int n, const float *xyz, int m, const float *tri1, const float *tri2,
                    const float *tri3, float *dist, int *point, int *index)
{
    tri_ctx c = {b, n, m, xyz, tri1, tri2, tri3, dist, point, index};
    parfor((long)b * n, tri_rows, &c);
}

/* Same search without the tail skip (checker for GEOM_FLAG_EXACT_TAIL). */
static void tri_rows_exact(long begin, long end, void *vctx)
{
    TRI_CTX_UNPACK
    for (long row = begin; row < end; row++) {
        {
            const int i = (int)(row / n), j = (int)(row % n);
            const float *p = xyz + ((size_t)i * n + j) * 3;
            float best = 10000.0f;
            int best_opt = 0, best_i = 0;
            for (int k = 0; k < m; k++) {
                size_t t = ((size_t)i * m + k) * 3;
                int option;
                float d = oracle_tri_pair(p, tri1 + t, tri2 + t, tri3 + t, &option);
                if (k == 0 || d < best) {
                    best = d;
                    best_opt = option;
                    best_i = k;
                }
            }
            dist[(size_t)i * n + j] = best;
            point[(size_t)i * n + j] = best_opt;
            index[(size_t)i * n + j] = best_i;
        }
    }
}

void oracle_tri_fwd_exact(int b, int n, const float *xyz, int m, const float *tri1,
                          const float *tri2, const float *tri3, float *dist, int *point, int *index)
{
    tri_ctx c = {b, n, m, xyz, tri1, tri2, tri3, dist, point, index};
    parfor((long)b * n, tri_rows_exact, &c);
}
