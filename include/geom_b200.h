/*
 * geom_b200.h -- C ABI of libgeom_b200.so: the B200 (sm_100a) implementation of the GEOMetrics
 * geometric hot path.  Plain pointers and sizes only; no torch / ATen types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository root).  Conventions shared by all entry points:
 *   - all pointers are DEVICE pointers on the current CUDA device unless stated otherwise;
 *   - tensors are dense row-major fp32 / int32, exactly the layouts the reference passes
 *     (points (B,N,3); per-face vertices (B,F,3); indices (B,N));
 *   - the caller allocates every output and every workspace; nothing is retained across calls and calls are
 *     thread-safe.  Process-wide state is limited to: one atomic bit per (kernel, device) recording that the kernel
 *     was opted in to > 48 KB of dynamic shared memory, and the debug knobs of geom_debug_set (tests / tuning only,
 *     all zero = library defaults; nothing is read from the environment);
 *   - work is enqueued on `stream` (a cudaStream_t; 0 = legacy default stream, which is what the
 *     reference launchers use, chamfer_distance.cu:67-68 / tri_distance.cu:223) and the call
 *     returns without synchronising;
 *   - return value: 0 = success, <0 = GEOM_ERR_* (argument errors), >0 = a cudaError_t value.
 *     (The reference launchers are declared `int` in the .cpp but defined `void` in the .cu and
 *     only printf CUDA errors: chamfer_distance.cpp:4 vs chamfer_distance.cu:57,70-72.)
 */
#ifndef GEOM_B200_H
#define GEOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *geom_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GEOM_API __attribute__((visibility("default")))
#else
#define GEOM_API
#endif

#define GEOM_ABI_VERSION 1

#define GEOM_OK 0
#define GEOM_ERR_INVALID_ARG (-1) /* null pointer, negative size, unsupported size */
#define GEOM_ERR_WORKSPACE (-2)   /* workspace null or smaller than *_workspace_bytes() */
#define GEOM_ERR_UNSUPPORTED_DEVICE (-3)

/* flags */
#define GEOM_FLAG_NONE 0u
/* Scan every candidate.  Default (flag clear) is bit-compatible with the reference kernels, which
 * skip the last (m mod 512) mod 4 candidates of the final 512-tile (chamfer_distance.cu:31-33,
 * tri_distance.cu:129,134) and, for Chamfer, zero every output when that final tile holds 1-3
 * candidates (chamfer_distance.cu:29-30,47). */
#define GEOM_FLAG_EXACT_TAIL 1u
/* Chamfer only: use the simple one-direction-at-a-time kernel instead of the fused two-way one. */
#define GEOM_FLAG_ONEWAY 2u
/* Chamfer and TriDistance: tensor-core filtered exact search (tcgen05 bf16 contraction as a conservative filter,
 * exact fp32 re-evaluation of the survivors); same results, bit for bit.  Needs a workspace of
 * geom_chamfer_tensor_workspace_bytes().  With neither GEOM_FLAG_TENSOR nor GEOM_FLAG_FP32 the library picks:
 * the tensor-core search when the workspace passed is large enough for it (and the sizes allow), the
 * CUDA-core search otherwise. */
#define GEOM_FLAG_TENSOR 4u
/* Chamfer and TriDistance: force the CUDA-core brute-force search. */
#define GEOM_FLAG_FP32 8u

GEOM_API int geom_abi_version(void);
/* Static string for a status returned by any geom_* call. */
GEOM_API const char *geom_status_string(int status);
/* SM count and compute capability of the current device (the library is built for sm_100a only). */
GEOM_API int geom_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Optional per-kernel device timing, thread-local to the calling thread (the only state the library
 * keeps).  Between geom_prof_begin() and geom_prof_end() every kernel this thread launches through
 * the API is bracketed by CUDA events on its launch stream; geom_prof_end() synchronises on them and
 * returns, per kernel id, the summed device milliseconds and the launch count (arrays of
 * GEOM_PROF_KERNELS entries, either may be NULL).  Used by bench.py for the roofline figures. */
#define GEOM_PROF_KERNELS 26
GEOM_API int geom_prof_begin(void);
GEOM_API int geom_prof_end(float *ms_per_kernel, int *launches_per_kernel);
GEOM_API const char *geom_prof_kernel_name(int id);

/* Debug knobs for tests and tuning sweeps (process-wide, atomic).  value 0 restores the library default.  Production
 * callers never need them. */
#define GEOM_DEBUG_TC_EPS_FACTOR 0      /* multiplies the Chamfer tensor-core filter margin (margin study) */
#define GEOM_DEBUG_TRI_PR 1             /* CUDA-core TriDistance: points per lane (1, 2, 4) ... */
#define GEOM_DEBUG_TRI_CHUNK 2          /* ... and triangles per work unit (64..512); both must be set */
#define GEOM_DEBUG_TRI_TC_EPS_FACTOR 3  /* multiplies the TriDistance tensor-core filter margin */
#define GEOM_DEBUG_CHAMFER_VARIANT 4    /* CUDA-core Chamfer: kernel variant + 1, */
#define GEOM_DEBUG_CHAMFER_STRIP 5      /*   strip length, */
#define GEOM_DEBUG_CHAMFER_TILE_COLS 6  /*   tile width */
#define GEOM_DEBUG_TC_CHUNKS 7          /* Chamfer tensor-core search: CTAs per row block (1..8; candidate split) */
#define GEOM_DEBUG_KNOBS 8
GEOM_API int geom_debug_set(int knob, float value);
/* Test hook: dt (128, pad128(m)) row-major = the approximate squared distances of the first min(n,128) points of xyz1
 * (n,3) against xyz2 (m,3) exactly as the tensor-core filter computes them (bf16 split operands, tcgen05.mma, fp32
 * accumulation in TMEM), *rmax = R (largest |coordinate|), both in DEVICE memory.  tests/test_gpu_chamfer.py asserts the
 * filter's error bound |dt - d| <= 2.5e-4 R^2 / 3 on it.  workspace >= geom_chamfer_tensor_workspace_bytes(1, 128, m). */
GEOM_API int geom_debug_tc_dump(int n, const float *xyz1, int m, const float *xyz2, float *dt, float *rmax,
                       void *workspace, size_t workspace_bytes, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer nearest-neighbour search.
 * Replaces: ChamferDistanceKernelLauncher(b,n,xyz,m,xyz2,result,result_i,result2,result2_i)
 *           chamfer_distance/chamfer_distance.cu:57-73 (bound from chamfer_distance.cpp:15-27).
 * For every point j of xyz1 (B,n,3): idx1[j] = argmin_k ||xyz2[k]-xyz1[j]||^2, dist1[j] = that
 * squared distance; and the same with roles swapped into (dist2, idx2) of shape (B,m).
 * Ties: lowest candidate index (chamfer_distance.cu:39,47).  Arithmetic per pair:
 * fma(dz,dz, fma(dx,dx, dy*dy)) with d* = candidate - query, as nvcc compiles :35-38.
 * dist1 / dist2 may be NULL (the Python caller discards them, chamfer_distance.py:31).
 * NaN / Inf coordinates get exactly what the reference kernel gives them (the first candidate of every 512-tile is
 * taken unconditionally, :39, tiles merge with a strict '>', :47: a NaN distance at a tile start voids that tile and a
 * NaN at candidate 0 sticks): every path detects non-finite inputs on the device and redoes the rows with a
 * transcription of the reference loop (slow, exact).
 * ------------------------------------------------------------------------------------------- */
GEOM_API size_t geom_chamfer_workspace_bytes(int b, int n, int m);
/* workspace for the tensor-core search (packed bf16 operands: 64 B per padded point; >= the plain workspace);
 * 16-byte aligned.  Passing a workspace of this size lets flags = 0 select the tensor-core search. */
GEOM_API size_t geom_chamfer_tensor_workspace_bytes(int b, int n, int m);
GEOM_API int geom_chamfer_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                     float *dist1, int *idx1, float *dist2, int *idx2,
                     void *workspace, size_t workspace_bytes, unsigned flags, geom_stream_t stream);

/* Host-buffer form of geom_chamfer_fwd: every pointer is HOST memory (pinned memory gives full
 * copy/compute overlap; pageable memory works, staged by the driver).  The batch is cut into at most `chunks`
 * slices (<= 8; 0 = default 4; the library sizes them): slice k+1 is copied to the device while slice k is searched and slice
 * k-1's results are copied back, on three internal streams.  Synchronous: returns when all results
 * are in host memory.  h_dist1 / h_dist2 may be NULL.  A geom_host_ctx owns the streams, events and a
 * device arena reused across calls; create one per calling thread, on the device it should use (a call made while
 * another device is current returns GEOM_ERR_INVALID_ARG).
 * This is the end-to-end path bench.py reports as "e2e". */
typedef struct geom_host_ctx geom_host_ctx;
/* The slice plan geom_chamfer_fwd_host uses (pure host logic, no device access): slice k covers batch items
 * lo_of[k]..lo_of[k+1]-1; lo_of must hold 9 ints; returns the number of slices (0 on invalid arguments).  With the
 * tensor-core search the slices are sized in whole waves of its one-CTA-per-SM grid (1, 2, 3, 3 ... waves). */
GEOM_API int geom_chamfer_host_plan(int b, int n, int m, int chunks, unsigned flags, int sm_count, int *lo_of);
GEOM_API int geom_host_ctx_create(geom_host_ctx **out);
GEOM_API int geom_host_ctx_destroy(geom_host_ctx *ctx);
GEOM_API int geom_chamfer_fwd_host(geom_host_ctx *ctx, int b, int n, const float *h_xyz1, int m,
                          const float *h_xyz2, float *h_dist1, int *h_idx1, float *h_dist2, int *h_idx2,
                          int chunks, unsigned flags);

/* ---------------------------------------------------------------------------------------------
 * Chamfer loss on the indices (forward sums and backward scatter).
 * Replaces the torch chain after chamfer_dist() in utils.py:402-417 (batch_point_to_point) and
 * utils.py:451-462 (batch_point_to_surface) -- index offsetting, index_select, (a-b)**2, sum, mean --
 * and its autograd backward (index_add); the legacy C form is nnd_backward,
 * old_GEOMetrics/chamfer_distance/src/my_lib.c:49-111.
 *   sums[0] = sum_{i,j} ||xyz2[i, idx1[i,j]] - xyz1[i,j]||^2      (B*n terms)
 *   sums[1] = sum_{i,k} ||xyz1[i, idx2[i,k]] - xyz2[i,k]||^2      (B*m terms)
 * accumulated in double, stored as fp32, deterministic.  dist1 (B,n) / dist2 (B,m) optionally
 * receive the per-point terms (NULL to skip).  The caller turns sums into means (and all-reduces
 * them across ranks when the batch is sharded).
 * Backward: for L = weights[0]*sums[0] + weights[1]*sums[1] (weights = 2 floats in DEVICE memory,
 * so that an upstream autograd scalar needs no host sync) writes dL/dxyz1 into grad1 (B,n,3) and
 * dL/dxyz2 into grad2 (B,m,3); either may be NULL.  Indices are constants, as in the reference.
 * ------------------------------------------------------------------------------------------- */
GEOM_API size_t geom_chamfer_loss_workspace_bytes(int b, int n, int m);
GEOM_API int geom_chamfer_loss_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                          const int *idx1, const int *idx2, float *dist1, float *dist2, float *sums,
                          void *workspace, size_t workspace_bytes, geom_stream_t stream);
/* The same with the F1 counts of utils.py:424-436 / :489-500 fused in: count1[i] = number of rows j of
 * item i of xyz1 with sqrt(sum((s*xyz2[idx1[j]] - s*xyz1[j])**2)) <= threshold (s = f1_scale; the reference uses s = 0.57,
 * threshold = 1e-2), count2[i] likewise for xyz2's rows; (B,) int32 each, overwritten.  The caller turns them into
 * recall / precision (count / num) and F1 -- on the device, so nothing synchronises until the value is read. */
GEOM_API int geom_chamfer_loss_f1_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                             const int *idx1, const int *idx2, float *dist1, float *dist2, float *sums,
                             float f1_scale, float f1_threshold, int *count1, int *count2,
                             void *workspace, size_t workspace_bytes, geom_stream_t stream);
GEOM_API int geom_chamfer_loss_bwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                          const int *idx1, const int *idx2, const float *weights,
                          float *grad1, float *grad2, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Point-to-triangle search.
 * Replaces: TriDistanceKernelLauncher(b,n,xyz,m,tri1,tri2,tri3,dist,point,index)
 *           tri_distance/tri_distance.cu:213-228 (bound from tri_distance.cpp:16-30).
 * For every point of xyz (B,n,3) over the m triangles (tri1,tri2,tri3 = per-face vertices A,B,C,
 * each (B,m,3)): dist = squared distance to the closest triangle as the reference kernel ranks
 * them, point = region code 0..6 (0 interior, 1/2/3 vertex A/B/C, 4/5/6 edge AB/BC/CA),
 * index = face index.  Arithmetic and quirks (option-6 direction vector, NaN tile poisoning,
 * best=10000 initialisation) follow tri_distance.cu:124-206 as compiled; see DESIGN.md.
 * dist may be NULL.  workspace >= geom_tri_workspace_bytes(b, n), 8-byte aligned.
 * Two interchangeable implementations return identical bits: the CUDA-core brute force (tri.cu) and the
 * tensor-core filtered search (tri_tc.cu: a bf16 tcgen05 contraction over bounding spheres lower-bounds every
 * (point, triangle) pair, only the pairs that can still win are evaluated with the reference arithmetic).  As for
 * Chamfer: GEOM_FLAG_TENSOR / GEOM_FLAG_FP32 force one; with neither the filtered search runs when the workspace
 * holds geom_tri_tensor_workspace_bytes(b, n, m) (16-byte aligned) and b*n*m >= 2^24.
 * ------------------------------------------------------------------------------------------- */
GEOM_API size_t geom_tri_workspace_bytes(int b, int n);
GEOM_API size_t geom_tri_tensor_workspace_bytes(int b, int n, int m);
GEOM_API int geom_tri_fwd(int b, int n, const float *xyz, int m, const float *tri1, const float *tri2,
                 const float *tri3, float *dist, int *point, int *index,
                 void *workspace, size_t workspace_bytes, unsigned flags, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Point-to-surface loss on the search result (forward sum, backward scatter to the vertices).
 * Replaces batch_point_to_surface's gathers (utils.py:467-479) + calc_point_to_line
 * (utils.py:506-550) + classes edge / Plane (utils.py:553-587) and their autograd backward.
 * xyz (B,n,3) ground-truth points; verts (B,V,3) predicted vertices; faces (F,3) int64 shared by the
 * batch (adj_info['faces']); option / index (B,n) int32 from geom_tri_fwd.
 *   sum[0] = sum over all B*n points of |c - p|^2, c = the closest point the option selects
 * (vertex / edge projection / plane projection; the CA edge uses its own direction as the Python
 * does).  dist (B,n) and closest (B,n,3) are optional outputs (NULL to skip).
 * Backward: grad_verts (B,V,3) = d(weight[0] * sum)/dverts; weight = 1 float in DEVICE memory.
 * ------------------------------------------------------------------------------------------- */
GEOM_API size_t geom_tri_loss_workspace_bytes(int b, int n);
GEOM_API int geom_tri_loss_fwd(int b, int n, const float *xyz, int V, const float *verts, int F,
                      const int64_t *faces, const int *option, const int *index, float *dist,
                      float *closest, float *sum, void *workspace, size_t workspace_bytes,
                      geom_stream_t stream);
GEOM_API int geom_tri_loss_bwd(int b, int n, const float *xyz, int V, const float *verts, int F,
                      const int64_t *faces, const int *option, const int *index, const float *weight,
                      float *grad_verts, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Differentiable face sampler.
 * Replaces batch_sample (utils.py:590-633) except its random draws, which stay with the caller's
 * generator (torch.multinomial over the areas, two uniform vectors).
 * geom_face_areas: areas (B,F) = 0.5*|cross(v0-v1, v1-v2)| per face (utils.py:596-601), divided by
 *   the per-mesh sum when normalise != 0 (utils.py:602).
 * geom_sample_fwd: points[b,s] = (1-u) v0 + u(1-v) v1 + u v v2 with (v0,v1,v2) the vertices of face
 *   choices[b*num+s] of mesh b, u = sqrt(r1), v = r2 (utils.py:614-631); bit-identical to the torch
 *   expression for identical (choices, r1, r2).
 * geom_sample_bwd: grad_verts (B,V,3) = scatter-add of the barycentric-weighted grad_points (the
 *   index_add autograd runs for utils.py:624-626); grad_verts is overwritten.
 * ------------------------------------------------------------------------------------------- */
GEOM_API int geom_face_areas(int b, int V, const float *verts, int F, const int64_t *faces, float *areas,
                    int normalise, geom_stream_t stream);
GEOM_API int geom_sample_fwd(int b, int V, const float *verts, int F, const int64_t *faces, int num,
                    const int64_t *choices, const float *r1, const float *r2, float *points,
                    geom_stream_t stream);
GEOM_API int geom_sample_bwd(int b, int V, int F, const int64_t *faces, int num, const int64_t *choices,
                    const float *r1, const float *r2, const float *grad_points, float *grad_verts,
                    geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Vertex feature pooling.
 * Replaces the gather part of batched_pooling (utils.py:331-389): for every vertex and every feature map
 * (NCHW fp32, square dim x dim), the reference's bilinear combination of the 4 texels around
 * (clamp(xs*dim, 0, dim-1), clamp(ys*dim, 0, dim-1)), bit-identical to the torch expression for identical
 * projected coordinates xs, ys (B,V) -- including its zero result when a clamped coordinate is integral
 * (floor == ceil makes both weights 0, utils.py:341-346).  The projection itself (batch_camera_info and
 * utils.py:319-330, a few elementwise ops on (B,V) tensors) stays with the caller.
 * geom_pool_fwd: out[b, v, col0 + c] for c over the concatenated channels of the n_maps (<= 8) maps; `out` is a
 *   (B, V, ld) buffer, so the caller can gather straight into the block input next to the vertex positions
 *   (models.py:239 concatenates them).  maps / channels / dims are HOST arrays of n_maps entries.
 * geom_pool_bwd: grad_maps[k] += scatter-add of the weighted grad_out (caller zero-initialises; entries may be
 *   NULL), grad_xs / grad_ys (B,V) = gradient through the bilinear weights (either may be NULL); floor, ceil and
 *   clamp carry torch's gradients (0, 0, pass-through inside [0, dim-1]).
 * ------------------------------------------------------------------------------------------- */
/* Edge-length regulariser (replaces batch_calc_edge, utils.py:636-651).
 * geom_edge_loss_fwd: *sum (device double, overwritten) = sum over the B*F faces of |p2-p1|^2 + |p3-p1|^2 + |p2-p3|^2;
 *   the reference's value is sum / (3*B*F).
 * geom_edge_loss_bwd: grad_verts (B,V,3, overwritten) = *grad_scalar * scale * d(sum)/d(verts); grad_scalar is a
 *   DEVICE scalar (the upstream gradient), scale the host-side 1/(3*B*F). */
GEOM_API int geom_edge_loss_fwd(int b, int V, const float *verts, int F, const int64_t *faces, double *sum, geom_stream_t stream);
GEOM_API int geom_edge_loss_bwd(int b, int V, const float *verts, int F, const int64_t *faces, const float *grad_scalar,
                       float scale, float *grad_verts, geom_stream_t stream);
/* geom_project_fwd / _bwd: the projection feeding the pooling (batch_camera_info, utils.py:283-312, and
 * utils.py:319-330) in one launch: img_info (B,3) rows are (azimuth deg, elevation deg, distance); xs, ys (B,V) are
 * the image coordinates as fractions of the image size.  Same expressions and order as the reference; equal to the
 * torch expression to fp32 rounding (torch's kernels may contract differently).  _bwd: grad_verts (B,V,3) overwritten
 * (img_info is data and gets no gradient); grad_xs / grad_ys may be NULL. */
GEOM_API int geom_project_fwd(int b, int V, const float *verts, const float *img_info, float *xs, float *ys,
                     geom_stream_t stream);
GEOM_API int geom_project_bwd(int b, int V, const float *verts, const float *img_info, const float *grad_xs,
                     const float *grad_ys, float *grad_verts, geom_stream_t stream);
GEOM_API int geom_pool_fwd(int b, int V, int n_maps, const float *const *maps, const int *channels, const int *dims,
                  const float *xs, const float *ys, float *out, int ld, int col0, geom_stream_t stream);
GEOM_API int geom_pool_bwd(int b, int V, int n_maps, const float *const *maps, float *const *grad_maps,
                  const int *channels, const int *dims, const float *xs, const float *ys, const float *grad_out,
                  int ld, int col0, float *grad_xs, float *grad_ys, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 0N-GCN layer.
 * Replaces the forward of ZERON_GCN (layers.py:34-41), Batch_Image_ZERON_GCNGCN (layers.py:107-116),
 * BatchZERON_GCN (layers.py:143-152) and the linear part of GCNMax / BatchGCNMax (layers.py:61-72,
 * 176-185):   support = X W ;  Y = [ adj . support[:, :ksplit] | support[:, ksplit:] ] + bias
 * with ksplit = out//3 or out//10 chosen by the caller.  The reference multiplies by a dense (V,V)
 * adjacency; here it is CSR, built once from that dense matrix:
 *   geom_csr_count(V, adj, transpose, row_cnt, row_ptr): row_ptr[0..V] (exclusive scan of the per-row
 *     non-zero counts; row_cnt is a V-int scratch).  The caller reads row_ptr[V] = nnz, allocates
 *     col_idx / val, and calls geom_csr_fill.  transpose != 0 builds the CSR of adj^T (backward).
 * geom_zeron_gcn_fwd: X (B,V,in), W (in,out), bias (out) or NULL, Y (B,V,out); flags bit 0 fuses
 *   max(.,0); flags bit 1 (GEOM_GCN_TENSOR) runs the feature GEMM on the tensor cores (tcgen05, three TF32
 *   MMAs per product term = fp32-level accuracy; 16 <= out <= 256, else the CUDA-core GEMM is used);
 *   workspace >= geom_zeron_gcn_workspace_bytes(B,V,ksplit) -- or geom_zeron_gcn_tensor_workspace_bytes
 *   (B,V,in,out,ksplit) with GEOM_GCN_TENSOR -- 16-byte aligned.
 * geom_zeron_gcn_bwd_smooth: dS (B,V,out) = [ adj^T . dY[:, :ksplit] | dY[:, ksplit:] ] from the CSR
 *   of adj^T; the caller finishes with the two plain GEMMs dX = dS W^T, dW = X^T dS.
 * fp32 throughout (fp32 accumulate); differs from the dense reference only by summation order.
 * ------------------------------------------------------------------------------------------- */
GEOM_API int geom_csr_count(int V, const float *adj, int transpose, int *row_cnt, int *row_ptr,
                   geom_stream_t stream);
GEOM_API int geom_csr_fill(int V, const float *adj, int transpose, const int *row_ptr, int *col_idx,
                  float *val, geom_stream_t stream);
/* Y (B,V,C) = A . X (B,V,C) for a CSR matrix A (V,V) shared by the batch: the dense
 * `torch.matmul(adj_orig, positions)` of the Laplacian regulariser (utils.py:654-662). */
GEOM_API int geom_csr_spmm(int b, int V, int C, const int *row_ptr, const int *col_idx, const float *val,
                  const float *X, float *Y, geom_stream_t stream);
#define GEOM_GCN_RELU 1
#define GEOM_GCN_TENSOR 2
GEOM_API size_t geom_zeron_gcn_workspace_bytes(int b, int V, int ksplit);
GEOM_API size_t geom_zeron_gcn_tensor_workspace_bytes(int b, int V, int in_features, int out_features, int ksplit);
GEOM_API int geom_zeron_gcn_fwd(int b, int V, int in_features, int out_features, int ksplit, const float *X,
                       const float *W, const float *bias, const int *row_ptr, const int *col_idx,
                       const float *val, float *Y, int flags, void *workspace, size_t workspace_bytes,
                       geom_stream_t stream);
GEOM_API int geom_zeron_gcn_bwd_smooth(int b, int V, int out_features, int ksplit, const int *t_row_ptr,
                              const int *t_col_idx, const float *t_val, const float *dY, float *dS,
                              geom_stream_t stream);

/* Whole backward of the layer (autograd of layers.py:107-116), every GEMM in this library:
 *   dS = [adj^T . dY[:, :ksplit] | dY[:, ksplit:]],  db = column sums of dY,  dW = X^T dS,  dX = dS W^T.
 * dX / dW / db may be NULL (not needed).  flags: GEOM_GCN_TENSOR runs dX on the tensor cores when 16 <= in <= 256.
 * workspace >= geom_zeron_gcn_bwd_workspace_bytes, 16-byte aligned.  dW is reduced over M in slices whose partial
 * tiles are added in a fixed order (deterministic). */
GEOM_API size_t geom_zeron_gcn_bwd_workspace_bytes(int b, int V, int in_features, int out_features);
GEOM_API int geom_zeron_gcn_bwd(int b, int V, int in_features, int out_features, int ksplit, const float *X, const float *W,
                       const float *dY, const int *t_row_ptr, const int *t_col_idx, const float *t_val, float *dX,
                       float *dW, float *db, int flags, void *workspace, size_t workspace_bytes, geom_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm1d(V) + ReLU + averaged residual of the deformation block.
 * Replaces, per stage of BatchMeshDeformationBlock.forward (models.py:241-294),  x = F.relu(bnK(x))  and, after the
 * second layer of a stage,  features = (features + x) / 2  -- torch's batch_norm / relu / add / div kernels and their
 * backward.  nn.BatchNorm1d(verts) on a (B,V,C) tensor normalises per VERTEX over the B*C values (models.py:222-235).
 * fwd: training != 0 uses batch statistics (biased variance) and updates running_mean / running_var (momentum,
 *   unbiased variance) and *batches_tracked (int64, may be NULL) as torch does; training == 0 uses the running
 *   statistics.  out = relu(bn(x)) or, with res != NULL, (res + relu(bn(x))) / 2; res may be the leading C columns of a
 *   wider (B,V,ld_res) tensor.  save_mean / save_invstd (V) are kept for the backward.
 * bwd: dx (B,V,C); dres = dout / 2 (NULL to skip); dgamma / dbeta (V, NULL to skip).
 * ------------------------------------------------------------------------------------------- */
GEOM_API int geom_bn_relu_res_fwd(int b, int V, int C, const float *x, const float *res, int ld_res, const float *gamma,
                         const float *beta, float *running_mean, float *running_var, long long *batches_tracked,
                         int training, float momentum, float eps, float *save_mean, float *save_invstd, float *out,
                         geom_stream_t stream);
GEOM_API int geom_bn_relu_res_bwd(int b, int V, int C, const float *x, const float *dout, const float *gamma, const float *beta,
                         const float *save_mean, const float *save_invstd, int training, int has_res, float *dx,
                         float *dres, float *dgamma, float *dbeta, geom_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GEOM_B200_H */
