// tri_common.cuh -- point-to-triangle arithmetic shared by tri.cu (CUDA-core brute force) and tri_tc.cu (tensor-core
// filtered search): every value is produced by the rounding sequence of the reference binary
// (tri_distance/tri_distance.cu:19-91,140-191 as compiled).
#pragma once

#include "common.cuh"

namespace geom {
namespace tri {

constexpr int kRecVec = 10;    // float4 per full triangle record
constexpr int kRefTile = 512;  // reference tile (tri_distance.cu:106)

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z)
{
    V3 v;
    v.x = x;
    v.y = y;
    v.z = z;
    return v;
}
__device__ __forceinline__ V3 vsub(V3 a, V3 b)
{
    return v3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z));
}
// S1: a.x*b.x + a.y*b.y + a.z*b.z as compiled: FMUL(y) -> FFMA(x) -> FFMA(z)
__device__ __forceinline__ float dot_s1(V3 a, V3 b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
// S2: cross product as compiled: left product inside the FFMA, right product rounded first
__device__ __forceinline__ V3 cross_s2(V3 X, V3 Y)
{
    return v3(__fmaf_rn(X.y, Y.z, -__fmul_rn(X.z, Y.y)),
              __fmaf_rn(X.z, Y.x, -__fmul_rn(X.x, Y.z)),
              __fmaf_rn(X.x, Y.y, -__fmul_rn(X.y, Y.x)));
}
__device__ __forceinline__ bool in_range(float a) { return (a <= 1.0f) && (a >= 0.0f); }
__device__ __forceinline__ V3 xyz(float4 v) { return v3(v.x, v.y, v.z); }
__device__ __forceinline__ V3 sel3(bool c, V3 a, V3 b) { return v3(c ? a.x : b.x, c ? a.y : b.y, c ? a.z : b.z); }
__device__ __forceinline__ V3 fma3(V3 d, float t, V3 s)
{
    return v3(__fmaf_rn(d.x, t, s.x), __fmaf_rn(d.y, t, s.y), __fmaf_rn(d.z, t, s.z));
}

// if (c) out = d * t + s, as predicated FFMAs (ptxas folds the bool -> predicate round trip).
__device__ __forceinline__ void pfma1(float &out, bool c, float d, float t, float s)
{
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %4, 0;\n\t@q fma.rn.f32 %0, %1, %2, %3;\n\t}"
        : "+f"(out)
        : "f"(d), "f"(t), "f"(s), "r"((int)c));
}
__device__ __forceinline__ void pfma3(V3 &out, bool c, V3 d, float t, V3 s)
{
    pfma1(out.x, c, d.x, t, s.x);
    pfma1(out.y, c, d.y, t, s.y);
    pfma1(out.z, c, d.z, t, s.z);
}

// Operand range in which Markstein's correction is exact: every intermediate stays a normal number.
constexpr float kDivLo = 1.0842021724855044e-19f;  // 2^-63
constexpr float kDivHi = 9.2233720368547758e+18f;  // 2^63
__device__ __forceinline__ bool div_safe_den(float b) { return b >= kDivLo && b <= kDivHi; }

// Triangle record, 10 x float4:
//  0: A.xyz | |AB|^2          1: B.xyz | |BC|^2          2: C.xyz | |CA|^2
//  3: AB.xyz | rn(1/|AB|^2)   4: BC.xyz | rn(1/|BC|^2)   5: CA.xyz | rn(1/|CA|^2)
//  6: cross(N,AB) | fast-division flag   7: cross(N,BC)   8: cross(N,CA)   9: N / |N|
__device__ __forceinline__ void build_record(float4 *rec, V3 A, V3 B, V3 C)
{
    const V3 AB = vsub(B, A), BC = vsub(C, B), CA = vsub(A, C);
    const V3 N = cross_s2(vsub(A, B), CA);  // cross(A-B, A-C), tri_distance.cu:148
    const V3 xab = cross_s2(N, AB), xbc = cross_s2(N, BC), xca = cross_s2(N, CA);
    const float inv = __frcp_rn(__fsqrt_rn(dot_s1(N, N)));  // normalize(), tri_distance.cu:77-81
    const float lab = dot_s1(AB, AB), lbc = dot_s1(BC, BC), lca = dot_s1(CA, CA);
    const bool safe = div_safe_den(lab) && div_safe_den(lbc) && div_safe_den(lca);
    rec[0] = make_float4(A.x, A.y, A.z, lab);
    rec[1] = make_float4(B.x, B.y, B.z, lbc);
    rec[2] = make_float4(C.x, C.y, C.z, lca);
    rec[3] = make_float4(AB.x, AB.y, AB.z, __frcp_rn(lab));
    rec[4] = make_float4(BC.x, BC.y, BC.z, __frcp_rn(lbc));
    rec[5] = make_float4(CA.x, CA.y, CA.z, __frcp_rn(lca));
    rec[6] = make_float4(xab.x, xab.y, xab.z, safe ? 1.0f : 0.0f);
    rec[7] = make_float4(xbc.x, xbc.y, xbc.z, 0.f);
    rec[8] = make_float4(xca.x, xca.y, xca.z, 0.f);
    rec[9] = make_float4(__fmul_rn(N.x, inv), __fmul_rn(N.y, inv), __fmul_rn(N.z, inv), 0.f);
}

// One (point, triangle) evaluation with the reference's cascade, tri_distance.cu:150-191.
// Used by the epilogue (option of the winner) and for the first candidate of a reference tile.
__device__ __forceinline__ float tri_pair_exact(V3 p, const float4 *rec, float &option)
{
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2];
    const V3 A = xyz(r0), B = xyz(r1), C = xyz(r2);
    const V3 AB = xyz(rec[3]), BC = xyz(rec[4]), CA = xyz(rec[5]);
    const V3 pA = vsub(p, A), pB = vsub(p, B), pC = vsub(p, C);
    const float uab = __fdiv_rn(dot_s1(pA, AB), r0.w);
    const float uca = __fdiv_rn(dot_s1(pC, CA), r2.w);
    const float ubc = __fdiv_rn(dot_s1(pB, BC), r1.w);
    V3 p2;
    if (uca > 1.0f && uab < 0.0f) {
        p2 = A;
        option = 1.0f;
    } else if (uab > 1.0f && ubc < 0.0f) {
        p2 = B;
        option = 2.0f;
    } else if (ubc > 1.0f && uca < 0.0f) {
        p2 = C;
        option = 3.0f;
    } else if (in_range(uab) && !(dot_s1(xyz(rec[6]), pA) > 0.0f)) {
        p2 = fma3(AB, uab, A);
        option = 4.0f;
    } else if (in_range(ubc) && !(dot_s1(xyz(rec[7]), pB) > 0.0f)) {
        p2 = fma3(BC, ubc, B);
        option = 5.0f;
    } else if (in_range(uca) && !(dot_s1(xyz(rec[8]), pC) > 0.0f)) {
        p2 = fma3(AB, uca, C);  // sic: AB's direction with the CA parameter, tri_distance.cu:180
        option = 6.0f;
    } else {
        const V3 nn = xyz(rec[9]);
        const float dist = dot_s1(pA, nn);
        p2 = v3(__fmaf_rn(-nn.x, dist, p.x), __fmaf_rn(-nn.y, dist, p.y), __fmaf_rn(-nn.z, dist, p.z));
        option = 0.0f;
    }
    const V3 df = vsub(p, p2);
    return dot_s1(df, df);
}

// a / b with y = rn(1/b): correctly rounded (== __fdiv_rn) for operands in the safe range.
__device__ __forceinline__ float div_markstein(float a, float b, float y)
{
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-b, q, a);
    return __fmaf_rn(r, y, q);
}

struct TriRegs {  // one triangle record in registers (broadcast loads from the warp's slab)
    V3 A, B, C, AB, BC, CA, xab, xbc, xca, nn;
    float lab, lbc, lca, rab, rbc, rca;
    bool safe;
};
__device__ __forceinline__ TriRegs load_record(const float4 *rec)
{
    TriRegs t;
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4], r5 = rec[5], r6 = rec[6];
    t.A = xyz(r0), t.B = xyz(r1), t.C = xyz(r2), t.AB = xyz(r3), t.BC = xyz(r4), t.CA = xyz(r5);
    t.lab = r0.w, t.lbc = r1.w, t.lca = r2.w, t.rab = r3.w, t.rbc = r4.w, t.rca = r5.w;
    t.xab = xyz(r6), t.xbc = xyz(rec[7]), t.xca = xyz(rec[8]), t.nn = xyz(rec[9]);
    t.safe = r6.w != 0.0f;
    return t;
}

// Branch-free evaluation: the same value tri_pair_exact returns, bit for bit.
__device__ __forceinline__ float tri_pair_fast(V3 p, const TriRegs &t)
{
    const V3 pA = vsub(p, t.A), pB = vsub(p, t.B), pC = vsub(p, t.C);
    const float nab = dot_s1(pA, t.AB), nbc = dot_s1(pB, t.BC), nca = dot_s1(pC, t.CA);
    float uab, ubc, uca;
    const float lo = fminf(fminf(fabsf(nab), fabsf(nbc)), fabsf(nca));
    const float hi = fmaxf(fmaxf(fabsf(nab), fabsf(nbc)), fabsf(nca));
    if (t.safe && lo >= kDivLo && hi <= kDivHi) {
        uab = div_markstein(nab, t.lab, t.rab);
        ubc = div_markstein(nbc, t.lbc, t.rbc);
        uca = div_markstein(nca, t.lca, t.rca);
    } else {  // zero / denormal / huge operands, degenerate triangles: the real thing
        uab = __fdiv_rn(nab, t.lab);
        ubc = __fdiv_rn(nbc, t.lbc);
        uca = __fdiv_rn(nca, t.lca);
    }
    const bool c1 = (uca > 1.0f) && (uab < 0.0f);
    const bool c2 = (uab > 1.0f) && (ubc < 0.0f);
    const bool c3 = (ubc > 1.0f) && (uca < 0.0f);
    const bool e4 = in_range(uab) && !(dot_s1(t.xab, pA) > 0.0f);
    const bool e5 = in_range(ubc) && !(dot_s1(t.xbc, pB) > 0.0f);
    const bool e6 = in_range(uca) && !(dot_s1(t.xca, pC) > 0.0f);
    // The cascade (vertex A, B, C, edge AB, BC, CA(sic), plane) as predicated overwrites in reverse
    // priority order: no select chains, no predicate algebra; all arithmetic stays on the FMA pipe.
    // Vertex regions are the edge form with parameter 0: S + D*0 == S (D is finite).
    const float dist = dot_s1(pA, t.nn);
    V3 p2 = v3(__fmaf_rn(-t.nn.x, dist, p.x), __fmaf_rn(-t.nn.y, dist, p.y), __fmaf_rn(-t.nn.z, dist, p.z));
    pfma3(p2, e6, t.AB, uca, t.C);  // the CA region uses AB's direction (tri_distance.cu:180)
    pfma3(p2, e5, t.BC, ubc, t.B);
    pfma3(p2, e4, t.AB, uab, t.A);
    pfma3(p2, c3, t.AB, 0.0f, t.C);
    pfma3(p2, c2, t.BC, 0.0f, t.B);
    pfma3(p2, c1, t.AB, 0.0f, t.A);
    const V3 df = vsub(p, p2);
    return dot_s1(df, df);
}

}  // namespace tri
}  // namespace geom
