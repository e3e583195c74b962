// poly_iou.cuh -- quadrilateral IoU device functions (compile with -fmad=false).
//
//  * poly_iou_f32: float triangle-fan polygon IoU with the exact operation order of devPolyIoU
//    (python/jdet/ops/nms_poly.py:17-133).  NOTE the reference works on absolute coordinates, so its
//    result carries cancellation noise that the keep decision depends on; it is therefore mirrored
//    literally (no centring, no early reject).
//  * merge_pair_suppress: float64 predicate of the merge stage -- hbb prefilter of
//    python/jdet/data/devkits/result_merge.py:91-100 followed by iou_poly
//    (python/jdet/ops/nms_poly.py:247-252; Shapely's intersection restated as convex clipping).
#pragma once
#include <cuda_runtime.h>

namespace rsdet {

__device__ __forceinline__ int sigf(float d) { return ((double)d > 1e-8) - ((double)d < -1e-8); }
__device__ __forceinline__ bool pt_eq(float2 a, float2 b) { return sigf(a.x - b.x) == 0 && sigf(a.y - b.y) == 0; }
__device__ __forceinline__ float cross3(float2 o, float2 a, float2 b) {
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

__device__ inline float poly_area_f32(float2* ps, int n) {  // nms_poly.py:44-50
    ps[n] = ps[0];
    float res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return res * 0.5f;
}

__device__ inline int line_cross(float2 a, float2 b, float2 c, float2 d, float2& p) {  // :51-60
    float s1 = cross3(a, b, c), s2 = cross3(a, b, d);
    if (sigf(s1) == 0 && sigf(s2) == 0) return 2;
    if (sigf(s2 - s1) == 0) return 0;
    p.x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p.y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

__device__ inline void polygon_cut(float2* p, int& n, float2 a, float2 b, float2* pp) {  // :61-75
    int m = 0;
    p[n] = p[0];
    for (int i = 0; i < n; i++) {
        int si = sigf(cross3(a, b, p[i]));
        if (si > 0) pp[m++] = p[i];
        if (si != sigf(cross3(a, b, p[i + 1]))) line_cross(a, b, p[i], p[i + 1], pp[m++]);
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !pt_eq(pp[i], pp[i - 1])) p[n++] = pp[i];
    while (n > 1 && pt_eq(p[n - 1], p[0])) n--;
}

__device__ inline float tri_intersect_area(float2 a, float2 b, float2 c, float2 d) {  // :79-96
    float2 o = make_float2(0.f, 0.f);
    int s1 = sigf(cross3(o, a, b));
    int s2 = sigf(cross3(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0f;
    if (s1 == -1) { float2 t = a; a = b; b = t; }
    if (s2 == -1) { float2 t = c; c = d; d = t; }
    float2 p[10], pp[10];
#pragma unroll
    for (int i = 0; i < 10; i++) { p[i] = o; pp[i] = o; }  // the reference leaves pp uninitialised
    p[1] = a;
    p[2] = b;
    int n = 3;
    polygon_cut(p, n, o, c, pp);
    polygon_cut(p, n, c, d, pp);
    polygon_cut(p, n, d, o, pp);
    float res = fabsf(poly_area_f32(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

// devPolyIoU, nms_poly.py:113-133.  p, q: 8 floats each.
__device__ inline float poly_iou_f32(const float* __restrict__ p, const float* __restrict__ q) {
    float2 ps1[6], ps2[6];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ps1[i] = make_float2(p[2 * i], p[2 * i + 1]);
        ps2[i] = make_float2(q[2 * i], q[2 * i + 1]);
    }
    if (poly_area_f32(ps1, 4) < 0) { float2 t = ps1[0]; ps1[0] = ps1[3]; ps1[3] = t; t = ps1[1]; ps1[1] = ps1[2]; ps1[2] = t; }
    if (poly_area_f32(ps2, 4) < 0) { float2 t = ps2[0]; ps2[0] = ps2[3]; ps2[3] = t; t = ps2[1]; ps2[1] = ps2[2]; ps2[2] = t; }
    ps1[4] = ps1[0];
    ps2[4] = ps2[0];
    float inter = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_intersect_area(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]);
    float uni = fabsf(poly_area_f32(ps1, 4)) + fabsf(poly_area_f32(ps2, 4)) - inter;
    if (uni == 0) return (inter + 1) / (uni + 1);
    return inter / uni;
}

// ---------------------------------------------------------------------------------- float64 merge
struct __align__(16) MBox {
    double p[8];            // quad, scene coordinates
    double x1, y1, x2, y2;  // hbb (result_merge.py:68-71)
};

__device__ inline double signed_area_d(const double2* p, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) {
        int j = (i + 1 == n) ? 0 : i + 1;
        s += p[i].x * p[j].y - p[j].x * p[i].y;
    }
    return 0.5 * s;
}

__device__ inline double convex_quad_intersection_area(const double* a8, const double* b8) {
    double2 A[4], B[4], buf1[12], buf2[12];
#pragma unroll
    for (int i = 0; i < 4; i++) { A[i] = make_double2(a8[2 * i], a8[2 * i + 1]); B[i] = make_double2(b8[2 * i], b8[2 * i + 1]); }
    if (signed_area_d(A, 4) < 0) { double2 t = A[1]; A[1] = A[3]; A[3] = t; }
    if (signed_area_d(B, 4) < 0) { double2 t = B[1]; B[1] = B[3]; B[3] = t; }
    double2* in = buf1;
    double2* out = buf2;
    int n = 4;
#pragma unroll
    for (int i = 0; i < 4; i++) in[i] = A[i];
    for (int e = 0; e < 4 && n > 0; e++) {
        double2 c0 = B[e], c1 = B[(e + 1) & 3];
        double ex = c1.x - c0.x, ey = c1.y - c0.y;
        int m = 0;
        for (int i = 0; i < n; i++) {
            double2 P = in[i], Q = in[(i + 1 == n) ? 0 : i + 1];
            double sp = ex * (P.y - c0.y) - ey * (P.x - c0.x);
            double sq = ex * (Q.y - c0.y) - ey * (Q.x - c0.x);
            if (sp >= 0) out[m++] = P;
            if ((sp > 0 && sq < 0) || (sp < 0 && sq > 0)) {
                double t = sp / (sp - sq);
                out[m++] = make_double2(P.x + t * (Q.x - P.x), P.y + t * (Q.y - P.y));
            }
        }
        double2* tmp = in; in = out; out = tmp;
        n = m;
    }
    if (n < 3) return 0.0;
    return fabs(signed_area_d(in, n));
}

__device__ __forceinline__ bool merge_hbb_overlap(const MBox& a, const MBox& b) {
    double xx1 = fmax(a.x1, b.x1), yy1 = fmax(a.y1, b.y1);
    double xx2 = fmin(a.x2, b.x2), yy2 = fmin(a.y2, b.y2);
    double w = fmax(0.0, xx2 - xx1), h = fmax(0.0, yy2 - yy1);
    return w * h > 0.0;  // hbb_ovr > 0  (the +1 areas keep the denominator positive)
}

__device__ inline double iou_poly_d(const MBox& a, const MBox& b) {
    double2 A[4], B[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { A[i] = make_double2(a.p[2 * i], a.p[2 * i + 1]); B[i] = make_double2(b.p[2 * i], b.p[2 * i + 1]); }
    double inter = convex_quad_intersection_area(a.p, b.p);
    double uni = fabs(signed_area_d(A, 4)) + fabs(signed_area_d(B, 4)) - inter;
    return inter / fmax(uni, 0.01);
}

}  // namespace rsdet
