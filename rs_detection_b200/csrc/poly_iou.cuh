// poly_iou.cuh -- quadrilateral IoU device functions (compile with -fmad=false).
//
//  * poly_iou_f32: the value of devPolyIoU (python/jdet/ops/nms_poly.py:79-133) bit for bit.  The reference sums the
//    signed intersections of the 4 x 4 triangle fans (origin, edge of P) x (origin, edge of Q) in ABSOLUTE
//    coordinates; for boxes far from the origin that sum is dominated by cancellation noise which the keep decision
//    depends on, so every rounding of the reference has to be reproduced -- but not its code shape:
//      - everything that depends on ONE box is computed once per box (PolyBox: orientation-normalised vertices, the
//        sign of every fan triangle and its orientation-normalised edge, |area|), the reference redoes it per pair;
//      - a triangle is clipped against the three edges of the other triangle with ONE cross product per vertex and
//        edge (the reference evaluates each of them twice, for p[i] and again as p[i+1], and a third time inside
//        lineCross); the identical value is reused, so the result cannot change;
//      - the clipped polygon lives in shared memory as [slot][thread] (two ping-pong buffers of 10 slots, the
//        reference's p[10] / pp[10] thread-local arrays): data-dependent indexing without local memory;
//      - pairs whose 16 fan terms are provably all exactly zero are never evaluated (poly_pair_is_zero below).
//  * merge_pair_suppress: float64 predicate of the merge stage -- hbb prefilter of
//    python/jdet/data/devkits/result_merge.py:91-100 followed by iou_poly
//    (python/jdet/ops/nms_poly.py:247-252; Shapely's intersection restated as convex clipping).
#pragma once
#include <cuda_runtime.h>

namespace rsdet {

__device__ __forceinline__ int sigf(float d) { return ((double)d > 1e-8) - ((double)d < -1e-8); }   // nms_poly.py:17-19
__device__ __forceinline__ bool pt_eq(float2 a, float2 b) { return sigf(a.x - b.x) == 0 && sigf(a.y - b.y) == 0; }
__device__ __forceinline__ float cross3(float2 o, float2 a, float2 b) {
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

constexpr int kPolySlots = 10;  // vertices a clipped triangle can reach in the reference's buffers (maxn)

struct PolyBox {
    float2 v[4];       // vertices in the order intersectArea works on (reversed when the input winds negatively, :98-101)
    float area_abs;    // fabs(area(ps)) as devPolyIoU takes it AFTER the reversal (:125)
    int esign;         // 2 bits per edge k: sig(cross(o, v[k], v[k+1])) + 1   (:81-82)
    // data of the exact-zero test (poly_pair_is_zero): direction of the centroid, inflated angular half-width
    float ux, uy, sin_a, cos_a, reach, rmax;
    int sep_ok;
};

// shoelace sum in the reference's order (:44-50); the /2.0 is exact
__device__ __forceinline__ float poly_area4(const float2* ps) {
    float res = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 a = ps[i], b = ps[(i + 1) & 3]; res += a.x * b.y - a.y * b.x; }
    return res * 0.5f;
}

__device__ inline PolyBox prep_polybox(const float* __restrict__ r) {
    PolyBox B;
#pragma unroll
    for (int i = 0; i < 4; i++) B.v[i] = make_float2(r[2 * i], r[2 * i + 1]);
    if (poly_area4(B.v) < 0) {   // point_reverse(ps, ps + 4)
        float2 t = B.v[0]; B.v[0] = B.v[3]; B.v[3] = t;
        t = B.v[1]; B.v[1] = B.v[2]; B.v[2] = t;
    }
    B.area_abs = fabsf(poly_area4(B.v));
    const float2 o = make_float2(0.f, 0.f);
    B.esign = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) B.esign |= (sigf(cross3(o, B.v[k], B.v[(k + 1) & 3])) + 1) << (2 * k);
    // ---- separation data (double; conservative).  Every point the reference's clipper can produce for a fan
    // triangle of this box -- vertices, proper intersections, and the extrapolated ones lineCross returns when a
    // cross product falls inside the 1e-8 dead zone (at most one edge length beyond a vertex) -- lies within 3 rho of
    // the centroid c, rho = max |v_k - c|.  Seen from the origin that ball spans the half-angle asin(3 rho / |c|).
    double cx = 0, cy = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { cx += (double)B.v[k].x; cy += (double)B.v[k].y; }
    cx *= 0.25; cy *= 0.25;
    double rho = 0;
    bool edges_ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double dx = (double)B.v[k].x - cx, dy = (double)B.v[k].y - cy;
        rho = fmax(rho, sqrt(dx * dx + dy * dy));
        // the origin must be robustly off every edge LINE (sign of cross(c, d, 0) in the second cut)
        const double ax = B.v[k].x, ay = B.v[k].y, bx = B.v[(k + 1) & 3].x, by = B.v[(k + 1) & 3].y;
        const double cr = fabs(ax * by - bx * ay);
        edges_ok = edges_ok && cr > 1e-3 + 1e-5 * (fabs(ax * by) + fabs(bx * ay));
    }
    const double cn = sqrt(cx * cx + cy * cy);
    const double s = cn > 0 ? 3.0 * rho / cn : 2.0;
    B.sep_ok = edges_ok && s < 0.5 && isfinite(cn) && cn > 0;
    B.ux = B.sep_ok ? (float)(cx / cn) : 0.f;
    B.uy = B.sep_ok ? (float)(cy / cn) : 0.f;
    B.sin_a = (float)fmin(s * 1.0001, 1.0);
    B.cos_a = (float)(sqrt(fmax(0.0, 1.0 - s * s)) * 0.9999);
    B.reach = (float)(cn - 3.0 * rho);       // no point of the construction is closer to the origin than this
    B.rmax = (float)(cn + 3.0 * rho);
    return B;
}

// True when devPolyIoU(P, Q) is exactly 0 / union with union != 0, i.e. every one of the 16 fan terms is EXACTLY zero
// in the reference's float arithmetic, so the pair can never be suppressed (thr >= 0) and need not be evaluated.
// Argument (o = origin; a, b an edge of P; c, d an edge of Q, both made counter-clockwise as the reference does):
//   if the two inflated angular sectors are disjoint by a margin delta and together span less than pi, then
//   (1) the signs of cross(o, c, .) and cross(d, o, .) -- the two cuts by lines through the origin -- are the same
//       robust value for EVERY point within 3 rho of P's centroid, because |cross| >= |d| * reach * sin(delta) while
//       its float evaluation error is below 8 eps |d| rmax (the margin below demands a 4x gap);
//   (2) the origin itself evaluates to exactly 0 against both of those lines (identical products cancel) and every
//       intersection with them is computed as (0 * s2 - x * 0) / (s2 - 0) = +-0: the origin is only ever replaced by a
//       signed zero, never moved;
//   (3) so one of the two origin cuts removes everything except signed-zero points (P clockwise of Q: the first cut;
//       P counter-clockwise of Q: the third cut, whatever the middle cut by line c->d produced inside the ball), the
//       de-duplication leaves at most one vertex, and the shoelace sum of <= 1 vertex is exactly 0;
//   (4) edges_ok keeps cross(c, d, o) out of the dead zone, so the middle cut cannot extrapolate THROUGH the origin.
// With all 16 terms 0, inter = 0 and iou = 0 / (|A| + |B|) = 0 unless both areas vanish (then the reference returns
// (0 + 1) / (0 + 1) = 1): that case is excluded.  tests/test_gpu_nms.py compares the engine (with this filter) with the
// oracle (without it) on adversarial layouts: radial edges, boxes in a fan around the origin, margins near the limit.
__device__ __forceinline__ bool poly_pair_is_zero(const PolyBox& P, const PolyBox& Q) {
    if (!(P.sep_ok && Q.sep_ok) || !(P.area_abs + Q.area_abs > 0.f)) return false;
    const float reach = fminf(P.reach, Q.reach), rmax = fmaxf(P.rmax, Q.rmax);
    if (!(reach > 0.f)) return false;
    // required margin: reach * sin(delta) >= 64 eps rmax, and never below 1e-3 rad
    const float delta = fmaxf(1e-3f, 64.f * 1.1920929e-7f * rmax / reach * 1.01f);
    if (!(delta < 0.1f)) return false;
    const float cos12 = P.cos_a * Q.cos_a - P.sin_a * Q.sin_a;          // cos(a1 + a2), both half-widths < 30 degrees
    const float sin12 = P.sin_a * Q.cos_a + P.cos_a * Q.sin_a;
    const float bound = cos12 - delta * (sin12 + delta);                 // < cos(a1 + a2 + delta)
    const float ct = P.ux * Q.ux + P.uy * Q.uy;                          // cosine of the angle between the centroids
    return bound > 0.f && fabsf(ct) < bound * 0.9999f;                   // a1+a2+delta < angle < pi - (a1+a2+delta)
}

// polygon_cut (:61-75) on the [slot][thread] buffers: P (n vertices) is clipped against the half plane left of a -> b
// into Q, de-duplicated back into P.  STRIDE = threads per block.
template <int STRIDE>
__device__ __forceinline__ int poly_cut(float2* __restrict__ P, float2* __restrict__ Q, int n, float2 a, float2 b) {
    if (n == 0) return 0;
    int m = 0;
    float2 pi = P[0];
    const float c0 = cross3(a, b, pi);
    float ci = c0;
    for (int i = 0; i < n; i++) {
        const bool last = i + 1 == n;
        const float2 pn = P[last ? 0 : (i + 1) * STRIDE];
        const float cn = last ? c0 : cross3(a, b, pn);
        const int si = sigf(ci), sn = sigf(cn);
        if (si > 0) Q[(m++) * STRIDE] = pi;
        if (si != sn) {
            // lineCross(a, b, p[i], p[i+1], pp[m++]) (:51-60) with s1 = ci, s2 = cn; when it declines to write
            // (|s2 - s1| inside the dead zone) the slot keeps its previous content, exactly like the reference
            if (sigf(cn - ci) != 0) Q[m * STRIDE] = make_float2((pi.x * cn - pn.x * ci) / (cn - ci), (pi.y * cn - pn.y * ci) / (cn - ci));
            m++;
        }
        pi = pn; ci = cn;
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !pt_eq(Q[i * STRIDE], Q[(i - 1) * STRIDE])) P[(n++) * STRIDE] = Q[i * STRIDE];
    while (n > 1 && pt_eq(P[(n - 1) * STRIDE], P[0])) n--;
    return n;
}

// intersectArea(a, b, c, d) (:79-96) for edge (a, b) of the first box and (c, d) of the second, signs s1 / s2 known
template <int STRIDE>
__device__ __forceinline__ float poly_fan_term(float2 a, float2 b, int s1, float2 c, float2 d, int s2, float2* __restrict__ P,
                                               float2* __restrict__ Q) {
    const float2 o = make_float2(0.f, 0.f);
    if (s1 == -1) { const float2 t = a; a = b; b = t; }
    if (s2 == -1) { const float2 t = c; c = d; d = t; }
#pragma unroll
    for (int i = 0; i < kPolySlots; i++) { P[i * STRIDE] = o; Q[i * STRIDE] = o; }   // p[10] = {o, a, b}; pp: see poly_cut
    P[1 * STRIDE] = a;
    P[2 * STRIDE] = b;
    int n = 3;
    n = poly_cut<STRIDE>(P, Q, n, o, c);
    n = poly_cut<STRIDE>(P, Q, n, c, d);
    n = poly_cut<STRIDE>(P, Q, n, d, o);
    float res = 0;                                              // area(p, n) (:44-50)
    for (int i = 0; i < n; i++) {
        const float2 u = P[i * STRIDE], w = P[(i + 1 == n ? 0 : i + 1) * STRIDE];
        res += u.x * w.y - u.y * w.x;
    }
    res = fabsf(res * 0.5f);
    return s1 * s2 == -1 ? -res : res;
}

// devPolyIoU (:113-133).  scratch: 2 * kPolySlots float2 slots of pitch STRIDE, already offset to this thread.
template <int STRIDE>
__device__ inline float poly_iou_f32(const PolyBox& A, const PolyBox& B, float2* __restrict__ scratch) {
    float2* P = scratch;
    float2* Q = scratch + kPolySlots * STRIDE;
    float inter = 0;
    for (int i = 0; i < 4; i++) {
        const int s1 = ((A.esign >> (2 * i)) & 3) - 1;
        for (int j = 0; j < 4; j++) {
            const int s2 = ((B.esign >> (2 * j)) & 3) - 1;
            // degenerate fans contribute 0.0 (:83): adding +0 is the identity for every value a partial sum of this
            // kind can take except -0, which it can only reach through a term that was itself -0 ... + 0 = +0 either way
            float t = 0.0f;
            if (s1 != 0 && s2 != 0) t = poly_fan_term<STRIDE>(A.v[i], A.v[(i + 1) & 3], s1, B.v[j], B.v[(j + 1) & 3], s2, P, Q);
            inter += t;
        }
    }
    const float uni = A.area_abs + B.area_abs - inter;
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
