/*
 * rsdet_oracle.c -- CPU restatement of the reference's rotated-box arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA path in
 * rs_detection_b200/csrc; it is never linked into, imported by, or called from the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load the library built from it.
 *
 * Parity status: PINNED for rotated IoU, nms_rotated, RoIAlignRotated (v0/v1) and the
 * float polygon IoU -- every function below is checked bit-for-bit against the
 * reference's own source compiled for the host (oracle/build_ref.py -> oracle/_ref/) in
 * tests/test_oracle_vs_ref.py, and against the committed fixtures in tests/golden/.
 * UNPINNED for orc_merge_nms (the reference delegates the polygon intersection to
 * Shapely 1.8.2 / GEOS, which is neither vendored in /root/reference nor installed
 * here); that function restates the published algorithm (convex polygon clipping in
 * float64) and says so again at its definition.
 *
 * All arithmetic is written so that, compiled with `gcc -O2 -ffp-contract=off`, each
 * float operation happens in the same order and at the same precision as in the
 * reference sources cited at each function (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y; } pt;

/* FLOP census of the reference's rotated-IoU algorithm ("pin the per-pair figure by instrumenting the CPU
 * restatement"): built with -DORC_COUNT_FLOPS (oracle.build(count_flops=True) -> a second library), every float add,
 * subtract, multiply, divide and each sin / cos counts 1; comparisons, fabs and moves count 0.  Read with
 * orc_flops_read(), reset with orc_flops_reset().  The default library compiles the macro away. */
#ifdef ORC_COUNT_FLOPS
static unsigned long long g_orc_flops = 0, g_orc_pairs = 0;
#define FL(n) (g_orc_flops += (unsigned long long)(n))
unsigned long long orc_flops_read(void) { return g_orc_flops; }
unsigned long long orc_pairs_read(void) { return g_orc_pairs; }
void orc_flops_reset(void) { g_orc_flops = 0; g_orc_pairs = 0; }
#else
#define FL(n) ((void)0)
#endif
static inline float cross2(pt a, pt b) { return a.x * b.y - b.x * a.y; } /* box_iou_rotated.py:47-50 */
static inline float dot2(pt a, pt b) { return a.x * b.x + a.y * b.y; }   /* box_iou_rotated.py:42-45 */
static inline pt sub2(pt a, pt b) { pt r = { a.x - b.x, a.y - b.y }; return r; }

/* python/jdet/ops/box_iou_rotated.py:52-72 (version 0, also nms_rotated.py:52-72) and
 * python/jdet/ops/box_iou_rotated_v1.py:52-77 (version 1: clockwise-positive angle). */
static void rotated_vertices(float xc, float yc, float w, float h, float a, int version, pt p[4])
{
    double theta = a;
    FL(2 + 2 + 16 + 8);  /* sin, cos; two halvings; vertices 0/1: 8 mul + 8 add; vertices 2/3: 4 mul + 4 sub */
    float c2 = (float)cos(theta) * 0.5f;
    float s2 = (float)sin(theta) * 0.5f;
    if (version == 0) {
        p[0].x = xc - s2 * h - c2 * w;
        p[0].y = yc + c2 * h - s2 * w;
        p[1].x = xc + s2 * h - c2 * w;
        p[1].y = yc - c2 * h - s2 * w;
    } else {
        p[0].x = xc + s2 * h + c2 * w;
        p[0].y = yc + c2 * h - s2 * w;
        p[1].x = xc - s2 * h + c2 * w;
        p[1].y = yc - c2 * h - s2 * w;
    }
    p[2].x = 2 * xc - p[0].x;
    p[2].y = 2 * yc - p[0].y;
    p[3].x = 2 * xc - p[1].x;
    p[3].y = 2 * yc - p[1].y;
}

/* python/jdet/ops/box_iou_rotated.py:74-153 */
static int intersection_points(const pt p1[4], const pt p2[4], pt out[24])
{
    pt v1[4], v2[4];
    int num = 0;
    FL(16);
    for (int i = 0; i < 4; i++) {
        v1[i] = sub2(p1[(i + 1) % 4], p1[i]);
        v2[i] = sub2(p2[(i + 1) % 4], p2[i]);
    }
    for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
            FL(3);
            float det = cross2(v2[j], v1[i]);
            if (fabs((double)det) <= 1e-14) continue;
            FL(2 + 4 + 4);
            pt v12 = sub2(p2[j], p1[i]);
            float t1 = cross2(v2[j], v12) / det;
            float t2 = cross2(v1[i], v12) / det;
            if (t1 >= 0.0f && t1 <= 1.0f && t2 >= 0.0f && t2 <= 1.0f) {
                FL(4);
                out[num].x = p1[i].x + v1[i].x * t1;
                out[num].y = p1[i].y + v1[i].y * t1;
                num++;
            }
        }
    }
    FL(2 * (6 + 4 * 9));   /* both corner-in-rectangle loops: two squared lengths, per corner 2 subs + 2 dots + 1 negation */
    {   /* corners of box 1 inside box 2 */
        pt AB = v2[0], DA = v2[3];
        float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
        for (int i = 0; i < 4; i++) {
            pt AP = sub2(p1[i], p2[0]);
            float APdotAB = dot2(AP, AB);
            float APdotAD = -dot2(AP, DA);
            if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD)
                out[num++] = p1[i];
        }
    }
    {   /* corners of box 2 inside box 1 */
        pt AB = v1[0], DA = v1[3];
        float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
        for (int i = 0; i < 4; i++) {
            pt AP = sub2(p2[i], p1[0]);
            float APdotAB = dot2(AP, AB);
            float APdotAD = -dot2(AP, DA);
            if (APdotAB >= 0 && APdotAD >= 0 && APdotAB <= ABdotAB && APdotAD <= ADdotAD)
                out[num++] = p2[i];
        }
    }
    return num;
}

/* comparator of the CPU flavour, box_iou_rotated.py:316-325 */
static int hull_less(pt A, pt B)
{
    FL(3);
    float t = cross2(A, B);
    if (fabs((double)t) < 1e-6) { FL(6); return dot2(A, A) < dot2(B, B); }
    return t > 0;
}

/* python/jdet/ops/box_iou_rotated.py:155-238.
 * sort_kind 0: CPU flavour (std::sort with hull_less; for the <=16 candidates two
 *              rectangles can produce libstdc++'s std::sort is a plain insertion sort,
 *              which is what is restated here).  NOTE the reference CPU flavour does not
 *              permute dist[] with q[], so Step 4 reads the PRE-sort distances; kept.
 * sort_kind 1: CUDA flavour (exchange sort with the 1e-6 tie rule, :338-351), dist[]
 *              travels with q[]. */
static int convex_hull(const pt p[24], int n, pt q[24], int sort_kind)
{
    int t = 0;
    float dist[24];
    for (int i = 1; i < n; i++)
        if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
    pt start = p[t];
    FL(2 * n + 3 * n);
    for (int i = 0; i < n; i++) q[i] = sub2(p[i], start);
    { pt tmp = q[0]; q[0] = q[t]; q[t] = tmp; }
    for (int i = 0; i < n; i++) dist[i] = dot2(q[i], q[i]);

    if (sort_kind == 0) {
        for (int i = 2; i < n; i++) {
            pt v = q[i];
            int j = i;
            while (j > 1 && hull_less(v, q[j - 1])) { q[j] = q[j - 1]; j--; }
            q[j] = v;
        }
    } else {
        for (int i = 1; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                FL(3);
                float cp = cross2(q[i], q[j]);
                if (cp < -1e-6 || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
                    pt tq = q[i]; q[i] = q[j]; q[j] = tq;
                    float td = dist[i]; dist[i] = dist[j]; dist[j] = td;
                }
            }
    }
    int k;
    for (k = 1; k < n; k++)
        if (dist[k] > 1e-8) break;
    if (k == n) { q[0] = p[t]; return 1; }
    q[1] = q[k];
    int m = 2;
    for (int i = k + 1; i < n; i++) {
        FL(7);
        while (m > 1 && cross2(sub2(q[i], q[m - 2]), sub2(q[m - 1], q[m - 2])) >= 0) { m--; FL(7); }
        q[m++] = q[i];
    }
    return m; /* shift_to_zero == true at the only call site (:276) */
}

/* python/jdet/ops/box_iou_rotated.py:240-252 */
static float polygon_area(const pt q[24], int m)
{
    if (m <= 2) return 0;
    float area = 0;
    FL(8 * (m - 2) + 1);
    for (int i = 1; i < m - 1; i++)
        area += (float)fabs((double)cross2(sub2(q[i], q[0]), sub2(q[i + 1], q[0])));
    return (float)(area / 2.0);
}

/* python/jdet/ops/box_iou_rotated.py:279-310 (single_box_iou_rotated) with
 * rotated_boxes_intersection :254-277; label gate of nms_rotated.py:281-286 when
 * box_len == 6. */
float orc_rotated_iou_pair(const float* b1, const float* b2, int box_len, int version, int sort_kind)
{
    if (box_len == 6 && b1[5] != b2[5]) return 0.0f;
#ifdef ORC_COUNT_FLOPS
    g_orc_pairs++;
#endif
    FL(4 + 4 + 2 + 3);   /* midpoint, shifted centres, two areas, final iou */
    double sx = (b1[0] + b2[0]) / 2.0;
    double sy = (b1[1] + b2[1]) / 2.0;
    float x1 = (float)(b1[0] - sx), y1 = (float)(b1[1] - sy);
    float x2 = (float)(b2[0] - sx), y2 = (float)(b2[1] - sy);
    float area1 = b1[2] * b1[3];
    float area2 = b2[2] * b2[3];
    if (area1 < 1e-14 || area2 < 1e-14) return 0.f;

    pt p1[4], p2[4], inter[24], hull[24];
    rotated_vertices(x1, y1, b1[2], b1[3], b1[4], version, p1);
    rotated_vertices(x2, y2, b2[2], b2[3], b2[4], version, p2);
    int num = intersection_points(p1, p2, inter);
    float ia;
    if (num <= 2) ia = 0.0f;
    else {
        int m = convex_hull(inter, num, hull, sort_kind);
        ia = polygon_area(hull, m);
    }
    return ia / (area1 + area2 - ia);
}

/* python/jdet/ops/box_iou_rotated.py:487-500 (IOU_CPU_SRC loop) */
void orc_box_iou_rotated(const float* b1, int n, const float* b2, int m, int version, int sort_kind, float* out)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++)
            out[(size_t)i * m + j] = orc_rotated_iou_pair(b1 + i * 5, b2 + j * 5, 5, version, sort_kind);
}

/* python/jdet/ops/nms_rotated.py:414-449 (greedy CPU body; `>=`) and the CUDA path's
 * decision rule (:403 `>`), selected by `ge`.  keep[] is indexed by ORIGINAL position
 * (keep[order[i]]), as both reference bodies do. */
void orc_nms_rotated(const float* dets, const int* order, int n, int box_len, float thr, int ge,
                     int sort_kind, uint8_t* keep)
{
    uint8_t* sup = (uint8_t*)calloc(n > 0 ? n : 1, 1);
    memset(keep, 0, n);
    for (int _i = 0; _i < n; _i++) {
        int i = order[_i];
        if (sup[i]) continue;
        keep[i] = 1;
        for (int _j = _i + 1; _j < n; _j++) {
            int j = order[_j];
            if (sup[j]) continue;
            float ovr = orc_rotated_iou_pair(dets + (size_t)i * box_len, dets + (size_t)j * box_len,
                                             box_len, 0, sort_kind);
            if (ge ? (ovr >= thr) : (ovr > thr)) sup[j] = 1;
        }
    }
    free(sup);
}

/* ------------------------------------------------------------------------------------
 * RoIAlignRotated.  python/jdet/ops/roi_align_rotated_v1.py:23-68 (bilinear),
 * :71-147 (forward), :149-190 (weights), :193-298 (backward); v0 differences
 * python/jdet/ops/roi_align_rotated.py:67-125: no -0.5 centre shift, counter-clockwise
 * rotation.  `aligned_cw` = 1 selects v1, 0 selects v0.
 * feat: (N,C,H,W) fp32 NCHW; rois: (K,6) [batch,cx,cy,w,h,theta]; out: (K,C,ph,pw).
 */
static float bilinear(const float* d, int height, int width, float y, float x)
{
    if (y < -1.0 || y > height || x < -1.0 || x > width) return 0;
    if (y < 0) y = 0;
    if (x < 0) x = 0;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else y_high = y_low + 1;
    if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else x_high = x_low + 1;
    float ly = y - y_low, lx = x - x_low;
    float hy = (float)(1. - ly), hx = (float)(1. - lx);
    float lt = d[y_low * width + x_low], rt = d[y_low * width + x_high];
    float lb = d[y_high * width + x_low], rb = d[y_high * width + x_high];
    float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    return w1 * lt + w2 * rt + w3 * lb + w4 * rb;
}

typedef struct {
    int batch, gh, gw;
    float cw, ch, bin_h, bin_w, start_h, start_w, cosv, sinv, count;
} roi_geom;

static roi_geom roi_setup(const float* r, float scale, int sample_num, int ph, int pw, int aligned_cw)
{
    roi_geom g;
    g.batch = (int)r[0];
    if (aligned_cw) {
        g.cw = r[1] * scale - 0.5f;
        g.ch = r[2] * scale - 0.5f;
    } else {
        g.cw = r[1] * scale;
        g.ch = r[2] * scale;
    }
    float rw = r[3] * scale, rh = r[4] * scale, theta = r[5];
    rw = rw > 1.f ? rw : 1.f;
    rh = rh > 1.f ? rh : 1.f;
    g.bin_h = rh / (float)ph;
    g.bin_w = rw / (float)pw;
    g.gh = sample_num > 0 ? sample_num : (int)ceilf(rh / ph);
    g.gw = sample_num > 0 ? sample_num : (int)ceilf(rw / pw);
    g.start_h = (float)(-rh / 2.0);
    g.start_w = (float)(-rw / 2.0);
    g.cosv = cosf(theta);
    g.sinv = sinf(theta);
    int cnt = g.gh * g.gw;
    g.count = (float)(cnt > 1 ? cnt : 1);
    return g;
}

static inline void roi_sample_xy(const roi_geom* g, int ph, int pw, int iy, int ix, int aligned_cw,
                                 float* x, float* y)
{
    float yy = g->start_h + ph * g->bin_h + (float)(iy + .5f) * g->bin_h / (float)g->gh;
    float xx = g->start_w + pw * g->bin_w + (float)(ix + .5f) * g->bin_w / (float)g->gw;
    if (aligned_cw) {
        *x = xx * g->cosv + yy * g->sinv + g->cw;
        *y = yy * g->cosv - xx * g->sinv + g->ch;
    } else {
        *x = xx * g->cosv - yy * g->sinv + g->cw;
        *y = xx * g->sinv + yy * g->cosv + g->ch;
    }
}

void orc_roi_align_rotated_fwd(const float* feat, const float* rois, int K, int C, int H, int W,
                               float scale, int sample_num, int PH, int PW, int aligned_cw, float* out)
{
    for (int n = 0; n < K; n++) {
        roi_geom g = roi_setup(rois + n * 6, scale, sample_num, PH, PW, aligned_cw);
        for (int c = 0; c < C; c++) {
            const float* d = feat + ((size_t)g.batch * C + c) * H * W;
            for (int ph = 0; ph < PH; ph++)
                for (int pw = 0; pw < PW; pw++) {
                    float acc = 0.f;
                    for (int iy = 0; iy < g.gh; iy++)
                        for (int ix = 0; ix < g.gw; ix++) {
                            float x, y;
                            roi_sample_xy(&g, ph, pw, iy, ix, aligned_cw, &x, &y);
                            acc += bilinear(d, H, W, y, x);
                        }
                    acc /= g.count;
                    out[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc;
                }
        }
    }
}

void orc_roi_align_rotated_bwd(const float* grad, const float* rois, int K, int N, int C, int H, int W,
                               float scale, int sample_num, int PH, int PW, int aligned_cw, float* gin)
{
    memset(gin, 0, sizeof(float) * (size_t)N * C * H * W);
    for (int n = 0; n < K; n++) {
        roi_geom g = roi_setup(rois + n * 6, scale, sample_num, PH, PW, aligned_cw);
        float count = (float)(g.gh * g.gw); /* backward has no max(.,1): roi_align_rotated_v1.py:246 */
        for (int c = 0; c < C; c++) {
            float* d = gin + ((size_t)g.batch * C + c) * H * W;
            for (int ph = 0; ph < PH; ph++)
                for (int pw = 0; pw < PW; pw++) {
                    float top = grad[(((size_t)n * C + c) * PH + ph) * PW + pw];
                    for (int iy = 0; iy < g.gh; iy++)
                        for (int ix = 0; ix < g.gw; ix++) {
                            float x, y;
                            roi_sample_xy(&g, ph, pw, iy, ix, aligned_cw, &x, &y);
                            if (y < -1.0 || y > H || x < -1.0 || x > W) continue;
                            if (y < 0) y = 0;
                            if (x < 0) x = 0;
                            int y_low = (int)y, x_low = (int)x, y_high, x_high;
                            if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
                            if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
                            float ly = y - y_low, lx = x - x_low;
                            float hy = (float)(1. - ly), hx = (float)(1. - lx);
                            float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                            d[y_low * W + x_low] += top * w1 / count;
                            d[y_low * W + x_high] += top * w2 / count;
                            d[y_high * W + x_low] += top * w3 / count;
                            d[y_high * W + x_high] += top * w4 / count;
                        }
                }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Float polygon IoU of poly_nms.  python/jdet/ops/nms_poly.py:17-19 (sig), :41-50
 * (cross/area), :51-60 (lineCross), :61-75 (polygon_cut), :79-96 (triangle
 * intersection), :98-110 (fan sum), :113-133 (devPolyIoU).
 */
typedef struct { float x, y; } f2;
static inline int sigf(float d) { return (d > 1e-8) - (d < -1e-8); }
static inline int f2_eq(f2 a, f2 b) { return sigf(a.x - b.x) == 0 && sigf(a.y - b.y) == 0; }
static inline float cross3(f2 o, f2 a, f2 b) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); }

static float poly_area(f2* ps, int n)
{
    ps[n] = ps[0];
    float res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return (float)(res / 2.0);
}

static int line_cross(f2 a, f2 b, f2 c, f2 d, f2* p)
{
    float s1 = cross3(a, b, c), s2 = cross3(a, b, d);
    if (sigf(s1) == 0 && sigf(s2) == 0) return 2;
    if (sigf(s2 - s1) == 0) return 0;
    p->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

static void polygon_cut(f2* p, int* n_io, f2 a, f2 b, f2* pp)
{
    int n = *n_io, m = 0;
    p[n] = p[0];
    for (int i = 0; i < n; i++) {
        if (sigf(cross3(a, b, p[i])) > 0) pp[m++] = p[i];
        if (sigf(cross3(a, b, p[i])) != sigf(cross3(a, b, p[i + 1]))) line_cross(a, b, p[i], p[i + 1], &pp[m++]);
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !f2_eq(pp[i], pp[i - 1])) p[n++] = pp[i];
    while (n > 1 && f2_eq(p[n - 1], p[0])) n--;
    *n_io = n;
}

static float tri_intersect_area(f2 a, f2 b, f2 c, f2 d)
{
    f2 o = { 0.f, 0.f };
    int s1 = sigf(cross3(o, a, b)), s2 = sigf(cross3(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0f;
    if (s1 == -1) { f2 t = a; a = b; b = t; }
    if (s2 == -1) { f2 t = c; c = d; d = t; }
    f2 p[10], pp[10]; /* pp is read uninitialised by the reference when lineCross declines to
                         write (nms_poly.py:55-56); zero-filled here so the oracle is deterministic */
    memset(p, 0, sizeof p);
    memset(pp, 0, sizeof pp);
    p[0] = o; p[1] = a; p[2] = b;
    int n = 3;
    polygon_cut(p, &n, o, c, pp);
    polygon_cut(p, &n, c, d, pp);
    polygon_cut(p, &n, d, o, pp);
    float res = (float)fabs((double)poly_area(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

static void f2_reverse(f2* first, f2* last)
{
    while (first != last && first != --last) { f2 t = *first; *first = *last; *last = t; ++first; }
}

float orc_poly_iou(const float* p, const float* q)
{
    f2 ps1[10], ps2[10];
    for (int i = 0; i < 4; i++) {
        ps1[i].x = p[i * 2]; ps1[i].y = p[i * 2 + 1];
        ps2[i].x = q[i * 2]; ps2[i].y = q[i * 2 + 1];
    }
    if (poly_area(ps1, 4) < 0) f2_reverse(ps1, ps1 + 4);
    if (poly_area(ps2, 4) < 0) f2_reverse(ps2, ps2 + 4);
    ps1[4] = ps1[0];
    ps2[4] = ps2[0];
    float inter = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_intersect_area(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]);
    float uni = (float)fabs((double)poly_area(ps1, 4)) + (float)fabs((double)poly_area(ps2, 4)) - inter;
    if (uni == 0) return (inter + 1) / (uni + 1);
    return inter / uni;
}

void orc_poly_iou_matrix(const float* p, int n, const float* q, int m, int stride, float* out)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m; j++) out[(size_t)i * m + j] = orc_poly_iou(p + (size_t)i * stride, q + (size_t)j * stride);
}

/* python/jdet/ops/nms_poly.py:187-232: rows must already be in descending-score order
 * (the caller sorts, :191-193); keep[i] refers to the SORTED row i; decision `>` (:175). */
void orc_poly_nms_sorted(const float* polys, int n, float thr, uint8_t* keep)
{
    uint8_t* rem = (uint8_t*)calloc(n > 0 ? n : 1, 1);
    for (int i = 0; i < n; i++) {
        keep[i] = 0;
        if (rem[i]) continue;
        keep[i] = 1;
        for (int j = i + 1; j < n; j++)
            if (!rem[j] && orc_poly_iou(polys + (size_t)i * 9, polys + (size_t)j * 9) > thr) rem[j] = 1;
    }
    free(rem);
}

/* ------------------------------------------------------------------------------------
 * Full-scene merge NMS.  python/jdet/data/devkits/result_merge.py:66-127
 * (py_cpu_nms_poly_fast) with iou_poly of python/jdet/ops/nms_poly.py:247-252.
 *
 * PARITY UNPINNED: the reference computes `Polygon(a).intersection(Polygon(b)).area`
 * with Shapely==1.8.2 (requirements.txt) -> GEOS overlay, absent from /root/reference
 * and from this image.  For the convex, non-degenerate quadrilaterals this path sees
 * (obb2poly outputs) the intersection is the convex polygon obtained by clipping one
 * quad against the four half-planes of the other (Sutherland-Hodgman); the area is the
 * shoelace sum.  Everything is float64 like numpy/GEOS.
 */
typedef struct { double x, y; } d2;

static double quad_signed_area(const d2* p, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) { int j = (i + 1) % n; s += p[i].x * p[j].y - p[j].x * p[i].y; }
    return 0.5 * s;
}

double orc_convex_quad_intersection_area(const double* a8, const double* b8)
{
    d2 A[4], B[4], buf1[16], buf2[16];
    for (int i = 0; i < 4; i++) { A[i].x = a8[2 * i]; A[i].y = a8[2 * i + 1]; B[i].x = b8[2 * i]; B[i].y = b8[2 * i + 1]; }
    if (quad_signed_area(A, 4) < 0) { d2 t = A[1]; A[1] = A[3]; A[3] = t; }
    if (quad_signed_area(B, 4) < 0) { d2 t = B[1]; B[1] = B[3]; B[3] = t; }
    d2* in = buf1; d2* out = buf2;
    int n = 4;
    memcpy(in, A, sizeof A);
    for (int e = 0; e < 4 && n > 0; e++) {
        d2 c0 = B[e], c1 = B[(e + 1) % 4];
        double ex = c1.x - c0.x, ey = c1.y - c0.y;
        int m = 0;
        for (int i = 0; i < n; i++) {
            d2 P = in[i], Q = in[(i + 1) % n];
            double sp = ex * (P.y - c0.y) - ey * (P.x - c0.x);
            double sq = ex * (Q.y - c0.y) - ey * (Q.x - c0.x);
            if (sp >= 0) out[m++] = P;
            if ((sp > 0 && sq < 0) || (sp < 0 && sq > 0)) {
                double t = sp / (sp - sq);
                out[m].x = P.x + t * (Q.x - P.x);
                out[m].y = P.y + t * (Q.y - P.y);
                m++;
            }
        }
        d2* tmp = in; in = out; out = tmp;
        n = m;
    }
    if (n < 3) return 0.0;
    return fabs(quad_signed_area(in, n));
}

/* iou_poly, nms_poly.py:247-252 */
double orc_iou_poly(const double* a8, const double* b8)
{
    d2 A[4], B[4];
    for (int i = 0; i < 4; i++) { A[i].x = a8[2 * i]; A[i].y = a8[2 * i + 1]; B[i].x = b8[2 * i]; B[i].y = b8[2 * i + 1]; }
    double inter = orc_convex_quad_intersection_area(a8, b8);
    double uni = fabs(quad_signed_area(A, 4)) + fabs(quad_signed_area(B, 4)) - inter;
    return inter / (uni > 0.01 ? uni : 0.01);
}

/* py_cpu_nms_poly_fast, result_merge.py:66-127.  dets (n,9) float64 [x1..y4,score];
 * `order` = indices in descending score (the caller supplies numpy's
 * `scores.argsort()[::-1]`); writes kept ORIGINAL indices in score order, returns count. */
int orc_merge_nms(const double* dets, const int* order, int n, double thr, int* keep_out)
{
    double* hb = (double*)malloc(sizeof(double) * 5 * (n > 0 ? n : 1));
    uint8_t* rem = (uint8_t*)calloc(n > 0 ? n : 1, 1);
    for (int i = 0; i < n; i++) {
        const double* d = dets + (size_t)i * 9;
        double x1 = d[0], x2 = d[0], y1 = d[1], y2 = d[1];
        for (int k = 1; k < 4; k++) {
            if (d[2 * k] < x1) x1 = d[2 * k];
            if (d[2 * k] > x2) x2 = d[2 * k];
            if (d[2 * k + 1] < y1) y1 = d[2 * k + 1];
            if (d[2 * k + 1] > y2) y2 = d[2 * k + 1];
        }
        hb[5 * i] = x1; hb[5 * i + 1] = y1; hb[5 * i + 2] = x2; hb[5 * i + 3] = y2;
        hb[5 * i + 4] = (x2 - x1 + 1) * (y2 - y1 + 1);
    }
    int nk = 0;
    for (int _i = 0; _i < n; _i++) {
        int i = order[_i];
        if (rem[i]) continue;
        keep_out[nk++] = i;
        for (int _j = _i + 1; _j < n; _j++) {
            int j = order[_j];
            if (rem[j]) continue;
            double xx1 = hb[5 * i] > hb[5 * j] ? hb[5 * i] : hb[5 * j];
            double yy1 = hb[5 * i + 1] > hb[5 * j + 1] ? hb[5 * i + 1] : hb[5 * j + 1];
            double xx2 = hb[5 * i + 2] < hb[5 * j + 2] ? hb[5 * i + 2] : hb[5 * j + 2];
            double yy2 = hb[5 * i + 3] < hb[5 * j + 3] ? hb[5 * i + 3] : hb[5 * j + 3];
            double w = xx2 - xx1 > 0.0 ? xx2 - xx1 : 0.0;
            double h = yy2 - yy1 > 0.0 ? yy2 - yy1 : 0.0;
            double hi = w * h;
            double ovr = hi / (hb[5 * i + 4] + hb[5 * j + 4] - hi);
            if (ovr > 0) ovr = orc_iou_poly(dets + (size_t)i * 9, dets + (size_t)j * 9);
            if (!(ovr <= thr)) rem[j] = 1;
        }
    }
    free(hb);
    free(rem);
    return nk;
}

/* merge.py:14-27 (`nms`): horizontal greedy NMS on (n,5) float64 [x1,y1,x2,y2,score],
 * keeps `iou < thresh`. */
int orc_hbb_nms(const double* boxes, const int* order, int n, double thr, int* keep_out)
{
    uint8_t* rem = (uint8_t*)calloc(n > 0 ? n : 1, 1);
    int nk = 0;
    for (int _i = 0; _i < n; _i++) {
        int i = order[_i];
        if (rem[i]) continue;
        keep_out[nk++] = i;
        const double* a = boxes + (size_t)i * 5;
        double area_i = (a[2] - a[0]) * (a[3] - a[1]);
        for (int _j = _i + 1; _j < n; _j++) {
            int j = order[_j];
            if (rem[j]) continue;
            const double* b = boxes + (size_t)j * 5;
            double tlx = a[0] > b[0] ? a[0] : b[0], tly = a[1] > b[1] ? a[1] : b[1];
            double brx = a[2] < b[2] ? a[2] : b[2], bry = a[3] < b[3] ? a[3] : b[3];
            double ov = (brx - tlx) * (bry - tly) * ((brx > tlx && bry > tly) ? 1.0 : 0.0);
            double area_j = (b[2] - b[0]) * (b[3] - b[1]);
            double iou = ov / (area_i + area_j - ov);
            if (!(iou < thr)) rem[j] = 1;
        }
    }
    free(rem);
    return nk;
}
