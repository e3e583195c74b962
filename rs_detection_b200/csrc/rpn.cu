// rpn.cu -- SURVEY 8(f) rank 2: the proposal stage of the oriented RPN for one image, on the device end to end.
//
// Replaces OrientedRPNHead._get_bboxes_single (python/jdet/models/roi_heads/oriented_rpn_head.py:136-216):
// per level permute + sigmoid + argsort + gathers, MidpointOffsetCoder.decode (models/boxes/coder.py:383-433,
// with rectpoly2obb / regular_obb, ops/bbox_transforms.py:577-599, 509-519), the size filter, obb2hbb
// (:626-632), the level-offset trick and jt.nms.  ~120 small Jittor kernels + 5 argsorts + host-visible
// boolean gathers in the reference; here: 1 score kernel, 1 radix sort (all levels, key = level | score),
// 1 decode kernel (also reduces the coordinate range), 1 offset kernel, the bitmask NMS engine, 1 gather.
// Compile with -fmad=false: the reference evaluates every multiply / add as its own elementwise kernel.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "nms_engine.cuh"

namespace rsdet {

struct RpnLevels {
    int num_levels, A, cls_per_anchor;
    int H[8], W[8];
    int n[8];         // H*W*A
    int off[9];       // prefix of n
    int take[8];      // min(n, nms_pre) (or n when nms_pre <= 0)
    int sorted[8];    // level goes through the score sort (n > nms_pre > 0)
    int cand_off[9];  // prefix of take
    const float* cls[8];
    const float* reg[8];
    const float* anc[8];
};

__device__ __forceinline__ uint32_t desc_key32(float s) {
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~u;
}
__device__ __forceinline__ int find_level(const int* pref, int L, int i) {
    int l = 0;
    while (l + 1 < L && i >= pref[l + 1]) l++;
    return l;
}

// scores in (h, w, a) order (oriented_rpn_head.py:165-177) + sort keys (level << 32 | descending score)
__global__ void rpn_score_kernel(RpnLevels lv, float* __restrict__ score, unsigned long long* __restrict__ key, int* __restrict__ idx,
                                 int* __restrict__ scalars) {
    const int total = lv.off[lv.num_levels];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        scalars[0] = 0x7f800000;   // min as ordered int (+inf)
        scalars[1] = (int)0x807fffff;  // max: fkey(-inf)
        scalars[2] = 0;            // live candidates
    }
    if (i >= total) return;
    const int l = find_level(lv.off, lv.num_levels, i);
    const int j = i - lv.off[l];
    const int a = j % lv.A, pix = j / lv.A;
    const size_t hw = (size_t)lv.H[l] * lv.W[l];
    float s;
    if (lv.cls_per_anchor == 1) {
        s = 1.f / (1.f + expf(-lv.cls[l][(size_t)a * hw + pix]));
    } else {
        const float x0 = lv.cls[l][(size_t)(2 * a) * hw + pix], x1 = lv.cls[l][(size_t)(2 * a + 1) * hw + pix];
        const float m = fmaxf(x0, x1);
        const float e0 = expf(x0 - m), e1 = expf(x1 - m);
        s = e1 / (e0 + e1);
    }
    score[i] = s;
    key[i] = ((unsigned long long)l << 32) | desc_key32(s);
    idx[i] = j;
}

// order-preserving int image of a float for atomicMin/atomicMax
__device__ __forceinline__ int fkey(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float fkey_inv(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float floor_mod_f(float a, float b) { return a - floorf(a / b) * b; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

struct RpnCoder {
    float mean[6], stdv[6];
    float max_ratio;
    float min_size;
};

__global__ void rpn_decode_kernel(RpnLevels lv, RpnCoder cd, const float* __restrict__ score, const int* __restrict__ idx_sorted,
                                  float* __restrict__ obb, float* __restrict__ hbb, float* __restrict__ cscore,
                                  int32_t* __restrict__ clevel, int* __restrict__ scalars) {
    const int ncand = lv.cand_off[lv.num_levels];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    float lo = INFINITY, hi = -INFINITY;
    int live = 0;
    if (c < ncand) {
        const int l = find_level(lv.cand_off, lv.num_levels, c);
        const int r = c - lv.cand_off[l];
        const int j = lv.sorted[l] ? idx_sorted[lv.off[l] + r] : r;
        const float s = score[lv.off[l] + j];
        const int a = j % lv.A, pix = j / lv.A;
        const size_t hw = (size_t)lv.H[l] * lv.W[l];
        float d[6];
#pragma unroll
        for (int k = 0; k < 6; k++) d[k] = lv.reg[l][(size_t)(a * 6 + k) * hw + pix] * cd.stdv[k] + cd.mean[k];
        const float* an = lv.anc[l] + (size_t)j * 4;
        const float dw = clampf(d[2], -cd.max_ratio, cd.max_ratio), dh = clampf(d[3], -cd.max_ratio, cd.max_ratio);
        const float px = (an[0] + an[2]) * 0.5f, py = (an[1] + an[3]) * 0.5f;
        const float pw = an[2] - an[0], ph = an[3] - an[1];
        const float gw = pw * expf(dw), gh = ph * expf(dh);
        const float gx = px + pw * d[0], gy = py + ph * d[1];
        const float x1 = gx - gw * 0.5f, y1 = gy - gh * 0.5f, x2 = gx + gw * 0.5f, y2 = gy + gh * 0.5f;
        const float da = clampf(d[4], -0.5f, 0.5f), db = clampf(d[5], -0.5f, 0.5f);
        const float ga = gx + da * gw, ga_ = gx - da * gw, gb = gy + db * gh, gb_ = gy - db * gh;
        // polys = [ga,y1, x2,gb, _ga,y2, x1,_gb], stretched from the centre to the longer diagonal (coder.py:421-430)
        float qx[4] = {ga - gx, x2 - gx, ga_ - gx, x1 - gx};
        float qy[4] = {y1 - gy, gb - gy, y2 - gy, gb_ - gy};
        float dl[4], dmax = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            dl[k] = sqrtf(qx[k] * qx[k] + qy[k] * qy[k]);
            dmax = fmaxf(dmax, dl[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float f = dmax / dl[k];
            qx[k] = qx[k] * f + gx;
            qy[k] = qy[k] * f + gy;
        }
        // rectpoly2obb (bbox_transforms.py:577-599)
        const float theta = atan2f(-(qy[1] - qy[0]), qx[1] - qx[0]);
        const float Cos = cosf(theta), Sin = sinf(theta);
        const float mx = (((qx[0] + qx[1]) + qx[2]) + qx[3]) / 4.f, my = (((qy[0] + qy[1]) + qy[2]) + qy[3]) / 4.f;
        float rx0 = INFINITY, rx1 = -INFINITY, ry0 = INFINITY, ry1 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float ux = qx[k] - mx, uy = qy[k] - my;
            const float rx = ux * Cos + uy * (-Sin), ry = ux * Sin + uy * Cos;
            rx0 = fminf(rx0, rx); rx1 = fmaxf(rx1, rx);
            ry0 = fminf(ry0, ry); ry1 = fmaxf(ry1, ry);
        }
        const float w = rx1 - rx0, h = ry1 - ry0;
        // regular_obb (:509-519): arithmetic select, then regular_theta
        const float g = w > h ? 1.f : 0.f;
        const float wr = w * g + h * (1.f - g), hr = h * g + w * (1.f - g);
        float tr = theta * g + (theta + 1.57079632679489661923f) * (1.f - g);
        tr = floor_mod_f(tr - (-1.57079632679489661923f), 3.14159265358979323846f) + (-1.57079632679489661923f);
        const bool ok = cd.min_size < 0.f || (wr > cd.min_size && hr > cd.min_size);
        float* o = obb + (size_t)c * 5;
        o[0] = mx; o[1] = my; o[2] = wr; o[3] = hr; o[4] = tr;
        // obb2hbb (:626-632)
        const float C2 = cosf(tr), S2 = sinf(tr);
        const float xb = fabsf(wr / 2 * C2) + fabsf(hr / 2 * S2), yb = fabsf(wr / 2 * S2) + fabsf(hr / 2 * C2);
        float* hb = hbb + (size_t)c * 4;
        hb[0] = mx - xb; hb[1] = my - yb; hb[2] = mx + xb; hb[3] = my + yb;
        cscore[c] = ok ? s : -INFINITY;
        clevel[c] = ok ? l : 0x7fffffff;
        if (ok) {
            lo = fminf(fminf(hb[0], hb[1]), fminf(hb[2], hb[3]));
            hi = fmaxf(fmaxf(hb[0], hb[1]), fmaxf(hb[2], hb[3]));
            live = 1;
        }
    }
    // hproposals.max() - hproposals.min() over the surviving rows (:209)
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        live += __shfl_xor_sync(0xffffffffu, live, o);
    }
    if ((threadIdx.x & 31) == 0 && live) {
        atomicMin(&scalars[0], fkey(lo));
        atomicMax(&scalars[1], fkey(hi));
        atomicAdd(&scalars[2], live);
    }
}

// offsets = level * (max_coordinate + 1); hproposals += offsets (:210-211)
__global__ void rpn_offset_kernel(float* __restrict__ hbb, const int32_t* __restrict__ clevel, int ncand, const int* __restrict__ scalars) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand || clevel[c] == 0x7fffffff) return;
    const float range = fkey_inv(scalars[1]) - fkey_inv(scalars[0]);
    const float off = (float)clevel[c] * (range + 1.f);
    float* hb = hbb + (size_t)c * 4;
    hb[0] += off; hb[1] += off; hb[2] += off; hb[3] += off;
}

__global__ void rpn_output_kernel(const int64_t* __restrict__ keep, const int32_t* __restrict__ num_keep, int nms_post,
                                  const float* __restrict__ obb, const float* __restrict__ cscore, float* __restrict__ dets,
                                  int32_t* __restrict__ num_dets) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = min(*num_keep, nms_post);
    if (r == 0) *num_dets = m;
    if (r >= nms_post) return;
    float* o = dets + (size_t)r * 6;
    if (r < m) {
        const int e = (int)keep[r];
#pragma unroll
        for (int k = 0; k < 5; k++) o[k] = obb[(size_t)e * 5 + k];
        o[5] = cscore[e];
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) o[k] = 0.f;
    }
}

static bool rpn_levels(const rsdet_rpn_cfg* cfg, RpnLevels& lv) {
    if (!cfg || cfg->num_levels < 1 || cfg->num_levels > 8 || cfg->num_anchors < 1 || cfg->nms_post < 1) return false;
    lv.num_levels = cfg->num_levels;
    lv.A = cfg->num_anchors;
    lv.cls_per_anchor = cfg->use_sigmoid ? 1 : 2;
    lv.off[0] = 0;
    lv.cand_off[0] = 0;
    for (int l = 0; l < cfg->num_levels; l++) {
        if (cfg->height[l] < 1 || cfg->width[l] < 1) return false;
        long long n = (long long)cfg->height[l] * cfg->width[l] * cfg->num_anchors;
        if (n > (1 << 24)) return false;
        lv.H[l] = cfg->height[l];
        lv.W[l] = cfg->width[l];
        lv.n[l] = (int)n;
        lv.sorted[l] = cfg->nms_pre > 0 && n > cfg->nms_pre;
        lv.take[l] = lv.sorted[l] ? cfg->nms_pre : (int)n;
        lv.off[l + 1] = lv.off[l] + lv.n[l];
        lv.cand_off[l + 1] = lv.cand_off[l] + lv.take[l];
    }
    return lv.cand_off[cfg->num_levels] <= (1 << 18);
}

constexpr size_t kRpnCubBytes = 16u << 20;

}  // namespace rsdet

using namespace rsdet;

extern "C" int rsdet_rpn_num_candidates(const rsdet_rpn_cfg* cfg) {
    RpnLevels lv;
    return rpn_levels(cfg, lv) ? lv.cand_off[lv.num_levels] : -1;
}

extern "C" size_t rsdet_rpn_proposals_workspace_bytes(const rsdet_rpn_cfg* cfg) {
    RpnLevels lv;
    if (!rpn_levels(cfg, lv)) return 0;
    const size_t total = (size_t)lv.off[lv.num_levels], nc = (size_t)lv.cand_off[lv.num_levels];
    size_t b = 0;
    b += ws_bytes<float>(total) + 2 * ws_bytes<unsigned long long>(total) + 2 * ws_bytes<int>(total);
    b += ws_bytes<int>(64) + align256(kRpnCubBytes);
    b += ws_bytes<float>(nc * 5) + ws_bytes<float>(nc * 4) + ws_bytes<float>(nc) + ws_bytes<int32_t>(nc);
    b += ws_bytes<int64_t>(nc) + ws_bytes<int32_t>(64);
    b += nms_ws_bytes(RSDET_NMS_HBB_P1, (int)nc);
    return b;
}

extern "C" int rsdet_rpn_proposals(const rsdet_rpn_cfg* cfg, const float* const* cls_scores, const float* const* bbox_preds,
                                   const float* const* anchors, float* dets, int32_t* num_dets, float* cand_obb, float* cand_hbb,
                                   float* cand_score, int32_t* cand_level, void* workspace, size_t workspace_bytes, void* stream) {
    RpnLevels lv;
    if (!rpn_levels(cfg, lv) || !cls_scores || !bbox_preds || !anchors || !dets || !num_dets) return RSDET_EINVAL;
    if (workspace_bytes < rsdet_rpn_proposals_workspace_bytes(cfg) || !workspace) return RSDET_EWORKSPACE;
    for (int l = 0; l < lv.num_levels; l++) {
        if (!cls_scores[l] || !bbox_preds[l] || !anchors[l]) return RSDET_EINVAL;
        lv.cls[l] = cls_scores[l];
        lv.reg[l] = bbox_preds[l];
        lv.anc[l] = anchors[l];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int total = lv.off[lv.num_levels], nc = lv.cand_off[lv.num_levels];
    Workspace ws(workspace, workspace_bytes);
    float* score = ws.take<float>(total);
    unsigned long long* keyA = ws.take<unsigned long long>(total);
    unsigned long long* keyB = ws.take<unsigned long long>(total);
    int* idxA = ws.take<int>(total);
    int* idxB = ws.take<int>(total);
    int* scalars = ws.take<int>(64);
    void* cub_tmp = ws.take<char>(kRpnCubBytes);
    float* obb = ws.take<float>((size_t)nc * 5);
    float* hbb = ws.take<float>((size_t)nc * 4);
    float* cscore = ws.take<float>(nc);
    int32_t* clevel = ws.take<int32_t>(nc);
    int64_t* keep = ws.take<int64_t>(nc);
    int32_t* num_keep = ws.take<int32_t>(64);
    const size_t nms_bytes = nms_ws_bytes(RSDET_NMS_HBB_P1, nc);
    void* nms_ws = ws.take<char>(nms_bytes);

    rpn_score_kernel<<<ceil_div(total, 256), 256, 0, st>>>(lv, score, keyA, idxA, scalars);
    count_launch();
    bool any_sorted = false;
    for (int l = 0; l < lv.num_levels; l++) any_sorted |= lv.sorted[l] != 0;
    const int* idx_sorted = idxA;
    if (any_sorted) {
        int level_bits = 1;
        while ((1 << level_bits) < lv.num_levels) level_bits++;
        cub::DoubleBuffer<unsigned long long> k(keyA, keyB);
        cub::DoubleBuffer<int> v(idxA, idxB);
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, k, v, total, 0, 32 + level_bits, st);
        if (need > kRpnCubBytes) return RSDET_EWORKSPACE;
        cub::DeviceRadixSort::SortPairs(cub_tmp, need, k, v, total, 0, 32 + level_bits, st);
        idx_sorted = v.Current();
        count_launch(6);
    }
    RpnCoder cd;
    for (int k = 0; k < 6; k++) { cd.mean[k] = cfg->means[k]; cd.stdv[k] = cfg->stds[k]; }
    cd.max_ratio = (float)fabs(log((double)cfg->wh_ratio_clip));
    cd.min_size = cfg->min_bbox_size;
    rpn_decode_kernel<<<ceil_div(nc, 128), 128, 0, st>>>(lv, cd, score, idx_sorted, obb, hbb, cscore, clevel, scalars);
    rpn_offset_kernel<<<ceil_div(nc, 256), 256, 0, st>>>(hbb, clevel, nc, scalars);
    count_launch(2);
    NmsArgs a{};
    a.kind = RSDET_NMS_HBB_P1;
    a.dets = hbb;
    a.scores = cscore;
    a.labels = clevel;
    a.n_max = nc;
    a.n_dev = scalars + 2;
    a.thr = cfg->nms_thresh;
    a.keep_score_idx = keep;
    a.num_keep = num_keep;
    int rc = nms_run(a, nms_ws, nms_bytes, st);
    if (rc != RSDET_OK) return rc;
    rpn_output_kernel<<<ceil_div(cfg->nms_post, 256), 256, 0, st>>>(keep, num_keep, cfg->nms_post, obb, cscore, dets, num_dets);
    count_launch();
    if (cand_obb) cudaMemcpyAsync(cand_obb, obb, sizeof(float) * 5 * nc, cudaMemcpyDeviceToDevice, st);
    if (cand_hbb) cudaMemcpyAsync(cand_hbb, hbb, sizeof(float) * 4 * nc, cudaMemcpyDeviceToDevice, st);
    if (cand_score) cudaMemcpyAsync(cand_score, cscore, sizeof(float) * nc, cudaMemcpyDeviceToDevice, st);
    if (cand_level) cudaMemcpyAsync(cand_level, clevel, sizeof(int32_t) * nc, cudaMemcpyDeviceToDevice, st);
    return cuda_status();
}
