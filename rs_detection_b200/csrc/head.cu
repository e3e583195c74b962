// head.cu -- SURVEY 8(f) rank 1: the test-time tail of the Oriented R-CNN head fused into two launches.
//
// Replaces, per image, OrientedHead.get_bboxes + get_results
// (python/jdet/models/roi_heads/oriented_head.py:498-536, 279-305) with the decode of
// OrientedDeltaXYWHTCoder (python/jdet/models/boxes/coder.py:477-514; regular_theta / regular_obb
// python/jdet/ops/bbox_transforms.py:501-519) and obb2poly (:612-623): softmax -> delta decode -> rescale ->
// score threshold (background = LAST column) -> obb2poly -> compaction, in the reference's row-major
// (roi, class) order.  ~25 small Jittor kernels, a boolean-mask gather and a nonzero() in the reference.
// Compile with -fmad=false (the reference runs each multiply / add as a separate elementwise kernel).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace rsdet {

struct HeadParams {
    float mean[5], stdv[5];
    float max_ratio;       // |log(wh_ratio_clip)|
    float inv_scale[4];    // unused (division kept exact); scale below
    float scale[4];
    int rescale;
    float score_thresh;
    int apply_softmax;
    int agnostic;          // bbox_pred is (K,5) shared by all classes
};

__device__ __forceinline__ float floor_mod(float a, float b) { return a - floorf(a / b) * b; }
// bbox_transforms.py:501-507, mode '180', start = -pi/2
__device__ __forceinline__ float regular_theta(float t) {
    const float start = -1.57079632679489661923f, cycle = 3.14159265358979323846f;
    return floor_mod(t - start, cycle) + start;
}

// one warp per RoI row; lanes over classes (strided)
template <bool EMIT>
__global__ void head_results_kernel(const float* __restrict__ rois, const float* __restrict__ cls, const float* __restrict__ pred, int K,
                                    int C, HeadParams P, int* __restrict__ counts, const int* __restrict__ offsets,
                                    float* __restrict__ dets, long long* __restrict__ labels) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= K) return;
    const float* s = cls + (size_t)row * (C + 1);
    // softmax over C+1 logits (nn.softmax: exp(x - max) / sum)
    float mx = -INFINITY;
    for (int c = lane; c <= C; c += 32) mx = fmaxf(mx, s[c]);
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    if (P.apply_softmax) {
        for (int c = lane; c <= C; c += 32) sum += expf(s[c] - mx);
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    int base = EMIT ? offsets[row] : 0;
    int total = 0;
    const float* r = rois + (size_t)row * 5;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        float score = 0.f;
        bool valid = false;
        if (c < C) {
            score = P.apply_softmax ? expf(s[c] - mx) / sum : s[c];
            valid = score > P.score_thresh;
        }
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (EMIT && valid) {
            const float* d = pred + (size_t)row * (P.agnostic ? 5 : 5 * C) + (P.agnostic ? 0 : 5 * c);
            float dx = d[0] * P.stdv[0] + P.mean[0], dy = d[1] * P.stdv[1] + P.mean[1];
            float dw = d[2] * P.stdv[2] + P.mean[2], dh = d[3] * P.stdv[3] + P.mean[3];
            float dt = d[4] * P.stdv[4] + P.mean[4];
            dw = fminf(fmaxf(dw, -P.max_ratio), P.max_ratio);
            dh = fminf(fmaxf(dh, -P.max_ratio), P.max_ratio);
            const float px = r[0], py = r[1], pw = r[2], ph = r[3], pt = r[4];
            const float cs = cosf(-pt), sn = sinf(-pt);
            float gx = dx * pw * cs - dy * ph * sn + px;
            float gy = dx * pw * sn + dy * ph * cs + py;
            float gw = pw * expf(dw), gh = ph * expf(dh);
            float gt = regular_theta(dt + pt);
            // regular_obb (:509-519)
            const bool wide = gw > gh;
            float w = wide ? gw : gh, h = wide ? gh : gw;
            float th = regular_theta(wide ? gt : gt + 1.57079632679489661923f);
            if (P.rescale) { gx = gx / P.scale[0]; gy = gy / P.scale[1]; w = w / P.scale[2]; h = h / P.scale[3]; }
            // obb2poly (:612-623)
            const float Cos = cosf(th), Sin = sinf(th);
            const float v1x = w / 2 * Cos, v1y = -w / 2 * Sin, v2x = -h / 2 * Sin, v2y = -h / 2 * Cos;
            const int pos = base + total + __popc(m & ((1u << lane) - 1));
            float* o = dets + (size_t)pos * 9;
            o[0] = gx + v1x + v2x; o[1] = gy + v1y + v2y;
            o[2] = gx + v1x - v2x; o[3] = gy + v1y - v2y;
            o[4] = gx - v1x - v2x; o[5] = gy - v1y - v2y;
            o[6] = gx - v1x + v2x; o[7] = gy - v1y + v2y;
            o[8] = score;
            labels[pos] = c;
        }
        total += __popc(m);
    }
    if (!EMIT && lane == 0) counts[row] = total;
}

__global__ void head_total_kernel(const int* __restrict__ counts, const int* __restrict__ offsets, int K, int32_t* __restrict__ out_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out_count = K > 0 ? offsets[K - 1] + counts[K - 1] : 0;
}

}  // namespace rsdet

using namespace rsdet;

extern "C" size_t rsdet_oriented_head_results_workspace_bytes(int k) {
    size_t n = (size_t)(k > 0 ? k : 1);
    return 2 * ws_bytes<int>(n) + align256(64 * 1024 + 16 * n);
}

extern "C" int rsdet_oriented_head_results(const float* rois5, const float* cls_score, const float* bbox_pred, int k, int num_classes,
                                           int reg_class_agnostic, const float* means5_host, const float* stds5_host,
                                           float wh_ratio_clip, const float* scale_factor4_host, float score_thresh,
                                           int apply_softmax, float* out_dets, int64_t* out_labels, int32_t* out_count,
                                           void* workspace, size_t workspace_bytes, void* stream) {
    if (k < 0 || num_classes < 1 || !out_count || !means5_host || !stds5_host || !(wh_ratio_clip > 0.f)) return RSDET_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 0) {
        cudaMemsetAsync(out_count, 0, sizeof(int32_t), st);
        return cuda_status();
    }
    if (!rois5 || !cls_score || !bbox_pred || !out_dets || !out_labels) return RSDET_EINVAL;
    if (workspace_bytes < rsdet_oriented_head_results_workspace_bytes(k)) return RSDET_EWORKSPACE;
    Workspace ws(workspace, workspace_bytes);
    int* counts = ws.take<int>(k);
    int* offsets = ws.take<int>(k);
    size_t cub_bytes = 64 * 1024 + 16 * (size_t)k;
    void* cub_tmp = ws.take<char>(cub_bytes);
    HeadParams P;
    for (int i = 0; i < 5; i++) { P.mean[i] = means5_host[i]; P.stdv[i] = stds5_host[i]; }
    P.max_ratio = (float)fabs(log((double)wh_ratio_clip));
    P.rescale = scale_factor4_host != nullptr;
    for (int i = 0; i < 4; i++) { P.scale[i] = P.rescale ? scale_factor4_host[i] : 1.f; P.inv_scale[i] = 1.f; }
    P.score_thresh = score_thresh;
    P.apply_softmax = apply_softmax;
    P.agnostic = reg_class_agnostic;
    const int grid = ceil_div(k, 8);
    head_results_kernel<false><<<grid, 256, 0, st>>>(rois5, cls_score, bbox_pred, k, num_classes, P, counts, nullptr, nullptr, nullptr);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, counts, offsets, k, st);
    if (need > cub_bytes) return RSDET_EWORKSPACE;
    cub::DeviceScan::ExclusiveSum(cub_tmp, need, counts, offsets, k, st);
    head_results_kernel<true><<<grid, 256, 0, st>>>(rois5, cls_score, bbox_pred, k, num_classes, P, counts, offsets, out_dets,
                                                   (long long*)out_labels);
    head_total_kernel<<<1, 32, 0, st>>>(counts, offsets, k, out_count);
    count_launch(5);
    return cuda_status();
}
