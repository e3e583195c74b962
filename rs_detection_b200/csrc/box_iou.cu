// box_iou.cu -- pairwise rotated IoU matrix + MaxIoU assignment (compile with -fmad=false).
//
// Replaces box_iou_rotated_cuda_kernel (python/jdet/ops/box_iou_rotated.py:413-461, v1 copy
// box_iou_rotated_v1.py:418-466) and the Python body of MaxIoUAssigner.assign_wrt_overlaps
// (python/jdet/models/boxes/assigner.py:111-170).
//
// Kernel plan (FP32-ALU bound; no tensor cores -- nothing here is a contraction):
//   prep  : one thread per box -> RBox (double sin/cos once per box, not once per pair)
//   tiles : 64x64 IoU tile per CTA iteration, 128 threads.  Stage 1 runs the bounding-circle test on
//           all 4096 pairs (coalesced zero stores for the rejects; hits in a register bit mask, compacted
//           once per tile); stage 2 runs the separating-axis test densely on the survivors; stage 3 hands
//           ONE remaining pair to each thread, so the heavy clipper runs on dense warps instead of the
//           reference's 1-in-10 active lanes.
#include "common.cuh"
#include "rotated_iou.cuh"

namespace rsdet {

constexpr int kTile = 64;
constexpr int kIouThreads = 128;

__global__ void prep_rbox_kernel(const float* __restrict__ boxes, int n, int version, int zero_tiny, RBox* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + (size_t)i * 5;
    RBox r = prep_rbox(b, version);
    // box_iou_rotated_v1.py:515-522: boxes with a side < 1e-3 give IoU 0 with everything
    if (zero_tiny && fminf(b[2], b[3]) < 0.001f) r.r = -1e30f;
    out[i] = r;
}

__global__ void __launch_bounds__(kIouThreads)
box_iou_tiles_kernel(const RBox* __restrict__ rows, int n1, const RBox* __restrict__ cols, int n2, float* __restrict__ out) {
    __shared__ RBox s_row[kTile];
    __shared__ RBox s_col[kTile];
    __shared__ unsigned short s_queue[kTile * kTile];
    constexpr int kQ2 = 1024;
    __shared__ unsigned short s_queue2[kQ2];
    __shared__ float2 s_pts[24 * kIouThreads];
    __shared__ int s_wsum[kIouThreads / 32];
    __shared__ int s_count;
    __shared__ int s_count2;

    const int tiles_x = ceil_div(n2, kTile);
    const int tiles_y = ceil_div(n1, kTile);
    const long long total = (long long)tiles_x * tiles_y;
    const int tid = threadIdx.x, lane = tid & 31;

    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int r0 = (int)(t / tiles_x) * kTile, c0 = (int)(t % tiles_x) * kTile;
        const int nr = min(kTile, n1 - r0), nc = min(kTile, n2 - c0);
        __syncthreads();  // previous iteration done with smem
        if (tid < kTile) {
            if (tid < nr) s_row[tid] = rows[r0 + tid];
        } else {
            int c = tid - kTile;
            if (c < nc) {
                s_col[c] = cols[c0 + c];
            }
        }
        __syncthreads();

        // stage 1: bounding circles on all 4096 pairs.  Thread = (column, row parity); hits go to a register
        // bit mask, the rejects get their zeros (a warp covers 32 consecutive columns of one row: coalesced),
        // and the hits are compacted once per tile (a ballot + atomic per 32 pairs costs more than the test).
        {
            const int c = tid & 63, rhalf = tid >> 6;
            unsigned hits = 0u;
            if (c < nc) {
                const float cx = s_col[c].x, cy = s_col[c].y, cr = s_col[c].r;
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const int r = 2 * k + rhalf;
                    if (r < nr) {
                        const float rs = s_row[r].r + cr, dx = s_row[r].x - cx, dy = s_row[r].y - cy;
                        const bool cand = (rs >= 0.f) && !(dx * dx + dy * dy > rs * rs);  // == rbox_may_overlap
                        hits |= (cand ? 1u : 0u) << k;
                        if (!cand) out[(size_t)(r0 + r) * n2 + c0 + c] = 0.f;
                    }
                }
            }
            const int mine = __popc(hits);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) s_wsum[tid >> 5] = incl;
            __syncthreads();
            int pos = incl - mine;
            for (int w = 0; w < (tid >> 5); w++) pos += s_wsum[w];
            if (tid == kIouThreads - 1) s_count = pos + mine;
            while (hits) {
                const int k = __ffs(hits) - 1;
                hits &= hits - 1;
                s_queue[pos++] = (unsigned short)(((2 * k + rhalf) << 6) | c);
            }
        }
        __syncthreads();

        // stage 2 (dense over the circle survivors): separating axes in both box frames -> exact zeros;
        // stage 3: one surviving pair per thread through the exact clipper
        const int cnt = s_count;
        for (int q0 = 0; q0 < cnt; q0 += kQ2) {
            const int qn = min(kQ2, cnt - q0);
            if (tid == 0) s_count2 = 0;
            __syncthreads();
            for (int qi = tid; qi < ((qn + 31) & ~31); qi += kIouThreads) {
                int p = 0;
                bool live = qi < qn;
                if (live) {
                    p = s_queue[q0 + qi];
                    live = rbox_inter_upper_bound(s_row[p >> 6], s_col[p & 63]) > 0.f;
                    if (!live) out[(size_t)(r0 + (p >> 6)) * n2 + c0 + (p & 63)] = 0.f;
                }
                const unsigned m = __ballot_sync(0xffffffffu, live);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_count2, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (live) s_queue2[base + __popc(m & ((1u << lane) - 1))] = (unsigned short)p;
                }
            }
            __syncthreads();
            const int cnt2 = s_count2;
            for (int qi = tid; qi < cnt2; qi += kIouThreads) {
                const int p = s_queue2[qi];
                const int r = p >> 6, c = p & 63;
                out[(size_t)(r0 + r) * n2 + c0 + c] = rotated_iou_pair<kIouThreads>(s_row[r], s_col[c], s_pts + tid);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ assignment
// row pass: per-gt max / argmax over proposals (one warp per gt)
__global__ void assign_row_max_kernel(const float* __restrict__ ov, int G, int n, float* __restrict__ gt_max, int* __restrict__ gt_argmax) {
    int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (g >= G) return;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        float v = ov[(size_t)g * n + j];
        if (v > best) { best = v; bi = j; }
    }
    for (int o = 16; o; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { gt_max[g] = best; gt_argmax[g] = bi == 0x7fffffff ? 0 : bi; }
}

// column pass: assigner.py:124-168
__global__ void assign_col_kernel(const float* __restrict__ ov, int G, int n, float pos_thr, float neg_lo, float neg_hi,
                                  float min_pos_iou, int match_low_quality, int assign_all,
                                  const float* __restrict__ gt_max, const int* __restrict__ gt_argmax,
                                  const int32_t* __restrict__ gt_labels, int32_t labels_fill,
                                  int32_t* __restrict__ gt_inds, float* __restrict__ max_ov, int32_t* __restrict__ labels) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float best = -INFINITY;
    int bi = 0;
    int lowq = 0;
    for (int g = 0; g < G; g++) {
        float v = ov[(size_t)g * n + j];
        if (v > best) { best = v; bi = g; }
        if (match_low_quality) {
            float gm = gt_max[g];
            if (gm >= min_pos_iou && (assign_all ? (v == gm) : (gt_argmax[g] == j))) lowq = g + 1;
        }
    }
    int a = -1;
    if (best >= neg_lo && best < neg_hi) a = 0;
    if (best >= pos_thr) a = bi + 1;
    if (lowq) a = lowq;
    gt_inds[j] = a;
    if (max_ov) max_ov[j] = best;
    if (labels) labels[j] = (a > 0 && gt_labels) ? gt_labels[a - 1] : labels_fill;
}

}  // namespace rsdet

using namespace rsdet;

extern "C" size_t rsdet_box_iou_rotated_workspace_bytes(int n1, int n2) {
    return ws_bytes<RBox>(n1 > 0 ? n1 : 0) + ws_bytes<RBox>(n2 > 0 ? n2 : 0);
}

extern "C" int rsdet_box_iou_rotated(const float* boxes1, int n1, const float* boxes2, int n2, int version,
                                        int zero_tiny, float* ious, void* workspace, size_t workspace_bytes, void* stream) {
    if (n1 < 0 || n2 < 0 || (version != 0 && version != 1)) return RSDET_EINVAL;
    if (n1 == 0 || n2 == 0) return RSDET_OK;
    if (!boxes1 || !boxes2 || !ious) return RSDET_EINVAL;
    if ((long long)n1 * n2 > (1ll << 40)) return RSDET_ELIMIT;
    Workspace ws(workspace, workspace_bytes);
    RBox* r1 = ws.take<RBox>(n1);
    RBox* r2 = ws.take<RBox>(n2);
    if (!ws.ok()) return RSDET_EWORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    prep_rbox_kernel<<<ceil_div(n1, 256), 256, 0, st>>>(boxes1, n1, version, zero_tiny, r1);
    prep_rbox_kernel<<<ceil_div(n2, 256), 256, 0, st>>>(boxes2, n2, version, zero_tiny, r2);
    long long tiles = (long long)ceil_div(n1, kTile) * ceil_div(n2, kTile);
    int grid = (int)(tiles < (long long)kNumSMs * 6 ? tiles : (long long)kNumSMs * 6);
    box_iou_tiles_kernel<<<grid, kIouThreads, 0, st>>>(r1, n1, r2, n2, ious);
    count_launch(3);
    return cuda_status();
}

extern "C" size_t rsdet_assign_workspace_bytes(int num_gts) {
    return ws_bytes<float>(num_gts > 0 ? num_gts : 0) + ws_bytes<int>(num_gts > 0 ? num_gts : 0);
}

extern "C" int rsdet_assign_wrt_overlaps(const float* overlaps, int num_gts, int n, float pos_iou_thr, float neg_lo,
                                            float neg_hi, float min_pos_iou, int match_low_quality, int gt_max_assign_all,
                                            const int32_t* gt_labels, int32_t labels_fill, int32_t* assigned_gt_inds,
                                            float* max_overlaps, int32_t* assigned_labels, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    if (num_gts <= 0 || n <= 0) return RSDET_EINVAL;  // assigner.py:91-92,121-122 raise ValueError
    if (!overlaps || !assigned_gt_inds) return RSDET_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    float* gt_max = nullptr;
    int* gt_arg = nullptr;
    if (match_low_quality) {
        Workspace ws(workspace, workspace_bytes);
        gt_max = ws.take<float>(num_gts);
        gt_arg = ws.take<int>(num_gts);
        if (!ws.ok()) return RSDET_EWORKSPACE;
        assign_row_max_kernel<<<ceil_div(num_gts, 8), 256, 0, st>>>(overlaps, num_gts, n, gt_max, gt_arg);
        count_launch();
    }
    assign_col_kernel<<<ceil_div(n, 128), 128, 0, st>>>(overlaps, num_gts, n, pos_iou_thr, neg_lo, neg_hi, min_pos_iou,
                                                         match_low_quality, gt_max_assign_all, gt_max, gt_arg, gt_labels,
                                                         labels_fill, assigned_gt_inds, max_overlaps, assigned_labels);
    count_launch();
    return cuda_status();
}
