// roi_align.cu -- multi-level RoIAlignRotated forward / backward for sm_100a.
//
// Replaces ROIAlignRotatedForward / ROIAlignBackward (python/jdet/ops/roi_align_rotated_v1.py:71-147,
// 193-298; v0: roi_align_rotated.py:60-126, 164-254) AND the Python level loop around them
// (python/jdet/models/roi_extractors/oriented_single_level.py:91-114): RoI extension, level mapping,
// four per-level launches, boolean gathers and masked scatter-adds become ONE launch.
//
// The op is a gather (fwd) / scatter (bwd): ~0.6 FLOP per byte, HBM/L2 bound.  Plan:
//   * features are read channels-last (NHWC): one bilinear tap = one contiguous C*4-byte row, read as
//     coalesced 16-byte vectors (a warp covers 512 contiguous bytes).  NCHW inputs (Jittor's layout)
//     are transposed once per call by a tiled transpose into the workspace; callers that keep a
//     channels-last pyramid skip it (cfg.channels_last).
//   * one CTA per RoI.  Phase A: the RoI's sampling grid (pooled_h*pooled_w*grid^2 samples -> 4 tap
//     offsets + 4 weights each) is computed ONCE by the first threads and staged in shared memory; the
//     reference recomputes it, sin/cos included, in every one of its K*C*49 threads.
//   * Phase B: thread = (channel quad, bin group); 16 independent 16-byte loads in flight per bin.
//   * the (C,7,7) output block of a RoI is contiguous in the reference layout; results are staged in
//     shared memory in a [k][bin][quad+1] layout (conflict-free both ways) and written with streaming
//     16-byte stores, so the 50 KB block leaves the SM fully coalesced and does not evict features
//     from L2.
//   * backward: same tap table; gradients are accumulated channels-last with 16-byte vector
//     reductions (red.global.add.v4.f32: one L2 atomic op per 4 channels instead of the reference's 4
//     scalar atomicAdd), then transposed back to NCHW, which also produces the dense zero-filled
//     output the reference gets from cudaMemsetAsync.
//
// Sample coordinates are computed with explicitly un-contracted IEEE operations in the reference's
// order: the validity test `y < -1 || y > H` is discontinuous, so coordinates must not drift.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace rsdet {

constexpr int kRoiThreads = 256;
constexpr int kMaxSamples = 1024;  // fast path: pooled_h*pooled_w*grid_h*grid_w

struct LevelSet {
    const float* feat[RSDET_MAX_LEVELS];  // channels-last maps
    float* grad[RSDET_MAX_LEVELS];
    int H[RSDET_MAX_LEVELS], W[RSDET_MAX_LEVELS];
    float scale[RSDET_MAX_LEVELS];
    int num_levels, batch, C;
    int PH, PW, sampling_ratio, version;
    float extend_w, extend_h, finest_scale;
    int dbg_skip_main;  // profiling aid (RSDET_ROI_DBG_SKIP_MAIN=1, RSDET_TUNING builds only): tap lists are built, then treated as empty
};

// ---------------------------------------------------------------------------------- transposes
// (N, C, HW) <-> (N, HW, C), 64x64 tiles through padded shared memory.
struct TransposeJob {
    const float* src[RSDET_MAX_LEVELS];
    float* dst[RSDET_MAX_LEVELS];
    int HW[RSDET_MAX_LEVELS];
    int tile_begin[RSDET_MAX_LEVELS + 1];  // prefix of tiles over levels
    int num_levels, N, C;
};

// 64(channels) x 64(pixels) tiles, 16-byte global accesses on both sides: a thread reads a float4 along the
// source's contiguous dimension, scatters it into a padded shared tile, and writes a float4 along the
// destination's contiguous dimension.  Falls back to scalar accesses at ragged edges / unaligned bases.
template <bool TO_NHWC>
__device__ __forceinline__ void transpose_tile(const TransposeJob& job, int t, float (*tile)[65]) {
    int l = 0;
    while (l + 1 < job.num_levels && t >= job.tile_begin[l + 1]) l++;
    t -= job.tile_begin[l];
    const int HW = job.HW[l], C = job.C;
    const int tiles_hw = ceil_div(HW, 64), tiles_c = ceil_div(C, 64);
    const int n = t / (tiles_hw * tiles_c);
    const int r = t % (tiles_hw * tiles_c);
    const int hw0 = (r % tiles_hw) * 64, c0 = (r / tiles_hw) * 64;
    const float* src = job.src[l] + (size_t)n * C * HW;
    float* dst = job.dst[l] + (size_t)n * C * HW;
    const int tid = threadIdx.x;
    // source: rows of `SR` index, contiguous along `SC`; destination the other way round
    const int srcRows = TO_NHWC ? C : HW, srcCols = TO_NHWC ? HW : C;      // src[row * srcCols + col]
    const int r0 = TO_NHWC ? c0 : hw0, q0 = TO_NHWC ? hw0 : c0;            // tile origin (row, col) in the source
    const bool vec_in = (srcCols & 3) == 0 && ((size_t)src & 15) == 0;
    const bool vec_out = (srcRows & 3) == 0 && ((size_t)dst & 15) == 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int idx = tid + 256 * k;            // 1024 float4 slots: 64 rows x 16 float4
        const int rr = idx >> 4, cc = (idx & 15) * 4;
        const int gr = r0 + rr, gc = q0 + cc;
        if (gr < srcRows) {
            if (vec_in && gc + 3 < srcCols) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)gr * srcCols + gc));
                tile[rr][cc] = v.x; tile[rr][cc + 1] = v.y; tile[rr][cc + 2] = v.z; tile[rr][cc + 3] = v.w;
            } else {
                for (int e = 0; e < 4; e++)
                    if (gc + e < srcCols) tile[rr][cc + e] = __ldg(src + (size_t)gr * srcCols + gc + e);
            }
        }
    }
    __syncthreads();
    // destination: dst[col * srcRows + row]; a float4 covers 4 consecutive source rows of one source column
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int idx = tid + 256 * k;
        const int cc = idx >> 4, rr = (idx & 15) * 4;  // cc = source column (dest row), rr = first of 4 source rows
        const int gc = q0 + cc, gr = r0 + rr;
        if (gc < srcCols) {
            if (vec_out && gr + 3 < srcRows) {
                const float4 v = make_float4(tile[rr][cc], tile[rr + 1][cc], tile[rr + 2][cc], tile[rr + 3][cc]);
                *reinterpret_cast<float4*>(dst + (size_t)gc * srcRows + gr) = v;
            } else {
                for (int e = 0; e < 4; e++)
                    if (gr + e < srcRows) dst[(size_t)gc * srcRows + gr + e] = tile[rr + e][cc];
            }
        }
    }
}

template <bool TO_NHWC>
__global__ void __launch_bounds__(256) transpose_kernel(TransposeJob job) {
    __shared__ float tile[64][65];
    transpose_tile<TO_NHWC>(job, blockIdx.x, tile);
}

struct LevelSet;
struct RoiGeom;
struct PrepArgs {                 // order block riding along with the NCHW->NHWC transpose (see roi_order_block)
    const float* rois;
    int K;
    const RoiGeom* geoms;
    int* order;
};
static int launch_transpose(bool to_nhwc, const float* const* src, float* const* dst, const int* H, const int* W, int L, int N,
                            int C, cudaStream_t st, const PrepArgs* prep = nullptr);

// ---------------------------------------------------------------------------------- geometry
struct RoiGeom {
    int batch, level, gh, gw;
    float cw, ch, bin_h, bin_w, start_h, start_w, cosv, sinv;
};

// oriented_single_level.py:73-89 (roi_rescale), :53-71 (map_roi_levels); roi_align_rotated_v1.py:85-120
__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ r, const LevelSet& L) {
    RoiGeom g;
    g.batch = (int)r[0];
    float w = __fmul_rn(L.extend_w, r[3]);
    float h = __fmul_rn(L.extend_h, r[4]);
    int lvl = 0;
    if (L.num_levels > 1) {
        float s = sqrtf(__fmul_rn(w, h));
        float t = floorf(log2f(__fadd_rn(__fdiv_rn(s, L.finest_scale), 1e-6f)));
        t = fminf(fmaxf(t, 0.f), (float)(L.num_levels - 1));
        lvl = (int)t;
    }
    g.level = lvl;
    const float sc = L.scale[lvl];
    if (L.version == 1) {
        g.cw = __fsub_rn(__fmul_rn(r[1], sc), 0.5f);
        g.ch = __fsub_rn(__fmul_rn(r[2], sc), 0.5f);
    } else {
        g.cw = __fmul_rn(r[1], sc);
        g.ch = __fmul_rn(r[2], sc);
    }
    float rw = fmaxf(__fmul_rn(w, sc), 1.f);
    float rh = fmaxf(__fmul_rn(h, sc), 1.f);
    g.bin_h = __fdiv_rn(rh, (float)L.PH);
    g.bin_w = __fdiv_rn(rw, (float)L.PW);
    g.gh = L.sampling_ratio > 0 ? L.sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)L.PH));
    g.gw = L.sampling_ratio > 0 ? L.sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)L.PW));
    g.start_h = -rh * 0.5f;
    g.start_w = -rw * 0.5f;
    g.cosv = cosf(r[5]);
    g.sinv = sinf(r[5]);
    return g;
}

__device__ __forceinline__ void sample_xy(const RoiGeom& g, int version, int ph, int pw, int iy, int ix, float& x, float& y) {
    float yy = __fadd_rn(__fadd_rn(g.start_h, __fmul_rn((float)ph, g.bin_h)),
                         __fdiv_rn(__fmul_rn((float)iy + .5f, g.bin_h), (float)g.gh));
    float xx = __fadd_rn(__fadd_rn(g.start_w, __fmul_rn((float)pw, g.bin_w)),
                         __fdiv_rn(__fmul_rn((float)ix + .5f, g.bin_w), (float)g.gw));
    if (version == 1) {  // clockwise-positive, roi_align_rotated_v1.py:133-134
        x = __fadd_rn(__fadd_rn(__fmul_rn(xx, g.cosv), __fmul_rn(yy, g.sinv)), g.cw);
        y = __fadd_rn(__fsub_rn(__fmul_rn(yy, g.cosv), __fmul_rn(xx, g.sinv)), g.ch);
    } else {  // roi_align_rotated.py:116-117
        x = __fadd_rn(__fsub_rn(__fmul_rn(xx, g.cosv), __fmul_rn(yy, g.sinv)), g.cw);
        y = __fadd_rn(__fadd_rn(__fmul_rn(xx, g.sinv), __fmul_rn(yy, g.cosv)), g.ch);
    }
}

struct Taps {
    int o[4];    // pixel offsets (y*W+x) of lt, rt, lb, rb; all 0 when the sample is out of range
    float w[4];  // bilinear weights; all 0 when out of range
};

// bilinear_interpolate(_gradient): roi_align_rotated_v1.py:23-68, 149-190
__device__ __forceinline__ Taps make_taps(int H, int W, float y, float x) {
    Taps t;
    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
        t.o[0] = t.o[1] = t.o[2] = t.o[3] = 0;
        t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
        return t;
    }
    if (y < 0) y = 0;
    if (x < 0) x = 0;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
    float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
    float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    t.o[0] = y_low * W + x_low;  t.o[1] = y_low * W + x_high;
    t.o[2] = y_high * W + x_low; t.o[3] = y_high * W + x_high;
    t.w[0] = __fmul_rn(hy, hx); t.w[1] = __fmul_rn(hy, lx); t.w[2] = __fmul_rn(ly, hx); t.w[3] = __fmul_rn(ly, lx);
    return t;
}

// base + off*16 as ONE mad.wide.u32 (the compiler otherwise builds the 64-bit address in 4 instructions)
__device__ __forceinline__ const float* tap_ptr(const float4* base, unsigned off16) {
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 16, %2;" : "=l"(a) : "r"(off16), "l"((unsigned long long)base));
    return reinterpret_cast<const float*>(a);
}
__device__ __forceinline__ float4 ldg_nc_v4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_cs_v4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg_cs_v4(const float* p) {
    float4 r;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// shared -> global bulk copy (cp.async.bulk, SASS UBLKCP) with an evict-first L2 policy: the output stream must
// not push the feature pyramid out of L2.  Returns once the source has been read (shared memory may be reused
// or the CTA may exit); global visibility follows at kernel end.
__device__ __forceinline__ void bulk_store_evict_first(float* dst, const float* src_smem, unsigned bytes) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(src_smem)), "r"(bytes), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// shared-memory layout of the fast path
//   taps   : nsamp * Taps (32 B each)
//   stage  : [4][nbins][Q+1] floats   (k = channel within quad, Q = quads per CTA chunk)
// merged tap lists: bin pitch cap + 1 entries of 8 bytes, region rounded to 16 bytes (float4 staging follows)
__host__ __device__ inline size_t list_bytes(size_t nbins, size_t cap) { return (8 * nbins * (cap + 1) + 15) & ~(size_t)15; }
__host__ __device__ inline int quads_per_chunk(int C) { return (C / 4) < 64 ? (C / 4) : 64; }

// ---------------------------------------------------------------------------------- per-RoI geometry
// sin/cos/log2/sqrt and six divisions per RoI: done by K parallel threads here instead of by thread 0 of
// every RoI's CTA (a ~1.5 us serial chain in front of each CTA's barrier).
__global__ void roi_geometry_kernel(LevelSet L, const float* __restrict__ rois, int K, RoiGeom* __restrict__ geoms,
                                    int32_t* __restrict__ levels_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    RoiGeom g = roi_geometry(rois + (size_t)i * 6, L);
    geoms[i] = g;
    if (levels_out) levels_out[i] = g.level;
}

// ---------------------------------------------------------------------------------- processing order
// RoIs are independent, so the ORDER in which CTAs take them is free.  A one-CTA counting sort buckets
// them by (level, 128-pixel image cell): CTAs that run at the same time then read the same few MB of
// one feature level, which keeps the gather inside L2 (the 200 MB output stream would otherwise push
// the 89 MB pyramid out between re-reads).  Output positions are untouched (roi index = output row).
constexpr int kCellShift = 7;   // 128-pixel cells in image coordinates
constexpr int kCellsPerAxis = 16;
constexpr int kBuckets = RSDET_MAX_LEVELS * kCellsPerAxis * kCellsPerAxis;

__global__ void __launch_bounds__(1024) roi_order_kernel(LevelSet L, const float* __restrict__ rois, const RoiGeom* __restrict__ geoms,
                                                         int K, int* __restrict__ order, int32_t* __restrict__ levels_out) {
    __shared__ int s_hist[kBuckets];
    __shared__ int s_warp[32];
    const int tid = threadIdx.x;
    for (int i = tid; i < kBuckets; i += 1024) s_hist[i] = 0;
    __syncthreads();
    auto bucket_of = [&](int i, int& lvl) {
        const float* r = rois + (size_t)i * 6;
        lvl = geoms[i].level;
        int cx = min(max((int)r[1] >> kCellShift, 0), kCellsPerAxis - 1);
        int cy = min(max((int)r[2] >> kCellShift, 0), kCellsPerAxis - 1);
        // boustrophedon rows: neighbouring buckets are neighbouring cells
        if (cy & 1) cx = kCellsPerAxis - 1 - cx;
        return (lvl * kCellsPerAxis + cy) * kCellsPerAxis + cx;
    };
    for (int i = tid; i < K; i += 1024) {
        int lvl;
        int bkt = bucket_of(i, lvl);
        if (levels_out) levels_out[i] = lvl;
        atomicAdd(&s_hist[bkt], 1);
    }
    __syncthreads();
    // exclusive scan of kBuckets (=2048) counters: 2 per thread
    int a0 = s_hist[2 * tid], a1 = s_hist[2 * tid + 1];
    int sum = a0 + a1, x = sum;
    const int lane = tid & 31, w = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += y; }
        s_warp[lane] = v;
    }
    __syncthreads();
    int excl = (w ? s_warp[w - 1] : 0) + x - sum;
    s_hist[2 * tid] = excl;
    s_hist[2 * tid + 1] = excl + a0;
    __syncthreads();
    for (int i = tid; i < K; i += 1024) {
        int lvl;
        int bkt = bucket_of(i, lvl);
        order[atomicAdd(&s_hist[bkt], 1)] = i;
    }
}

// The locality order of a call (what roi_order_kernel does) as ONE block of any size, so that for NCHW callers it can
// ride along as an extra block of the NCHW->NHWC transpose: the order kernel is a 7.8 us single-CTA latency chain in
// front of every gather, the transpose of a 1024^2 pyramid keeps the other SMs busy for 30 us anyway.  (Doing the
// geometry in the same block as well was measured and dropped: sin/cos/log2 of 4000 RoIs on one SM take longer than
// the two launches they replace.)  geoms[] must have been written by roi_geometry_kernel.  s_hist: kBuckets ints.
constexpr int kPrepMaxRois = 16384;
template <int THREADS>
__device__ __forceinline__ void roi_order_block(const float* __restrict__ rois, int K, const RoiGeom* __restrict__ geoms,
                                                int* __restrict__ order, int* s_hist, int* s_warp) {
    const int tid = threadIdx.x;
    for (int i = tid; i < kBuckets; i += THREADS) s_hist[i] = 0;
    __syncthreads();
    auto bucket_of = [&](const float* r, int lvl) {
        int cx = min(max((int)r[1] >> kCellShift, 0), kCellsPerAxis - 1);
        int cy = min(max((int)r[2] >> kCellShift, 0), kCellsPerAxis - 1);
        if (cy & 1) cx = kCellsPerAxis - 1 - cx;                      // boustrophedon rows: neighbouring buckets = neighbouring cells
        return (lvl * kCellsPerAxis + cy) * kCellsPerAxis + cx;
    };
    for (int i = tid; i < K; i += THREADS) {
        const float* r = rois + (size_t)i * 6;
        atomicAdd(&s_hist[bucket_of(r, geoms[i].level)], 1);
    }
    __syncthreads();
    // exclusive scan of kBuckets counters, kBuckets / THREADS consecutive ones per thread
    constexpr int PER = kBuckets / THREADS;
    int a[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { a[k] = s_hist[PER * tid + k]; sum += a[k]; }
    int x = sum;
    const int lane = tid & 31, w = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int v = lane < THREADS / 32 ? s_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += y; }
        s_warp[lane] = v;
    }
    __syncthreads();
    int excl = (w ? s_warp[w - 1] : 0) + x - sum;
#pragma unroll
    for (int k = 0; k < PER; k++) { s_hist[PER * tid + k] = excl; excl += a[k]; }
    __syncthreads();
    for (int i = tid; i < K; i += THREADS) {
        const float* r = rois + (size_t)i * 6;
        order[atomicAdd(&s_hist[bucket_of(r, geoms[i].level)], 1)] = i;
    }
}

// NCHW callers: the pyramid transpose with the order block riding along as block 0
__global__ void __launch_bounds__(256) transpose_prep_kernel(TransposeJob job, const float* __restrict__ rois, int K,
                                                             const RoiGeom* __restrict__ geoms, int* __restrict__ order) {
    __shared__ float tile[64][65];
    static_assert(sizeof(float) * 64 * 65 >= sizeof(int) * (kBuckets + 32), "the prep block borrows the transpose tile");
    if (blockIdx.x == 0) {
        int* s_hist = reinterpret_cast<int*>(&tile[0][0]);
        roi_order_block<256>(rois, K, geoms, order, s_hist, s_hist + kBuckets);
        return;
    }
    transpose_tile<true>(job, blockIdx.x - 1, tile);
}

static int launch_transpose(bool to_nhwc, const float* const* src, float* const* dst, const int* H, const int* W, int L, int N,
                            int C, cudaStream_t st, const PrepArgs* prep) {
    TransposeJob job;
    job.num_levels = L; job.N = N; job.C = C;
    int total = 0;
    for (int l = 0; l < L; l++) {
        job.src[l] = src[l]; job.dst[l] = dst[l]; job.HW[l] = H[l] * W[l];
        job.tile_begin[l] = total;
        total += N * ceil_div(job.HW[l], 64) * ceil_div(C, 64);
    }
    job.tile_begin[L] = total;
    if (prep && to_nhwc) {
        transpose_prep_kernel<<<total + 1, 256, 0, st>>>(job, prep->rois, prep->K, prep->geoms, prep->order);
        count_launch();
        return cuda_status();
    }
    if (total == 0) return RSDET_OK;
    if (to_nhwc) transpose_kernel<true><<<total, 256, 0, st>>>(job);
    else transpose_kernel<false><<<total, 256, 0, st>>>(job);
    count_launch();
    return cuda_status();
}

// geometry (+ locality order for calls of >= 256 RoIs) of a call
static void launch_prep(const LevelSet& L, const float* rois, int K, RoiGeom* geoms, int* order, int32_t* levels_out, cudaStream_t st,
                        bool order_rides_with_transpose = false) {
    roi_geometry_kernel<<<ceil_div(K, 128), 128, 0, st>>>(L, rois, K, geoms, levels_out);
    count_launch();
    if (order && !order_rides_with_transpose) {
        roi_order_kernel<<<1, 1024, 0, st>>>(L, rois, geoms, K, order, nullptr);
        count_launch();
    }
}

// ---------------------------------------------------------------------------------- tap lists
// Shared by forward and backward.  Measured on B200 (profiles/README.md): the gather is bound by L2
// round trips and L1 load wavefronts -- every bilinear tap is a 512-byte warp load and a RoI issues 784
// of them per 128 channels, ~9x its own output -- not by HBM.  So the sampling grid is computed once per
// RoI (A1) AND merged per bin (A2): taps of the bin's samples that land on the same pixel are summed
// into one (offset, weight) entry and zero-weight taps are dropped; on the benchmark proposals 16 taps
// collapse to ~9.  Offsets are in 16-byte units of the channels-last map (one mad.wide per address).
//   s_list [nbins*(cap+1)] int2 {offset, weight bits} (bin pitch cap + 1: conflict-free per-bin lanes),
//   s_cnt [nbins], tmp: 3*nbins*(cap+1) words of scratch.
//   unit/base: stored offset = base + pixel * unit (unit = C/4 for float4 addressing, 1 for TMA row indices)
__device__ __forceinline__ void build_tap_lists(const RoiGeom& g, const LevelSet& L, int H, int W, int2* s_list, int* s_cnt,
                                                float* tmp, int unit, int base) {
    const int tid = threadIdx.x;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int nsamp = nbins * spb, cap = 4 * spb, ntaps = nbins * cap;
    const int C4 = unit;
    int* s_off = reinterpret_cast<int*>(tmp);
    float* s_w = tmp + ntaps + nbins;      // room for the cap + 1 pitch
    float* s_wsum = tmp + 2 * (ntaps + nbins);
    // A1: one thread per sample.  Scratch pitch per bin = cap + 1 words when the per-bin merge below reads it
    // with one lane per bin (lane stride 17 words: conflict-free; 16 would be a 16-way bank conflict).
    const int tp = cap == 16 ? 17 : cap;
    for (int s = tid; s < nsamp; s += kRoiThreads) {
        int b = s / spb, q = s % spb;
        int ph = b / L.PW, pw = b % L.PW, iy = q / g.gw, ix = q % g.gw;
        float x, y;
        sample_xy(g, L.version, ph, pw, iy, ix, x, y);
        const Taps t = make_taps(H, W, y, x);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            s_off[b * tp + q * 4 + k] = base + t.o[k] * C4;
            s_w[b * tp + q * 4 + k] = t.w[k];
        }
    }
    __syncthreads();
    if (cap == 16) {
        // A2 (2x2 sampling grid, the Oriented R-CNN setting): ONE LANE PER BIN, everything in registers.  The
        // warp-cooperative variant it replaces (below, still used for cap == 4) spent 32 shuffles per 16 taps:
        // 4.1 M SHFL per 4000 RoIs = 6.9 M of the kernel's 29.8 M L1 wavefronts (shuffles share the data pipe
        // of the gather loads, profiles/README.md).  Same arithmetic: a tap is a LEADER if its weight is
        // non-zero and no earlier live tap of the bin hits the same pixel; a leader adds the weights of its
        // later duplicates in tap order; leaders are compacted in tap order.
        for (int b = tid; b < nbins; b += kRoiThreads) {
            int o[16];
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) { o[j] = s_off[b * 17 + j]; w[j] = s_w[b * 17 + j]; }
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float acc = w[j];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (i <= j) continue;                            // constant trip count: stays in registers
                    const bool same = o[i] == o[j] && w[j] != 0.f;   // w[j] == 0: dead or already absorbed
                    acc += same ? w[i] : 0.f;
                    w[i] = same ? 0.f : w[i];
                }
                if (w[j] != 0.f) s_list[b * 17 + pos++] = make_int2(o[j], __float_as_int(acc));
            }
            s_cnt[b] = pos;
        }
        __syncthreads();
        return;
    }
    if (cap == 4) {
        // A2 (fast): cap lanes per bin, everything in registers.  Each lane walks the cap lanes of its bin with
        // warp shuffles (uniform trip count): it is a LEADER if no earlier live lane hits the same pixel, and it
        // sums the weights of all live lanes on its pixel in lane order; a ballot of the leaders gives each
        // one its slot in the bin's compacted list.
        const int lane = tid & 31;
        const int rounds = (ntaps + kRoiThreads - 1) / kRoiThreads;
        for (int rd = 0; rd < rounds; rd++) {
            const int t = rd * kRoiThreads + tid;
            const bool in = t < ntaps;
            const int b = in ? t / cap : 0, j = in ? t - b * cap : 0;
            const int o = in ? s_off[t] : -1;
            const float w = in ? s_w[t] : 0.f;
            const bool live = in && w != 0.f;
            const int seg = lane & ~(cap - 1);                        // first lane of my bin inside the warp
            bool leader = live;
            float wsum = 0.f;
            for (int i = 0; i < cap; i++) {                           // same trip count on every lane
                const int oi = __shfl_sync(0xffffffffu, o, seg + i);
                const float wi = __shfl_sync(0xffffffffu, w, seg + i);
                const bool same = live && wi != 0.f && oi == o;
                if (same && i < (lane - seg)) leader = false;
                if (same) wsum += wi;
            }
            const unsigned leaders = __ballot_sync(0xffffffffu, leader);
            const unsigned segmask = ((1u << cap) - 1u) << seg;       // cap is 4 or 16 here
            if (leader) s_list[b * (cap + 1) + __popc(leaders & segmask & ((1u << lane) - 1u))] = make_int2(o, __float_as_int(wsum));
            if (in && j == 0) s_cnt[b] = __popc(leaders & segmask);
        }
        __syncthreads();
        return;
    }
    // A2a: one thread per tap: a tap is a LEADER if its weight is non-zero and no earlier tap of the bin
    // hits the same pixel; a leader collects the weights of its later duplicates (fixed order).
    for (int t = tid; t < ntaps; t += kRoiThreads) {
        const int b = t / cap, j = t - b * cap;
        const int* ob = s_off + b * cap;
        const float* wb = s_w + b * cap;
        const int o = ob[j];
        float w = wb[j];
        bool leader = w != 0.f;
        for (int i = 0; i < j && leader; i++) leader = !(ob[i] == o && wb[i] != 0.f);
        if (leader)
            for (int i = j + 1; i < cap; i++)
                if (ob[i] == o) w += wb[i];
        s_wsum[t] = leader ? w : 0.f;
    }
    __syncthreads();
    // A2b: compaction (position = number of leaders before me in my bin)
    for (int t = tid; t < ntaps; t += kRoiThreads) {
        const int b = t / cap, j = t - b * cap;
        const float* ws = s_wsum + b * cap;
        const float w = ws[j];
        int pos = 0;
        for (int i = 0; i < j; i++) pos += ws[i] != 0.f;
        if (w != 0.f) s_list[b * (cap + 1) + pos] = make_int2(s_off[t], __float_as_int(w));
        if (j == cap - 1) s_cnt[b] = pos + (w != 0.f);
    }
    __syncthreads();
}

// Conflict-free access to the [c][bin] staging block.  A warp touches channels 4*lane + i of one bin at a
// time; with all lanes on the same component i the 32 addresses fall on 8 banks (4-way conflict: 7.2 M of the
// forward kernel's 34 M L1 wavefronts in profiles/r1_d).  Lane octet o = lane/8 therefore handles component
// (i + o) & 3 at step i: for an odd bin count the four octets land on the four residues mod 4 and the 32
// banks are distinct.  rot4 rotates a quad left by o so that step i finds its value in slot i.
__device__ __forceinline__ float4 rot4(float4 v, int o) {
    if (o & 1) v = make_float4(v.y, v.z, v.w, v.x);
    if (o & 2) v = make_float4(v.z, v.w, v.x, v.y);
    return v;
}
__device__ __forceinline__ float4 unrot4(float4 v, int o) { return rot4(v, (4 - o) & 3); }

// ---------------------------------------------------------------------------------- forward (fast)
// One CTA per (RoI, chunk of up to 64 channel quads), thread = (channel quad(s), bin group).  QPT = 2:
// the second channel quad is an immediate +512 B off the same address.
template <int QPT>
__global__ void __launch_bounds__(kRoiThreads, 4)
roi_align_fwd_kernel(LevelSet L, const float* __restrict__ rois, const int* __restrict__ order, const RoiGeom* __restrict__ geoms,
                     int K, float* __restrict__ out, int32_t* __restrict__ levels_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int roi = order ? order[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;  // samples per bin (fast path: fixed grid)
    const int cap = 4 * spb;                               // taps per bin before merging
    const int C = L.C;
    const int Q = quads_per_chunk(C);
    const int chunk0 = blockIdx.y * Q * 4;                 // first channel of this chunk
    const int Qc = min(Q, (C - chunk0) / 4);               // quads in this chunk
    // smem: [merged lists: nbins*cap int2][counts: nbins int, padded][staging [c][bin] | phase-A scratch]
    int2* s_list = reinterpret_cast<int2*>(smem_raw);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + list_bytes(nbins, cap));
    float* s_stage = reinterpret_cast<float*>(smem_raw + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~15));
    __shared__ RoiGeom s_g;

    if (!geoms) {
        if (tid == 0) {
            s_g = roi_geometry(rois + (size_t)roi * 6, L);
            if (levels_out && blockIdx.y == 0) levels_out[roi] = s_g.level;
        }
        __syncthreads();
    }
    const RoiGeom g = geoms ? geoms[roi] : s_g;
    const int H = L.H[g.level], W = L.W[g.level];
    if (L.dbg_skip_main < 2) build_tap_lists(g, L, H, W, s_list, s_cnt, s_stage, C >> 2, 0);   // the scratch is dead after its final barrier
    if (L.dbg_skip_main) { for (int b = tid; b < nbins; b += kRoiThreads) s_cnt[b] = 0; __syncthreads(); }

    const int lanes = QPT == 2 ? 32 : Qc;                  // threads per bin group
    const int groups = kRoiThreads / lanes;
    const int cq = tid % lanes, grp = tid / lanes;
    const int oct = (cq >> 3) & 3;
    const float4* __restrict__ feat =
        reinterpret_cast<const float4*>(L.feat[g.level] + (size_t)g.batch * H * W * C + chunk0) + cq;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)max(spb, 1), inv_count = 1.f / count;
    if (grp < groups) {
        for (int b = grp; b < nbins; b += groups) {
            float4 acc[QPT];
#pragma unroll
            for (int u = 0; u < QPT; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int2* lp = s_list + b * (cap + 1);
            const int cnt = s_cnt[b];
            for (int e = 0; e < cnt; e += 4) {
                int2 en[4];
#pragma unroll
                for (int k = 0; k < 4; k++) en[k] = lp[min(e + k, cnt - 1)];  // warp-uniform broadcast reads
                float4 v[QPT][4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (e + k < cnt) {
#pragma unroll
                        for (int u = 0; u < QPT; u++) v[u][k] = ldg_nc_v4(tap_ptr(feat, (unsigned)en[k].x) + u * 128);
                    }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (e + k < cnt) {
                        const float wt = __int_as_float(en[k].y);
#pragma unroll
                        for (int u = 0; u < QPT; u++) {
                            acc[u].x = fmaf(wt, v[u][k].x, acc[u].x);
                            acc[u].y = fmaf(wt, v[u][k].y, acc[u].y);
                            acc[u].z = fmaf(wt, v[u][k].z, acc[u].z);
                            acc[u].w = fmaf(wt, v[u][k].w, acc[u].w);
                        }
                    }
            }
#pragma unroll
            for (int u = 0; u < QPT; u++) {
                // output_val /= count (:143); a power-of-two count makes the reciprocal multiply exact
                if (pow2) { acc[u].x *= inv_count; acc[u].y *= inv_count; acc[u].z *= inv_count; acc[u].w *= inv_count; }
                else { acc[u].x /= count; acc[u].y /= count; acc[u].z /= count; acc[u].w /= count; }
                const int c0 = (cq + u * 32) * 4;
                const float4 r = rot4(acc[u], oct);
                s_stage[(c0 + ((0 + oct) & 3)) * nbins + b] = r.x;
                s_stage[(c0 + ((1 + oct) & 3)) * nbins + b] = r.y;
                s_stage[(c0 + ((2 + oct) & 3)) * nbins + b] = r.z;
                s_stage[(c0 + ((3 + oct) & 3)) * nbins + b] = r.w;
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // the staged block IS the output block of this (roi, chunk): one bulk copy shared -> global (TMA) when it is
    // 16-byte aligned, else a coalesced copy with streaming stores
    float* __restrict__ dst = out + ((size_t)roi * C + chunk0) * nbins;
    const int total = Qc * 4 * nbins;
    if ((((size_t)roi * C + chunk0) * nbins & 3) == 0 && (total & 3) == 0 && (((size_t)out) & 15) == 0) {
        if (tid == 0) bulk_store_evict_first(dst, s_stage, (unsigned)total * 4u);
    } else {
        for (int e = tid; e < total; e += kRoiThreads) __stcs(dst + e, s_stage[e]);
    }
}

#ifdef RSDET_TUNING   // measured alternatives of the forward gather (A/B builds only, see roi_path_choice)
// ---------------------------------------------------------------------------------- forward (row windows)
// The default forward path for the Oriented R-CNN geometry (7x7 bins, 2x2 samples per bin, C % 256 == 0).
//
// What bounds the gather on B200 is the path from L2 to the SM, not HBM: the bin-major kernel above pulls ~441
// merged taps x 1 KB per RoI through it although they cover only ~223 DISTINCT feature pixels -- neighbouring bins
// share the one-pixel border of their bilinear footprints.  Here a warp owns one bin ROW and walks it in WINDOWS of
// two adjacent bins (w, w+1): every distinct pixel of the row is loaded once, in the first window that uses it,
// and applied to both bins of the window with two weights; the accumulator of bin w+1 is carried into window w+1
// as its first bin.  Accumulators are therefore statically indexed (no per-contribution control flow), a row needs
// ~44 pixel loads instead of ~63, and a pixel feeding three or more consecutive bins (tiny RoIs) simply appears in
// a later window again.  Rows r and r+1 run side by side in neighbouring warps, so their shared border hits in L1.
//
// List construction is warp-local and deterministic:
//   A1  one thread per sample: 4 (pixel, weight) taps in the reference's float arithmetic, bounding box of the
//       RoI's tap pixels;
//   per row warp: bitmap of the row's pixels inside that box (shared-memory atomicOr) -> rank = prefix popcount =
//       raster order; weight table wt[pixel][bin] accumulated in four rounds (round q = sample q of every bin, so a
//       (pixel, bin) cell receives at most one add per round: plain read-modify-write, fixed order); per pixel the
//       windows are chosen greedily over its bin mask and window lists are laid out with ballots (pixel order).
// RoIs whose box exceeds the 4096-pixel bitmap (long diagonal ones) skip the dedupe: one entry per tap.
constexpr int kPxRows = 7, kPxCols = 7;
constexpr int kPxBins = kPxRows * kPxCols;
constexpr int kPxMaxTaps = 16 * kPxCols;             // taps of one bin row
constexpr int kPxEntPitch = kPxMaxTaps + 3 * kPxCols + 3;  // entries per row: <= one per tap, window starts 4-aligned (batches of 4) -> 136
constexpr int kPxBmWords = 128;                      // dedupe bitmap: boxes of up to 4096 pixels
constexpr int kPxPixPad = 128;                       // per-tap / per-pixel arrays (<= 112 taps per row)
constexpr int kPxWtRows = 88;                        // distinct pixels per row handled by the shared-pixel lists (1 % of rows have more)

struct alignas(16) PxLists {                         // what the gather reads: one RoI
    unsigned pix[kPxRows][kPxEntPitch];              // pixel index (y * W + x) per entry
    float wa[kPxRows][kPxEntPitch];                  // weight for the window's first bin
    float wb[kPxRows][kPxEntPitch];                  // weight for the window's second bin
    int wbeg[kPxRows][8], wcnt[kPxRows][8];
};
constexpr int kPxTapPitch = 20;                      // taps of a bin: 16 + 4 pad words (the rounds read bin * 20 + q * 4 + k: 28 banks)
struct alignas(16) PxTaps {                          // A1 output: (y << 16 | x, weight) per tap, tap = bin * 20 + sample * 4 + k
    int key[kPxBins * kPxTapPitch];
    float w[kPxBins * kPxTapPitch];
};
struct alignas(16) PxRowScratch {                    // per-row build scratch
    unsigned bm[kPxBmWords];
    int wpre[kPxBmWords];
    float wt[kPxWtRows * 8];                         // [pixel rank][bin column]; rows with more distinct pixels go DIRECT
    unsigned pid[kPxPixPad];                         // its pixel index
    unsigned char rk[kPxPixPad];                     // pixel rank of every tap of the row
};

// windows chosen for a pixel feeding the bins in `m`: lowest uncovered bin w opens window (w, w+1)
__device__ __forceinline__ unsigned px_windows(unsigned m) {
    unsigned e = 0;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        if (m) { const int w = __ffs(m) - 1; e |= 1u << w; m &= ~(3u << w); }
    }
    return e;
}

__device__ __forceinline__ unsigned ws_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool ws_mbar_try(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(ws_smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// `sleep_ns` > 0: back off between polls (a role that is far ahead must not burn the issue slots the others need)
__device__ __forceinline__ void ws_mbar_wait(unsigned long long* b, unsigned parity, unsigned sleep_ns = 0) {
    while (!ws_mbar_try(b, parity))
        if (sleep_ns) __nanosleep(sleep_ns);
}

// A1 for sample s (= bin * 4 + q): the four bilinear taps of make_taps, kept as (x, y) keys; updates the caller's
// bounding box of touched pixels.  (A clamped sample repeats a pixel, but the repeated tap then has weight exactly 0.)
__device__ __forceinline__ void px_sample(const RoiGeom& g, int version, int H, int W, int s, PxTaps& T, int& x0, int& x1,
                                          int& y0, int& y1) {
    const int b = s >> 2, q = s & 3;
    const int ph = b / kPxCols, pw = b - ph * kPxCols, iy = q >> 1, ix = q & 1;
    float x, y;
    sample_xy(g, version, ph, pw, iy, ix, x, y);
    int kx[2] = {0, 0}, ky[2] = {0, 0};
    float wx[2] = {0.f, 0.f}, wy[2] = {0.f, 0.f};
    if (!(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W)) {
        if (y < 0) y = 0;
        if (x < 0) x = 0;
        int yl = (int)y, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
        wy[0] = __fsub_rn(1.f, ly); wy[1] = ly; wx[0] = __fsub_rn(1.f, lx); wx[1] = lx;
        kx[0] = xl; kx[1] = xh; ky[0] = yl; ky[1] = yh;
        x0 = min(x0, xl); x1 = max(x1, xh); y0 = min(y0, yl); y1 = max(y1, yh);
    }
    int4 kk;
    float4 ww;
    kk.x = (ky[0] << 16) | kx[0]; kk.y = (ky[0] << 16) | kx[1]; kk.z = (ky[1] << 16) | kx[0]; kk.w = (ky[1] << 16) | kx[1];
    ww.x = __fmul_rn(wy[0], wx[0]); ww.y = __fmul_rn(wy[0], wx[1]); ww.z = __fmul_rn(wy[1], wx[0]); ww.w = __fmul_rn(wy[1], wx[1]);
    *reinterpret_cast<int4*>(&T.key[b * kPxTapPitch + q * 4]) = kk;
    *reinterpret_cast<float4*>(&T.w[b * kPxTapPitch + q * 4]) = ww;
}

// One warp builds the window lists of bin row `row` (warp-local: only __syncwarp inside).
__device__ __forceinline__ void px_build_row(const PxTaps& T, PxRowScratch& R, PxLists& Lst, int row, int lane, int W,
                                             const int* box) {
    const int px0 = box[0], py0 = box[2];
    const int pwid = box[1] - px0 + 1, phgt = box[3] - py0 + 1;
    const bool dedupe = pwid > 0 && pwid * phgt <= 32 * kPxBmWords;
    int key[4], loc[4], rank[4];
    float wgt[4];
    bool val[4];
    // lane's taps t = lane + 32 i: bin column 2 i + (lane >> 4), sample (lane >> 2) & 3 (the same for all four)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int t = lane + 32 * i;
        const bool in = t < kPxMaxTaps;
        const int ti = (row * kPxCols + (t >> 4)) * kPxTapPitch + (t & 15);
        key[i] = in ? T.key[ti] : 0;
        wgt[i] = in ? T.w[ti] : 0.f;
        val[i] = wgt[i] != 0.f;
        // bit position of the pixel in the box bitmap.  Any fixed bijection gives a deterministic pixel order; this
        // one sends raster neighbours to different words (word = index mod 128), because the taps of a bin row sit
        // in a few adjacent pixel rows and a raster bitmap would serialise their atomicOr on the same words
        const int ras = ((key[i] >> 16) - py0) * pwid + ((key[i] & 0xffff) - px0);
        loc[i] = ((ras & (kPxBmWords - 1)) << 5) | (ras >> 7);
        R.bm[lane + 32 * i] = 0u;
    }
    __syncwarp();
    int npix = kPxWtRows + 1;
    if (dedupe) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (val[i]) atomicOr(&R.bm[loc[i] >> 5], 1u << (loc[i] & 31));
        __syncwarp();
        const uint4 m = *reinterpret_cast<const uint4*>(R.bm + 4 * lane);
        const int c0 = __popc(m.x), c1 = __popc(m.y), c2 = __popc(m.z), c3 = __popc(m.w);
        int x = c0 + c1 + c2 + c3;
        const int sum = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        npix = __shfl_sync(0xffffffffu, x, 31);
        const int ex = x - sum;
        *reinterpret_cast<int4*>(R.wpre + 4 * lane) = make_int4(ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i++)
            rank[i] = val[i] ? R.wpre[loc[i] >> 5] + __popc(R.bm[loc[i] >> 5] & ((1u << (loc[i] & 31)) - 1u)) : 0;
    }
    if (npix > kPxWtRows) {
        // DIRECT mode -- the box exceeds the bitmap (long diagonal RoI) or the row touches more than 88 distinct
        // pixels (large bins: little to share): one entry per tap, in window (bin, bin + 1) with the second weight 0
        int acc = 0;
#pragma unroll
        for (int w = 0; w < kPxCols; w++) {
            const int i = w >> 1;                                     // taps of bin column w: slot i, lane half w & 1
            const bool has = val[i] && (lane >> 4) == (w & 1);
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) {
                const int idx = acc + __popc(bal & ((1u << lane) - 1u));
                Lst.pix[row][idx] = (unsigned)((key[i] >> 16) * W + (key[i] & 0xffff));
                Lst.wa[row][idx] = wgt[i];
                Lst.wb[row][idx] = 0.f;
            }
            if (lane == 0) { Lst.wbeg[row][w] = acc; Lst.wcnt[row][w] = __popc(bal); }
            acc += (__popc(bal) + 3) & ~3;
        }
        return;
    }
    // pixel index per rank; tap -> rank table for the rounds below
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (val[i]) {
            R.pid[rank[i]] = (unsigned)((key[i] >> 16) * W + (key[i] & 0xffff));
            R.rk[lane + 32 * i] = (unsigned char)rank[i];
        }
    // weight table wt[pixel][bin]: zero the live rows, then four rounds.  Round q adds sample q of every bin
    // (lane = bin column * 4 + tap): a (pixel, bin) cell receives at most one add per round, so the plain
    // read-modify-write is race-free and the summation order is fixed.
    for (int e = lane; e < npix * 2; e += 32) reinterpret_cast<float4*>(R.wt)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    {
        const int bc = lane >> 2, k = lane & 3;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (lane < 4 * kPxCols) {
                const float wq = T.w[(row * kPxCols + bc) * kPxTapPitch + q * 4 + k];
                if (wq != 0.f) R.wt[R.rk[bc * 16 + q * 4 + k] * 8 + bc] += wq;
            }
            __syncwarp();
        }
    }
    // window lists: count, lay out (4-aligned starts), fill -- all in pixel (rank) order.  The bins a pixel feeds are
    // the non-zero cells of its weight row (weights are positive, sums cannot cancel).  npix <= 88: three pixels per lane.
    constexpr int kChunks = (kPxWtRows + 31) / 32;
    unsigned ew[kChunks], pm[kChunks];
    float wv[kChunks][8];
    int cnt[kPxCols];
#pragma unroll
    for (int w = 0; w < kPxCols; w++) cnt[w] = 0;
#pragma unroll
    for (int c = 0; c < kChunks; c++) {
        const int r = c * 32 + lane;
        const bool live = r < npix;
        const float4 lo = live ? *reinterpret_cast<const float4*>(R.wt + r * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 hi = live ? *reinterpret_cast<const float4*>(R.wt + r * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        wv[c][0] = lo.x; wv[c][1] = lo.y; wv[c][2] = lo.z; wv[c][3] = lo.w;
        wv[c][4] = hi.x; wv[c][5] = hi.y; wv[c][6] = hi.z; wv[c][7] = 0.f;
        pm[c] = 0u;
#pragma unroll
        for (int w = 0; w < kPxCols; w++) pm[c] |= wv[c][w] != 0.f ? 1u << w : 0u;
        ew[c] = px_windows(pm[c]);
#pragma unroll
        for (int w = 0; w < kPxCols; w++) cnt[w] += __popc(__ballot_sync(0xffffffffu, (ew[c] >> w) & 1u));
    }
    int beg[kPxCols];
    {
        int acc = 0;
#pragma unroll
        for (int w = 0; w < kPxCols; w++) { beg[w] = acc; acc += (cnt[w] + 3) & ~3; }
    }
    if (lane < kPxCols) {
        int bsel = 0, csel = 0;
#pragma unroll
        for (int w = 0; w < kPxCols; w++) if (lane == w) { bsel = beg[w]; csel = cnt[w]; }
        Lst.wbeg[row][lane] = bsel;
        Lst.wcnt[row][lane] = csel;
    }
#pragma unroll
    for (int c = 0; c < kChunks; c++) {
        const int r = c * 32 + lane;
        const unsigned P = r < npix ? R.pid[r] : 0u;
#pragma unroll
        for (int w = 0; w < kPxCols; w++) {
            const bool has = (ew[c] >> w) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, has);
            if (has) {
                const int idx = beg[w] + __popc(bal & ((1u << lane) - 1u));
                Lst.pix[row][idx] = P;
                Lst.wa[row][idx] = wv[c][w];
                Lst.wb[row][idx] = ((pm[c] >> (w + 1)) & 1u) ? wv[c][w + 1] : 0.f;
            }
            beg[w] += __popc(bal);
        }
    }
}

// One warp gathers bin row `row`: windows (w, w+1), PB pixels (2 x PB LDG.128) in flight per thread; finished bins go
// to the [c][bin] staging block (conflict-free component rotation, see rot4).  `feat` already points at this lane's
// first channel quad of the RoI's image.
template <int PB>   // pixels per load batch; kPxEntPitch is sized for 4
__device__ __forceinline__ void px_gather_row(const PxLists& Lst, const float* __restrict__ feat, unsigned rowbytes,
                                              float* __restrict__ stage, int row, int lane,
                                              unsigned long long* stage_free = nullptr, unsigned free_parity = 0) {
    const int oct = (lane >> 3) & 3;
    float4 A0 = make_float4(0.f, 0.f, 0.f, 0.f), A1 = A0;
    for (int w = 0; w < kPxCols; w++) {
        float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;
        const int beg = Lst.wbeg[row][w], cnt = Lst.wcnt[row][w];
        for (int e = 0; e < cnt; e += PB) {
            unsigned pxs[PB];
            float was[PB], wbs[PB];
#pragma unroll
            for (int h = 0; h < PB / 4; h++) {
                const uint4 px = *reinterpret_cast<const uint4*>(&Lst.pix[row][beg + e + 4 * h]);
                const float4 fa = *reinterpret_cast<const float4*>(&Lst.wa[row][beg + e + 4 * h]);
                const float4 fb = *reinterpret_cast<const float4*>(&Lst.wb[row][beg + e + 4 * h]);
                pxs[4 * h] = px.x; pxs[4 * h + 1] = px.y; pxs[4 * h + 2] = px.z; pxs[4 * h + 3] = px.w;
                was[4 * h] = fa.x; was[4 * h + 1] = fa.y; was[4 * h + 2] = fa.z; was[4 * h + 3] = fa.w;
                wbs[4 * h] = fb.x; wbs[4 * h + 1] = fb.y; wbs[4 * h + 2] = fb.z; wbs[4 * h + 3] = fb.w;
            }
            const int rem = cnt - e;
            float4 v[PB][2];
#pragma unroll
            for (int k = 0; k < PB; k++)
                if (k < rem) {
                    unsigned long long ad;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(ad) : "r"(pxs[k]), "r"(rowbytes), "l"((unsigned long long)feat));
                    v[k][0] = ldg_nc_v4(reinterpret_cast<const float*>(ad));
                    v[k][1] = ldg_nc_v4(reinterpret_cast<const float*>(ad) + 128);
                }
#pragma unroll
            for (int k = 0; k < PB; k++)
                if (k < rem) {
                    const float a = was[k], b = wbs[k];
                    A0.x = fmaf(a, v[k][0].x, A0.x); A0.y = fmaf(a, v[k][0].y, A0.y); A0.z = fmaf(a, v[k][0].z, A0.z); A0.w = fmaf(a, v[k][0].w, A0.w);
                    A1.x = fmaf(a, v[k][1].x, A1.x); A1.y = fmaf(a, v[k][1].y, A1.y); A1.z = fmaf(a, v[k][1].z, A1.z); A1.w = fmaf(a, v[k][1].w, A1.w);
                    B0.x = fmaf(b, v[k][0].x, B0.x); B0.y = fmaf(b, v[k][0].y, B0.y); B0.z = fmaf(b, v[k][0].z, B0.z); B0.w = fmaf(b, v[k][0].w, B0.w);
                    B1.x = fmaf(b, v[k][1].x, B1.x); B1.y = fmaf(b, v[k][1].y, B1.y); B1.z = fmaf(b, v[k][1].z, B1.z); B1.w = fmaf(b, v[k][1].w, B1.w);
                }
        }
        if (w == 0 && stage_free) ws_mbar_wait(stage_free, free_parity, 100);   // the previous block has left shared memory
        const int b = row * kPxCols + w;       // bin (row, w) is complete
#pragma unroll
        for (int u = 0; u < 2; u++) {
            float4 r = u ? A1 : A0;
            r.x *= 0.25f; r.y *= 0.25f; r.z *= 0.25f; r.w *= 0.25f;      // /count, count = 4 samples (exact)
            r = rot4(r, oct);
            const int c0 = (lane + u * 32) * 4;
            stage[(c0 + ((0 + oct) & 3)) * kPxBins + b] = r.x;
            stage[(c0 + ((1 + oct) & 3)) * kPxBins + b] = r.y;
            stage[(c0 + ((2 + oct) & 3)) * kPxBins + b] = r.z;
            stage[(c0 + ((3 + oct) & 3)) * kPxBins + b] = r.w;
        }
        A0 = B0; A1 = B1;
    }
}

#ifdef RSDET_PROF
__device__ unsigned long long* g_px_prof = nullptr;   // set through rsdet_tuning_set_prof (profiling builds only)
#endif

// ---- one CTA per RoI (7 warps = 7 bin rows; build, barrier, gather, bulk store).  Kept for small calls and as the
// A/B reference of the persistent kernel below.
struct PxSmem {
    PxLists lists;
    int box[4];                                      // x0, x1, y0, y1 of the RoI's tap pixels
    union alignas(16) {
        float stage[256 * kPxBins];                  // [c][bin] = the RoI's output block
        struct { PxTaps taps; PxRowScratch row[kPxRows]; } b;
    } u;
};

// blockDim = (32, 7): threadIdx.y is the warp = bin row and is known to be warp-uniform (list addresses live in
// uniform registers, loop bounds are uniform branches).
template <int PB>
__global__ void __block_size__((32, kPxRows, 1)) __maxnreg__(PB == 8 ? 128 : 96)
roi_align_fwd_px_kernel(LevelSet L, const int* __restrict__ order, const RoiGeom* __restrict__ geoms, int K,
                        float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PxSmem& S = *reinterpret_cast<PxSmem*>(smem_raw);
    const int roi = order ? order[blockIdx.x] : blockIdx.x;
    const int lane = threadIdx.x, row = threadIdx.y, tid = row * 32 + lane;
    const int C = L.C;
    const int chunk0 = blockIdx.y * 256;
    const RoiGeom g = geoms[roi];
    const int H = L.H[g.level], W = L.W[g.level];

#ifdef RSDET_PROF
    long long tk0 = clock64(), tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0;
#endif
    if (tid == 0) { S.box[0] = 0x7fffffff; S.box[1] = -1; S.box[2] = 0x7fffffff; S.box[3] = -1; }
    __syncthreads();
    {
        int x0 = 0x7fffffff, x1 = -1, y0 = 0x7fffffff, y1 = -1;
        if (tid < kPxBins * 4) px_sample(g, L.version, H, W, tid, S.u.b.taps, x0, x1, y0, y1);
        x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);     // REDUX: off the LSU data pipe
        y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
        if (lane == 0 && x1 >= 0) { atomicMin(&S.box[0], x0); atomicMax(&S.box[1], x1); atomicMin(&S.box[2], y0); atomicMax(&S.box[3], y1); }
    }
    __syncthreads();
#ifdef RSDET_PROF
    tk1 = clock64();
#endif
    const float* __restrict__ feat_img = L.feat[g.level] + (size_t)g.batch * H * W * C + chunk0;
    px_build_row(S.u.b.taps, S.u.b.row[row], S.lists, row, lane, W, S.box);
#ifdef RSDET_PROF
    tk2 = clock64();
#endif
    __syncthreads();   // the build scratch becomes the staging block
#ifdef RSDET_PROF
    tk3 = clock64();
#endif
    px_gather_row<PB>(S.lists, feat_img + lane * 4, (unsigned)C * 4u, S.u.stage, row, lane);
#ifdef RSDET_PROF
    tk4 = clock64();
#endif
    // the staged block IS the RoI's output block: one bulk copy shared -> global through the async proxy (TMA),
    // which keeps the 50 KB read-out and the stores off the LSU data pipe the gather is bound by
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
#ifdef RSDET_PROF
    const long long tk5 = clock64();
#endif
    if (tid == 0) {
        float* dst = out + ((size_t)roi * C + chunk0) * kPxBins;            // 256 * 49 floats: 16-byte aligned
        bulk_store_evict_first(dst, S.u.stage, 256u * kPxBins * 4u);
    }
#ifdef RSDET_PROF
    if (g_px_prof && lane == 0) {   // per-phase cycle sums over all row warps (tools/roi_sweep.py --prof)
        atomicAdd(&g_px_prof[0], (unsigned long long)(tk1 - tk0)); atomicAdd(&g_px_prof[1], (unsigned long long)(tk2 - tk1));
        atomicAdd(&g_px_prof[2], (unsigned long long)(tk3 - tk2)); atomicAdd(&g_px_prof[3], (unsigned long long)(tk4 - tk3));
        atomicAdd(&g_px_prof[4], (unsigned long long)(tk5 - tk4)); atomicAdd(&g_px_prof[5], (unsigned long long)(clock64() - tk5));
        atomicAdd(&g_px_prof[6], 1ull);
    }
#endif
}

// ---- persistent, warp-specialised form (the default).  Each CTA loops over RoIs i = blockIdx.x, + gridDim.x, ...
// of the locality order with three roles that only meet at mbarriers:
//   builders  (warps 0-7):  sampling grid + window lists (one warp per bin row) of RoI i+1 into the other list buffer, while
//   gatherers (warps 8-14): one warp per bin row stream RoI i's pixels (nothing but loads + FMAs + staging stores),
//   storer    (warp 15):    hands the finished 50 KB block to the TMA (bulk copy shared -> global) and frees it.
// Register budget is rebalanced with setmaxnreg (builders 48, gather group 80 per thread).  In the one-CTA-per-RoI
// kernel above the list construction (~40 % of a CTA's life) overlaps other CTAs' gathers only by chance; here
// the memory pipe of an SM always has its gather warps issuing.
constexpr int kWsBuilders = 8, kWsGatherWarps = 8;   // 7 row builders (+1 that only helps with the sampling grid); gather group = 7 row warps + the storer
constexpr int kWsWarps = kWsBuilders + kWsGatherWarps;
constexpr int kWsRegsLaunch = 64, kWsRegsBuild = 48, kWsRegsGather = 80;  // 256*48 + 256*80 = 512*64

struct WsSmem {
    PxLists lists[2];
    const float* feat[2];                            // image base of the RoI in buffer b (level, batch, chunk applied)
    int box[4];
    int pad_[2];
    unsigned long long full[2], empty[2], stage_full, stage_free;   // mbarriers
    PxTaps taps;
    PxRowScratch row[kPxRows];
    alignas(16) float stage[256 * kPxBins];
};

template <int PB>
__global__ void __block_size__((32, kWsWarps, 1)) __maxnreg__(kWsRegsLaunch)
roi_align_fwd_ws_kernel(LevelSet L, const int* __restrict__ order, const RoiGeom* __restrict__ geoms, int K,
                        float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem& S = *reinterpret_cast<WsSmem*>(smem_raw);
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int C = L.C;
    const int chunk0 = blockIdx.y * 256;
    if (warp == 0 && lane == 0) {
        ws_mbar_init(&S.full[0], kPxRows); ws_mbar_init(&S.full[1], kPxRows);
        ws_mbar_init(&S.empty[0], kPxRows); ws_mbar_init(&S.empty[1], kPxRows);
        ws_mbar_init(&S.stage_full, kPxRows); ws_mbar_init(&S.stage_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, step = gridDim.x;

    if (warp < kWsBuilders) {
        // ------------------------------------------------------------------ builders
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsRegsBuild));
        const int btid = warp * 32 + lane;
        int it = 0;
        for (int i = first; i < K; i += step, it++) {
            const int buf = it & 1;
            const int roi = order ? order[i] : i;
            const RoiGeom g = geoms[roi];
            const int H = L.H[g.level], W = L.W[g.level];
            if (it >= 2) ws_mbar_wait(&S.empty[buf], ((it >> 1) - 1) & 1, 500);   // the gatherers are done with this buffer
            const float* __restrict__ feat_img = L.feat[g.level] + (size_t)g.batch * H * W * C + chunk0;
            if (btid == 0) {
                S.box[0] = 0x7fffffff; S.box[1] = -1; S.box[2] = 0x7fffffff; S.box[3] = -1;
                S.feat[buf] = feat_img;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kWsBuilders * 32) : "memory");   // also: previous RoI's rows have read the taps
            {
                int x0 = 0x7fffffff, x1 = -1, y0 = 0x7fffffff, y1 = -1;
                for (int sidx = btid; sidx < kPxBins * 4; sidx += kWsBuilders * 32) px_sample(g, L.version, H, W, sidx, S.taps, x0, x1, y0, y1);
                x0 = __reduce_min_sync(0xffffffffu, x0); x1 = __reduce_max_sync(0xffffffffu, x1);
                y0 = __reduce_min_sync(0xffffffffu, y0); y1 = __reduce_max_sync(0xffffffffu, y1);
                if (lane == 0 && x1 >= 0) { atomicMin(&S.box[0], x0); atomicMax(&S.box[1], x1); atomicMin(&S.box[2], y0); atomicMax(&S.box[3], y1); }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kWsBuilders * 32) : "memory");
            if (warp < kPxRows) {
                px_build_row(S.taps, S.row[warp], S.lists[buf], warp, lane, W, S.box);
                __syncwarp();
                if (lane == 0) ws_mbar_arrive(&S.full[buf]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsRegsGather));
        const int row = warp - kWsBuilders;
        if (row < kPxRows) {
            // -------------------------------------------------------------- gatherers
            int it = 0;
            for (int i = first; i < K; i += step, it++) {
                const int buf = it & 1;
                ws_mbar_wait(&S.full[buf], (it >> 1) & 1, 100);
                const float* __restrict__ feat = S.feat[buf] + lane * 4;
                px_gather_row<PB>(S.lists[buf], feat, (unsigned)C * 4u, S.stage, row, lane, it >= 1 ? &S.stage_free : nullptr,
                                  (unsigned)((it - 1) & 1));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) { ws_mbar_arrive(&S.empty[buf]); ws_mbar_arrive(&S.stage_full); }
            }
        } else if (lane == 0) {
            // -------------------------------------------------------------- storer
            int it = 0;
            for (int i = first; i < K; i += step, it++) {
                ws_mbar_wait(&S.stage_full, it & 1, 200);
                const int roi = order ? order[i] : i;
                float* dst = out + ((size_t)roi * C + chunk0) * kPxBins;
                bulk_store_evict_first(dst, S.stage, 256u * kPxBins * 4u);   // returns once the block has been read
                ws_mbar_arrive(&S.stage_free);
            }
        }
    }
}

#endif  // RSDET_TUNING (row-window kernels)

// cudaFuncSetAttribute is per device: remember what was set for each one (a process may drive several GPUs)
static void set_dyn_smem(const void* func, size_t bytes) {
    struct Slot { const void* f; size_t b; };
    static Slot slots[64][16];
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64) { cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); return; }
    for (int i = 0; i < 16; i++) {
        Slot& s = slots[dev][i];
        if (s.f == func || s.f == nullptr) {
            if (s.f == func && s.b >= bytes) return;
            s.f = func; s.b = bytes;
            cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            return;
        }
    }
    cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Forward path: 1 = bin-major register kernel + TMA bulk store (the default and the only path of the shipped library:
// fastest measured, profiles/README.md "round 2"), and -- in RSDET_TUNING builds only, selected through RSDET_ROI_PATH
// for A/B measurements -- 0 = row windows, persistent warp-specialised, 3 = row windows, one CTA per RoI, 2 = TMA gather4.
static int roi_path_choice() {
#ifdef RSDET_TUNING
    if (const char* e = getenv("RSDET_ROI_PATH")) return atoi(e);
#endif
    return 1;
}

#ifdef RSDET_TUNING
static bool px_path_ok(const rsdet_roi_align_cfg* c) {
    if (c->pooled_h != kPxRows || c->pooled_w != kPxCols || c->sampling_ratio != 2 || c->channels % 256 != 0) return false;
    for (int l = 0; l < c->num_levels; l++)   // pixel keys are (y << 16 | x); pixel indices 32-bit
        if (c->height[l] >= 32768 || c->width[l] >= 65536 || (long long)c->height[l] * c->width[l] >= (1ll << 31)) return false;
    return true;
}
#endif

// ---------------------------------------------------------------------------------- forward (TMA gather4)
// Same decomposition (one CTA per RoI, merged tap lists, warp = bin group), but the feature rows are no
// longer pulled through the LSU into registers: each group of 4 merged taps is ONE Blackwell
// `cp.async.bulk.tensor.2d ... tile::gather4` (4 arbitrary pixel rows x 256 channels = 4 KB) issued by a
// single lane into a per-warp ring in shared memory and tracked by an mbarrier.  What this buys: bytes in
// flight are bounded by shared memory (8 warps x 3 stages x 4 KB = 96 KB per CTA, 2 CTAs per SM) instead
// of by the register file (8 x 16 B per thread), so the L2 round trips (~1 us loaded) overlap instead of
// serialising; address generation and the 784 x 64 vector loads per RoI leave the instruction stream.
// Results stay in registers until the ring is idle, then the ring is reused as the [c][bin] staging area.
constexpr int kTmaStages = 3;
#ifndef RSDET_BULK_ROWS
#define RSDET_BULK_ROWS 0
#endif
constexpr bool kBulkRows = RSDET_BULK_ROWS != 0;
constexpr int kTmaMaxSlots = 8;  // bins per warp (ceil(nbins / 8) <= 8 -> nbins <= 64)

struct TmaMaps { CUtensorMap m[RSDET_MAX_LEVELS]; };

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kRoiThreads, 2)
roi_align_fwd_tma_kernel(LevelSet L, const __grid_constant__ TmaMaps maps, const float* __restrict__ rois,
                         const int* __restrict__ order, int K, float* __restrict__ out, int32_t* __restrict__ levels_out) {
    // (aligned by hand: an __align__(1024) on the extern array would pad EVERY kernel of this translation unit
    //  by 1 KB of static shared memory, which costs the register path its fourth CTA per SM)
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int roi = order ? order[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int cap = 4 * spb;
    const int C = L.C;                                     // == 256 * gridDim.y
    const int chunk0 = blockIdx.y * 256;
    // smem: [ring 8 warps x kTmaStages x 4 KB | staging [256][nbins] | phase-A scratch][lists][counts][mbarriers]
    const size_t ring_bytes = (size_t)8 * kTmaStages * 4096;
    const size_t big = max(ring_bytes, max((size_t)256 * nbins * 4, (size_t)12 * nbins * (cap + 1)));
    float* s_stage = reinterpret_cast<float*>(smem_raw);
    int2* s_list = reinterpret_cast<int2*>(smem_raw + big);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + big + list_bytes(nbins, cap));
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw + big + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~15));
    __shared__ RoiGeom s_g;

    if (tid == 0) {
        s_g = roi_geometry(rois + (size_t)roi * 6, L);
        if (levels_out && !order && blockIdx.y == 0) levels_out[roi] = s_g.level;
    }
    if (tid < 8 * kTmaStages) mbar_init(s_bar + tid, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const RoiGeom g = s_g;
    const int H = L.H[g.level], W = L.W[g.level];
    build_tap_lists(g, L, H, W, s_list, s_cnt, s_stage, 1, g.batch * H * W);  // entries = global pixel-row index
    if (L.dbg_skip_main) { for (int b = tid; b < nbins; b += kRoiThreads) s_cnt[b] = 0; __syncthreads(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // scratch (generic writes) -> TMA writes

    const CUtensorMap* map = &maps.m[g.level];
    unsigned char* ring = smem_raw + (size_t)warp * kTmaStages * 4096;
    unsigned long long* bar = s_bar + warp * kTmaStages;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)max(spb, 1), inv_count = 1.f / count;

    // issue cursor (warp-uniform): next (bin, first entry) to fetch
    int ib = warp, ie = 0, istage = 0, inflight = 0;
    auto issue = [&]() {
        while (ib < nbins && ie >= s_cnt[ib]) { ib += 8; ie = 0; }
        if (ib >= nbins) return;
        if (kBulkRows) {
            // variant: four 1-D bulk copies (one per pixel row) instead of one gather4
            const int cnt = s_cnt[ib];
            const int nrow = min(4, cnt - ie);
            if (lane == 0) mbar_expect_tx(bar + istage, 1024u * nrow);
            __syncwarp();
            if (lane < nrow) {
                const int r = s_list[ib * (cap + 1) + ie + lane].x;
                const float* src = L.feat[g.level] + (size_t)r * C + chunk0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(ring + istage * 4096 + lane * 1024)), "l"(src), "r"(1024u), "r"(smem_u32(bar + istage)) : "memory");
            }
        } else if (lane == 0) {
            const int cnt = s_cnt[ib];
            const int2* lp = s_list + ib * (cap + 1);
            const int r0 = lp[ie].x, r1 = lp[min(ie + 1, cnt - 1)].x, r2 = lp[min(ie + 2, cnt - 1)].x, r3 = lp[min(ie + 3, cnt - 1)].x;
            mbar_expect_tx(bar + istage, 4096u);
            tma_gather4(ring + istage * 4096, map, chunk0, r0, r1, r2, r3, bar + istage);
        }
        ie += 4;
        istage = istage + 1 == kTmaStages ? 0 : istage + 1;
        inflight++;
    };
#pragma unroll
    for (int d = 0; d < kTmaStages; d++) issue();

    float4 res[kTmaMaxSlots][2];
    int cstage = 0;
    unsigned phase = 0;  // bit s = parity to wait for on stage s
#pragma unroll
    for (int slot = 0; slot < kTmaMaxSlots; slot++) {
        const int b = warp + 8 * slot;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (b < nbins) {
            const int cnt = s_cnt[b];
            const int2* lp = s_list + b * (cap + 1);
            for (int e = 0; e < cnt; e += 4) {
                float wt[4];
#pragma unroll
                for (int k = 0; k < 4; k++) wt[k] = e + k < cnt ? __int_as_float(lp[e + k].y) : 0.f;
                mbar_wait(bar + cstage, (phase >> cstage) & 1u);
                const float4* row = reinterpret_cast<const float4*>(ring + cstage * 4096) + lane;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (kBulkRows && e + k >= cnt) continue;  // that row was not fetched
                    const float4 v0 = row[k * 64], v1 = row[k * 64 + 32];
                    a0.x = fmaf(wt[k], v0.x, a0.x); a0.y = fmaf(wt[k], v0.y, a0.y); a0.z = fmaf(wt[k], v0.z, a0.z); a0.w = fmaf(wt[k], v0.w, a0.w);
                    a1.x = fmaf(wt[k], v1.x, a1.x); a1.y = fmaf(wt[k], v1.y, a1.y); a1.z = fmaf(wt[k], v1.z, a1.z); a1.w = fmaf(wt[k], v1.w, a1.w);
                }
                phase ^= 1u << cstage;
                cstage = cstage + 1 == kTmaStages ? 0 : cstage + 1;
                inflight--;
                __syncwarp();  // every lane has read the stage before lane 0 hands it back to the TMA unit
                issue();
            }
            if (pow2) { a0.x *= inv_count; a0.y *= inv_count; a0.z *= inv_count; a0.w *= inv_count;
                        a1.x *= inv_count; a1.y *= inv_count; a1.z *= inv_count; a1.w *= inv_count; }
            else { a0.x /= count; a0.y /= count; a0.z /= count; a0.w /= count;
                   a1.x /= count; a1.y /= count; a1.z /= count; a1.w /= count; }
        }
        res[slot][0] = a0;
        res[slot][1] = a1;
    }
    __syncthreads();  // all rings idle: reuse them as the staging area
#pragma unroll
    for (int slot = 0; slot < kTmaMaxSlots; slot++) {
        const int b = warp + 8 * slot;
        if (b < nbins) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int c0 = (lane + u * 32) * 4;
                s_stage[(c0 + 0) * nbins + b] = res[slot][u].x;
                s_stage[(c0 + 1) * nbins + b] = res[slot][u].y;
                s_stage[(c0 + 2) * nbins + b] = res[slot][u].z;
                s_stage[(c0 + 3) * nbins + b] = res[slot][u].w;
            }
        }
    }
    __syncthreads();
    float* __restrict__ dst = out + ((size_t)roi * C + chunk0) * nbins;
    const int total = 256 * nbins;
    if ((((size_t)roi * C + chunk0) * nbins & 3) == 0 && (total & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(s_stage);
        for (int e = tid; e < total / 4; e += kRoiThreads) stg_cs_v4(dst + (size_t)e * 4, s4[e]);
    } else {
        for (int e = tid; e < total; e += kRoiThreads) __stcs(dst + e, s_stage[e]);
    }
}

// ---------------------------------------------------------------------------------- backward (fast)
// Same CTA shape and tap lists.  The RoI's (C,7,7) gradient block is copied into shared memory as it
// lies in memory (coalesced 16-byte loads); every merged tap becomes ONE 16-byte vector reduction per
// channel quad (red.global.add.v4.f32) into the channels-last accumulator.
template <int QPT>
__global__ void __launch_bounds__(kRoiThreads, 4)
roi_align_bwd_kernel(LevelSet L, const float* __restrict__ rois, const int* __restrict__ order, const RoiGeom* __restrict__ geoms,
                     int K, const float* __restrict__ grad_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int roi = order ? order[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int cap = 4 * spb;
    const int C = L.C;
    const int Q = quads_per_chunk(C);
    const int chunk0 = blockIdx.y * Q * 4;
    const int Qc = min(Q, (C - chunk0) / 4);
    int2* s_list = reinterpret_cast<int2*>(smem_raw);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + list_bytes(nbins, cap));
    float* s_stage = reinterpret_cast<float*>(smem_raw + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~15));
    __shared__ RoiGeom s_g;

    if (!geoms) {
        if (tid == 0) s_g = roi_geometry(rois + (size_t)roi * 6, L);
        __syncthreads();
    }
    const RoiGeom g = geoms ? geoms[roi] : s_g;
    const int H = L.H[g.level], W = L.W[g.level];
    build_tap_lists(g, L, H, W, s_list, s_cnt, s_stage, C >> 2, 0);

    // stage this chunk's gradient block [c_local][bin] (= memory order)
    const float* __restrict__ src = grad_out + ((size_t)roi * C + chunk0) * nbins;
    const int total = Qc * 4 * nbins;
    if ((((size_t)roi * C + chunk0) * nbins & 3) == 0 && (total & 3) == 0) {
        float4* s4 = reinterpret_cast<float4*>(s_stage);
        for (int e = tid; e < total / 4; e += kRoiThreads) s4[e] = ldg_cs_v4(src + (size_t)e * 4);
    } else {
        for (int e = tid; e < total; e += kRoiThreads) s_stage[e] = __ldg(src + e);
    }
    __syncthreads();

    const int lanes = QPT == 2 ? 32 : Qc;
    const int groups = kRoiThreads / lanes;
    const int cq = tid % lanes, grp = tid / lanes;
    const int oct = (cq >> 3) & 3;
    float4* __restrict__ gfeat = reinterpret_cast<float4*>(L.grad[g.level] + (size_t)g.batch * H * W * C + chunk0) + cq;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)spb, inv_count = 1.f / count;  // no max(.,1) in backward (:246)
    if (grp < groups) {
        for (int b = grp; b < nbins; b += groups) {
            float4 top[QPT];
#pragma unroll
            for (int u = 0; u < QPT; u++) {
                const int c0 = (cq + u * 32) * 4;
                float4 r;
                r.x = s_stage[(c0 + ((0 + oct) & 3)) * nbins + b];
                r.y = s_stage[(c0 + ((1 + oct) & 3)) * nbins + b];
                r.z = s_stage[(c0 + ((2 + oct) & 3)) * nbins + b];
                r.w = s_stage[(c0 + ((3 + oct) & 3)) * nbins + b];
                top[u] = unrot4(r, oct);
            }
            const int2* lp = s_list + b * (cap + 1);
            const int cnt = s_cnt[b];
            for (int e = 0; e < cnt; e++) {
                const int2 en = lp[e];
                const float wk = __int_as_float(en.y);
                float* dst = const_cast<float*>(tap_ptr(reinterpret_cast<const float4*>(gfeat), (unsigned)en.x));
#pragma unroll
                for (int u = 0; u < QPT; u++) {
                    // g = top * w / count (:280-283); exact reciprocal when count is a power of two
                    float4 gv;
                    if (pow2) gv = make_float4(top[u].x * wk * inv_count, top[u].y * wk * inv_count, top[u].z * wk * inv_count, top[u].w * wk * inv_count);
                    else gv = make_float4(top[u].x * wk / count, top[u].y * wk / count, top[u].z * wk / count, top[u].w * wk / count);
                    red_add_v4(dst + u * 128, gv);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------- generic path
// Any pooled size / adaptive sampling grid / channel count; NCHW or NHWC directly.  One thread per
// output element like the reference, but with the level mapping fused.  Used only when the fast path's
// preconditions (C % 4 == 0, fixed sampling grid, <= kMaxSamples samples) do not hold.
struct GenericMaps {
    const float* feat[RSDET_MAX_LEVELS];
    float* grad[RSDET_MAX_LEVELS];
    int channels_last;
};

__device__ __forceinline__ size_t feat_index(int cl, int n, int c, int pix, int C, int HW) {
    return cl ? ((size_t)n * HW + pix) * C + c : ((size_t)n * C + c) * HW + pix;
}

template <bool BACKWARD>
__global__ void roi_align_generic_kernel(LevelSet L, GenericMaps M, const float* __restrict__ rois, int K,
                                         float* __restrict__ out_or_gradout, int32_t* __restrict__ levels_out) {
    const int nbins = L.PH * L.PW;
    const long long total = (long long)K * L.C * nbins;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int pw = (int)(idx % L.PW), ph = (int)((idx / L.PW) % L.PH);
        int c = (int)((idx / nbins) % L.C), n = (int)(idx / nbins / L.C);
        RoiGeom g = roi_geometry(rois + (size_t)n * 6, L);
        if (!BACKWARD && levels_out && c == 0 && ph == 0 && pw == 0) levels_out[n] = g.level;
        const int H = L.H[g.level], W = L.W[g.level];
        if (!BACKWARD) {
            const float count = (float)max(g.gh * g.gw, 1);
            float acc = 0.f;
            for (int iy = 0; iy < g.gh; iy++)
                for (int ix = 0; ix < g.gw; ix++) {
                    float x, y;
                    sample_xy(g, L.version, ph, pw, iy, ix, x, y);
                    Taps t = make_taps(H, W, y, x);
                    const float* f = M.feat[g.level];
                    float v = t.w[0] * f[feat_index(M.channels_last, g.batch, c, t.o[0], L.C, H * W)] +
                              t.w[1] * f[feat_index(M.channels_last, g.batch, c, t.o[1], L.C, H * W)] +
                              t.w[2] * f[feat_index(M.channels_last, g.batch, c, t.o[2], L.C, H * W)] +
                              t.w[3] * f[feat_index(M.channels_last, g.batch, c, t.o[3], L.C, H * W)];
                    acc += v;
                }
            out_or_gradout[idx] = acc / count;
        } else {
            const float count = (float)(g.gh * g.gw);
            const float top = out_or_gradout[idx];
            for (int iy = 0; iy < g.gh; iy++)
                for (int ix = 0; ix < g.gw; ix++) {
                    float x, y;
                    sample_xy(g, L.version, ph, pw, iy, ix, x, y);
                    Taps t = make_taps(H, W, y, x);
                    if (t.w[0] == 0.f && t.w[1] == 0.f && t.w[2] == 0.f && t.w[3] == 0.f) continue;
                    float* gf = M.grad[g.level];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        atomicAdd(gf + feat_index(M.channels_last, g.batch, c, t.o[k], L.C, H * W), top * t.w[k] / count);
                }
        }
    }
}

static bool fast_path_ok(const rsdet_roi_align_cfg* c) {
    for (int l = 0; l < c->num_levels; l++)  // tap offsets are 32-bit element offsets inside one image
        if ((long long)c->height[l] * c->width[l] * c->channels >= (1ll << 31)) return false;
    return c->channels % 4 == 0 && c->sampling_ratio > 0 &&
           c->pooled_h * c->pooled_w * c->sampling_ratio * c->sampling_ratio <= kMaxSamples;
}

static int check_cfg(const rsdet_roi_align_cfg* c) {
    if (!c) return RSDET_EINVAL;
    if (c->num_levels < 1 || c->num_levels > RSDET_MAX_LEVELS || c->batch < 1 || c->channels < 1) return RSDET_EINVAL;
    if (c->pooled_h < 1 || c->pooled_w < 1 || (c->version != 0 && c->version != 1)) return RSDET_EINVAL;
    for (int l = 0; l < c->num_levels; l++)
        if (c->height[l] < 1 || c->width[l] < 1) return RSDET_EINVAL;
    return RSDET_OK;
}

static LevelSet make_levels(const rsdet_roi_align_cfg* c) {
    LevelSet L;
    L.num_levels = c->num_levels; L.batch = c->batch; L.C = c->channels;
    L.PH = c->pooled_h; L.PW = c->pooled_w; L.sampling_ratio = c->sampling_ratio; L.version = c->version;
    L.extend_w = c->extend_w; L.extend_h = c->extend_h; L.finest_scale = c->finest_scale;
    L.dbg_skip_main = 0;
#ifdef RSDET_TUNING
    if (const char* e = getenv("RSDET_ROI_DBG_SKIP_MAIN")) L.dbg_skip_main = atoi(e);
#endif
    for (int l = 0; l < RSDET_MAX_LEVELS; l++) {
        L.feat[l] = nullptr; L.grad[l] = nullptr;
        L.H[l] = l < c->num_levels ? c->height[l] : 1;
        L.W[l] = l < c->num_levels ? c->width[l] : 1;
        L.scale[l] = l < c->num_levels ? c->spatial_scale[l] : 1.f;
    }
    return L;
}

static size_t fwd_smem_bytes(const rsdet_roi_align_cfg* c) {
    size_t nbins = (size_t)c->pooled_h * c->pooled_w;
    size_t ntaps = nbins * 4 * c->sampling_ratio * c->sampling_ratio;
    size_t stage = sizeof(float) * 4 * nbins * quads_per_chunk(c->channels), tmp = 12 * (ntaps + nbins);
    return list_bytes(nbins, ntaps / nbins) + ((nbins * 4 + 15) & ~(size_t)15) + (stage > tmp ? stage : tmp);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (librsdet links no libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D view [N*H*W pixel rows][C] of a channels-last map; box = one row of 256 channels (gather4 fetches 4 rows)
static bool make_row_map(CUtensorMap* m, const float* base, long long rows, int C) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
    cuuint32_t box[2] = {256u, 1u};
    cuuint32_t estr[2] = {1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static size_t tma_smem_bytes(const rsdet_roi_align_cfg* c) {
    size_t nbins = (size_t)c->pooled_h * c->pooled_w, cap = 4 * (size_t)c->sampling_ratio * c->sampling_ratio;
    size_t big = (size_t)8 * kTmaStages * 4096;
    if (256 * nbins * 4 > big) big = 256 * nbins * 4;
    if (12 * nbins * (cap + 1) > big) big = 12 * nbins * (cap + 1);
    return big + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~(size_t)15) + 8 * 8 * kTmaStages + 64 + 1024;
}

static bool tma_path_ok(const rsdet_roi_align_cfg* c) {
    // Opt-in (RSDET_ROI_TMA=1): measured 308 us per 4000-RoI tile against 200 us for the register path on
    // B200 (profiles/README.md, "TMA gather4 experiment") -- correct, but two 105 KB CTAs per SM expose the
    // tap-list and write-out phases that four 57 KB CTAs overlap.  Kept for the round-2 persistent-CTA rework.
    if (roi_path_choice() != 2) return false;
    if (c->channels % 256 != 0 || c->pooled_h * c->pooled_w > 8 * kTmaMaxSlots) return false;
    for (int l = 0; l < c->num_levels; l++)
        if ((long long)c->batch * c->height[l] * c->width[l] >= (1ll << 31)) return false;
    return tma_smem_bytes(c) <= 110 * 1024 && encode_tiled_fn() != nullptr;
}

}  // namespace rsdet

using namespace rsdet;

extern "C" size_t rsdet_roi_align_rotated_workspace_bytes(const rsdet_roi_align_cfg* cfg, int num_rois, int backward) {
    (void)backward;
    if (check_cfg(cfg) != RSDET_OK) return 0;
    size_t b = ws_bytes<int>(num_rois > 0 ? num_rois : 1) + ws_bytes<RoiGeom>(num_rois > 0 ? num_rois : 1);  // order, geometry
    if (cfg->channels_last || !fast_path_ok(cfg)) return b + 256;
    for (int l = 0; l < cfg->num_levels; l++)
        b += ws_bytes<float>((size_t)cfg->batch * cfg->channels * cfg->height[l] * cfg->width[l]);
    return b + 256;
}

extern "C" int rsdet_roi_align_rotated_forward(const rsdet_roi_align_cfg* cfg, const float* const* feats_host, const float* rois,
                                               int num_rois, float* out, int32_t* levels_out, void* workspace,
                                               size_t workspace_bytes, void* stream) {
    int rc = check_cfg(cfg);
    if (rc != RSDET_OK) return rc;
    if (num_rois < 0) return RSDET_EINVAL;
    if (num_rois == 0) return RSDET_OK;
    if (!feats_host || !rois || !out) return RSDET_EINVAL;
    for (int l = 0; l < cfg->num_levels; l++)
        if (!feats_host[l]) return RSDET_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    LevelSet L = make_levels(cfg);
    if (!fast_path_ok(cfg)) {
        GenericMaps M;
        M.channels_last = cfg->channels_last;
        for (int l = 0; l < RSDET_MAX_LEVELS; l++) { M.feat[l] = l < cfg->num_levels ? feats_host[l] : nullptr; M.grad[l] = nullptr; }
        long long total = (long long)num_rois * cfg->channels * cfg->pooled_h * cfg->pooled_w;
        int grid = (int)(ceil_div_ll(total, 256) < (long long)kNumSMs * 16 ? ceil_div_ll(total, 256) : (long long)kNumSMs * 16);
        roi_align_generic_kernel<false><<<grid, 256, 0, st>>>(L, M, rois, num_rois, out, levels_out);
        count_launch();
        return cuda_status();
    }
    if (workspace_bytes < rsdet_roi_align_rotated_workspace_bytes(cfg, num_rois, 0)) return RSDET_EWORKSPACE;
    Workspace ws(workspace, workspace_bytes);
    int* order_ws = ws.take<int>(num_rois);
    RoiGeom* geoms = ws.take<RoiGeom>(num_rois);
    if (cfg->channels_last) {
        for (int l = 0; l < cfg->num_levels; l++) L.feat[l] = feats_host[l];
    } else {
        float* dst[RSDET_MAX_LEVELS];
        for (int l = 0; l < cfg->num_levels; l++) {
            dst[l] = ws.take<float>((size_t)cfg->batch * cfg->channels * cfg->height[l] * cfg->width[l]);
            L.feat[l] = dst[l];
        }
        if (!ws.ok()) return RSDET_EWORKSPACE;
    }
    // locality order (needs the int[K] slot at the start of the workspace; skipped for tiny calls)
    int* order = num_rois >= 256 && order_ws ? order_ws : nullptr;
    const bool ride = !cfg->channels_last && order && num_rois <= kPrepMaxRois;
    launch_prep(L, rois, num_rois, geoms, order, levels_out, st, ride);
    if (!cfg->channels_last) {
        // NCHW -> NHWC copy of the pyramid; the locality-order block rides along as block 0 of the same launch
        const PrepArgs prep = {rois, num_rois, geoms, order};
        float* dst[RSDET_MAX_LEVELS];
        for (int l = 0; l < cfg->num_levels; l++) dst[l] = const_cast<float*>(L.feat[l]);
        rc = launch_transpose(true, feats_host, dst, cfg->height, cfg->width, cfg->num_levels, cfg->batch, cfg->channels, st,
                              ride ? &prep : nullptr);
        if (rc != RSDET_OK) return rc;
    }
    size_t smem = fwd_smem_bytes(cfg);
    set_dyn_smem((const void*)roi_align_fwd_kernel<1>, smem);
    set_dyn_smem((const void*)roi_align_fwd_kernel<2>, smem);
#ifdef RSDET_TUNING
    if (px_path_ok(cfg) && (roi_path_choice() == 0 || roi_path_choice() == 3)) {
        const int chunks = cfg->channels / 256;
        if (roi_path_choice() == 3) {   // one CTA per RoI
            dim3 pgrid(num_rois, chunks), pblock(32, kPxRows, 1);
            set_dyn_smem((const void*)roi_align_fwd_px_kernel<4>, sizeof(PxSmem));
            roi_align_fwd_px_kernel<4><<<pgrid, pblock, sizeof(PxSmem), st>>>(L, order, geoms, num_rois, out);
        } else {                        // persistent, warp-specialised: two CTAs per SM
            const int ctas = num_rois < 2 * kNumSMs ? num_rois : 2 * kNumSMs;
            dim3 pgrid(ctas, chunks), pblock(32, kWsWarps, 1);
            set_dyn_smem((const void*)roi_align_fwd_ws_kernel<4>, sizeof(WsSmem));
            roi_align_fwd_ws_kernel<4><<<pgrid, pblock, sizeof(WsSmem), st>>>(L, order, geoms, num_rois, out);
        }
        count_launch();
        return cuda_status();
    }
#endif
    if (tma_path_ok(cfg)) {
        TmaMaps maps;
        bool ok = true;
        for (int l = 0; l < cfg->num_levels && ok; l++)
            ok = make_row_map(&maps.m[l], L.feat[l], (long long)cfg->batch * cfg->height[l] * cfg->width[l], cfg->channels);
        for (int l = cfg->num_levels; l < RSDET_MAX_LEVELS; l++) maps.m[l] = maps.m[0];
        if (ok) {
            const size_t tsmem = tma_smem_bytes(cfg);
            set_dyn_smem((const void*)roi_align_fwd_tma_kernel, tsmem);
            dim3 tgrid(num_rois, cfg->channels / 256);
            roi_align_fwd_tma_kernel<<<tgrid, kRoiThreads, tsmem, st>>>(L, maps, rois, order, num_rois, out, levels_out);
            count_launch();
            return cuda_status();
        }
    }
    const int Q = quads_per_chunk(cfg->channels);
    dim3 grid(num_rois, ceil_div(cfg->channels / 4, Q));
    if (cfg->channels % 256 == 0)
        roi_align_fwd_kernel<2><<<grid, kRoiThreads, smem, st>>>(L, rois, order, geoms, num_rois, out, nullptr);
    else
        roi_align_fwd_kernel<1><<<grid, kRoiThreads, smem, st>>>(L, rois, order, geoms, num_rois, out, nullptr);
    count_launch();
    return cuda_status();
}

extern "C" int rsdet_roi_align_rotated_backward(const rsdet_roi_align_cfg* cfg, const float* grad_out, const float* rois,
                                                int num_rois, float* const* grad_feats_host, void* workspace,
                                                size_t workspace_bytes, void* stream) {
    int rc = check_cfg(cfg);
    if (rc != RSDET_OK) return rc;
    if (num_rois < 0 || !grad_feats_host) return RSDET_EINVAL;
    for (int l = 0; l < cfg->num_levels; l++)
        if (!grad_feats_host[l]) return RSDET_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    LevelSet L = make_levels(cfg);
    const bool fast = fast_path_ok(cfg);
    const bool direct = cfg->channels_last || !fast;  // accumulate straight into the caller's buffers
    float* acc[RSDET_MAX_LEVELS];
    if (workspace_bytes < rsdet_roi_align_rotated_workspace_bytes(cfg, num_rois, 1)) return RSDET_EWORKSPACE;
    Workspace ws(workspace, workspace_bytes);
    int* order_ws = ws.take<int>(num_rois > 0 ? num_rois : 1);
    RoiGeom* geoms = ws.take<RoiGeom>(num_rois > 0 ? num_rois : 1);
    if (direct) {
        for (int l = 0; l < cfg->num_levels; l++) acc[l] = grad_feats_host[l];
    } else {
        for (int l = 0; l < cfg->num_levels; l++)
            acc[l] = ws.take<float>((size_t)cfg->batch * cfg->channels * cfg->height[l] * cfg->width[l]);
        if (!ws.ok()) return RSDET_EWORKSPACE;
    }
    for (int l = 0; l < cfg->num_levels; l++) {
        cudaMemsetAsync(acc[l], 0, sizeof(float) * (size_t)cfg->batch * cfg->channels * cfg->height[l] * cfg->width[l], st);
        L.grad[l] = acc[l];
    }
    count_launch(cfg->num_levels);
    if (num_rois > 0) {
        if (!grad_out || !rois) return RSDET_EINVAL;
        if (!fast) {
            GenericMaps M;
            M.channels_last = cfg->channels_last;
            for (int l = 0; l < RSDET_MAX_LEVELS; l++) { M.feat[l] = nullptr; M.grad[l] = l < cfg->num_levels ? acc[l] : nullptr; }
            long long total = (long long)num_rois * cfg->channels * cfg->pooled_h * cfg->pooled_w;
            int grid = (int)(ceil_div_ll(total, 256) < (long long)kNumSMs * 16 ? ceil_div_ll(total, 256) : (long long)kNumSMs * 16);
            roi_align_generic_kernel<true><<<grid, 256, 0, st>>>(L, M, rois, num_rois, const_cast<float*>(grad_out), nullptr);
        } else {
            size_t smem = fwd_smem_bytes(cfg);
            set_dyn_smem((const void*)roi_align_bwd_kernel<1>, smem);
            set_dyn_smem((const void*)roi_align_bwd_kernel<2>, smem);
            int* order = num_rois >= 256 ? order_ws : nullptr;
            launch_prep(L, rois, num_rois, geoms, order, nullptr, st);
            int Q = quads_per_chunk(cfg->channels);
            dim3 grid(num_rois, ceil_div(cfg->channels / 4, Q));
            if (cfg->channels % 256 == 0) roi_align_bwd_kernel<2><<<grid, kRoiThreads, smem, st>>>(L, rois, order, geoms, num_rois, grad_out);
            else roi_align_bwd_kernel<1><<<grid, kRoiThreads, smem, st>>>(L, rois, order, geoms, num_rois, grad_out);
        }
        count_launch();
    }
    if (!direct) {
        rc = launch_transpose(false, acc, grad_feats_host, cfg->height, cfg->width, cfg->num_levels, cfg->batch, cfg->channels, st);
        if (rc != RSDET_OK) return rc;
    }
    return cuda_status();
}

#ifdef RSDET_PROF
// profiling builds only: device buffer of 8 counters filled by roi_align_fwd_px_kernel (per-phase cycle sums)
extern "C" int rsdet_tuning_set_prof(unsigned long long* dev_counters) {
    return (int)cudaMemcpyToSymbol(g_px_prof, &dev_counters, sizeof(dev_counters));
}
#endif

extern "C" int rsdet_nchw_to_nhwc(const float* src, int n, int c, int h, int w, float* dst, void* stream) {
    if (!src || !dst || n < 1 || c < 1 || h < 1 || w < 1) return RSDET_EINVAL;
    const float* s[1] = {src};
    float* d[1] = {dst};
    return launch_transpose(true, s, d, &h, &w, 1, n, c, (cudaStream_t)stream);
}

extern "C" int rsdet_nhwc_to_nchw(const float* src, int n, int c, int h, int w, float* dst, void* stream) {
    if (!src || !dst || n < 1 || c < 1 || h < 1 || w < 1) return RSDET_EINVAL;
    const float* s[1] = {src};
    float* d[1] = {dst};
    return launch_transpose(false, s, d, &h, &w, 1, n, c, (cudaStream_t)stream);
}
