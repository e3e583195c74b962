// A/B-measurement kernels of the RoIAlignRotated forward gather and their host dispatch.  Included by roi_align.cu inside
// namespace rsdet in RSDET_TUNING builds only (RSDET_TUNING=1 python -m rs_detection_b200.build --force); the shipped
// library contains none of this.  Measurements: profiles/README.md "Round 2".
#pragma once

// ---------------------------------------------------------------------------------- forward (7x7, 256-bit loads)
// A/B only (RSDET_ROI_V8 = 1 / 2): measured 137.6 us (LDS.128 list reads alone) and 142.1 us (+ 256-bit loads) per tile against
// 138.4 us for roi_align_fwd77_kernel<7> in the same build -- neither is a gain, so the shipped kernel stays as it is.
// Same decomposition as roi_align_fwd77_kernel with two Blackwell-specific changes to the gather loop:
//   * V8: a tap is ONE 256-bit load per lane (ld.global.nc.v8.f32, SASS LDG.E.ENL2.256; sm_100+): lane l owns channels
//     8l..8l+7 and a warp instruction covers the pixel's whole 1 KB row -- half the LSU instructions of the 2 x LDG.128 form;
//   * the bin lists have a pitch of 18 entries (16-byte aligned), so the four entries of a batch are two LDS.128
//     broadcasts instead of four LDS.64, and the count sits in the pad entry (no s_cnt array).
// Staging: lane l = 4g + r stores component (j + g) & 7 of its eight channels at step j: word address (8l + m) * 49 + b
// falls in bank (8r + 17m + b) mod 32, which is distinct for all 32 (r, m) pairs of a warp.  Arithmetic and tap order are
// those of roi_align_fwd77_kernel; results are bit-identical.
constexpr size_t kStage77v8Offset = (8 * 49 * 18 + 127) & ~(size_t)127;
__device__ __forceinline__ void ldg_nc_v8(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
template <int WARPS, bool V8>
__global__ void __launch_bounds__(32 * WARPS, 4)
roi_align_fwd77v8_kernel(LevelSet L, const RoiGeom* __restrict__ gsorted, int K, float* __restrict__ out) {
    constexpr int NB = 49, PITCH = 18, THREADS = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int2* s_list = reinterpret_cast<int2*>(smem_raw);
    float* s_stage = reinterpret_cast<float*>(smem_raw + kStage77v8Offset);
    RoiGeom g = gsorted[blockIdx.x];
    const int roi = g.gh;                                   // processing-order record: gh carries the RoI index
    g.gh = 2; g.gw = 2;
    const int C = L.C, chunk0 = blockIdx.y * 256;
    const int H = L.H[g.level], W = L.W[g.level];
    build_tap_lists<THREADS, true, PITCH>(g, L, H, W, s_list, nullptr, s_stage, C >> 2, 0);

    // V8: lane owns 8 consecutive channels (32 bytes); else two quads 512 bytes apart as in roi_align_fwd77_kernel
    const float4* __restrict__ feat =
        reinterpret_cast<const float4*>(L.feat[g.level] + (size_t)g.batch * H * W * C + chunk0) + (V8 ? 2 * lane : lane);
    const int rotg = V8 ? (lane >> 2) : ((lane >> 3) & 3);
    auto stage = [&](float (&a)[8], int b) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] *= 0.25f;              // output_val /= count (:143), count = 4: exact
        if (V8) {
            float r[8];
            // r[j] = a[(j + rotg) & 7]
            if (rotg & 1) { const float t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = a[4]; a[4] = a[5]; a[5] = a[6]; a[6] = a[7]; a[7] = t; }
            if (rotg & 2) { const float t0 = a[0], t1 = a[1]; a[0] = a[2]; a[1] = a[3]; a[2] = a[4]; a[3] = a[5]; a[4] = a[6]; a[5] = a[7]; a[6] = t0; a[7] = t1; }
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] = (rotg & 4) ? a[(j + 4) & 7] : a[j];
            float* const sb = s_stage + lane * 8 * NB + b;
#pragma unroll
            for (int j = 0; j < 8; j++) sb[((j + rotg) & 7) * NB] = r[j];
        } else {
            const float4 r0 = rot4(make_float4(a[0], a[1], a[2], a[3]), rotg), r1 = rot4(make_float4(a[4], a[5], a[6], a[7]), rotg);
            float* const sb = s_stage + lane * 4 * NB + b;
            float* const q0 = sb + ((0 + rotg) & 3) * NB;
            float* const q1 = sb + ((1 + rotg) & 3) * NB;
            float* const q2 = sb + ((2 + rotg) & 3) * NB;
            float* const q3 = sb + ((3 + rotg) & 3) * NB;
            q0[0] = r0.x; q1[0] = r0.y; q2[0] = r0.z; q3[0] = r0.w;
            q0[128 * NB] = r1.x; q1[128 * NB] = r1.y; q2[128 * NB] = r1.z; q3[128 * NB] = r1.w;
        }
    };
    auto load = [&](unsigned off16, float (&v)[8]) {
        const float* q = tap_ptr(feat, off16);
        if (V8) ldg_nc_v8(q, v);
        else {
            const float4 x = ldg_nc_v4(q), y = ldg_nc_v4(q + 128);
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
        }
    };
    const int2* lp = s_list + warp * PITCH;
    int4 ea = reinterpret_cast<const int4*>(lp)[0], eb = reinterpret_cast<const int4*>(lp)[1];
    int cnt = lp[16].x;
#pragma unroll 1
    for (int b = warp; b < NB; b += WARPS) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = 0.f;
        const int nbin = min(b + WARPS, NB - 1);
        int ncnt = cnt;
        int e = 0;
#pragma unroll 1
        do {
            const bool p0 = e < cnt, p1 = e + 1 < cnt, p2 = e + 2 < cnt, p3 = e + 3 < cnt;
            float v0[8], v1[8], v2[8], v3[8];
            if (p0) load((unsigned)ea.x, v0);
            if (p1) load((unsigned)ea.z, v1);
            if (p2) load((unsigned)eb.x, v2);
            if (p3) load((unsigned)eb.z, v3);
            const float w0 = __int_as_float(ea.y), w1 = __int_as_float(ea.w), w2 = __int_as_float(eb.y), w3 = __int_as_float(eb.w);
            e += 4;
            const bool more = e < cnt;
            const int2* np = more ? lp + e : s_list + nbin * PITCH;
            if (!more) ncnt = s_list[nbin * PITCH + 16].x;
            ea = reinterpret_cast<const int4*>(np)[0]; eb = reinterpret_cast<const int4*>(np)[1];
            if (p0) {
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(w0, v0[j], acc[j]);
            }
            if (p1) {
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(w1, v1[j], acc[j]);
            }
            if (p2) {
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(w2, v2[j], acc[j]);
            }
            if (p3) {
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = fmaf(w3, v3[j], acc[j]);
            }
        } while (e < cnt);
        stage(acc, b);
        lp = s_list + nbin * PITCH;
        cnt = ncnt;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    float* __restrict__ dst = out + ((size_t)roi * C + chunk0) * NB;
    if ((((size_t)out) & 15) == 0 && ((C * NB) & 3) == 0) {
        if (tid == 0) bulk_store_evict_first(dst, s_stage, 256u * NB * 4u);
    } else {
        for (int i = tid; i < 256 * NB; i += THREADS) __stcs(dst + i, s_stage[i]);
    }
}

// ---------------------------------------------------------------------------------- forward (7x7, persistent, warp-specialised bin-major)
// A/B only (RSDET_ROI_V8 = 4).  Persistent CTA of GW gather warps + one builder warp, double-buffered in-place lists:
// the builder fetches the next RoI (global counter), builds its merged lists while the gather warps work on the current
// one, and owns the bulk store of the finished block.  Hand-offs are named barriers (bar.arrive / bar.sync, 288 threads):
//   FULL[p]  builder -> gatherers   lists + meta of the RoI in buffer p are published
//   DONE     gatherers -> builder   every gather warp has staged its bins of the current RoI
//   FREE     builder -> gatherers   the bulk store of the previous RoI has finished reading the staging block
__device__ unsigned g_roi77p_counter;   // A/B kernel only: single-stream measurements
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int GW, int MINB>
__global__ void __launch_bounds__(32 * (GW + 1), MINB)
roi_align_fwd77ws_kernel(LevelSet L, const RoiGeom* __restrict__ gsorted, int K, float* __restrict__ out) {
    constexpr int NB = 49, PITCH = 18, ALL = 32 * (GW + 1);
    constexpr int kMetaOff = 2 * 8 * NB * PITCH, kStageOff = (kMetaOff + 32 + 127) & ~127;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int2* s_lists = reinterpret_cast<int2*>(smem_raw);
    volatile int4* s_meta = reinterpret_cast<volatile int4*>(smem_raw + kMetaOff);      // {roi, level, batch, valid} per buffer
    float* s_stage = reinterpret_cast<float*>(smem_raw + kStageOff);
    const int C = L.C;
    if (warp == GW) {
        // ------------------------------------------------------------------ builder warp
        auto build = [&](int idx, int p) {
            RoiGeom g = gsorted[idx];
            const int roi = g.gh;
            g.gh = 2; g.gw = 2;
            const int H = L.H[g.level], W = L.W[g.level];
            int2* sl = s_lists + p * NB * PITCH;
#pragma unroll 1
            for (int s = lane; s < NB * 4; s += 32) {       // A1: raw taps into the bin's own list slots
                const int b = s >> 2, q = s & 3;
                const int ph = b / 7, pw = b - ph * 7, iy = q >> 1, ix = q & 1;
                float x, y;
                sample_xy(g, L.version, ph, pw, iy, ix, x, y);
                const Taps t = make_taps(H, W, y, x);
#pragma unroll
                for (int k = 0; k < 4; k++) sl[b * PITCH + q * 4 + k] = make_int2(t.o[k] * (C >> 2), __float_as_int(t.w[k]));
            }
            __syncwarp();
#pragma unroll 1
            for (int b = lane; b < NB; b += 32) {           // A2: merge in registers, write back in place
                int o[16];
                float w[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { const int2 e = sl[b * PITCH + j]; o[j] = e.x; w[j] = __int_as_float(e.y); }
                int pos = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float acc = w[j];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        if (i <= j) continue;
                        const bool same = o[i] == o[j] && w[j] != 0.f;
                        acc += same ? w[i] : 0.f;
                        w[i] = same ? 0.f : w[i];
                    }
                    if (w[j] != 0.f) sl[b * PITCH + pos++] = make_int2(o[j], __float_as_int(acc));
                }
                sl[b * PITCH + 16] = make_int2(pos, 0);
            }
            if (lane == 0) { s_meta[p].x = roi; s_meta[p].y = g.level; s_meta[p].z = g.batch; s_meta[p].w = 1; }
            __syncwarp();
        };
        int p = 0;
        build(blockIdx.x, 0);
        __threadfence_block();
        nb_arrive(1, ALL);
#pragma unroll 1
        for (;;) {
            int nidx = 0;
            if (lane == 0) nidx = (int)(atomicAdd(&g_roi77p_counter, 1u) + gridDim.x);
            nidx = __shfl_sync(0xffffffffu, nidx, 0);
            const bool nvalid = nidx < K;
            if (nvalid) build(nidx, p ^ 1);
            else if (lane == 0) s_meta[p ^ 1].w = 0;
            __threadfence_block();
            nb_arrive(1 + (p ^ 1), ALL);
            nb_sync(3, ALL);                                // the current RoI's block is complete in the staging area
            if (lane == 0) {
                const int roi = s_meta[p].x;
                unsigned long long pol;
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                             ::"l"(out + (size_t)roi * C * NB), "r"((unsigned)__cvta_generic_to_shared(s_stage)), "r"(256u * NB * 4u), "l"(pol) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            if (!nvalid) break;
            nb_arrive(4, ALL);                              // staging area free again
            p ^= 1;
        }
        return;
    }
    // ---------------------------------------------------------------------- gather warps
    const int oct = (lane >> 3) & 3;
    float* const sbase = s_stage + lane * 4 * NB;
    int p = 0;
    bool need_free = false;
#pragma unroll 1
    for (;;) {
        nb_sync(1 + p, ALL);
        const int roi_valid = s_meta[p].w;
        if (!roi_valid) break;
        const int level = s_meta[p].y, batch = s_meta[p].z;
        const int H = L.H[level], W = L.W[level];
        const int2* s_list = s_lists + p * NB * PITCH;
        const float4* __restrict__ feat = reinterpret_cast<const float4*>(L.feat[level] + (size_t)batch * H * W * C) + lane;
        auto stage = [&](float4 a0, float4 a1, int b) {
            a0.x *= 0.25f; a0.y *= 0.25f; a0.z *= 0.25f; a0.w *= 0.25f;
            a1.x *= 0.25f; a1.y *= 0.25f; a1.z *= 0.25f; a1.w *= 0.25f;
            const float4 r0 = rot4(a0, oct), r1 = rot4(a1, oct);
            float* const sb = sbase + b;
            float* const q0 = sb + ((0 + oct) & 3) * NB;
            float* const q1 = sb + ((1 + oct) & 3) * NB;
            float* const q2 = sb + ((2 + oct) & 3) * NB;
            float* const q3 = sb + ((3 + oct) & 3) * NB;
            q0[0] = r0.x; q1[0] = r0.y; q2[0] = r0.z; q3[0] = r0.w;
            q0[128 * NB] = r1.x; q1[128 * NB] = r1.y; q2[128 * NB] = r1.z; q3[128 * NB] = r1.w;
        };
#define RSDET_ACCW(P, WT, VA, VB)                                                                                       \
        if (P) {                                                                                                        \
            acc0.x = fmaf(WT, VA.x, acc0.x); acc0.y = fmaf(WT, VA.y, acc0.y); acc0.z = fmaf(WT, VA.z, acc0.z); acc0.w = fmaf(WT, VA.w, acc0.w); \
            acc1.x = fmaf(WT, VB.x, acc1.x); acc1.y = fmaf(WT, VB.y, acc1.y); acc1.z = fmaf(WT, VB.z, acc1.z); acc1.w = fmaf(WT, VB.w, acc1.w); \
        }
        const int2* lp = s_list + warp * PITCH;
        int4 ea = reinterpret_cast<const int4*>(lp)[0], eb = reinterpret_cast<const int4*>(lp)[1];
        int cnt = lp[16].x;
#pragma unroll 1
        for (int b = warp; b < NB; b += GW) {
            float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
            const int nbin = min(b + GW, NB - 1);
            int ncnt = cnt;
            int e = 0;
#pragma unroll 1
            do {
                const bool p0 = e < cnt, p1 = e + 1 < cnt, p2 = e + 2 < cnt, p3 = e + 3 < cnt;
                float4 v00, v01, v10, v11, v20, v21, v30, v31;
                if (p0) { const float* q = tap_ptr(feat, (unsigned)ea.x); v00 = ldg_nc_v4(q); v01 = ldg_nc_v4(q + 128); }
                if (p1) { const float* q = tap_ptr(feat, (unsigned)ea.z); v10 = ldg_nc_v4(q); v11 = ldg_nc_v4(q + 128); }
                if (p2) { const float* q = tap_ptr(feat, (unsigned)eb.x); v20 = ldg_nc_v4(q); v21 = ldg_nc_v4(q + 128); }
                if (p3) { const float* q = tap_ptr(feat, (unsigned)eb.z); v30 = ldg_nc_v4(q); v31 = ldg_nc_v4(q + 128); }
                const float w0 = __int_as_float(ea.y), w1 = __int_as_float(ea.w), w2 = __int_as_float(eb.y), w3 = __int_as_float(eb.w);
                e += 4;
                const bool more = e < cnt;
                const int2* np = more ? lp + e : s_list + nbin * PITCH;
                if (!more) ncnt = s_list[nbin * PITCH + 16].x;
                ea = reinterpret_cast<const int4*>(np)[0]; eb = reinterpret_cast<const int4*>(np)[1];
                RSDET_ACCW(p0, w0, v00, v01) RSDET_ACCW(p1, w1, v10, v11) RSDET_ACCW(p2, w2, v20, v21) RSDET_ACCW(p3, w3, v30, v31)
            } while (e < cnt);
            if (need_free) { nb_sync(4, ALL); need_free = false; }      // once per RoI and warp, before its first staging write
            stage(acc0, acc1, b);
            lp = s_list + nbin * PITCH;
            cnt = ncnt;
        }
#undef RSDET_ACCW
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __threadfence_block();
        nb_arrive(3, ALL);
        need_free = true;
        p ^= 1;
    }
}

// ---------------------------------------------------------------------------------- forward (7x7, persistent, pipelined lists)
// A/B only (RSDET_ROI_Q = 1 dynamic bins / 2 static columns): measured 126.8 / 125.4 us per tile against 119.2 us for
// roi_align_fwd77p_kernel -- the warps that merge the next item's lists become the stragglers of the current one
// (barrier stall 1.88 -> 2.46 per issue), so the shipped kernel keeps the list build in front of the gather.
// roi_align_fwd77p_kernel with the list build of item i+1 moved under the gather of item i:
//   * two list buffers; at the top of iteration i every thread computes the raw taps of item i+1 (A1, ~150 instructions),
//     after the barrier warps 0 and 1 merge them in place (A2, the ~600-instruction dependent chain that used to hold the
//     other six warps at a barrier) while the other warps already gather item i, then join;
//   * bins are handed out dynamically (a shared counter, one ATOMS per bin, taken at the start of a bin's last batch so
//     its latency lies under that batch's loads): the late joiners and the uneven tap counts balance out, and 49 bins no
//     longer have to be split 7 + 6 x 7 over eight warps;
//   * two CTA barriers per item instead of three.
// Same arithmetic and per-bin tap order: bit-identical results.
constexpr size_t kMeta77qOffset = 2 * 8 * 49 * 18;                                   // two list buffers
constexpr size_t kStage77qOffset = (kMeta77qOffset + 128 + 127) & ~(size_t)127;     // + meta block
template <int WARPS, int MINB, int MODE>   // MODE 0: dynamic bins, A2 on warps 0-1; 1: static columns, A2 on warps 6-7 (which share column 6)
__global__ void __launch_bounds__(32 * WARPS, MINB)
roi_align_fwd77q_kernel(LevelSet L, const RoiGeom* __restrict__ gsorted, int K, int chunks, float* __restrict__ out,
                        unsigned* __restrict__ counter) {
    constexpr int NB = 49, PITCH = 18, THREADS = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int2* s_lists = reinterpret_cast<int2*>(smem_raw);
    volatile int* s_next = reinterpret_cast<volatile int*>(smem_raw + kMeta77qOffset);            // item after the next one
    int* s_binctr = reinterpret_cast<int*>(smem_raw + kMeta77qOffset + 4);
    const RoiGeom* s_geom = reinterpret_cast<const RoiGeom*>(smem_raw + kMeta77qOffset + 16);     // record of the NEXT item
    volatile int4* s_meta = reinterpret_cast<volatile int4*>(smem_raw + kMeta77qOffset + 64);     // per buffer: {roi, chunk0, level, batch}
    float* s_stage = reinterpret_cast<float*>(smem_raw + kStage77qOffset);
    const int C = L.C, items = K * chunks;
    const int oct = (lane >> 3) & 3;
    float* const sbase = s_stage + lane * 4 * NB;

    // raw taps of item `idx` (record in s_geom) into list buffer `buf`, one thread per sample
    auto a1 = [&](int idx, int buf) {
        RoiGeom g = *s_geom;
        const int roi = g.gh;                  // processing-order record: gh carries the RoI index
        g.gh = 2; g.gw = 2;
        const int H = L.H[g.level], W = L.W[g.level];
        int2* sl = s_lists + buf * NB * PITCH;
        if (tid < NB * 4) {
            const int b = tid >> 2, q = tid & 3;
            const int ph = b / 7, pw = b - ph * 7, iy = q >> 1, ix = q & 1;
            float x, y;
            sample_xy(g, L.version, ph, pw, iy, ix, x, y);
            const Taps t = make_taps(H, W, y, x);
#pragma unroll
            for (int k = 0; k < 4; k++) sl[b * PITCH + q * 4 + k] = make_int2(t.o[k] * (C >> 2), __float_as_int(t.w[k]));
        }
        if (tid == THREADS - 2) { s_meta[buf].x = roi; s_meta[buf].y = (idx % chunks) * 256; s_meta[buf].z = g.level; s_meta[buf].w = g.batch; }
    };
    // merge in place, one lane per bin (same arithmetic and order as build_tap_lists)
    auto a2 = [&](int buf) {
        const int t2 = MODE == 1 ? tid - 6 * 32 : tid;
        if (t2 >= 0 && t2 < NB) {
            int2* sl = s_lists + buf * NB * PITCH;
            const int b = t2;
            int o[16];
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; j++) { const int2 e = sl[b * PITCH + j]; o[j] = e.x; w[j] = __int_as_float(e.y); }
            int pos = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float acc = w[j];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (i <= j) continue;
                    const bool same = o[i] == o[j] && w[j] != 0.f;
                    acc += same ? w[i] : 0.f;
                    w[i] = same ? 0.f : w[i];
                }
                if (w[j] != 0.f) sl[b * PITCH + pos++] = make_int2(o[j], __float_as_int(acc));
            }
            sl[b * PITCH + 16] = make_int2(pos, 0);
        }
    };
    auto load_record = [&](int idx) {          // three 16-byte async copies into s_geom
        if (tid < 3 && idx < items) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s_geom) + 16u * tid),
                         "l"(reinterpret_cast<const char*>(gsorted + idx / chunks) + 16 * tid) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };

    // prologue: lists of the first item, record of the second
    int idx = blockIdx.x;
    if (idx >= items) return;
    load_record(idx);
    if (tid < 3) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    a1(idx, 0);
    if (tid == THREADS - 1) *s_next = (int)(atomicAdd(counter, 1u) + gridDim.x);
    __syncthreads();
    int nidx = *s_next;
    a2(0);
    load_record(nidx);
    if (tid < 3) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    int buf = 0;
#pragma unroll 1
    while (true) {
        const bool have_next = nidx < items;
        if (have_next) a1(nidx, buf ^ 1);
        if (tid == THREADS - 1) {      // a thread without A1 / A2 work
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous block has left the staging area
            *s_binctr = 0;
            if (have_next) *s_next = (int)(atomicAdd(counter, 1u) + gridDim.x);
        }
        __syncthreads();
        const int nnidx = have_next ? *s_next : items;
        load_record(nnidx);
        if (have_next && (MODE == 1 ? warp >= 6 : warp < 2)) a2(buf ^ 1);

        const int roi = s_meta[buf].x, chunk0 = s_meta[buf].y, level = s_meta[buf].z, batch = s_meta[buf].w;
        const int H = L.H[level], W = L.W[level];
        const int2* s_list = s_lists + buf * NB * PITCH;
        const float4* __restrict__ feat = reinterpret_cast<const float4*>(L.feat[level] + (size_t)batch * H * W * C + chunk0) + lane;
        auto stage = [&](float4 a0, float4 a1_, int b) {
            a0.x *= 0.25f; a0.y *= 0.25f; a0.z *= 0.25f; a0.w *= 0.25f;   // output_val /= count (:143), count = 4: exact
            a1_.x *= 0.25f; a1_.y *= 0.25f; a1_.z *= 0.25f; a1_.w *= 0.25f;
            const float4 r0 = rot4(a0, oct), r1 = rot4(a1_, oct);
            float* const sb = sbase + b;
            float* const q0 = sb + ((0 + oct) & 3) * NB;
            float* const q1 = sb + ((1 + oct) & 3) * NB;
            float* const q2 = sb + ((2 + oct) & 3) * NB;
            float* const q3 = sb + ((3 + oct) & 3) * NB;
            q0[0] = r0.x; q1[0] = r0.y; q2[0] = r0.z; q3[0] = r0.w;
            q0[128 * NB] = r1.x; q1[128 * NB] = r1.y; q2[128 * NB] = r1.z; q3[128 * NB] = r1.w;
        };
        auto grab = [&]() {                    // next unclaimed bin of this item (>= NB: none left)
            int b = 0;
            if (lane == 0) b = atomicAdd(s_binctr, 1);
            return __shfl_sync(0xffffffffu, b, 0);
        };
#define RSDET_ACCQ(P, WT, VA, VB)                                                                                       \
        if (P) {                                                                                                        \
            acc0.x = fmaf(WT, VA.x, acc0.x); acc0.y = fmaf(WT, VA.y, acc0.y); acc0.z = fmaf(WT, VA.z, acc0.z); acc0.w = fmaf(WT, VA.w, acc0.w); \
            acc1.x = fmaf(WT, VB.x, acc1.x); acc1.y = fmaf(WT, VB.y, acc1.y); acc1.z = fmaf(WT, VB.z, acc1.z); acc1.w = fmaf(WT, VB.w, acc1.w); \
        }
        const int bend = MODE == 1 ? (warp == 6 ? 28 : NB) : NB;
        auto next_bin = [&](int cur) { if (MODE == 1) return cur + 7 < bend ? cur + 7 : NB; return grab(); };
        int b = MODE == 1 ? (warp == 7 ? 34 : warp) : grab();
        if (b < NB) {
            const int2* lp = s_list + b * PITCH;
            int4 ea = reinterpret_cast<const int4*>(lp)[0], eb = reinterpret_cast<const int4*>(lp)[1];
            int cnt = lp[16].x;
#pragma unroll 1
            while (b < NB) {
                float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
                int nbin = NB, ncnt = 0;
                int e = 0;
#pragma unroll 1
                do {
                    const bool p0 = e < cnt, p1 = e + 1 < cnt, p2 = e + 2 < cnt, p3 = e + 3 < cnt;
                    float4 v00, v01, v10, v11, v20, v21, v30, v31;
                    if (p0) { const float* q = tap_ptr(feat, (unsigned)ea.x); v00 = ldg_nc_v4(q); v01 = ldg_nc_v4(q + 128); }
                    if (p1) { const float* q = tap_ptr(feat, (unsigned)ea.z); v10 = ldg_nc_v4(q); v11 = ldg_nc_v4(q + 128); }
                    if (p2) { const float* q = tap_ptr(feat, (unsigned)eb.x); v20 = ldg_nc_v4(q); v21 = ldg_nc_v4(q + 128); }
                    if (p3) { const float* q = tap_ptr(feat, (unsigned)eb.z); v30 = ldg_nc_v4(q); v31 = ldg_nc_v4(q + 128); }
                    const float w0 = __int_as_float(ea.y), w1 = __int_as_float(ea.w), w2 = __int_as_float(eb.y), w3 = __int_as_float(eb.w);
                    e += 4;
                    const int2* np;
                    if (e < cnt) np = lp + e;
                    else {                     // last batch of this bin: claim the next one while the loads fly
                        nbin = next_bin(b);
                        const int nb_ = min(nbin, NB - 1);
                        np = s_list + nb_ * PITCH;
                        ncnt = np[16].x;
                    }
                    ea = reinterpret_cast<const int4*>(np)[0]; eb = reinterpret_cast<const int4*>(np)[1];
                    RSDET_ACCQ(p0, w0, v00, v01) RSDET_ACCQ(p1, w1, v10, v11) RSDET_ACCQ(p2, w2, v20, v21) RSDET_ACCQ(p3, w3, v30, v31)
                } while (e < cnt);
                stage(acc0, acc1, b);
                b = nbin;
                lp = s_list + min(nbin, NB - 1) * PITCH;
                cnt = ncnt;
            }
        }
#undef RSDET_ACCQ
        if (tid < 3) asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == THREADS - 1) {      // the thread that will wait for it
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         ::"l"(out + ((size_t)roi * C + chunk0) * NB), "r"((unsigned)__cvta_generic_to_shared(s_stage)), "r"(256u * NB * 4u), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (!have_next) break;
        nidx = nnidx;
        buf ^= 1;
    }
    if (tid == THREADS - 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------- forward (channel-split passes)
// Same tap lists and the same warp = bin gather, but a 256-channel chunk is produced in NP passes of 256/NP channels:
// the staging block shrinks to 50/NP KB, so four resident CTAs leave most of the SM's 256 KB to L1 -- the pixel rows
// that neighbouring bins of a RoI share (a bin's 3x3..4x4 footprint overlaps the next bin's by a row / column of
// pixels) are then served by L1 instead of the L2 -> SM fabric, whose ~6300 B/clk chip-wide cap bounds the bin-major
// kernel.  A warp works on BPW bins at once (one 16-byte load per lane and tap: 32 lanes = 128 channels), so that the
// loads in flight per thread stay at 4 taps x BPW.  The block of each pass leaves as its own TMA bulk store; the next
// pass starts gathering while that store still reads the staging block and only waits for it before its first
// staging write.
template <int NP, int BPW>
__global__ void __launch_bounds__(kRoiThreads, 4)
roi_align_fwd_split_kernel(LevelSet L, const float* __restrict__ rois, const int* __restrict__ order, const RoiGeom* __restrict__ geoms,
                           int K, float* __restrict__ out) {
    static_assert(256 / NP == 128, "one 16-byte load per lane: 128 channels per pass");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int roi = order ? order[blockIdx.x] : blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kRoiThreads / 32;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int cap = 4 * spb;
    const int C = L.C;
    const int chunk0 = blockIdx.y * 256;
    int2* s_list = reinterpret_cast<int2*>(smem_raw);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + list_bytes(nbins, cap));
    float* s_stage = reinterpret_cast<float*>(smem_raw + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~15));
    const RoiGeom g = geoms[roi];
    const int H = L.H[g.level], W = L.W[g.level];
    build_tap_lists(g, L, H, W, s_list, s_cnt, s_stage, C >> 2, 0);
    const int oct = (lane >> 3) & 3;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)max(spb, 1), inv_count = 1.f / count;
    const int CP = 256 / NP;                                // channels per pass
    const bool bulk_ok = ((size_t)out & 15) == 0 && ((CP * nbins) & 3) == 0 && (((size_t)C * nbins) & 3) == 0;
#pragma unroll 1
    for (int pass = 0; pass < NP; pass++) {
        const float4* __restrict__ feat =
            reinterpret_cast<const float4*>(L.feat[g.level] + (size_t)g.batch * H * W * C + chunk0 + pass * CP) + lane;
        bool first = pass > 0;                              // the previous pass's bulk store may still read the block
#pragma unroll 1
        for (int b0 = warp; b0 < nbins; b0 += kWarps * BPW) {
            float4 acc[BPW];
            int cnt[BPW], cmax = 0;
#pragma unroll
            for (int u = 0; u < BPW; u++) {
                acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int b = b0 + u * kWarps;
                cnt[u] = b < nbins ? s_cnt[b] : 0;
                cmax = max(cmax, cnt[u]);
            }
            for (int e = 0; e < cmax; e += 4) {
                float4 v[BPW][4];
#pragma unroll
                for (int u = 0; u < BPW; u++) {
                    const int2* lp = s_list + (b0 + u * kWarps) * (cap + 1);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (e + k < cnt[u]) v[u][k] = ldg_nc_v4(tap_ptr(feat, (unsigned)lp[e + k].x));
                }
#pragma unroll
                for (int u = 0; u < BPW; u++) {
                    const int2* lp = s_list + (b0 + u * kWarps) * (cap + 1);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (e + k < cnt[u]) {
                            const float wt = __int_as_float(lp[e + k].y);
                            acc[u].x = fmaf(wt, v[u][k].x, acc[u].x);
                            acc[u].y = fmaf(wt, v[u][k].y, acc[u].y);
                            acc[u].z = fmaf(wt, v[u][k].z, acc[u].z);
                            acc[u].w = fmaf(wt, v[u][k].w, acc[u].w);
                        }
                }
            }
            if (first) {                                    // every warp owns a first bin (kWarps <= nbins is checked by the host)
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                first = false;
            }
#pragma unroll
            for (int u = 0; u < BPW; u++) {
                const int b = b0 + u * kWarps;
                if (b >= nbins) break;
                if (pow2) { acc[u].x *= inv_count; acc[u].y *= inv_count; acc[u].z *= inv_count; acc[u].w *= inv_count; }
                else { acc[u].x /= count; acc[u].y /= count; acc[u].z /= count; acc[u].w /= count; }
                const int c0 = lane * 4;
                const float4 r = rot4(acc[u], oct);
                s_stage[(c0 + ((0 + oct) & 3)) * nbins + b] = r.x;
                s_stage[(c0 + ((1 + oct) & 3)) * nbins + b] = r.y;
                s_stage[(c0 + ((2 + oct) & 3)) * nbins + b] = r.z;
                s_stage[(c0 + ((3 + oct) & 3)) * nbins + b] = r.w;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        float* __restrict__ dst = out + ((size_t)roi * C + chunk0 + pass * CP) * nbins;
        const int total = CP * nbins;
        if (bulk_ok) {
            if (tid == 0) {
                unsigned long long pol;
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                             ::"l"(dst), "r"((unsigned)__cvta_generic_to_shared(s_stage)), "r"((unsigned)total * 4u), "l"(pol) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (pass == NP - 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            for (int e = tid; e < total; e += kRoiThreads) __stcs(dst + e, s_stage[e]);
            __syncthreads();
        }
    }
}

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

static bool px_path_ok(const rsdet_roi_align_cfg* c) {
    if (c->pooled_h != kPxRows || c->pooled_w != kPxCols || c->sampling_ratio != 2 || c->channels % 256 != 0) return false;
    for (int l = 0; l < c->num_levels; l++)   // pixel keys are (y << 16 | x); pixel indices 32-bit
        if (c->height[l] >= 32768 || c->width[l] >= 65536 || (long long)c->height[l] * c->width[l] >= (1ll << 31)) return false;
    return true;
}

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

// ---------------------------------------------------------------------------------- forward (tap-list records + persistent gather)
// The tap lists of a RoI do not depend on the features, and building them inside the gather kernel costs each CTA
// ~8 000 of its ~30 000 cycles -- not because of its instructions but because its shared-memory operations queue in
// the SM's single L1 FIFO behind the other CTAs' gather loads (profiles/README.md).  So: `roi_lists_kernel` builds the
// merged lists of every RoI at full speed into a global record ([lists][counts][level, batch], L2-resident: 27 MB per
// 4000 RoIs), and `roi_align_fwd_rec_kernel` is a gather-only loop: each CTA walks RoIs p = blockIdx.x, + gridDim.x, ...
// of the locality order, the NEXT RoI's record arrives by one TMA bulk copy (mbarrier-tracked, double-buffered) while
// the current one is gathered, and the output block leaves by a TMA bulk store that the next RoI's gather overlaps
// (it waits for the store to have READ the staging block only before its first staging write).  With gridDim.x = K the
// same kernel is the one-CTA-per-RoI form.  NP = 2 produces the 256 channels in two 128-channel passes (25 KB staging:
// four CTAs per SM with both record buffers).
__host__ __device__ inline size_t rec_cnt_bytes(size_t nbins) { return ((nbins + 2) * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t rec_bytes(size_t nbins, size_t cap) { return list_bytes(nbins, cap) + rec_cnt_bytes(nbins); }
__host__ __device__ inline size_t rec_stride(size_t nbins, size_t cap) { return (rec_bytes(nbins, cap) + 127) & ~(size_t)127; }

__global__ void __launch_bounds__(kRoiThreads)
roi_lists_kernel(LevelSet L, const RoiGeom* __restrict__ geoms, int K, unsigned char* __restrict__ records) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int roi = blockIdx.x, tid = threadIdx.x;
    const int nbins = L.PH * L.PW, cap = 4 * L.sampling_ratio * L.sampling_ratio;
    int2* s_list = reinterpret_cast<int2*>(smem_raw);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + list_bytes(nbins, cap));
    float* tmp = reinterpret_cast<float*>(smem_raw + rec_bytes(nbins, cap));
    const RoiGeom g = geoms[roi];
    build_tap_lists(g, L, L.H[g.level], L.W[g.level], s_list, s_cnt, tmp, L.C >> 2, 0);
    if (tid == 0) { s_cnt[nbins] = g.level; s_cnt[nbins + 1] = g.batch; }
    __syncthreads();
    const int n16 = (int)(rec_bytes(nbins, cap) >> 4);
    float4* dst = reinterpret_cast<float4*>(records + (size_t)roi * rec_stride(nbins, cap));
    const float4* src = reinterpret_cast<const float4*>(smem_raw);
    for (int i = tid; i < n16; i += kRoiThreads) dst[i] = src[i];   // unused list slots carry stale shared memory: never read
}

template <int NP>
__global__ void __launch_bounds__(kRoiThreads, NP == 2 ? 4 : 3)
roi_align_fwd_rec_kernel(LevelSet L, const int* __restrict__ order, const unsigned char* __restrict__ records, int K,
                         float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kRoiThreads / 32;
    constexpr int QPT = NP == 2 ? 1 : 2;                    // 16-byte loads per lane and tap
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int cap = 4 * spb;
    const int C = L.C;
    const int chunk0 = blockIdx.y * 256;
    const unsigned rbytes = (unsigned)rec_bytes(nbins, cap);
    const size_t rstride = rec_stride(nbins, cap);
    float* s_stage = reinterpret_cast<float*>(smem_raw + 2 * rbytes);
    const int G = gridDim.x;
    int p = blockIdx.x;
    if (p >= K) return;
    int r0 = order ? order[p] : p;
    int r1 = p + G < K ? (order ? order[p + G] : p + G) : -1;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar[0], rbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_raw)), "l"(records + (size_t)r0 * rstride), "r"(rbytes), "r"(smem_u32(&s_bar[0])) : "memory");
    }
    __syncthreads();
    const int oct = (lane >> 3) & 3;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)max(spb, 1), inv_count = 1.f / count;
    constexpr int CP = 256 / NP;
    const bool bulk_ok = ((size_t)out & 15) == 0 && ((CP * nbins) & 3) == 0 && (((size_t)C * nbins) & 3) == 0;
    bool store_pending = false;                              // a bulk store may still be reading the staging block
#pragma unroll 1
    for (int it = 0; p < K; it++, p += G) {
        const int buf = it & 1;
        const int r2 = p + 2 * G < K ? (order ? order[p + 2 * G] : p + 2 * G) : -1;   // consumed next iteration
        if (tid == 0 && r1 >= 0) {                           // every warp left buffer buf^1 at the barrier that ended iteration it-1
            mbar_expect_tx(&s_bar[buf ^ 1], rbytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem_raw + (buf ^ 1) * rbytes)), "l"(records + (size_t)r1 * rstride), "r"(rbytes), "r"(smem_u32(&s_bar[buf ^ 1])) : "memory");
        }
        mbar_wait(&s_bar[buf], (it >> 1) & 1);
        const int2* s_list = reinterpret_cast<const int2*>(smem_raw + buf * rbytes);
        const int* s_cnt = reinterpret_cast<const int*>(smem_raw + buf * rbytes + list_bytes(nbins, cap));
        const int level = s_cnt[nbins], batch = s_cnt[nbins + 1];
        const size_t img = (size_t)batch * L.H[level] * L.W[level] * C;
#pragma unroll 1
        for (int pass = 0; pass < NP; pass++) {
            const float4* __restrict__ feat = reinterpret_cast<const float4*>(L.feat[level] + img + chunk0 + pass * CP) + lane;
            bool first = store_pending;
#pragma unroll 1
            for (int b = warp; b < nbins; b += kWarps) {
                float4 acc[QPT];
#pragma unroll
                for (int u = 0; u < QPT; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int2* lp = s_list + b * (cap + 1);
                const int cnt = s_cnt[b];
                for (int e = 0; e < cnt; e += 4) {
                    int2 en[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) en[k] = lp[min(e + k, cnt - 1)];
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
                if (first) {                                 // every warp owns a first bin (kWarps <= nbins, checked by the host)
                    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncthreads();
                    first = false;
                }
#pragma unroll
                for (int u = 0; u < QPT; u++) {
                    if (pow2) { acc[u].x *= inv_count; acc[u].y *= inv_count; acc[u].z *= inv_count; acc[u].w *= inv_count; }
                    else { acc[u].x /= count; acc[u].y /= count; acc[u].z /= count; acc[u].w /= count; }
                    const int c0 = (lane + u * 32) * 4;
                    const float4 r = rot4(acc[u], oct);
                    s_stage[(c0 + ((0 + oct) & 3)) * nbins + b] = r.x;
                    s_stage[(c0 + ((1 + oct) & 3)) * nbins + b] = r.y;
                    s_stage[(c0 + ((2 + oct) & 3)) * nbins + b] = r.z;
                    s_stage[(c0 + ((3 + oct) & 3)) * nbins + b] = r.w;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            float* __restrict__ dst = out + ((size_t)r0 * C + chunk0 + pass * CP) * nbins;
            const int total = CP * nbins;
            if (bulk_ok) {
                if (tid == 0) {
                    unsigned long long pol;
                    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                                 ::"l"(dst), "r"(smem_u32(s_stage)), "r"((unsigned)total * 4u), "l"(pol) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
            } else {
                for (int e = tid; e < total; e += kRoiThreads) __stcs(dst + e, s_stage[e]);
                __syncthreads();
            }
        }
        r0 = r1; r1 = r2;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------- forward (balanced tap streams)
// What bounds the gather (tools/l2_bandwidth.cu, tools/l2_mix.cu, profiles/README.md "round 2, second session"): not
// L2 -> SM bandwidth (this access pattern reaches 16-20 TB/s next to the output stream, the bin-major kernel pulls
// 11), but the number of dependent load -> FMA round trips of the slowest warp of a CTA, each ~1 500 cycles: a warp
// owning bins with 9 merged taps issues batches of 4 + 4 + 1, and warp 0 owns 7 of the 49 bins.  The records of
// `roi_streams_kernel` therefore hold, per RoI, EIGHT TAP STREAMS of (nearly) equal length: the bins are dealt to the
// warps in contiguous runs that split the prefix sum of the tap counts evenly, each stream is the concatenation of
// its bins' merged taps padded to a multiple of four with zero-weight repeats of its last tap, and an entry is
//   x = offset (16-byte units, a multiple of C/4 >= 64) | bin,   y = weight, sign bit set on the last tap of a bin.
// The gather warp walks its stream in full batches of four (eight 16-byte loads in flight per thread, no predicates)
// and, when an entry carries the sign bit, scales the accumulators and writes them to the bin's staging slots.  Sums
// are formed in the same tap order as in the bin-major kernel: results are bit-identical.
constexpr int kStreamHdrBytes = 128;   // int wstart[9], level, batch, nzero, pad[4]; then 64 bytes: bins without taps
__host__ __device__ inline size_t stream_rec_bytes(size_t nbins, size_t cap) { return kStreamHdrBytes + 8 * (nbins * cap + 32); }
__host__ __device__ inline size_t stream_rec_stride(size_t nbins, size_t cap) { return (stream_rec_bytes(nbins, cap) + 127) & ~(size_t)127; }

__global__ void __launch_bounds__(kRoiThreads)
roi_streams_kernel(LevelSet L, const RoiGeom* __restrict__ geoms, int K, unsigned char* __restrict__ records) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_pos[64], s_load[9], s_wstart[9], s_nzero;
    const int roi = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int nbins = L.PH * L.PW, cap = 4 * L.sampling_ratio * L.sampling_ratio;   // nbins <= 64, cap == 16 (host check)
    constexpr int kWarps = kRoiThreads / 32;
    // smem: [record: header | streams][bin lists][counts][phase-A scratch]
    int* hdr = reinterpret_cast<int*>(smem_raw);
    unsigned char* zero_bins = smem_raw + 64;
    int2* stream = reinterpret_cast<int2*>(smem_raw + kStreamHdrBytes);
    const size_t rbytes = stream_rec_bytes(nbins, cap);
    int2* s_list = reinterpret_cast<int2*>(smem_raw + rbytes);
    int* s_cnt = reinterpret_cast<int*>(smem_raw + rbytes + list_bytes(nbins, cap));
    float* tmp = reinterpret_cast<float*>(smem_raw + rbytes + list_bytes(nbins, cap) + ((nbins * 4 + 15) & ~15));
    const RoiGeom g = geoms[roi];
    build_tap_lists(g, L, L.H[g.level], L.W[g.level], s_list, s_cnt, tmp, L.C >> 2, 0);
    if (tid < 32) {
        // bins 'lane' and 'lane + 32': exclusive prefix sums of the tap counts, owner warp by the midpoint rule
        const int c0 = lane < nbins ? s_cnt[lane] : 0, c1 = lane + 32 < nbins ? s_cnt[lane + 32] : 0;
        int i0 = c0, i1 = c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, i0, d), b = __shfl_up_sync(0xffffffffu, i1, d);
            if (lane >= d) { i0 += a; i1 += b; }
        }
        const int t0 = __shfl_sync(0xffffffffu, i0, 31);
        i1 += t0;
        const int T = __shfl_sync(0xffffffffu, i1, 31);
        const int p0 = i0 - c0, p1 = i1 - c1;
        const int o0 = T > 0 ? min(kWarps - 1, (2 * p0 + c0) * (kWarps / 2) / T) : 0;
        const int o1 = T > 0 ? min(kWarps - 1, (2 * p1 + c1) * (kWarps / 2) / T) : 0;
        int first = 0, wstart = 0, my_first0 = 0, my_first1 = 0, my_ws0 = 0, my_ws1 = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const int ld = __reduce_add_sync(0xffffffffu, (o0 == w ? c0 : 0) + (o1 == w ? c1 : 0));
            if (o0 == w) { my_first0 = first; my_ws0 = wstart; }
            if (o1 == w) { my_first1 = first; my_ws1 = wstart; }
            if (lane == 0) { s_load[w] = ld; s_wstart[w] = wstart; }
            first += ld;
            wstart += (ld + 3) & ~3;
        }
        if (lane == 0) s_wstart[kWarps] = wstart;
        if (lane < nbins) s_pos[lane] = my_ws0 + p0 - my_first0;
        if (lane + 32 < nbins) s_pos[lane + 32] = my_ws1 + p1 - my_first1;
        // bins without taps (every sample outside the map): the gather writes their zeros explicitly
        const unsigned z0 = __ballot_sync(0xffffffffu, lane < nbins && c0 == 0), z1 = __ballot_sync(0xffffffffu, lane + 32 < nbins && c1 == 0);
        if (lane < nbins && c0 == 0) zero_bins[__popc(z0 & ((1u << lane) - 1u))] = (unsigned char)lane;
        if (lane + 32 < nbins && c1 == 0) zero_bins[__popc(z0) + __popc(z1 & ((1u << lane) - 1u))] = (unsigned char)(lane + 32);
        if (lane == 0) s_nzero = __popc(z0) + __popc(z1);
    }
    __syncthreads();
    for (int idx = tid; idx < nbins * cap; idx += kRoiThreads) {
        const int b = idx / cap, j = idx - b * cap, c = s_cnt[b];
        if (j < c) {
            const int2 e = s_list[b * (cap + 1) + j];
            stream[s_pos[b] + j] = make_int2(e.x | b, j == c - 1 ? (int)((unsigned)e.y | 0x80000000u) : e.y);
        }
    }
    if (tid < 16) {
        if (tid <= kWarps) hdr[tid] = s_wstart[tid];
        else if (tid == 9) hdr[9] = g.level;
        else if (tid == 10) hdr[10] = g.batch;
        else if (tid == 11) hdr[11] = s_nzero;
        else hdr[tid] = 0;
    }
    __syncthreads();
    if (tid < kWarps * 4) {   // padding: zero-weight repeats of the stream's last tap (same address: an L1 hit)
        const int w = tid >> 2, k = tid & 3, ld = s_load[w];
        if (ld > 0 && ld + k < ((ld + 3) & ~3)) {
            const int2 last = stream[s_wstart[w] + ld - 1];
            stream[s_wstart[w] + ld + k] = make_int2(last.x, 0);
        }
    }
    __syncthreads();
    const int n16 = (int)(rbytes >> 4);
    float4* dst = reinterpret_cast<float4*>(records + (size_t)roi * stream_rec_stride(nbins, cap));
    const float4* src = reinterpret_cast<const float4*>(smem_raw);
    for (int i = tid; i < n16; i += kRoiThreads) dst[i] = src[i];   // slots past the streams carry stale shared memory: never read
}

// Asks L2 to fetch the channels-last pyramid ahead of the gather (cp.async.bulk.prefetch.L2: returns at once, the
// fills run in the background while the tap streams are built).  A first-touch DRAM miss inside a gather batch costs
// the warp a ~2 000-cycle round trip instead of ~900; with the pyramid (89 MB for a 1024^2 tile) L2-resident the
// gather's dependent chains only ever wait for L2.
struct PrefetchJob { const float* base[RSDET_MAX_LEVELS]; unsigned long long bytes[RSDET_MAX_LEVELS]; int levels; };
__global__ void l2_prefetch_kernel(PrefetchJob job) {
    constexpr unsigned long long kPiece = 32768;
    if (threadIdx.x != 0) return;
    for (int l = 0; l < job.levels; l++) {
        const unsigned long long n = (job.bytes[l] + kPiece - 1) / kPiece;
        for (unsigned long long i = blockIdx.x; i < n; i += gridDim.x) {
            const unsigned long long off = i * kPiece;
            const unsigned len = (unsigned)((job.bytes[l] - off < kPiece ? job.bytes[l] - off : kPiece) & ~15ull);
            if (len) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)job.base[l] + off), "r"(len) : "memory");
        }
    }
}

template <int NBUF>   // record buffers per CTA: 2 = the next RoI's record is prefetched (3 CTAs/SM), 1 = four CTAs/SM
__global__ void __launch_bounds__(kRoiThreads, NBUF == 2 ? 3 : 4)
roi_align_fwd_flat_kernel(LevelSet L, const int* __restrict__ order, const unsigned char* __restrict__ records, int K,
                          float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kRoiThreads / 32;
    const int nbins = L.PH * L.PW;
    const int spb = L.sampling_ratio * L.sampling_ratio;
    const int cap = 4 * spb;
    const int C = L.C;
    const int chunk0 = blockIdx.y * 256;
    const unsigned rbytes = (unsigned)stream_rec_bytes(nbins, cap);
    const size_t rstride = stream_rec_stride(nbins, cap);
    float* s_stage = reinterpret_cast<float*>(smem_raw + NBUF * rbytes);
    const int G = gridDim.x;
    int p = blockIdx.x;
    if (p >= K) return;
    int r0 = order ? order[p] : p;
    int r1 = p + G < K ? (order ? order[p + G] : p + G) : -1;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&s_bar[0], rbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem_raw)), "l"(records + (size_t)r0 * rstride), "r"(rbytes), "r"(smem_u32(&s_bar[0])) : "memory");
    }
    __syncthreads();
    const int oct = (lane >> 3) & 3;
    const bool pow2 = (spb & (spb - 1)) == 0;
    const float count = (float)max(spb, 1), inv_count = 1.f / count;
    const unsigned unit_mask = ~(unsigned)((C >> 2) - 1) | ~63u;   // offsets are multiples of C/4 (>= 64): the low 6 bits carry the bin
    const bool bulk_ok = ((size_t)out & 15) == 0 && (((size_t)C * nbins) & 3) == 0;
    bool store_pending = false;                              // a bulk store may still be reading the staging block
#pragma unroll 1
    for (int it = 0; p < K; it++, p += G) {
        const int buf = NBUF == 2 ? (it & 1) : 0;
        const int r2 = p + 2 * G < K ? (order ? order[p + 2 * G] : p + 2 * G) : -1;   // consumed next iteration
        if (NBUF == 2 && tid == 0 && r1 >= 0) {              // every warp left buffer buf^1 at the barrier that ended iteration it-1
            mbar_expect_tx(&s_bar[buf ^ 1], rbytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem_raw + (buf ^ 1) * rbytes)), "l"(records + (size_t)r1 * rstride), "r"(rbytes),
                           "r"(smem_u32(&s_bar[buf ^ 1])) : "memory");
        }
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
        const long long tf0 = clock64();
#endif
        mbar_wait(&s_bar[buf], NBUF == 2 ? ((it >> 1) & 1) : (it & 1));
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
        const long long tf1 = clock64();
        long long tf_drain = 0;
#endif
        const int* hdr = reinterpret_cast<const int*>(smem_raw + buf * rbytes);
        const unsigned char* zero_bins = smem_raw + buf * rbytes + 64;
        const int2* stream = reinterpret_cast<const int2*>(smem_raw + buf * rbytes + kStreamHdrBytes);
        const int level = hdr[9], batch = hdr[10], nzero = hdr[11];
        const float4* __restrict__ feat =
            reinterpret_cast<const float4*>(L.feat[level] + (size_t)batch * L.H[level] * L.W[level] * C + chunk0) + lane;
        bool need_sync = store_pending;
        auto drain = [&]() {                                 // once per RoI and warp, between the first batch's loads and its FMAs
            if (need_sync) {
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
                const long long td = clock64();
#endif
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                need_sync = false;
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
                tf_drain = clock64() - td;
#endif
            }
        };
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        const int beg = hdr[warp], end = hdr[warp + 1];
#pragma unroll 1
        for (int e = beg; e < end; e += 4) {
            const int4 q0 = *reinterpret_cast<const int4*>(stream + e), q1 = *reinterpret_cast<const int4*>(stream + e + 2);
            const int ex[4] = {q0.x, q0.z, q1.x, q1.z}, ey[4] = {q0.y, q0.w, q1.y, q1.w};
            float4 v[2][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float* ptr = tap_ptr(feat, (unsigned)ex[k] & unit_mask);
                v[0][k] = ldg_nc_v4(ptr);
                v[1][k] = ldg_nc_v4(ptr + 128);
            }
            drain();
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float wt = __int_as_float(ey[k] & 0x7fffffff);
                acc0.x = fmaf(wt, v[0][k].x, acc0.x); acc0.y = fmaf(wt, v[0][k].y, acc0.y);
                acc0.z = fmaf(wt, v[0][k].z, acc0.z); acc0.w = fmaf(wt, v[0][k].w, acc0.w);
                acc1.x = fmaf(wt, v[1][k].x, acc1.x); acc1.y = fmaf(wt, v[1][k].y, acc1.y);
                acc1.z = fmaf(wt, v[1][k].z, acc1.z); acc1.w = fmaf(wt, v[1][k].w, acc1.w);
                if (ey[k] < 0) {                             // last tap of bin b: output_val /= count (:143), into the [c][bin] block
                    const int b = ex[k] & 63;
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        float4 a = u ? acc1 : acc0;
                        if (pow2) { a.x *= inv_count; a.y *= inv_count; a.z *= inv_count; a.w *= inv_count; }
                        else { a.x /= count; a.y /= count; a.z /= count; a.w /= count; }
                        const int c0 = (lane + u * 32) * 4;
                        const float4 r = rot4(a, oct);
                        s_stage[(c0 + ((0 + oct) & 3)) * nbins + b] = r.x;
                        s_stage[(c0 + ((1 + oct) & 3)) * nbins + b] = r.y;
                        s_stage[(c0 + ((2 + oct) & 3)) * nbins + b] = r.z;
                        s_stage[(c0 + ((3 + oct) & 3)) * nbins + b] = r.w;
                    }
                    acc0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    acc1 = acc0;
                }
            }
        }
        drain();                                             // warps without taps
        for (int z = warp; z < nzero; z += kWarps) {
            const int b = zero_bins[z];
#pragma unroll
            for (int i = 0; i < 8; i++) s_stage[(lane * 8 + i) * nbins + b] = 0.f;
        }
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
        const long long tf2 = clock64();
#endif
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                     // the block is complete; the record buffer is free
#if defined(RSDET_PROF) && defined(RSDET_TUNING)
        if (g_px_prof && lane == 0) {   // per warp: record wait, gather (without the drain wait), drain wait, wait for the slowest warp, batches
            atomicAdd(&g_px_prof[0], (unsigned long long)(tf1 - tf0)); atomicAdd(&g_px_prof[1], (unsigned long long)(tf2 - tf1 - tf_drain));
            atomicAdd(&g_px_prof[2], (unsigned long long)tf_drain); atomicAdd(&g_px_prof[3], (unsigned long long)(clock64() - tf2));
            atomicAdd(&g_px_prof[4], (unsigned long long)((end - beg) >> 2)); atomicAdd(&g_px_prof[6], 1ull);
        }
#endif
        float* __restrict__ dst = out + ((size_t)r0 * C + chunk0) * nbins;
        const int total = 256 * nbins;
        if (NBUF == 1 && tid == 0 && r1 >= 0) {              // single buffer: the next record travels while the block is stored
            mbar_expect_tx(&s_bar[0], rbytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem_raw)), "l"(records + (size_t)r1 * rstride), "r"(rbytes), "r"(smem_u32(&s_bar[0])) : "memory");
        }
        if (bulk_ok) {
            if (tid == 0) {
                unsigned long long pol;
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                             ::"l"(dst), "r"(smem_u32(s_stage)), "r"((unsigned)total * 4u), "l"(pol) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            store_pending = true;
        } else {
            for (int e = tid; e < total; e += kRoiThreads) __stcs(dst + e, s_stage[e]);
            __syncthreads();
        }
        r0 = r1; r1 = r2;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

static bool stream_path_ok(const rsdet_roi_align_cfg* c) {   // 16 taps per bin before merging, bins in 6 bits, a 16-byte-aligned output block
    return c->channels % 256 == 0 && c->sampling_ratio == 2 && c->pooled_h * c->pooled_w <= 64 && (c->pooled_h * c->pooled_w) % 4 == 1;
}
static bool split_path_ok(const rsdet_roi_align_cfg* c) {
    return c->channels % 256 == 0 && c->pooled_h * c->pooled_w >= kRoiThreads / 32;
}
static size_t split_smem_bytes(const rsdet_roi_align_cfg* c) {
    size_t nbins = (size_t)c->pooled_h * c->pooled_w;
    size_t ntaps = nbins * 4 * c->sampling_ratio * c->sampling_ratio;
    size_t stage = sizeof(float) * nbins * 128, tmp = 12 * (ntaps + nbins);
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

// Host dispatch of the A/B paths (RSDET_ROI_PATH): returns true when one of them took the call (*rc = its status).
static bool tuning_forward(const rsdet_roi_align_cfg* cfg, const LevelSet& L, Workspace& ws, const float* rois, const int* order,
                           const RoiGeom* geoms, int num_rois, float* out, int32_t* levels_out, cudaStream_t st, int* rc) {
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
        { *rc = cuda_status(); return true; }
    }
    if (stream_path_ok(cfg) && (roi_path_choice() == 7 || roi_path_choice() == 8)) {
        const size_t nbins = (size_t)cfg->pooled_h * cfg->pooled_w, capz = 4 * (size_t)cfg->sampling_ratio * cfg->sampling_ratio;
        unsigned char* records = ws.take<unsigned char>((size_t)num_rois * stream_rec_stride(nbins, capz));
        if (!ws.ok()) { *rc = RSDET_EWORKSPACE; return true; }
        const size_t lsmem = stream_rec_bytes(nbins, capz) + list_bytes(nbins, capz) + ((nbins * 4 + 15) & ~(size_t)15) + 12 * (nbins * capz + nbins);
        set_dyn_smem((const void*)roi_streams_kernel, lsmem);
        if (getenv("RSDET_ROI_L2PF"))
        {
            PrefetchJob pj;
            pj.levels = cfg->num_levels;
            for (int l = 0; l < cfg->num_levels; l++) {
                pj.base[l] = L.feat[l];
                pj.bytes[l] = (unsigned long long)cfg->batch * cfg->height[l] * cfg->width[l] * cfg->channels * sizeof(float);
            }
            l2_prefetch_kernel<<<kNumSMs * 2, 32, 0, st>>>(pj);
        }
        roi_streams_kernel<<<num_rois, kRoiThreads, lsmem, st>>>(L, geoms, num_rois, records);
        int per_sm = 0;
        if (const char* e = getenv("RSDET_ROI_PERSIST")) per_sm = atoi(e);   // CTAs per SM of the persistent grid; 0: one CTA per RoI
        const int ctas = per_sm > 0 && num_rois > per_sm * kNumSMs ? per_sm * kNumSMs : num_rois;
        dim3 rgrid(ctas, cfg->channels / 256);
        if (roi_path_choice() == 7) {
            const size_t rsmem = 2 * stream_rec_bytes(nbins, capz) + sizeof(float) * nbins * 256;
            set_dyn_smem((const void*)roi_align_fwd_flat_kernel<2>, rsmem);
            roi_align_fwd_flat_kernel<2><<<rgrid, kRoiThreads, rsmem, st>>>(L, order, records, num_rois, out);
        } else {
            const size_t rsmem = stream_rec_bytes(nbins, capz) + sizeof(float) * nbins * 256;
            set_dyn_smem((const void*)roi_align_fwd_flat_kernel<1>, rsmem);
            roi_align_fwd_flat_kernel<1><<<rgrid, kRoiThreads, rsmem, st>>>(L, order, records, num_rois, out);
        }
        count_launch(2);
        { *rc = cuda_status(); return true; }
    }
    if (split_path_ok(cfg) && (roi_path_choice() == 5 || roi_path_choice() == 6)) {
        const size_t nbins = (size_t)cfg->pooled_h * cfg->pooled_w, capz = 4 * (size_t)cfg->sampling_ratio * cfg->sampling_ratio;
        unsigned char* records = ws.take<unsigned char>((size_t)num_rois * rec_stride(nbins, capz));
        if (!ws.ok()) { *rc = RSDET_EWORKSPACE; return true; }
        const size_t lsmem = rec_bytes(nbins, capz) + 12 * (nbins * capz + nbins);
        set_dyn_smem((const void*)roi_lists_kernel, lsmem);
        roi_lists_kernel<<<num_rois, kRoiThreads, lsmem, st>>>(L, geoms, num_rois, records);
        int per_sm = 0;
        if (const char* e = getenv("RSDET_ROI_PERSIST")) per_sm = atoi(e);   // CTAs per SM of the persistent grid; 0: one CTA per RoI
        const int ctas = per_sm > 0 && num_rois > per_sm * kNumSMs ? per_sm * kNumSMs : num_rois;
        dim3 rgrid(ctas, cfg->channels / 256);
        if (roi_path_choice() == 5) {
            const size_t rsmem = 2 * rec_bytes(nbins, capz) + sizeof(float) * nbins * 128;
            set_dyn_smem((const void*)roi_align_fwd_rec_kernel<2>, rsmem);
            roi_align_fwd_rec_kernel<2><<<rgrid, kRoiThreads, rsmem, st>>>(L, order, records, num_rois, out);
        } else {
            const size_t rsmem = 2 * rec_bytes(nbins, capz) + sizeof(float) * nbins * 256;
            set_dyn_smem((const void*)roi_align_fwd_rec_kernel<1>, rsmem);
            roi_align_fwd_rec_kernel<1><<<rgrid, kRoiThreads, rsmem, st>>>(L, order, records, num_rois, out);
        }
        count_launch(2);
        { *rc = cuda_status(); return true; }
    }
    if (split_path_ok(cfg) && roi_path_choice() == 4) {
        const size_t ssmem = split_smem_bytes(cfg);
        dim3 sgrid(num_rois, cfg->channels / 256);
        int carve = -1;
        if (const char* e = getenv("RSDET_ROI_CARVE")) carve = atoi(e);
        const int bpw = getenv("RSDET_ROI_BPW") ? atoi(getenv("RSDET_ROI_BPW")) : 2;
        if (bpw == 1) {
            set_dyn_smem((const void*)roi_align_fwd_split_kernel<2, 1>, ssmem);
            if (carve >= 0) cudaFuncSetAttribute(roi_align_fwd_split_kernel<2, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            roi_align_fwd_split_kernel<2, 1><<<sgrid, kRoiThreads, ssmem, st>>>(L, rois, order, geoms, num_rois, out);
        } else {
            set_dyn_smem((const void*)roi_align_fwd_split_kernel<2, 2>, ssmem);
            if (carve >= 0) cudaFuncSetAttribute(roi_align_fwd_split_kernel<2, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            roi_align_fwd_split_kernel<2, 2><<<sgrid, kRoiThreads, ssmem, st>>>(L, rois, order, geoms, num_rois, out);
        }
        count_launch();
        { *rc = cuda_status(); return true; }
    }
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
            { *rc = cuda_status(); return true; }
        }
    }
    return false;
}
