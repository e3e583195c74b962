// SM -> L2 write path on B200: what does a 50 KB output block per "RoI" cost, alone and next to the gather loads?
// (measurement tool, not part of the library; see tools/l2_mix.cu for the combined loop)
//   flavour 0 st.global.v4   1 st.global.cs.v4   2 st.global.wt.v4   3 TMA bulk store   4 TMA bulk store + L2 evict_first
//   5 TMA bulk store in 7 pieces of 7 KB   6 st.global.v4 + L2::evict_first policy   7 st.global.L1::no_allocate.v4
// destination: every CTA owns `span` blocks of 50 KB and cycles through them (span 1: 30 MB in total, L2-resident and
// rewritten; span 16: 475 MB streamed).  `loads` = 0 / 1: with the 512 KB-per-RoI gather loop running in the same CTAs.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int FL, bool LOADS>
__global__ void __launch_bounds__(256, 4) k(const float4* __restrict__ buf, unsigned row_mask, int rois, int span, float* __restrict__ out, float* sink) {
    extern __shared__ __align__(16) float s_stage[];
    const int lane = threadIdx.x & 31;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    s = s * 2654435761u + 12345u;
    float4 acc = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int e = threadIdx.x; e < 12544; e += 256) s_stage[e] = (float)e;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int r = 0; r < rois; r++) {
        if (LOADS)
            for (int it = 0; it < 16; it++) {
                float4 v[8];
#pragma unroll
                for (int kk = 0; kk < 8; kk++) {
                    s = s * 1664525u + 1013904223u;
                    const unsigned row = (s >> 8) & row_mask;
                    const float4* p = buf + (size_t)row * 64 + ((s >> 7) & 1) * 32 + lane;
                    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[kk].x), "=f"(v[kk].y), "=f"(v[kk].z), "=f"(v[kk].w) : "l"(p));
                }
#pragma unroll
                for (int kk = 0; kk < 8; kk++) { acc.x += v[kk].x; acc.y += v[kk].y; acc.z += v[kk].z; acc.w += v[kk].w; }
            }
        __syncthreads();
        float* my_out = out + ((size_t)blockIdx.x * span + (r % span)) * 12544;
        if (FL <= 2 || FL >= 6) {
            for (int e = threadIdx.x; e < 3136; e += 256) {
                float* p = my_out + 4 * e;
                if (FL == 0) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
                if (FL == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
                if (FL == 2) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
                if (FL == 6) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w), "l"(pol) : "memory");
                if (FL == 7) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
            }
        } else if (threadIdx.x == 0) {
            if (FL == 3) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(my_out), "r"(smem_u32(s_stage)), "r"(50176u) : "memory");
            if (FL == 4) asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(my_out), "r"(smem_u32(s_stage)), "r"(50176u), "l"(pol) : "memory");
            if (FL == 5)
                for (int q = 0; q < 7; q++)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(my_out + q * 1792), "r"(smem_u32(s_stage + q * 1792)), "r"(7168u) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <int FL, bool LOADS>
static void run(const float4* buf, float* out, float* sink, int span) {
    const int grid = 148 * 4, rois = 16;
    cudaFuncSetAttribute(k<FL, LOADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) k<FL, LOADS><<<grid, 256, 50176>>>(buf, 65535, rois, span, out, sink);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; r++) k<FL, LOADS><<<grid, 256, 50176>>>(buf, 65535, rois, span, out, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double stored = (double)reps * grid * rois * 50176.0, loaded = LOADS ? (double)reps * grid * 8 * rois * 16 * 8 * 512.0 : 0.0;
    printf("flavour %d span %2d loads %d: %7.1f us per launch; stores %5.2f TB/s, loads %5.2f TB/s (%s)\n", FL, span, (int)LOADS, ms * 1e3 / reps,
           stored / (ms * 1e-3) / 1e12, loaded / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float4* buf; float* out; float* sink;
    cudaMalloc(&buf, (size_t)64 << 20);
    cudaMalloc(&out, (size_t)148 * 4 * 50176 * 16);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, (size_t)64 << 20);
    printf("per launch: 592 CTAs x 16 blocks of 50 KB stored (475 MB)%s\n", "; with loads: 4.97 GB gathered");
    {   // plain memset for scale
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemsetAsync(out, 0, (size_t)148 * 4 * 50176 * 16);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) cudaMemsetAsync(out, 0, (size_t)148 * 4 * 50176 * 16);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemsetAsync 475 MB: %.1f us = %.2f TB/s\n", ms * 100, 10 * 148.0 * 4 * 50176 * 16 / (ms * 1e-3) / 1e12);
    }
    for (int span : {1, 16}) {
        run<0, false>(buf, out, sink, span); run<1, false>(buf, out, sink, span); run<2, false>(buf, out, sink, span);
        run<3, false>(buf, out, sink, span); run<4, false>(buf, out, sink, span); run<5, false>(buf, out, sink, span);
        run<6, false>(buf, out, sink, span); run<7, false>(buf, out, sink, span);
    }
    run<0, true>(buf, out, sink, 0 + 1); run<0, true>(buf, out, sink, 16);
    run<1, true>(buf, out, sink, 16); run<3, true>(buf, out, sink, 16); run<4, true>(buf, out, sink, 16); run<6, true>(buf, out, sink, 16);
    return 0;
}
