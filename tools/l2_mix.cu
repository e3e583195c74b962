// What slows the RoI gather below the L2 -> SM ceiling?  The gather loop of tools/l2_bandwidth.cu (pseudo-random 1 KB
// rows of a window of a 64 MB buffer, 8 x LDG.E.128 in flight per thread, 4 CTAs x 8 warps per SM, a CTA barrier per
// "RoI" = 512 KB loaded) with an output block per RoI written next to it (measurement tool, not part of the library).
//   store mode 0 none   1 st.global.cs.v4   2 TMA bulk store (evict-first), waited for   3 TMA bulk store, waited for one RoI later
//   window: bytes of the buffer the rows are drawn from;  block: bytes stored per RoI;  span: output blocks a CTA cycles through
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE, int MLP>
__global__ void __launch_bounds__(256, 4) mix(const float4* __restrict__ buf, unsigned row_mask, int rois, int block_floats, int span,
                                              float* __restrict__ out, float* sink) {
    extern __shared__ __align__(16) float s_stage[];   // up to 12544 floats = 50 KB
    const int lane = threadIdx.x & 31;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    s = s * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = threadIdx.x; e < 12544; e += 256) s_stage[e] = (float)e;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    for (int r = 0; r < rois; r++) {
        for (int it = 0; it < 128 / MLP; it++) {
            float4 v[MLP];
#pragma unroll
            for (int k = 0; k < MLP; k++) {
                s = s * 1664525u + 1013904223u;
                const unsigned row = (s >> 8) & row_mask;
                const float4* p = buf + (size_t)row * 64 + ((s >> 7) & 1) * 32 + lane;
                asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(p));
            }
#pragma unroll
            for (int k = 0; k < MLP; k++) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
        }
        __syncthreads();
        float* my_out = out + ((size_t)blockIdx.x * span + (r % span)) * 12544;
        if (MODE == 1) {
            for (int e = threadIdx.x; e < block_floats / 4; e += 256)
                asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(my_out + 4 * e), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
        }
        if ((MODE == 2 || MODE == 3) && threadIdx.x == 0 && block_floats > 0) {
            if (MODE == 3) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                         ::"l"(my_out), "r"(smem_u32(s_stage)), "r"((unsigned)block_floats * 4u), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (MODE == 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (MODE == 3 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <int MODE, int MLP>
static void run(const float4* buf, float* out, float* sink, size_t window, int block_bytes, int span) {
    const int grid = 148 * 4, rois = 8;
    cudaFuncSetAttribute(mix<MODE, MLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50176);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned mask = (unsigned)(window / 1024) - 1;
    for (int i = 0; i < 2; i++) mix<MODE, MLP><<<grid, 256, 50176>>>(buf, mask, rois, block_bytes / 4, span, out, sink);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; r++) mix<MODE, MLP><<<grid, 256, 50176>>>(buf, mask, rois, block_bytes / 4, span, out, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loaded = (double)reps * grid * 8 * rois * 128 * 512.0, stored = (double)reps * grid * rois * (MODE ? block_bytes : 0);
    printf("store mode %d  in flight %2d  window %3zu MB  block %5d B  span %2d: %7.1f us per launch, loads %6.2f TB/s, stores %5.2f TB/s (%s)\n", MODE, MLP,
           window >> 20, MODE ? block_bytes : 0, span, ms * 1e3 / reps, loaded / (ms * 1e-3) / 1e12, stored / (ms * 1e-3) / 1e12,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float4* buf; float* out; float* sink;
    cudaMalloc(&buf, (size_t)64 << 20);
    cudaMalloc(&out, (size_t)148 * 4 * 50176 * 8);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, (size_t)64 << 20);
    printf("per launch: 592 CTAs x 8 'RoIs' x 512 KB loaded = 2.48 GB\n");
    for (size_t window : {(size_t)4 << 20, (size_t)16 << 20, (size_t)64 << 20}) {
        run<0, 8>(buf, out, sink, window, 0, 1);
        run<1, 8>(buf, out, sink, window, 50176, 1);
        run<2, 8>(buf, out, sink, window, 50176, 1);
        run<2, 8>(buf, out, sink, window, 50176, 8);
        run<3, 8>(buf, out, sink, window, 50176, 8);
    }
    for (int block : {6272, 12544, 25088, 50176}) run<3, 8>(buf, out, sink, (size_t)16 << 20, block, 8);
    run<0, 4>(buf, out, sink, (size_t)16 << 20, 0, 1);
    run<3, 4>(buf, out, sink, (size_t)16 << 20, 50176, 8);
    run<0, 16>(buf, out, sink, (size_t)16 << 20, 0, 1);
    run<3, 16>(buf, out, sink, (size_t)16 << 20, 50176, 8);
    return 0;
}
