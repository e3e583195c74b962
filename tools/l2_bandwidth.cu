// L2 -> SM read ceiling for the RoI gather's access pattern (measurement tool, not part of the library).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_bandwidth.cu -o /tmp/l2bw && /tmp/l2bw
//
// Every warp reads whole 512-byte pieces (one LDG.E.128 per lane) of pseudo-random 1 KB rows of a buffer that fits
// L2 (32 MB) or does not (2 GB), MLP loads in flight per thread, at several residencies.  The number the RoI kernel
// is compared with is the best L2-resident line: bytes / time as a multiple of what `roi_align_fwd_kernel` pulls
// through the same fabric (profiles/README.md, round 2).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MLP, bool NOALLOC>
__global__ void gather_rows(const float4* __restrict__ buf, unsigned row_mask, int iters, float* sink) {
    const int lane = threadIdx.x & 31;
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp id: the row sequence is per warp
    s = s * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int it = 0; it < iters; it++) {
        float4 v[MLP];
#pragma unroll
        for (int k = 0; k < MLP; k++) {
            s = s * 1664525u + 1013904223u;
            const unsigned row = (s >> 8) & row_mask;            // 1 KB rows = 64 float4
            const float4* p = buf + (size_t)row * 64 + ((s >> 7) & 1) * 32 + lane;
            if (NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(p));
            else asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(p));
        }
#pragma unroll
        for (int k = 0; k < MLP; k++) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) *sink = acc.x;
}

template <int MLP, bool NOALLOC>
static void run(const char* name, const float4* buf, size_t bytes, int ctas_per_sm, int threads, float* sink) {
    const unsigned rows = (unsigned)(bytes / 1024);
    const int iters = 4096 / MLP;
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_rows<MLP, NOALLOC><<<grid, threads>>>(buf, rows - 1, iters, sink);
    gather_rows<MLP, NOALLOC><<<grid, threads>>>(buf, rows - 1, iters, sink);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; r++) gather_rows<MLP, NOALLOC><<<grid, threads>>>(buf, rows - 1, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)reps * grid * (threads / 32) * iters * MLP * 512.0;
    printf("%-10s buffer %5zu MB  warps/SM %2d  loads in flight/thread %2d  %s: %7.2f TB/s\n", name, bytes >> 20,
           ctas_per_sm * threads / 32, MLP, NOALLOC ? "no_allocate" : "L1 default", total / (ms * 1e-3) / 1e12);
}

int main() {
    float4* buf;
    float* sink;
    const size_t big = (size_t)2 << 30;
    cudaMalloc(&buf, big);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, big);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SM clock (max) %d MHz\n", clk / 1000);
    // hot windows: every warp of the chip draws its rows from the same small region (what a locality-ordered RoI batch does)
    for (size_t bytes : {(size_t)256 << 10, (size_t)1 << 20, (size_t)4 << 20, (size_t)16 << 20}) {
        run<8, true>("L2-window", buf, bytes, 4, 256, sink);
        run<8, false>("L2-window", buf, bytes, 4, 256, sink);
    }
    for (size_t bytes : {(size_t)32 << 20, (size_t)64 << 20, big}) {
        const char* nm = bytes <= ((size_t)64 << 20) ? "L2" : "HBM";
        run<4, false>(nm, buf, bytes, 4, 256, sink);
        run<8, false>(nm, buf, bytes, 2, 256, sink);
        run<8, false>(nm, buf, bytes, 4, 256, sink);
        run<8, true>(nm, buf, bytes, 4, 256, sink);
        run<8, false>(nm, buf, bytes, 8, 256, sink);
        run<16, false>(nm, buf, bytes, 4, 256, sink);
        run<16, true>(nm, buf, bytes, 8, 256, sink);
    }
    return 0;
}
