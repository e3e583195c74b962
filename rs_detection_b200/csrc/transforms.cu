// transforms.cu -- box format conversions on the hot path + library bookkeeping.
//
// obb2poly / obb2hbb / poly2hbb: python/jdet/ops/bbox_transforms.py:612-623, 626-632, 602-609 (each is
// ~10 tiny elementwise Jittor kernels in the reference; one fused kernel here, op order preserved).
// poly2origpoly: python/jdet/data/devkits/result_merge.py:196-203.
// Compile with -fmad=false so a*b+c stays two IEEE operations like the reference's separate kernels.
#include "common.cuh"

namespace rsdet {

unsigned long long g_launches = 0;

__global__ void obb2poly_kernel(const float* __restrict__ obb, int n, float* __restrict__ poly) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* o = obb + (size_t)i * 5;
    float cx = o[0], cy = o[1], w = o[2], h = o[3], t = o[4];
    float Cos = cosf(t), Sin = sinf(t);
    float v1x = w / 2 * Cos, v1y = -w / 2 * Sin;
    float v2x = -h / 2 * Sin, v2y = -h / 2 * Cos;
    float* p = poly + (size_t)i * 8;
    p[0] = cx + v1x + v2x; p[1] = cy + v1y + v2y;
    p[2] = cx + v1x - v2x; p[3] = cy + v1y - v2y;
    p[4] = cx - v1x - v2x; p[5] = cy - v1y - v2y;
    p[6] = cx - v1x + v2x; p[7] = cy - v1y + v2y;
}

__global__ void obb2hbb_kernel(const float* __restrict__ obb, int n, float* __restrict__ hbb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* o = obb + (size_t)i * 5;
    float cx = o[0], cy = o[1], w = o[2], h = o[3], t = o[4];
    float Cos = cosf(t), Sin = sinf(t);
    float xb = fabsf(w / 2 * Cos) + fabsf(h / 2 * Sin);
    float yb = fabsf(w / 2 * Sin) + fabsf(h / 2 * Cos);
    float* p = hbb + (size_t)i * 4;
    p[0] = cx - xb; p[1] = cy - yb; p[2] = cx + xb; p[3] = cy + yb;
}

__global__ void poly2hbb_kernel(const float* __restrict__ poly, int n, int np, float* __restrict__ hbb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = poly + (size_t)i * np * 2;
    float x1 = q[0], x2 = q[0], y1 = q[1], y2 = q[1];
    for (int k = 1; k < np; k++) {
        x1 = fminf(x1, q[2 * k]); x2 = fmaxf(x2, q[2 * k]);
        y1 = fminf(y1, q[2 * k + 1]); y2 = fmaxf(y2, q[2 * k + 1]);
    }
    float* p = hbb + (size_t)i * 4;
    p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2;
}

__global__ void poly2origpoly_kernel(const double* __restrict__ polys, const double* __restrict__ offs, int n,
                                     double* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * 8) return;
    int i = e >> 3, k = e & 7;
    const double* o = offs + (size_t)i * 3;
    out[e] = (polys[e] + ((k & 1) ? o[1] : o[0])) / o[2];
}

}  // namespace rsdet

using namespace rsdet;

extern "C" int rsdet_version(void) { return 100; }

extern "C" unsigned long long rsdet_launch_count(void) { return g_launches; }

extern "C" const char* rsdet_error_string(int code) {
    switch (code) {
        case RSDET_OK: return "ok";
        case RSDET_EINVAL: return "invalid argument";
        case RSDET_EWORKSPACE: return "workspace too small";
        case RSDET_ELIMIT: return "size above the documented limit";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

extern "C" int rsdet_obb2poly(const float* obb, int n, float* poly, void* stream) {
    if (n < 0) return RSDET_EINVAL;
    if (n == 0) return RSDET_OK;
    if (!obb || !poly) return RSDET_EINVAL;
    obb2poly_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(obb, n, poly);
    count_launch();
    return cuda_status();
}

extern "C" int rsdet_obb2hbb(const float* obb, int n, float* hbb, void* stream) {
    if (n < 0) return RSDET_EINVAL;
    if (n == 0) return RSDET_OK;
    if (!obb || !hbb) return RSDET_EINVAL;
    obb2hbb_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(obb, n, hbb);
    count_launch();
    return cuda_status();
}

extern "C" int rsdet_poly2hbb(const float* poly, int n, int num_points, float* hbb, void* stream) {
    if (n < 0 || num_points < 1) return RSDET_EINVAL;
    if (n == 0) return RSDET_OK;
    if (!poly || !hbb) return RSDET_EINVAL;
    poly2hbb_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(poly, n, num_points, hbb);
    count_launch();
    return cuda_status();
}

extern "C" int rsdet_poly2origpoly(const double* polys, const double* offs, int n, double* out, void* stream) {
    if (n < 0) return RSDET_EINVAL;
    if (n == 0) return RSDET_OK;
    if (!polys || !offs || !out) return RSDET_EINVAL;
    poly2origpoly_kernel<<<ceil_div(n * 8, 256), 256, 0, (cudaStream_t)stream>>>(polys, offs, n, out);
    count_launch();
    return cuda_status();
}
