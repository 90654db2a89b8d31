// vecops.cu -- batched vector kernels for the device-resident SCG driver
// (vgpa_b200/batched_scg.py): per-problem dot products, AXPYs, direction updates and
// masked copies over rows of (B, n) device arrays.  One CTA per problem, fixed-order
// block reductions (bitwise reproducible).  These are the optimiser's own arithmetic
// (reference: src/numerics/optim_scg.py:137-274), not part of the free-energy path.
#include "../../include/vgpa_b200.h"
#include <cuda_runtime.h>

namespace {

constexpr int TH = 512;

__device__ __forceinline__ double block_sum(double v, double* sh)
{
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int o = TH >> 1; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double* sh)
{
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (int o = TH >> 1; o > 0; o >>= 1) {
        if (tid < o) sh[tid] = fmax(sh[tid], sh[tid + o]);
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

// out[0*B+p] = x.y   out[1*B+p] = x.z (z may be null)   out[2*B+p] = x.x
__global__ void __launch_bounds__(TH) bdot_kernel(long long n, const double* __restrict__ x, const double* __restrict__ y,
                                                  const double* __restrict__ z, long long stride, double* __restrict__ out, int B)
{
    __shared__ double sh[TH];
    const int p = blockIdx.x;
    const double* xp = x + p * stride;
    const double* yp = y + p * stride;
    const double* zp = z ? z + p * stride : nullptr;
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long i = threadIdx.x; i < n; i += TH) {
        const double xv = xp[i];
        a = fma(xv, yp[i], a);
        if (zp) b = fma(xv, zp[i], b);
        c = fma(xv, xv, c);
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        out[p] = a;
        out[B + p] = b;
        out[2 * B + p] = c;
    }
}

// out = y + a[p] * x
__global__ void __launch_bounds__(TH) baxpy_kernel(long long n, const double* __restrict__ a, const double* __restrict__ x,
                                                   const double* __restrict__ y, double* __restrict__ out, long long stride)
{
    const int p = blockIdx.x;
    const double ap = a[p];
    const double* xp = x + p * stride;
    const double* yp = y + p * stride;
    double* op = out + p * stride;
    for (long long i = threadIdx.x; i < n; i += TH) op[i] = yp[i] + ap * xp[i];
}

// direction update: mode 0 keep, 1: d = gamma d - g (Polak-Ribiere), 2: d = -g (restart)
__global__ void __launch_bounds__(TH) bdir_kernel(long long n, const int* __restrict__ mode, const double* __restrict__ gamma,
                                                  double* __restrict__ d, const double* __restrict__ g, long long stride)
{
    const int p = blockIdx.x, m = mode[p];
    if (m == 0) return;
    const double gm = gamma[p];
    double* dp = d + p * stride;
    const double* gp = g + p * stride;
    for (long long i = threadIdx.x; i < n; i += TH) dp[i] = (m == 1) ? (gm * dp[i]) - gp[i] : -gp[i];
}

// dst[p] = src[p] where mask[p] != 0
__global__ void __launch_bounds__(TH) bcopy_kernel(long long n, const int* __restrict__ mask, const double* __restrict__ src,
                                                   double* __restrict__ dst, long long stride)
{
    const int p = blockIdx.x;
    if (!mask[p]) return;
    const double* sp = src + p * stride;
    double* dp = dst + p * stride;
    for (long long i = threadIdx.x; i < n; i += TH) dp[i] = sp[i];
}

// out[p] = max |x|, out[B+p] = sum |x|
__global__ void __launch_bounds__(TH) bstats_kernel(long long n, const double* __restrict__ x, long long stride,
                                                    double* __restrict__ out, int B)
{
    __shared__ double sh[TH];
    const int p = blockIdx.x;
    const double* xp = x + p * stride;
    double m = 0.0, s = 0.0;
    for (long long i = threadIdx.x; i < n; i += TH) {
        const double v = fabs(xp[i]);
        m = fmax(m, v);
        s += v;
    }
    m = block_max(m, sh);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        out[p] = m;
        out[B + p] = s;
    }
}

int done(const char*) { return cudaGetLastError() == cudaSuccess ? VGPA_OK : VGPA_ECUDA; }

}  // namespace

extern "C" {

int vgpa_bdot(int B, int64_t n, const double* x, const double* y, const double* z, int64_t stride, double* out3B,
              void* stream)
{
    if (B < 1 || n < 1 || !x || !y || !out3B) return VGPA_EINVAL;
    bdot_kernel<<<B, TH, 0, static_cast<cudaStream_t>(stream)>>>(n, x, y, z, stride, out3B, B);
    return done("bdot");
}
int vgpa_baxpy(int B, int64_t n, const double* a, const double* x, const double* y, double* out, int64_t stride,
               void* stream)
{
    if (B < 1 || n < 1 || !a || !x || !y || !out) return VGPA_EINVAL;
    baxpy_kernel<<<B, TH, 0, static_cast<cudaStream_t>(stream)>>>(n, a, x, y, out, stride);
    return done("baxpy");
}
int vgpa_bdir(int B, int64_t n, const int32_t* mode, const double* gamma, double* d, const double* g, int64_t stride,
              void* stream)
{
    if (B < 1 || n < 1 || !mode || !gamma || !d || !g) return VGPA_EINVAL;
    bdir_kernel<<<B, TH, 0, static_cast<cudaStream_t>(stream)>>>(n, mode, gamma, d, g, stride);
    return done("bdir");
}
int vgpa_bcopy(int B, int64_t n, const int32_t* mask, const double* src, double* dst, int64_t stride, void* stream)
{
    if (B < 1 || n < 1 || !mask || !src || !dst) return VGPA_EINVAL;
    bcopy_kernel<<<B, TH, 0, static_cast<cudaStream_t>(stream)>>>(n, mask, src, dst, stride);
    return done("bcopy");
}
int vgpa_bstats(int B, int64_t n, const double* x, int64_t stride, double* out2B, void* stream)
{
    if (B < 1 || n < 1 || !x || !out2B) return VGPA_EINVAL;
    bstats_kernel<<<B, TH, 0, static_cast<cudaStream_t>(stream)>>>(n, x, stride, out2B, B);
    return done("bstats");
}

}  // extern "C"
