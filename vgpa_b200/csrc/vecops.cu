// vecops.cu -- batched vector kernels for the device-resident SCG driver
// (vgpa_b200/batched_scg.py): per-problem dot products, AXPYs, direction updates and
// masked copies over rows of (B, n) device arrays.  These are the optimiser's own arithmetic
// (reference: src/numerics/optim_scg.py:137-274), not part of the free-energy path, but at the
// Lorenz-96 shape (13 MB per row, ~15 passes per iteration) they cost as much as an evaluation, so
// they are laid out for HBM: a row is cut into gridDim.y slices so that small batches still fill the
// GPU, rows of problems that are no longer active are skipped, and reductions finish in a second
// kernel that adds the slice partials in a fixed order (bitwise reproducible for a given batch).
#include "../../include/vgpa_b200.h"
#include <algorithm>
#include <cuda_runtime.h>

namespace {

constexpr int TH = 256;
constexpr int MAX_SLICES = 64;

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

// elements [lo, hi) of the row handled by this CTA (slice blockIdx.y of gridDim.y)
__device__ __forceinline__ void slice_of(long long n, long long& lo, long long& hi)
{
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    lo = min(n, per * (long long)blockIdx.y);
    hi = min(n, lo + per);
}
__device__ __forceinline__ bool skipped(const int* active, int p) { return active != nullptr && active[p] == 0; }

// part[(k * B + p) * S + s], k = 0: x.y  1: x.z (z may be null)  2: x.x   over slice s of row p
__global__ void __launch_bounds__(TH) bdot_part_kernel(long long n, const double* __restrict__ x, const double* __restrict__ y,
                                                       const double* __restrict__ z, long long stride,
                                                       double* __restrict__ part, int B, const int* __restrict__ active)
{
    __shared__ double sh[TH];
    const int p = blockIdx.x, s = blockIdx.y, S = gridDim.y;
    if (skipped(active, p)) return;
    long long lo, hi;
    slice_of(n, lo, hi);
    const double* xp = x + p * stride;
    const double* yp = y + p * stride;
    const double* zp = z ? z + p * stride : nullptr;
    double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += TH) {
        const double xv = xp[i];
        a = fma(xv, yp[i], a);
        if (zp) b = fma(xv, zp[i], b);
        c = fma(xv, xv, c);
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        part[((long long)0 * B + p) * S + s] = a;
        part[((long long)1 * B + p) * S + s] = b;
        part[((long long)2 * B + p) * S + s] = c;
    }
}
// out[k * B + p] = sum over the S slice partials, in slice order; KMAX: 1 = max instead of sum for k == 0
template <bool FIRST_IS_MAX>
__global__ void finish_kernel(const double* __restrict__ part, double* __restrict__ out, int B, int S, int K,
                              const int* __restrict__ active)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= K * B) return;
    const int k = q / B, p = q - k * B;
    if (skipped(active, p)) return;
    const double* r = part + (long long)q * S;
    double v = r[0];
    for (int s = 1; s < S; ++s) v = (FIRST_IS_MAX && k == 0) ? fmax(v, r[s]) : v + r[s];
    out[q] = v;
}

// out = y + a[p] * x
__global__ void __launch_bounds__(TH) baxpy_kernel(long long n, const double* __restrict__ a, const double* __restrict__ x,
                                                   const double* __restrict__ y, double* __restrict__ out, long long stride,
                                                   const int* __restrict__ active)
{
    const int p = blockIdx.x;
    if (skipped(active, p)) return;
    long long lo, hi;
    slice_of(n, lo, hi);
    const double ap = a[p];
    const double* xp = x + p * stride;
    const double* yp = y + p * stride;
    double* op = out + p * stride;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += TH) op[i] = yp[i] + ap * xp[i];
}

// direction update: mode 0 keep, 1: d = gamma d - g (Polak-Ribiere), 2: d = -g (restart)
__global__ void __launch_bounds__(TH) bdir_kernel(long long n, const int* __restrict__ mode, const double* __restrict__ gamma,
                                                  double* __restrict__ d, const double* __restrict__ g, long long stride)
{
    const int p = blockIdx.x, m = mode[p];
    if (m == 0) return;
    long long lo, hi;
    slice_of(n, lo, hi);
    const double gm = gamma[p];
    double* dp = d + p * stride;
    const double* gp = g + p * stride;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += TH) dp[i] = (m == 1) ? (gm * dp[i]) - gp[i] : -gp[i];
}

// dst[p] = src[p] where mask[p] != 0
__global__ void __launch_bounds__(TH) bcopy_kernel(long long n, const int* __restrict__ mask, const double* __restrict__ src,
                                                   double* __restrict__ dst, long long stride)
{
    const int p = blockIdx.x;
    if (!mask[p]) return;
    long long lo, hi;
    slice_of(n, lo, hi);
    const double* sp = src + p * stride;
    double* dp = dst + p * stride;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += TH) dp[i] = sp[i];
}

// part[(k * B + p) * S + s], k = 0: max |x|  1: sum |x|
__global__ void __launch_bounds__(TH) bstats_part_kernel(long long n, const double* __restrict__ x, long long stride,
                                                         double* __restrict__ part, int B, const int* __restrict__ active)
{
    __shared__ double sh[TH];
    const int p = blockIdx.x, s = blockIdx.y, S = gridDim.y;
    if (skipped(active, p)) return;
    long long lo, hi;
    slice_of(n, lo, hi);
    const double* xp = x + p * stride;
    double m = 0.0, t = 0.0;
#pragma unroll 4
    for (long long i = lo + threadIdx.x; i < hi; i += TH) {
        const double v = fabs(xp[i]);
        m = fmax(m, v);
        t += v;
    }
    m = block_max(m, sh);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) {
        part[((long long)0 * B + p) * S + s] = m;
        part[((long long)1 * B + p) * S + s] = t;
    }
}

// slices per row: enough CTAs to fill the GPU at small B, at least 8192 elements per slice
int slices(int B, long long n)
{
    const long long want = (4 * 148 + B - 1) / B;
    return (int)std::max<long long>(1, std::min<long long>({want, (long long)MAX_SLICES, n / 8192}));
}

int done() { return cudaGetLastError() == cudaSuccess ? VGPA_OK : VGPA_ECUDA; }

// slice partials of the reductions: a small per-thread, per-device buffer that only grows (a
// stream-ordered allocation per call was measured: with the default pool it goes back to the driver
// at every synchronisation and dominated the optimiser)
double* partials(size_t count)
{
    thread_local double* buf = nullptr;
    thread_local size_t cap = 0;
    thread_local int dev = -1;
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return nullptr;
    if (cur != dev || count > cap) {
        if (buf != nullptr && cur == dev) cudaFree(buf);
        buf = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(count, 1 << 16);
        if (cudaMalloc(&buf, sizeof(double) * want) != cudaSuccess) return nullptr;
        cap = want;
        dev = cur;
    }
    return buf;
}

}  // namespace

extern "C" {

int vgpa_bdot(int B, int64_t n, const double* x, const double* y, const double* z, int64_t stride, double* out3B,
              const int32_t* active, void* stream)
{
    if (B < 1 || n < 1 || !x || !y || !out3B) return VGPA_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = slices(B, n);
    double* part = partials(3 * (size_t)B * S);
    if (part == nullptr) return VGPA_ECUDA;
    bdot_part_kernel<<<dim3(B, S), TH, 0, st>>>(n, x, y, z, stride, part, B, active);
    finish_kernel<false><<<(3 * B + 127) / 128, 128, 0, st>>>(part, out3B, B, S, 3, active);
    return done();
}
int vgpa_baxpy(int B, int64_t n, const double* a, const double* x, const double* y, double* out, int64_t stride,
               const int32_t* active, void* stream)
{
    if (B < 1 || n < 1 || !a || !x || !y || !out) return VGPA_EINVAL;
    baxpy_kernel<<<dim3(B, slices(B, n)), TH, 0, static_cast<cudaStream_t>(stream)>>>(n, a, x, y, out, stride, active);
    return done();
}
int vgpa_bdir(int B, int64_t n, const int32_t* mode, const double* gamma, double* d, const double* g, int64_t stride,
              void* stream)
{
    if (B < 1 || n < 1 || !mode || !gamma || !d || !g) return VGPA_EINVAL;
    bdir_kernel<<<dim3(B, slices(B, n)), TH, 0, static_cast<cudaStream_t>(stream)>>>(n, mode, gamma, d, g, stride);
    return done();
}
int vgpa_bcopy(int B, int64_t n, const int32_t* mask, const double* src, double* dst, int64_t stride, void* stream)
{
    if (B < 1 || n < 1 || !mask || !src || !dst) return VGPA_EINVAL;
    bcopy_kernel<<<dim3(B, slices(B, n)), TH, 0, static_cast<cudaStream_t>(stream)>>>(n, mask, src, dst, stride);
    return done();
}
int vgpa_bstats(int B, int64_t n, const double* x, int64_t stride, double* out2B, const int32_t* active, void* stream)
{
    if (B < 1 || n < 1 || !x || !out2B) return VGPA_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = slices(B, n);
    double* part = partials(2 * (size_t)B * S);
    if (part == nullptr) return VGPA_ECUDA;
    bstats_part_kernel<<<dim3(B, S), TH, 0, st>>>(n, x, stride, part, B, active);
    finish_kernel<true><<<(2 * B + 127) / 128, 128, 0, st>>>(part, out2B, B, S, 2, active);
    return done();
}

}  // extern "C"
