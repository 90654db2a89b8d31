// api.cu -- the C ABI of libvgpa_b200.so (include/vgpa_b200.h): handle management,
// chunked three-phase evaluation (forward sweep -> time-parallel energy -> backward
// sweep fused with the gradient), host staging with copy/compute overlap.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vgpa_b200.h"
#include "common.cuh"

using namespace vgpa;

namespace {

thread_local std::string g_create_error;

// Hand-over of large device blocks between handles (vgpa_scratch_cache): while it is enabled, the large blocks
// of a destroyed handle -- its trajectory scratch -- are kept and given to the next handle that asks for exactly
// the same size on the same device.  An ensemble optimised in resident sub-batches creates one evaluator per
// sub-batch, all of one shape; cudaFree of ~12 GB of scratch was measured at 0.02-0.4 s in a lone process and at
// up to 2 s when the processes of four or eight GPUs free at the same time (more than the optimisation of
// the sub-batch's tail).  Off by default: a destroyed handle returns its memory.
struct BlockCache {
    struct Blk { void* p; size_t bytes; int dev; };
    std::mutex mu;
    bool enabled = false;
    std::vector<Blk> blocks;
    static constexpr size_t MIN_BYTES = size_t(32) << 20;
    static constexpr size_t MAX_BLOCKS = 24;
    void* take(size_t n)
    {
        if (n < MIN_BYTES) return nullptr;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < blocks.size(); ++i)
            if (blocks[i].bytes == n && blocks[i].dev == dev) {
                void* p = blocks[i].p;
                blocks.erase(blocks.begin() + i);
                return p;
            }
        return nullptr;
    }
    bool put(void* p, size_t n)
    {
        if (n < MIN_BYTES) return false;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (!enabled || blocks.size() >= MAX_BLOCKS) return false;
        blocks.push_back({p, n, dev});
        return true;
    }
    long long purge()      // frees every cached block; returns the bytes released
    {
        std::vector<Blk> out;
        {
            std::lock_guard<std::mutex> lk(mu);
            out.swap(blocks);
        }
        int cur = 0;
        cudaGetDevice(&cur);
        long long total = 0;
        for (const Blk& b : out) {
            cudaSetDevice(b.dev);
            cudaFree(b.p);
            total += (long long)b.bytes;
        }
        cudaSetDevice(cur);
        return total;
    }
};
BlockCache g_block_cache;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t alloc(size_t n)
    {
        release();
        if (n == 0) n = 8;
        if (void* c = g_block_cache.take(n)) {
            p = c;
            bytes = n;
            return cudaSuccess;
        }
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaErrorMemoryAllocation && g_block_cache.purge() > 0) {   // the kept blocks are in the way
            cudaGetLastError();
            e = cudaMalloc(&p, n);
        }
        if (e == cudaSuccess) bytes = n;
        else p = nullptr;
        return e;
    }
    void release()
    {
        if (p && !g_block_cache.put(p, bytes)) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

int n_theta(int model) { return model == VGPA_MODEL_L63 ? 3 : 1; }

}  // namespace

struct vgpa_handle {
    vgpa_desc d{};
    Batch batch{};
    Scratch scratch{};
    int chunk = 0;
    long long n_x = 0;   // N * D * (D + 1)
    DevBuf theta, sigma, R, obs_t, obs_index, obs_y, m0, s0, E0, status;
    DevBuf sc_mt, sc_st, sc_dEm, sc_dEs, sc_esde;
    // second scratch lane: consecutive chunks run on two internal streams so that the
    // latency-bound energy kernel of one chunk overlaps the sweeps of the other
    DevBuf sc2_mt, sc2_st, sc2_dEm, sc2_dEs, sc2_esde;
    Scratch scratch2{};
    int lanes = 1;
    cudaStream_t s_lane[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    cudaEvent_t ev_done = nullptr;      // end of the last vgpa_eval_device (a call on ANOTHER stream waits for it: one scratch)
    // host-API staging: two slots of one chunk each
    DevBuf st_x[2], st_g[2], st_F;
    DevBuf plist2[2];                   // vgpa_set_active_list: compacted list of problem indices (device copies,
    int plist_flip = 0;                 //   alternating: the evaluation in flight may still read the previous one)
    cudaStream_t s_aux = nullptr;       //   non-blocking stream of the list upload
    cudaEvent_t plist_ev[2] = {nullptr, nullptr};   //   end of the last evaluation that read each copy
    bool plist_used[2] = {false, false};
    std::vector<int> plist_host;        // ... and the host copy (error messages name the problem, not the position)
    int n_list = -1;                    // < 0: no list, every problem
    cudaStream_t s_comp = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2]{}, ev_comp[2]{}, ev_d2h[2]{};
    cudaStream_t last_stream = nullptr;
    bool status_dirty = false;
    long long launches = 0;
    std::string err;
    // optional per-kernel timing (bench.py roofline)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev[4];
    size_t tev_used[4] = {0, 0, 0, 0};
    double t_ms[4] = {0, 0, 0, 0};
    long long t_n[4] = {0, 0, 0, 0};

    void tick(int kind, cudaStream_t st, bool begin)
    {
        if (!timing) return;
        if (begin) {
            if (tev_used[kind] == tev[kind].size()) {
                cudaEvent_t a, b;
                cudaEventCreate(&a);
                cudaEventCreate(&b);
                tev[kind].push_back({a, b});
            }
            cudaEventRecord(tev[kind][tev_used[kind]].first, st);
        } else {
            cudaEventRecord(tev[kind][tev_used[kind]].second, st);
            ++tev_used[kind];
        }
    }

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what)
    {
        return fail(VGPA_ECUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
    }
};

#define CK(call, what)                                         \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return h->cuda_fail(e__, what); \
    } while (0)

namespace {

// upload a per-problem array given with an element stride (0 = shared)
int upload(vgpa_handle* h, DevBuf& dst, const double* src, long long stride, long long len, int B,
           const double** dev_ptr, long long* dev_stride, const char* name)
{
    if (src == nullptr) return h->fail(VGPA_EINVAL, "descriptor field %s is NULL", name);
    if (stride != 0 && stride < len) return h->fail(VGPA_EINVAL, "%s_stride %lld < %lld", name, stride, len);
    const long long copies = (stride == 0) ? 1 : B;
    std::vector<double> host((size_t)copies * len);
    for (long long q = 0; q < copies; ++q) memcpy(&host[(size_t)q * len], src + q * stride, sizeof(double) * len);
    CK(dst.alloc(host.size() * sizeof(double)), "cudaMalloc(params)");
    CK(cudaMemcpy(dst.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy(params)");
    *dev_ptr = dst.as<double>();
    *dev_stride = (stride == 0) ? 0 : len;
    return VGPA_OK;
}

bool small_model(int model) { return model != VGPA_MODEL_L96; }

// one pass over problems [p0, p0 + count) with device buffers
void run_chunk(vgpa_handle* h, const double* d_x, long long xs, int want_grad, double* d_F, double* d_grad,
               long long gs, int p0, int count, const Extra& ex, cudaStream_t st, int lane = 0)
{
    Scratch sc = lane ? h->scratch2 : h->scratch;
    sc.status = h->status.as<int>() + p0;
    const Batch& b = h->batch;
    const bool small = small_model(b.model);
    h->tick(0, st, true);
    if (small && want_grad && launch_small_fused(b, d_x, xs, d_F, d_grad, gs, p0, count, ex, st)) {
        h->tick(0, st, false);      // the one launch is accounted to the first phase
        h->launches += 1;
        return;
    }
    if (small) launch_small_fwd(b, sc, d_x, xs, p0, count, st);
    else launch_l96_fwd(b, sc, d_x, xs, p0, count, st);
    h->tick(0, st, false);
    h->tick(1, st, true);
    if (small) launch_small_energy(b, sc, d_x, xs, p0, count, ex, st);
    else launch_l96_energy(b, sc, d_x, xs, p0, count, ex, st);
    h->tick(1, st, false);
    h->tick(2, st, true);
    launch_finalize(b, sc, d_F, p0, count, ex, st);
    h->tick(2, st, false);
    h->launches += 3;
    if (want_grad) {
        h->tick(3, st, true);
        if (small) launch_small_bwd(b, sc, d_x, xs, d_grad, gs, p0, count, ex, st);
        else launch_l96_bwd(b, sc, d_x, xs, d_grad, gs, p0, count, ex, st);
        h->tick(3, st, false);
        h->launches += 1;
    }
}

// The host-buffer entry points (vgpa_eval, vgpa_eval_full) always serve EVERY problem: the active set
// and the compacted list installed for vgpa_eval_device are suspended for the call (a masked problem
// would otherwise return whatever the staging buffers hold), and a device evaluation still pending on
// the caller's stream is waited for first (it shares the scratch and the status words).
struct HostCallScope {
    vgpa_handle* h;
    const int *active, *plist;
    explicit HostCallScope(vgpa_handle* h_) : h(h_), active(h_->batch.active), plist(h_->batch.plist)
    {
        if (h->status_dirty) cudaStreamSynchronize(h->last_stream);
        h->batch.active = nullptr;
        h->batch.plist = nullptr;
    }
    ~HostCallScope()
    {
        h->batch.active = active;
        h->batch.plist = plist;
    }
};

int check_status(vgpa_handle* h)
{
    if (!h->status_dirty) return VGPA_OK;
    h->status_dirty = false;
    std::vector<int> st(h->d.B);
    // Every caller has synchronised the stream that wrote the status words.  The copy runs on the handle's private
    // NON-BLOCKING stream: a plain cudaMemcpy goes through the legacy default stream, which waits for every blocking
    // stream of the process -- e.g. for the evaluation another host thread has in flight on another handle
    // (ShardedBatchedSCG with concurrent sub-batches: two handles serialised each other at every status check).
    if (h->s_aux == nullptr) CK(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking), "cudaStreamCreate");
    CK(cudaMemcpyAsync(st.data(), h->status.p, sizeof(int) * h->d.B, cudaMemcpyDeviceToHost, h->s_aux), "cudaMemcpyAsync(status)");
    CK(cudaStreamSynchronize(h->s_aux), "cudaStreamSynchronize(status)");
    for (int p = 0; p < h->d.B; ++p)
        if (st[p] != 0) {
            // under a compacted launch the status words are indexed by launch position
            const int prob = (h->n_list >= 0 && p < (int)h->plist_host.size()) ? h->plist_host[p] : p;
            return h->fail(VGPA_ENOTPD, "Matrix is not positive definite: S(t) of problem %d at time index %d",
                           prob, st[p] - 1);
        }
    return VGPA_OK;
}

}  // namespace

extern "C" {

const char* vgpa_version(void) { return "vgpa_b200 0.1 (sm_100a)"; }

const char* vgpa_last_error(const vgpa_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vgpa_create(const vgpa_desc* d, vgpa_handle** out)
{
    if (out) *out = nullptr;
    if (d == nullptr || out == nullptr) {
        g_create_error = "vgpa_create: NULL argument";
        return VGPA_EINVAL;
    }
    vgpa_handle* h = new vgpa_handle();
    h->d = *d;
    auto bail = [&](int rc) {
        g_create_error = h->err;
        vgpa_destroy(h);
        return rc;
    };
    // ---- validation (the ValueErrors of the reference constructors) ----
    if (d->model < 0 || d->model > 3) return bail(h->fail(VGPA_EINVAL, "Unknown stochastic model -> %d", d->model));
    if (d->method < 0 || d->method > 3)
        return bail(h->fail(VGPA_EINVAL, "Integration method is unknown -> %d", d->method));
    const int needD = (d->model == VGPA_MODEL_L63) ? 3 : (d->model == VGPA_MODEL_L96 ? 40 : 1);
    if (d->D != needD) return bail(h->fail(VGPA_EINVAL, "Wrong state dimension %d for model %d (need %d)", d->D, d->model, needD));
    if (d->N < 2) return bail(h->fail(VGPA_EINVAL, "Need at least two time points, got %d", d->N));
    if (d->M < 0 || d->B < 1) return bail(h->fail(VGPA_EINVAL, "Wrong sizes M=%d B=%d", d->M, d->B));
    if (!(d->dt > 0.0) || !(d->dt_model > 0.0))
        return bail(h->fail(VGPA_EINVAL, "Discrete time step should be strictly positive -> %g", d->dt));
    if (d->M > 0 && d->obs_t == nullptr) return bail(h->fail(VGPA_EINVAL, "obs_t is NULL"));
    for (int n = 0; n < d->M; ++n) {
        if (d->obs_t[n] < 0 || d->obs_t[n] >= d->N || (n > 0 && d->obs_t[n] <= d->obs_t[n - 1]))
            return bail(h->fail(VGPA_EINVAL, "obs_t must be sorted unique indices in [0, N)"));
        // gaussian_like.py:137-146 indexes the covariance by the observation ordinal n
        if (n >= d->N) return bail(h->fail(VGPA_EINVAL, "more observations than time points"));
    }
    {
        const long long cnt = (d->sigma_stride == 0) ? 1 : d->B;
        for (long long q = 0; d->sigma && q < cnt; ++q)
            for (int i = 0; i < d->D; ++i)
                if (!(d->sigma[q * d->sigma_stride + i] > 0.0))
                    return bail(h->fail(VGPA_EINVAL, "The diffusion noise value should be strictly positive"));
    }
    if (cudaSetDevice(d->device) != cudaSuccess)
        return bail(h->fail(VGPA_ECUDA, "cudaSetDevice(%d) failed: no usable CUDA device", d->device));
    {
        cudaDeviceProp prop;
        cudaError_t e = cudaGetDeviceProperties(&prop, d->device);
        if (e != cudaSuccess) return bail(h->cuda_fail(e, "cudaGetDeviceProperties"));
        if (prop.major != 10)
            return bail(h->fail(VGPA_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                                d->device, prop.major, prop.minor));
    }
    const int D = d->D, N = d->N, M = d->M, B = d->B;
    h->n_x = (long long)N * D * (D + 1);
    Batch& b = h->batch;
    b.model = d->model; b.method = d->method; b.D = D; b.N = N; b.M = M; b.B = B;
    b.dt = d->dt; b.dt_model = d->dt_model;
    int rc;
#define UP(field, len)                                                                              \
    if ((rc = upload(h, h->field, d->field, d->field##_stride, (len), B, &b.field, &b.field##_stride, \
                     #field)) != VGPA_OK)                                                           \
        return bail(rc);
    UP(theta, n_theta(d->model));
    UP(sigma, D);
    UP(R, D);
    UP(m0, D);
    UP(s0, (long long)D * D);
    UP(E0, 1);
    if (M > 0) {
        UP(obs_y, (long long)M * D);
    } else {
        b.obs_y = nullptr;
        b.obs_y_stride = 0;
    }
#undef UP
    {
        std::vector<long long> ot(std::max(M, 1), 0);
        std::vector<int> oi(N, -1);
        for (int n = 0; n < M; ++n) {
            ot[n] = d->obs_t[n];
            oi[d->obs_t[n]] = n;
        }
        cudaError_t e;
        if ((e = h->obs_t.alloc(ot.size() * sizeof(long long))) != cudaSuccess ||
            (e = h->obs_index.alloc(oi.size() * sizeof(int))) != cudaSuccess ||
            (e = cudaMemcpy(h->obs_t.p, ot.data(), ot.size() * sizeof(long long), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(h->obs_index.p, oi.data(), oi.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(h->cuda_fail(e, "upload(obs_t)"));
        b.obs_t = h->obs_t.as<long long>();
        b.obs_index = h->obs_index.as<int>();
    }
    // ---- scratch: trajectories for one chunk of problems ----
    const long long per = 8LL * N * (2LL * D + 2LL * D * D + 1);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    long long budget = d->scratch_bytes > 0 ? d->scratch_bytes : std::min<long long>(56LL << 30, (long long)(free_b * 0.35));
    long long chunk = std::max<long long>(1, std::min<long long>(B, budget / per));
    if (!small_model(d->model) && chunk < B) {
        if (d->scratch_bytes > 0) {
            // caller-given budget: whole waves (multiples of 6 * 148: 3 backward / 4 forward CTAs per SM) when it allows
            const long long wave = 148 * 6;
            if (chunk >= wave) chunk = (chunk / wave) * wave;
            else if (chunk >= 148) chunk = (chunk / 148) * 148;
        } else {
            // default: as FEW passes as the budget allows, of EQUAL size.  Every pass ends in a partial wave of the
            // sweeps (one CTA per problem, a 1001-step recurrence each: the last CTAs run alone), so two passes of
            // 2048 problems beat four of 888 plus one of 544 by 2.3 % on the 4096-problem shard (measured; 54 GB
            // of scratch instead of 23)
            const long long passes = (B + chunk - 1) / chunk;
            chunk = (B + passes - 1) / passes;
        }
    }
    if (const char* ce = getenv("VGPA_CHUNK")) {   // experiments: explicit problems per pass
        const long long c = atoll(ce);
        if (c >= 1) chunk = std::min<long long>(c, B);
    }
    h->chunk = (int)chunk;
    {
        cudaError_t e;
        if ((e = h->sc_mt.alloc(sizeof(double) * chunk * N * D)) != cudaSuccess ||
            (e = h->sc_st.alloc(sizeof(double) * chunk * N * D * D)) != cudaSuccess ||
            (e = h->sc_dEm.alloc(sizeof(double) * chunk * N * D)) != cudaSuccess ||
            (e = h->sc_dEs.alloc(sizeof(double) * chunk * N * D * D)) != cudaSuccess ||
            (e = h->sc_esde.alloc(sizeof(double) * chunk * N)) != cudaSuccess ||
            (e = h->status.alloc(sizeof(int) * B)) != cudaSuccess ||
            (e = cudaMemset(h->status.p, 0, sizeof(int) * B)) != cudaSuccess)
            return bail(h->cuda_fail(e, "cudaMalloc(scratch)"));
    }
    {
        const char* env = getenv("VGPA_LANES");
        int want = env ? atoi(env) : 1;   // VGPA_LANES=2: measured +2 % (18.6k -> 19.0k evals/s) for 2x scratch; off by default
                                          // (per-kernel timings of overlapping kernels would also stop being meaningful)
        h->lanes = (!small_model(d->model) && B > chunk && want >= 2) ? 2 : 1;
        if (h->lanes == 2) {
            cudaError_t e;
            if ((e = h->sc2_mt.alloc(sizeof(double) * chunk * N * D)) != cudaSuccess ||
                (e = h->sc2_st.alloc(sizeof(double) * chunk * N * D * D)) != cudaSuccess ||
                (e = h->sc2_dEm.alloc(sizeof(double) * chunk * N * D)) != cudaSuccess ||
                (e = h->sc2_dEs.alloc(sizeof(double) * chunk * N * D * D)) != cudaSuccess ||
                (e = h->sc2_esde.alloc(sizeof(double) * chunk * N)) != cudaSuccess ||
                (e = cudaStreamCreateWithFlags(&h->s_lane[0], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaStreamCreateWithFlags(&h->s_lane[1], cudaStreamNonBlocking)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_join[0], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_join[1], cudaEventDisableTiming)) != cudaSuccess) {
                cudaGetLastError();
                h->lanes = 1;   // not enough memory for a second lane: run the chunks back to back
                for (DevBuf* q : {&h->sc2_mt, &h->sc2_st, &h->sc2_dEm, &h->sc2_dEs, &h->sc2_esde}) q->release();
            }
        }
        h->scratch2.mt = h->sc2_mt.as<double>();
        h->scratch2.st = h->sc2_st.as<double>();
        h->scratch2.dEm = h->sc2_dEm.as<double>();
        h->scratch2.dEs = h->sc2_dEs.as<double>();
        h->scratch2.esde_t = h->sc2_esde.as<double>();
        h->scratch2.status = h->status.as<int>();
    }
    h->scratch.mt = h->sc_mt.as<double>();
    h->scratch.st = h->sc_st.as<double>();
    h->scratch.dEm = h->sc_dEm.as<double>();
    h->scratch.dEs = h->sc_dEs.as<double>();
    h->scratch.esde_t = h->sc_esde.as<double>();
    h->scratch.status = h->status.as<int>();
    {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) != cudaSuccess)
            return bail(h->cuda_fail(e, "cudaStreamCreate"));
        for (int q = 0; q < 2; ++q)
            if ((e = cudaEventCreateWithFlags(&h->ev_h2d[q], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_comp[q], cudaEventDisableTiming)) != cudaSuccess ||
                (e = cudaEventCreateWithFlags(&h->ev_d2h[q], cudaEventDisableTiming)) != cudaSuccess)
                return bail(h->cuda_fail(e, "cudaEventCreate"));
    }
    *out = h;
    return VGPA_OK;
}

void vgpa_destroy(vgpa_handle* h)
{
    if (!h) return;
    cudaSetDevice(h->d.device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&h->theta, &h->sigma, &h->R, &h->obs_t, &h->obs_index, &h->obs_y, &h->m0, &h->s0, &h->E0,
                      &h->status, &h->sc_mt, &h->sc_st, &h->sc_dEm, &h->sc_dEs, &h->sc_esde, &h->sc2_mt, &h->sc2_st,
                      &h->sc2_dEm, &h->sc2_dEs, &h->sc2_esde, &h->st_x[0],
                      &h->st_x[1], &h->st_g[0], &h->st_g[1], &h->st_F, &h->plist2[0], &h->plist2[1]})
        b->release();
    if (h->s_aux) cudaStreamDestroy(h->s_aux);
    for (int q = 0; q < 2; ++q)
        if (h->plist_ev[q]) cudaEventDestroy(h->plist_ev[q]);
    for (int q = 0; q < 2; ++q) {
        if (h->ev_h2d[q]) cudaEventDestroy(h->ev_h2d[q]);
        if (h->ev_comp[q]) cudaEventDestroy(h->ev_comp[q]);
        if (h->ev_d2h[q]) cudaEventDestroy(h->ev_d2h[q]);
    }
    for (int k = 0; k < 4; ++k)
        for (auto& pr : h->tev[k]) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    for (int q = 0; q < 2; ++q) {
        if (h->s_lane[q]) cudaStreamDestroy(h->s_lane[q]);
        if (h->ev_join[q]) cudaEventDestroy(h->ev_join[q]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    delete h;
}

int vgpa_eval_device(vgpa_handle* h, const double* d_x, int64_t x_stride, int want_grad, double* d_F,
                     double* d_grad, int64_t grad_stride, void* stream)
{
    if (!h) return VGPA_EINVAL;
    if (d_x == nullptr || d_F == nullptr) return h->fail(VGPA_EINVAL, "x / F is NULL");
    if (want_grad && d_grad == nullptr) return h->fail(VGPA_EINVAL, "want_grad set but grad is NULL");
    if (x_stride != 0 && x_stride < h->n_x) return h->fail(VGPA_EINVAL, "x_stride %lld < %lld", (long long)x_stride, h->n_x);
    if (want_grad && grad_stride < h->n_x && h->d.B > 1)
        return h->fail(VGPA_EINVAL, "grad_stride %lld < %lld", (long long)grad_stride, h->n_x);
    if (((uintptr_t)d_x & 15) || (x_stride & 1) || (want_grad && (((uintptr_t)d_grad & 15) || (grad_stride & 1))))
        return h->fail(VGPA_EINVAL, "device buffers must be 16-byte aligned with even strides");
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // The handle owns ONE scratch and one set of status words: an evaluation enqueued on another stream than the
    // previous one is ordered after it (same stream: stream order does it).
    if (h->ev_done != nullptr && h->status_dirty && st != h->last_stream)
        CK(cudaStreamWaitEvent(st, h->ev_done, 0), "cudaStreamWaitEvent");
    CK(cudaMemsetAsync(h->status.p, 0, sizeof(int) * h->d.B, st), "cudaMemsetAsync(status)");
    Extra ex{};
    const int total = (h->n_list >= 0 && h->batch.plist != nullptr) ? h->n_list : h->d.B;   // launch positions
    if (h->lanes == 2) {
        CK(cudaEventRecord(h->ev_fork, st), "cudaEventRecord");
        CK(cudaStreamWaitEvent(h->s_lane[0], h->ev_fork, 0), "cudaStreamWaitEvent");
        CK(cudaStreamWaitEvent(h->s_lane[1], h->ev_fork, 0), "cudaStreamWaitEvent");
        int ci = 0;
        for (int p0 = 0; p0 < total; p0 += h->chunk, ++ci) {
            const int count = std::min(h->chunk, total - p0);
            run_chunk(h, d_x, x_stride, want_grad, d_F, d_grad, grad_stride, p0, count, ex, h->s_lane[ci & 1], ci & 1);
        }
        for (int q = 0; q < 2; ++q) {
            CK(cudaEventRecord(h->ev_join[q], h->s_lane[q]), "cudaEventRecord");
            CK(cudaStreamWaitEvent(st, h->ev_join[q], 0), "cudaStreamWaitEvent");
        }
    } else {
        for (int p0 = 0; p0 < total; p0 += h->chunk) {
            const int count = std::min(h->chunk, total - p0);
            run_chunk(h, d_x, x_stride, want_grad, d_F, d_grad, grad_stride, p0, count, ex, st);
        }
    }
    CK(cudaGetLastError(), "kernel launch");
    if (h->batch.plist != nullptr && h->n_list >= 0) {   // this evaluation reads the current copy of the list
        const int q = h->plist_flip;
        if (h->plist_ev[q] == nullptr) CK(cudaEventCreateWithFlags(&h->plist_ev[q], cudaEventDisableTiming), "cudaEventCreate");
        CK(cudaEventRecord(h->plist_ev[q], st), "cudaEventRecord");
        h->plist_used[q] = true;
    }
    if (h->ev_done == nullptr) CK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming), "cudaEventCreate");
    CK(cudaEventRecord(h->ev_done, st), "cudaEventRecord");
    h->last_stream = st;
    h->status_dirty = true;
    return VGPA_OK;
}

int vgpa_set_active(vgpa_handle* h, const int32_t* d_active)
{
    if (!h) return VGPA_EINVAL;
    h->batch.active = d_active;
    return VGPA_OK;
}

int vgpa_set_active_list(vgpa_handle* h, const int32_t* list, int32_t n)
{
    if (!h) return VGPA_EINVAL;
    if (list == nullptr || n < 0) {
        h->n_list = -1;
        h->batch.plist = nullptr;
        h->plist_host.clear();
        return VGPA_OK;
    }
    if (h->batch.model != MODEL_L96)
        return h->fail(VGPA_EINVAL, "vgpa_set_active_list: compacted launches exist for the D = 40 kernels only "
                                    "(use vgpa_set_active for the small models)");
    if (n > h->d.B) return h->fail(VGPA_EINVAL, "vgpa_set_active_list: %d entries for %d problems", (int)n, h->d.B);
    for (int k = 0; k < n; ++k)
        if (list[k] < 0 || list[k] >= h->d.B)
            return h->fail(VGPA_EINVAL, "vgpa_set_active_list: entry %d = %d outside [0, %d)", k, (int)list[k], h->d.B);
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    // Upload on a private non-blocking stream into the buffer the evaluation in flight does NOT read: the call
    // waits for its own tiny copy only, never for the kernels already enqueued on the caller's stream.
    h->plist_flip ^= 1;
    DevBuf& buf = h->plist2[h->plist_flip];
    if (h->plist_used[h->plist_flip]) {    // an evaluation two list changes ago read this copy: it must be over
        CK(cudaEventSynchronize(h->plist_ev[h->plist_flip]), "cudaEventSynchronize");
        h->plist_used[h->plist_flip] = false;
    }
    const size_t bytes = sizeof(int) * (size_t)std::max(h->d.B, 1);
    if (buf.bytes < bytes) CK(buf.alloc(bytes), "cudaMalloc(list)");
    if (h->s_aux == nullptr) CK(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking), "cudaStreamCreate");
    if (n > 0) {
        CK(cudaMemcpyAsync(buf.p, list, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, h->s_aux), "cudaMemcpyAsync(list)");
        CK(cudaStreamSynchronize(h->s_aux), "cudaStreamSynchronize");
    }
    h->plist_host.assign(list, list + n);
    h->n_list = n;
    h->batch.plist = buf.as<int>();
    return VGPA_OK;
}

long long vgpa_scratch_cache(int enable)
{
    {
        std::lock_guard<std::mutex> lk(g_block_cache.mu);
        g_block_cache.enabled = enable != 0;
    }
    return enable ? 0 : g_block_cache.purge();
}

int vgpa_sync(vgpa_handle* h)
{
    if (!h) return VGPA_EINVAL;
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    CK(cudaStreamSynchronize(h->last_stream), "cudaStreamSynchronize");
    return check_status(h);
}

int vgpa_eval(vgpa_handle* h, const double* x, int64_t x_stride, int want_grad, double* F, double* grad,
              int64_t grad_stride)
{
    if (!h) return VGPA_EINVAL;
    if (x == nullptr || F == nullptr) return h->fail(VGPA_EINVAL, "x / F is NULL");
    if (want_grad && grad == nullptr) return h->fail(VGPA_EINVAL, "want_grad set but grad is NULL");
    if (x_stride != 0 && x_stride < h->n_x) return h->fail(VGPA_EINVAL, "x_stride %lld < %lld", (long long)x_stride, h->n_x);
    if (want_grad && h->d.B > 1 && grad_stride < h->n_x)
        return h->fail(VGPA_EINVAL, "grad_stride %lld < %lld", (long long)grad_stride, h->n_x);
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    HostCallScope scope(h);
    // host buffers: short passes, so that the H2D copy of pass c + 1, the kernels of pass c and the D2H copy of
    // pass c - 1 overlap and the two staging slots per direction stay small (D = 40: 64 problems = 0.84 GB each)
    const int B = h->d.B, C = small_model(h->batch.model) ? h->chunk : std::min(h->chunk, 64);
    const long long nx = h->n_x;
    const bool shared_x = (x_stride == 0);
    // staging (lazily sized): per slot one chunk of x rows and gradient rows
    const size_t xbytes = sizeof(double) * (size_t)nx * (shared_x ? 1 : C);
    const int nslots = (B > C) ? 2 : 1;
    for (int q = 0; q < nslots; ++q) {
        if (h->st_x[q].bytes < xbytes) CK(h->st_x[q].alloc(xbytes), "cudaMalloc(x staging)");
        if (want_grad && h->st_g[q].bytes < sizeof(double) * (size_t)nx * C)
            CK(h->st_g[q].alloc(sizeof(double) * (size_t)nx * C), "cudaMalloc(grad staging)");
    }
    if (h->st_F.bytes < sizeof(double) * B) CK(h->st_F.alloc(sizeof(double) * B), "cudaMalloc(F)");
    CK(cudaMemsetAsync(h->status.p, 0, sizeof(int) * B, h->s_comp), "cudaMemsetAsync(status)");
    Extra ex{};
    int ci = 0;
    for (int p0 = 0; p0 < B; p0 += C, ++ci) {
        const int count = std::min(C, B - p0);
        const int q = ci & 1;
        // slot q: its previous gradients must have left and its previous x must be consumed
        if (ci >= 2) {
            CK(cudaStreamWaitEvent(h->s_h2d, h->ev_comp[q], 0), "cudaStreamWaitEvent");
            CK(cudaStreamWaitEvent(h->s_comp, h->ev_d2h[q], 0), "cudaStreamWaitEvent");
        }
        if (!shared_x || ci == 0) {
            if (shared_x) {
                CK(cudaMemcpyAsync(h->st_x[0].p, x, sizeof(double) * nx, cudaMemcpyHostToDevice, h->s_h2d), "H2D x");
            } else if (x_stride == nx) {
                CK(cudaMemcpyAsync(h->st_x[q].p, x + (long long)p0 * x_stride, sizeof(double) * nx * count,
                                   cudaMemcpyHostToDevice, h->s_h2d), "H2D x");
            } else {
                CK(cudaMemcpy2DAsync(h->st_x[q].p, sizeof(double) * nx, x + (long long)p0 * x_stride,
                                     sizeof(double) * x_stride, sizeof(double) * nx, count,
                                     cudaMemcpyHostToDevice, h->s_h2d), "H2D x");
            }
        }
        CK(cudaEventRecord(h->ev_h2d[q], h->s_h2d), "cudaEventRecord");
        CK(cudaStreamWaitEvent(h->s_comp, h->ev_h2d[q], 0), "cudaStreamWaitEvent");
        // kernels index problems globally (p0 + lp): offset the staged pointers accordingly
        const double* dx = shared_x ? h->st_x[0].as<double>() : h->st_x[q].as<double>() - (long long)p0 * nx;
        double* dg = want_grad ? h->st_g[q].as<double>() - (long long)p0 * nx : nullptr;
        run_chunk(h, dx, shared_x ? 0 : nx, want_grad, h->st_F.as<double>(), dg, nx, p0, count, ex, h->s_comp);
        CK(cudaGetLastError(), "kernel launch");
        CK(cudaEventRecord(h->ev_comp[q], h->s_comp), "cudaEventRecord");
        if (want_grad) {
            CK(cudaStreamWaitEvent(h->s_d2h, h->ev_comp[q], 0), "cudaStreamWaitEvent");
            if (grad_stride == nx || count == 1) {
                CK(cudaMemcpyAsync(grad + (long long)p0 * grad_stride, h->st_g[q].p, sizeof(double) * nx * count,
                                   cudaMemcpyDeviceToHost, h->s_d2h), "D2H grad");
            } else {
                CK(cudaMemcpy2DAsync(grad + (long long)p0 * grad_stride, sizeof(double) * grad_stride, h->st_g[q].p,
                                     sizeof(double) * nx, sizeof(double) * nx, count, cudaMemcpyDeviceToHost,
                                     h->s_d2h), "D2H grad");
            }
            CK(cudaEventRecord(h->ev_d2h[q], h->s_d2h), "cudaEventRecord");
        }
    }
    CK(cudaMemcpyAsync(F, h->st_F.p, sizeof(double) * B, cudaMemcpyDeviceToHost, h->s_comp), "D2H F");
    CK(cudaStreamSynchronize(h->s_comp), "cudaStreamSynchronize");
    CK(cudaStreamSynchronize(h->s_d2h), "cudaStreamSynchronize");
    h->status_dirty = true;
    return check_status(h);
}

int vgpa_eval_full(vgpa_handle* h, int64_t problem, const double* x, const vgpa_full_out* out)
{
    if (!h) return VGPA_EINVAL;
    if (x == nullptr || out == nullptr) return h->fail(VGPA_EINVAL, "x / out is NULL");
    if (problem < 0 || problem >= h->d.B) return h->fail(VGPA_EINVAL, "problem index %lld out of range", (long long)problem);
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    HostCallScope scope(h);
    const long long nx = h->n_x, N = h->d.N, D = h->d.D, nv = N * D, nm = N * D * D;
    DevBuf dx, dg, dF, dl, dp, def, dedf, dparts;
    auto cleanup = [&]() { for (DevBuf* b : {&dx, &dg, &dF, &dl, &dp, &def, &dedf, &dparts}) b->release(); };
    cudaError_t e;
    if ((e = dx.alloc(sizeof(double) * nx)) != cudaSuccess || (e = dg.alloc(sizeof(double) * nx)) != cudaSuccess ||
        (e = dF.alloc(sizeof(double) * h->d.B)) != cudaSuccess || (e = dl.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dp.alloc(sizeof(double) * nm)) != cudaSuccess || (e = def.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dedf.alloc(sizeof(double) * nm)) != cudaSuccess || (e = dparts.alloc(sizeof(double) * 3)) != cudaSuccess ||
        (e = cudaMemcpy(dx.p, x, sizeof(double) * nx, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cleanup();
        return h->cuda_fail(e, "vgpa_eval_full setup");
    }
    Extra ex{};
    ex.lamt = dl.as<double>(); ex.psit = dp.as<double>(); ex.Efx = def.as<double>(); ex.Edf = dedf.as<double>();
    ex.parts = dparts.as<double>();
    cudaMemsetAsync(h->status.p, 0, sizeof(int) * h->d.B, h->s_comp);
    // kernels address x / grad / F by global problem index
    run_chunk(h, dx.as<double>() - problem * nx, nx, 1, dF.as<double>(), dg.as<double>() - problem * nx, nx,
              (int)problem, 1, ex, h->s_comp);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_comp);
    if (e != cudaSuccess) {
        cleanup();
        return h->cuda_fail(e, "vgpa_eval_full kernels");
    }
    auto back = [&](double* dst, const void* src, long long n) {
        if (dst && e == cudaSuccess) e = cudaMemcpy(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost);
    };
    back(out->F, dF.as<double>() + problem, 1);
    back(out->parts, dparts.p, 3);
    back(out->grad, dg.p, nx);
    back(out->mt, h->scratch.mt, nv);
    back(out->st, h->scratch.st, nm);
    back(out->lamt, dl.p, nv);
    back(out->psit, dp.p, nm);
    back(out->Efx, def.p, nv);
    back(out->Edf, dedf.p, nm);
    back(out->dEsde_dm, h->scratch.dEm, nv);
    back(out->dEsde_ds, h->scratch.dEs, nm);
    cleanup();
    if (e != cudaSuccess) return h->cuda_fail(e, "vgpa_eval_full copy-back");
    h->status_dirty = true;
    return check_status(h);
}

// ---- stand-alone sweeps (FwdOde.__call__ / BwdOde.__call__) -------------------------
static int check_sweep_args(int method, int D, int N, double dt)
{
    if (method < 0 || method > 3) { g_create_error = "Integration method is unknown"; return VGPA_EINVAL; }
    if (D != 1 && D != 3 && D != 40) { g_create_error = "unsupported state dimension (1, 3 or 40)"; return VGPA_EINVAL; }
    if (N < 2) { g_create_error = "need at least two time points"; return VGPA_EINVAL; }
    if (!(dt > 0.0)) { g_create_error = "Discrete time step should be strictly positive"; return VGPA_EINVAL; }
    return VGPA_OK;
}

int vgpa_solve_fwd(int device, int method, int D, int N, double dt, const double* A, const double* b,
                   const double* m0, const double* s0, const double* sigma, double* mt, double* st)
{
    int rc = check_sweep_args(method, D, N, dt);
    if (rc) return rc;
    if (!A || !b || !m0 || !s0 || !sigma || !mt || !st) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const long long nv = (long long)N * D, nm = (long long)N * D * D;
    DevBuf dx, dm0, ds0, dsig, dmt, dst;
    cudaError_t e;
    auto cleanup = [&]() { for (DevBuf* q : {&dx, &dm0, &ds0, &dsig, &dmt, &dst}) q->release(); };
    if ((e = dx.alloc(sizeof(double) * (nm + nv))) != cudaSuccess || (e = dm0.alloc(sizeof(double) * D)) != cudaSuccess ||
        (e = ds0.alloc(sizeof(double) * D * D)) != cudaSuccess || (e = dsig.alloc(sizeof(double) * D)) != cudaSuccess ||
        (e = dmt.alloc(sizeof(double) * nv)) != cudaSuccess || (e = dst.alloc(sizeof(double) * nm)) != cudaSuccess ||
        (e = cudaMemcpy(dx.p, A, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dx.as<double>() + nm, b, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dm0.p, m0, sizeof(double) * D, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(ds0.p, s0, sizeof(double) * D * D, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dsig.p, sigma, sizeof(double) * D, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_solve_fwd: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    Batch bt{};
    bt.model = (D == 40) ? MODEL_L96 : (D == 3 ? MODEL_L63 : MODEL_OU);
    bt.method = method; bt.D = D; bt.N = N; bt.B = 1; bt.dt = dt; bt.dt_model = dt;
    bt.sigma = dsig.as<double>(); bt.m0 = dm0.as<double>(); bt.s0 = ds0.as<double>();
    Scratch sc{};
    sc.mt = dmt.as<double>(); sc.st = dst.as<double>();
    if (D == 40) launch_l96_fwd(bt, sc, dx.as<double>(), 0, 0, 1, nullptr);
    else launch_small_fwd(bt, sc, dx.as<double>(), 0, 0, 1, nullptr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mt, dmt.p, sizeof(double) * nv, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(st, dst.p, sizeof(double) * nm, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA error in vgpa_solve_fwd: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    return VGPA_OK;
}

int vgpa_solve_bwd(int device, int method, int D, int N, double dt, const double* A, const double* dEm,
                   const double* dEs, const double* jm, const double* js, double* lam, double* psi)
{
    int rc = check_sweep_args(method, D, N, dt);
    if (rc) return rc;
    if (!A || !dEm || !dEs || !jm || !js || !lam || !psi) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const long long nv = (long long)N * D, nm = (long long)N * D * D;
    DevBuf dA, dgm, dgs, djm, djs, dl, dp;
    cudaError_t e;
    auto cleanup = [&]() { for (DevBuf* q : {&dA, &dgm, &dgs, &djm, &djs, &dl, &dp}) q->release(); };
    if ((e = dA.alloc(sizeof(double) * nm)) != cudaSuccess || (e = dgm.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dgs.alloc(sizeof(double) * nm)) != cudaSuccess || (e = djm.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = djs.alloc(sizeof(double) * nm)) != cudaSuccess || (e = dl.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dp.alloc(sizeof(double) * nm)) != cudaSuccess ||
        (e = cudaMemcpy(dA.p, A, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dgm.p, dEm, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dgs.p, dEs, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(djm.p, jm, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(djs.p, js, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_solve_bwd: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    launch_bwd_dense(method, D, N, dt, dA.as<double>(), dgm.as<double>(), dgs.as<double>(), djm.as<double>(),
                     djs.as<double>(), dl.as<double>(), dp.as<double>(), nullptr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(lam, dl.p, sizeof(double) * nv, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(psi, dp.p, sizeof(double) * nm, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA error in vgpa_solve_bwd: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    return VGPA_OK;
}

int vgpa_model_energy(int device, int model, int D, int N, double dt_model, const double* theta,
                      const double* sigma, const double* A, const double* b, const double* m, const double* S,
                      double* Esde, double* Ef, double* Edf, double* dEm, double* dEs, double* dEth, double* dEsig)
{
    if (model < 0 || model > 3) { g_create_error = "Unknown stochastic model"; return VGPA_EINVAL; }
    const int needD = (model == VGPA_MODEL_L63) ? 3 : (model == VGPA_MODEL_L96 ? 40 : 1);
    if (D != needD || N < 2 || !(dt_model > 0.0)) { g_create_error = "Wrong dimensions for model.energy"; return VGPA_EINVAL; }
    if (!theta || !sigma || !A || !b || !m || !S || !Esde || !Ef || !Edf || !dEm || !dEs) {
        g_create_error = "NULL argument";
        return VGPA_EINVAL;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const long long nv = (long long)N * D, nm = (long long)N * D * D;
    const int nth = n_theta(model);
    DevBuf dx, dth, dsig, dmt, dst, dgm, dgs, des, def, dedf, dparts, dzero, dF, dstat, dft, dfs, dhth, dhsig;
    auto cleanup = [&]() {
        for (DevBuf* q : {&dx, &dth, &dsig, &dmt, &dst, &dgm, &dgs, &des, &def, &dedf, &dparts, &dzero, &dF, &dstat,
                          &dft, &dfs, &dhth, &dhsig})
            q->release();
    };
    const bool hyper = (dEth != nullptr) || (dEsig != nullptr);
    const int nhth = (model == VGPA_MODEL_L96) ? D : nth;   // lorenz_96.py:431: one entry per state dimension
    const long long nsig = (D == 1) ? 1 : (long long)D * D;
    cudaError_t e;
    if ((e = dx.alloc(sizeof(double) * (nm + nv))) != cudaSuccess || (e = dth.alloc(sizeof(double) * nth)) != cudaSuccess ||
        (e = dsig.alloc(sizeof(double) * D)) != cudaSuccess || (e = dmt.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dst.alloc(sizeof(double) * nm)) != cudaSuccess || (e = dgm.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dgs.alloc(sizeof(double) * nm)) != cudaSuccess || (e = des.alloc(sizeof(double) * N)) != cudaSuccess ||
        (e = def.alloc(sizeof(double) * nv)) != cudaSuccess || (e = dedf.alloc(sizeof(double) * nm)) != cudaSuccess ||
        (e = dparts.alloc(sizeof(double) * 3)) != cudaSuccess || (e = dzero.alloc(sizeof(double))) != cudaSuccess ||
        (e = dF.alloc(sizeof(double))) != cudaSuccess || (e = dstat.alloc(sizeof(int))) != cudaSuccess ||
        (e = cudaMemset(dzero.p, 0, sizeof(double))) != cudaSuccess || (e = cudaMemset(dstat.p, 0, sizeof(int))) != cudaSuccess ||
        (e = cudaMemcpy(dx.p, A, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dx.as<double>() + nm, b, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dth.p, theta, sizeof(double) * nth, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dsig.p, sigma, sizeof(double) * D, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dmt.p, m, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dst.p, S, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_model_energy: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    Batch bt{};
    bt.model = model; bt.method = ODE_EULER; bt.D = D; bt.N = N; bt.M = 0; bt.B = 1;
    bt.dt = dt_model; bt.dt_model = dt_model;
    bt.theta = dth.as<double>(); bt.sigma = dsig.as<double>(); bt.R = dsig.as<double>();
    bt.E0 = dzero.as<double>();
    Scratch sc{};
    sc.mt = dmt.as<double>(); sc.st = dst.as<double>(); sc.dEm = dgm.as<double>(); sc.dEs = dgs.as<double>();
    sc.esde_t = des.as<double>(); sc.status = dstat.as<int>();
    Extra ex{};
    ex.Efx = def.as<double>(); ex.Edf = dedf.as<double>(); ex.parts = dparts.as<double>();
    if (model == VGPA_MODEL_L96) launch_l96_energy(bt, sc, dx.as<double>(), 0, 0, 1, ex, nullptr);
    else launch_small_energy(bt, sc, dx.as<double>(), 0, 0, 1, ex, nullptr);
    launch_finalize(bt, sc, dF.as<double>(), 0, 1, ex, nullptr);
    if (hyper) {
        if ((e = dft.alloc(sizeof(double) * (size_t)N * nhth)) != cudaSuccess || (e = dfs.alloc(sizeof(double) * nv)) != cudaSuccess ||
            (e = dhth.alloc(sizeof(double) * nhth)) != cudaSuccess || (e = dhsig.alloc(sizeof(double) * nsig)) != cudaSuccess) {
            cleanup();
            g_create_error = std::string("CUDA error in vgpa_model_energy: ") + cudaGetErrorString(e);
            return VGPA_ECUDA;
        }
        // parts[1] = Esde (written by finalize on the same stream)
        launch_hyper(model, D, N, dt_model, dth.as<double>(), dsig.as<double>(), dx.as<double>(), dmt.as<double>(),
                     dst.as<double>(), dparts.as<double>() + 1, dft.as<double>(), dfs.as<double>(),
                     dhth.as<double>(), dhsig.as<double>(), dstat.as<int>(), nullptr);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    double parts[3] = {0, 0, 0};
    int stat = 0;
    if (e == cudaSuccess) e = cudaMemcpy(parts, dparts.p, sizeof parts, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&stat, dstat.p, sizeof stat, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && dEth) e = cudaMemcpy(dEth, dhth.p, sizeof(double) * nhth, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && dEsig) e = cudaMemcpy(dEsig, dhsig.p, sizeof(double) * nsig, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(Ef, def.p, sizeof(double) * nv, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(Edf, dedf.p, sizeof(double) * nm, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dEm, dgm.p, sizeof(double) * nv, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(dEs, dgs.p, sizeof(double) * nm, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA error in vgpa_model_energy: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    if (stat != 0) {
        char buf[128];
        snprintf(buf, sizeof buf, "Matrix is not positive definite: S(t) at time index %d", stat - 1);
        g_create_error = buf;
        return VGPA_ENOTPD;
    }
    *Esde = parts[1];
    return VGPA_OK;
}

int vgpa_obs_energy(int device, int D, int N, int M, const int64_t* obs_t, const double* obs_y, const double* R,
                    const double* mt, const double* st, double* Eobs, double* jm, double* js, double* dr)
{
    if ((D != 1 && D != 3 && D != 40) || N < 2 || M < 0 || M > N) { g_create_error = "Wrong dimensions for the likelihood"; return VGPA_EINVAL; }
    if ((M > 0 && (!obs_t || !obs_y)) || !R || !mt || !st || !Eobs || !jm || !js) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    for (int n = 0; n < M; ++n)
        if (obs_t[n] < 0 || obs_t[n] >= N) { g_create_error = "observation index out of range"; return VGPA_EINVAL; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const long long nv = (long long)N * D, nm = (long long)N * D * D;
    DevBuf dot, doy, dR, dmt, dst, djm, djs, des, dparts, dzero, dF, ddr;
    auto cleanup = [&]() { for (DevBuf* q : {&dot, &doy, &dR, &dmt, &dst, &djm, &djs, &des, &dparts, &dzero, &dF, &ddr}) q->release(); };
    std::vector<long long> ot(std::max(M, 1), 0);
    for (int n = 0; n < M; ++n) ot[n] = obs_t[n];
    cudaError_t e;
    if ((e = dot.alloc(sizeof(long long) * ot.size())) != cudaSuccess || (e = doy.alloc(sizeof(double) * std::max(M, 1) * D)) != cudaSuccess ||
        (e = dR.alloc(sizeof(double) * D)) != cudaSuccess || (e = dmt.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = dst.alloc(sizeof(double) * nm)) != cudaSuccess || (e = djm.alloc(sizeof(double) * nv)) != cudaSuccess ||
        (e = djs.alloc(sizeof(double) * nm)) != cudaSuccess || (e = des.alloc(sizeof(double) * N)) != cudaSuccess ||
        (e = dparts.alloc(sizeof(double) * 3)) != cudaSuccess || (e = dzero.alloc(sizeof(double))) != cudaSuccess ||
        (e = dF.alloc(sizeof(double))) != cudaSuccess ||
        (e = cudaMemset(dzero.p, 0, sizeof(double))) != cudaSuccess || (e = cudaMemset(des.p, 0, sizeof(double) * N)) != cudaSuccess ||
        (e = cudaMemcpy(dot.p, ot.data(), sizeof(long long) * ot.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (M > 0 && (e = cudaMemcpy(doy.p, obs_y, sizeof(double) * M * D, cudaMemcpyHostToDevice)) != cudaSuccess) ||
        (e = cudaMemcpy(dR.p, R, sizeof(double) * D, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dmt.p, mt, sizeof(double) * nv, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(dst.p, st, sizeof(double) * nm, cudaMemcpyHostToDevice)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_obs_energy: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    Batch bt{};
    bt.model = (D == 1) ? MODEL_OU : (D == 3 ? MODEL_L63 : MODEL_L96);
    bt.D = D; bt.N = N; bt.M = M; bt.B = 1; bt.dt = 1.0; bt.dt_model = 1.0;
    bt.sigma = dR.as<double>(); bt.R = dR.as<double>(); bt.E0 = dzero.as<double>();
    bt.obs_t = dot.as<long long>(); bt.obs_y = doy.as<double>();
    Scratch sc{};
    sc.mt = dmt.as<double>(); sc.st = dst.as<double>(); sc.esde_t = des.as<double>();
    Extra ex{};
    ex.parts = dparts.as<double>();
    launch_finalize(bt, sc, dF.as<double>(), 0, 1, ex, nullptr);
    launch_jump_tables(D, N, M, dot.as<long long>(), doy.as<double>(), dR.as<double>(), dmt.as<double>(),
                       djm.as<double>(), djs.as<double>(), nullptr);
    if (dr != nullptr && D == 1) {
        if ((e = ddr.alloc(sizeof(double) * N)) != cudaSuccess) {
            cleanup();
            g_create_error = std::string("CUDA error in vgpa_obs_energy: ") + cudaGetErrorString(e);
            return VGPA_ECUDA;
        }
        launch_obs_dr(N, M, dot.as<long long>(), doy.as<double>(), dR.as<double>(), dmt.as<double>(), dst.as<double>(),
                      ddr.as<double>(), nullptr);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    double parts[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(parts, dparts.p, sizeof parts, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && dr != nullptr) {
        if (D == 1) e = cudaMemcpy(dr, ddr.p, sizeof(double) * N, cudaMemcpyDeviceToHost);
        else std::fill(dr, dr + (size_t)N * M * M, 0.0);   // the reference's n-D branch returns zeros (gaussian_like.py:226)
    }
    if (e == cudaSuccess) e = cudaMemcpy(jm, djm.p, sizeof(double) * nv, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(js, djs.p, sizeof(double) * nm, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) {
        g_create_error = std::string("CUDA error in vgpa_obs_energy: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    *Eobs = parts[2];
    return VGPA_OK;
}

static int init_check(vgpa_handle* h, const double* x, int64_t x_stride)
{
    if (x == nullptr) return h->fail(VGPA_EINVAL, "x is NULL");
    if (x_stride < h->n_x) return h->fail(VGPA_EINVAL, "x_stride %lld < %lld", (long long)x_stride, h->n_x);
    if (h->d.M < 1) return h->fail(VGPA_EINVAL, "initialization needs at least one observation");
    return VGPA_OK;
}

int vgpa_initialization(vgpa_handle* h, double t0, double* d_x, int64_t x_stride, void* stream)
{
    if (!h) return VGPA_EINVAL;
    if (int rc = init_check(h, d_x, x_stride)) return rc;
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = h->d.B, D = h->d.D, n = h->d.M + 2;
    DevBuf scratch, err;
    cudaError_t e;
    if ((e = scratch.alloc(sizeof(double) * (size_t)B * D * 2 * n)) != cudaSuccess || (e = err.alloc(sizeof(int))) != cudaSuccess ||
        (e = cudaMemsetAsync(err.p, 0, sizeof(int), st)) != cudaSuccess) {
        scratch.release(); err.release();
        return h->cuda_fail(e, "vgpa_initialization");
    }
    launch_initialization(h->batch, 0, B, t0, scratch.as<double>(), d_x, x_stride, err.as<int>(), st);
    h->launches += 2;
    int bad = 0;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, err.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // the scratch is released below
    scratch.release(); err.release();
    if (e != cudaSuccess) return h->cuda_fail(e, "vgpa_initialization");
    if (bad) return h->fail(VGPA_EINVAL, "`x` must be strictly increasing sequence: an observation sits at the first or last grid index");
    return VGPA_OK;
}

int vgpa_initialization_host(vgpa_handle* h, double t0, double* x, int64_t x_stride)
{
    if (!h) return VGPA_EINVAL;
    if (int rc = init_check(h, x, x_stride)) return rc;
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    const int B = h->d.B, D = h->d.D, n = h->d.M + 2;
    const long long nx = h->n_x;
    const int C = std::max(1, std::min(B, 64));
    DevBuf scratch, err, dx;
    auto cleanup = [&]() { scratch.release(); err.release(); dx.release(); };
    cudaError_t e;
    if ((e = scratch.alloc(sizeof(double) * (size_t)C * D * 2 * n)) != cudaSuccess || (e = err.alloc(sizeof(int))) != cudaSuccess ||
        (e = dx.alloc(sizeof(double) * (size_t)C * nx)) != cudaSuccess || (e = cudaMemset(err.p, 0, sizeof(int))) != cudaSuccess) {
        cleanup();
        return h->cuda_fail(e, "vgpa_initialization_host");
    }
    for (int p0 = 0; p0 < B && e == cudaSuccess; p0 += C) {
        const int count = std::min(C, B - p0);
        // kernels index problems globally: offset the staging pointer accordingly
        launch_initialization(h->batch, p0, count, t0, scratch.as<double>(), dx.as<double>() - (long long)p0 * nx, nx,
                              err.as<int>(), h->s_comp);
        h->launches += 2;
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(x + (long long)p0 * x_stride, sizeof(double) * x_stride, dx.p, sizeof(double) * nx,
                                  sizeof(double) * nx, count, cudaMemcpyDeviceToHost, h->s_comp);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_comp);
    }
    int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&bad, err.p, sizeof(int), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return h->cuda_fail(e, "vgpa_initialization_host");
    if (bad) return h->fail(VGPA_EINVAL, "`x` must be strictly increasing sequence: an observation sits at the first or last grid index");
    return VGPA_OK;
}

// ---- batched sample paths and observations (datagen.cu) -------------------------
static int model_dim(int model) { return model == VGPA_MODEL_L96 ? 40 : (model == VGPA_MODEL_L63 ? 3 : 1); }
static int traj_ntheta(int model) { return model == VGPA_MODEL_L63 ? 3 : (model == VGPA_MODEL_OU ? 2 : 1); }

static int check_traj_args(int model, int N, int B, double dt, const void* theta, const void* sigma,
                           const void* x_init, const void* z, const void* path, int64_t z_stride, int64_t path_stride)
{
    if (model < 0 || model > 3) { g_create_error = "Unknown stochastic model"; return VGPA_EINVAL; }
    const int D = model_dim(model);
    if (N < 1 || B < 1) { g_create_error = "need N >= 1 time points and B >= 1 paths"; return VGPA_EINVAL; }
    if (!(dt > 0.0)) { g_create_error = "Discrete time step should be strictly positive"; return VGPA_EINVAL; }
    if (!theta || !sigma || !z || !path) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    if (D == 1 && !x_init) { g_create_error = "x_init is required for the 1-D models"; return VGPA_EINVAL; }
    if ((z_stride != 0 && z_stride < (int64_t)N * D) || (B > 1 && path_stride < (int64_t)N * D)) {
        g_create_error = "z_stride / path_stride shorter than one path"; return VGPA_EINVAL;
    }
    return VGPA_OK;
}

int vgpa_make_trajectory_device(int device, int model, int N, int B, double dt, const double* d_theta, int64_t theta_stride,
                                const double* d_sigma, int64_t sigma_stride, const double* d_x_init, int64_t x_init_stride,
                                const double* d_z, int64_t z_stride, double* d_path, int64_t path_stride, void* stream)
{
    if (int rc = check_traj_args(model, N, B, dt, d_theta, d_sigma, d_x_init, d_z, d_path, z_stride, path_stride)) return rc;
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    TrajArgs a{model, model_dim(model), N, B, dt, d_theta, theta_stride, d_sigma, sigma_stride, d_x_init, x_init_stride,
               d_z, z_stride, d_path, path_stride};
    launch_trajectories(a, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_create_error = std::string("CUDA error in vgpa_make_trajectory_device: ") + cudaGetErrorString(e); return VGPA_ECUDA; }
    return VGPA_OK;
}

// copy `rows` rows of `len` elements, `stride` apart (0 = one shared row), to a compact device buffer
static cudaError_t stage_rows(DevBuf& dst, const double* src, int64_t stride, int64_t len, int rows, long long* dstride)
{
    const int r = stride == 0 ? 1 : rows;
    cudaError_t e = dst.alloc(sizeof(double) * (size_t)len * r);
    if (e != cudaSuccess) return e;
    *dstride = stride == 0 ? 0 : len;
    if (r == 1 || stride == len) return cudaMemcpy(dst.p, src, sizeof(double) * (size_t)len * r, cudaMemcpyHostToDevice);
    return cudaMemcpy2D(dst.p, sizeof(double) * len, src, sizeof(double) * stride, sizeof(double) * len, r, cudaMemcpyHostToDevice);
}

int vgpa_make_trajectory(int device, int model, int N, int B, double dt, const double* theta, int64_t theta_stride,
                         const double* sigma, int64_t sigma_stride, const double* x_init, int64_t x_init_stride,
                         const double* z, int64_t z_stride, double* path, int64_t path_stride)
{
    if (int rc = check_traj_args(model, N, B, dt, theta, sigma, x_init, z, path, z_stride, path_stride)) return rc;
    if ((theta_stride != 0 && theta_stride < traj_ntheta(model)) || (sigma_stride != 0 && sigma_stride < model_dim(model)) ||
        (x_init && x_init_stride != 0 && x_init_stride < model_dim(model))) {
        g_create_error = "theta_stride / sigma_stride / x_init_stride shorter than one entry (0 = shared)";
        return VGPA_EINVAL;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const int D = model_dim(model);
    const int64_t len = (int64_t)N * D;
    DevBuf dth, dsg, dxi, dz, dp;
    auto cleanup = [&]() { for (DevBuf* q : {&dth, &dsg, &dxi, &dz, &dp}) q->release(); };
    TrajArgs a{};
    a.model = model; a.D = D; a.N = N; a.B = B; a.dt = dt;
    cudaError_t e;
    if ((e = stage_rows(dth, theta, theta_stride, traj_ntheta(model), B, &a.theta_stride)) != cudaSuccess ||
        (e = stage_rows(dsg, sigma, sigma_stride, D, B, &a.sigma_stride)) != cudaSuccess ||
        (x_init && (e = stage_rows(dxi, x_init, x_init_stride, D, B, &a.x_init_stride)) != cudaSuccess) ||
        (e = stage_rows(dz, z, z_stride, len, B, &a.z_stride)) != cudaSuccess ||
        (e = dp.alloc(sizeof(double) * (size_t)len * B)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_make_trajectory: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    a.theta = dth.as<double>(); a.sigma = dsg.as<double>(); a.x_init = x_init ? dxi.as<double>() : nullptr;
    a.z = dz.as<double>(); a.path = dp.as<double>(); a.path_stride = len;
    launch_trajectories(a, nullptr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaMemcpy2D(path, sizeof(double) * (B > 1 ? path_stride : len), dp.p, sizeof(double) * len, sizeof(double) * len, B,
                         cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) { g_create_error = std::string("CUDA error in vgpa_make_trajectory: ") + cudaGetErrorString(e); return VGPA_ECUDA; }
    return VGPA_OK;
}

static int check_obs_args(int D, int N, int M, int B, const int64_t* obs_t /* host, or NULL = not checked */)
{
    if (D != 1 && D != 3 && D != 40) { g_create_error = "unsupported state dimension (1, 3 or 40)"; return VGPA_EINVAL; }
    if (N < 1 || M < 0 || B < 1) { g_create_error = "need N >= 1, M >= 0, B >= 1"; return VGPA_EINVAL; }
    for (int j = 0; obs_t && j < M; ++j)
        if (obs_t[j] < 0 || obs_t[j] >= N) { g_create_error = "observation index outside the time grid"; return VGPA_EINVAL; }
    return VGPA_OK;
}

int vgpa_collect_obs_device(int device, int D, int N, int M, int B, const int64_t* d_obs_t, const double* d_R, int64_t R_stride,
                            const double* d_path, int64_t path_stride, const double* d_xi, int64_t xi_stride,
                            double* d_obs_y, int64_t obs_y_stride, void* stream)
{
    if (int rc = check_obs_args(D, N, M, B, nullptr)) return rc;
    if (!d_obs_t || !d_R || !d_path || !d_xi || !d_obs_y) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    ObsArgs a{D, N, M, B, reinterpret_cast<const long long*>(d_obs_t), d_R, R_stride, d_path, path_stride, d_xi, xi_stride,
              d_obs_y, obs_y_stride};
    launch_collect_obs(a, static_cast<cudaStream_t>(stream));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_create_error = std::string("CUDA error in vgpa_collect_obs_device: ") + cudaGetErrorString(e); return VGPA_ECUDA; }
    return VGPA_OK;
}

int vgpa_collect_obs(int device, int D, int N, int M, int B, const int64_t* obs_t, const double* R, int64_t R_stride,
                     const double* path, int64_t path_stride, const double* xi, int64_t xi_stride, double* obs_y,
                     int64_t obs_y_stride)
{
    if (!obs_t || !R || !path || !xi || !obs_y) { g_create_error = "NULL argument"; return VGPA_EINVAL; }
    if (int rc = check_obs_args(D, N, M, B, obs_t)) return rc;
    if (M == 0) return VGPA_OK;
    if ((R_stride != 0 && R_stride < D) || (path_stride != 0 && path_stride < (int64_t)N * D) ||
        (xi_stride != 0 && xi_stride < (int64_t)M * D) || (B > 1 && obs_y_stride < (int64_t)M * D)) {
        g_create_error = "R_stride / path_stride / xi_stride / obs_y_stride shorter than one entry (0 = shared)";
        return VGPA_EINVAL;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_create_error = "cudaSetDevice failed"; return VGPA_ECUDA; }
    const int64_t plen = (int64_t)N * D, olen = (int64_t)M * D;
    DevBuf dt_, dR, dp, dx, dy;
    auto cleanup = [&]() { for (DevBuf* q : {&dt_, &dR, &dp, &dx, &dy}) q->release(); };
    ObsArgs a{};
    a.D = D; a.N = N; a.M = M; a.B = B;
    cudaError_t e;
    if ((e = dt_.alloc(sizeof(int64_t) * M)) != cudaSuccess ||
        (e = cudaMemcpy(dt_.p, obs_t, sizeof(int64_t) * M, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = stage_rows(dR, R, R_stride, D, B, &a.R_stride)) != cudaSuccess ||
        (e = stage_rows(dp, path, path_stride, plen, B, &a.path_stride)) != cudaSuccess ||
        (e = stage_rows(dx, xi, xi_stride, olen, B, &a.xi_stride)) != cudaSuccess ||
        (e = dy.alloc(sizeof(double) * (size_t)olen * B)) != cudaSuccess) {
        cleanup();
        g_create_error = std::string("CUDA error in vgpa_collect_obs: ") + cudaGetErrorString(e);
        return VGPA_ECUDA;
    }
    a.obs_t = dt_.as<long long>(); a.R = dR.as<double>(); a.path = dp.as<double>(); a.xi = dx.as<double>();
    a.obs_y = dy.as<double>(); a.obs_y_stride = olen;
    launch_collect_obs(a, nullptr);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess)
        e = cudaMemcpy2D(obs_y, sizeof(double) * (B > 1 ? obs_y_stride : olen), dy.p, sizeof(double) * olen, sizeof(double) * olen,
                         B, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) { g_create_error = std::string("CUDA error in vgpa_collect_obs: ") + cudaGetErrorString(e); return VGPA_ECUDA; }
    return VGPA_OK;
}

void* vgpa_host_alloc(int64_t bytes)
{
    void* p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void vgpa_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// Host copy on several threads: what a caller with an ordinary (pageable) x uses to fill a page-locked
// staging buffer; one core copies ~10 GB/s, which for a 13 MB x is a tenth of a single-problem evaluation.
void vgpa_host_copy(void* dst, const void* src, int64_t bytes, int threads)
{
    if (!dst || !src || bytes <= 0) return;
    const int64_t min_part = 1 << 20;
    int n = (int)std::min<int64_t>(std::max(threads, 1), (bytes + min_part - 1) / min_part);
    if (n <= 1) {
        std::memcpy(dst, src, (size_t)bytes);
        return;
    }
    const int64_t part = ((bytes + n - 1) / n + 63) & ~int64_t(63);
    std::vector<std::thread> pool;
    pool.reserve(n - 1);
    for (int i = 1; i < n; ++i) {
        const int64_t off = part * i;
        if (off >= bytes) break;
        const int64_t len = std::min(part, bytes - off);
        pool.emplace_back([=]() { std::memcpy((char*)dst + off, (const char*)src + off, (size_t)len); });
    }
    std::memcpy(dst, src, (size_t)std::min(part, bytes));
    for (auto& t : pool) t.join();
}

int vgpa_host_equal(const void* a, const void* b, int64_t bytes, int threads)
{
    if (bytes <= 0) return 1;
    if (!a || !b) return 0;
    const int64_t min_part = 1 << 20;
    const int n = (int)std::min<int64_t>(std::max(threads, 1), (bytes + min_part - 1) / min_part);
    if (n <= 1) return std::memcmp(a, b, (size_t)bytes) == 0;
    const int64_t part = ((bytes + n - 1) / n + 63) & ~int64_t(63);
    std::vector<int> same(n, 1);
    std::vector<std::thread> pool;
    pool.reserve(n - 1);
    for (int i = 1; i < n; ++i) {
        const int64_t off = part * i;
        if (off >= bytes) break;
        const int64_t len = std::min(part, bytes - off);
        int* out = &same[i];
        pool.emplace_back([=]() { *out = std::memcmp((const char*)a + off, (const char*)b + off, (size_t)len) == 0; });
    }
    same[0] = std::memcmp(a, b, (size_t)std::min(part, bytes)) == 0;
    for (auto& t : pool) t.join();
    for (int v : same)
        if (!v) return 0;
    return 1;
}

int vgpa_set_timing(vgpa_handle* h, int enable)
{
    if (!h) return VGPA_EINVAL;
    h->timing = enable != 0;
    return VGPA_OK;
}

int vgpa_get_timing(vgpa_handle* h, double ms[4], int64_t launches[4])
{
    if (!h || !ms || !launches) return VGPA_EINVAL;
    CK(cudaSetDevice(h->d.device), "cudaSetDevice");
    CK(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    for (int k = 0; k < 4; ++k) {
        for (size_t q = 0; q < h->tev_used[k]; ++q) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, h->tev[k][q].first, h->tev[k][q].second) == cudaSuccess) {
                h->t_ms[k] += t;
                h->t_n[k] += 1;
            }
        }
        h->tev_used[k] = 0;
        ms[k] = h->t_ms[k];
        launches[k] = h->t_n[k];
        h->t_ms[k] = 0;
        h->t_n[k] = 0;
    }
    return VGPA_OK;
}

int64_t vgpa_launch_count(const vgpa_handle* h) { return h ? h->launches : 0; }
int64_t vgpa_chunk_size(const vgpa_handle* h) { return h ? h->chunk : 0; }
int64_t vgpa_scratch_in_use(const vgpa_handle* h)
{
    if (!h) return 0;
    return (int64_t)(h->sc_mt.bytes + h->sc_st.bytes + h->sc_dEm.bytes + h->sc_dEs.bytes + h->sc_esde.bytes +
                     h->sc2_mt.bytes + h->sc2_st.bytes + h->sc2_dEm.bytes + h->sc2_dEs.bytes + h->sc2_esde.bytes);
}

}  // extern "C"
