// device.cu -- process-wide launch context: device selection, the library's stream, device
// scalars with a pinned mirror, growable scratch.
#include "device.hpp"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace smb200 {

std::atomic<uint64_t> g_launch_count{0};
bool g_debug_sync = getenv("SMB200_DEBUG_SYNC") != nullptr;

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) on the library's
// stream: growth and release are ordered with the kernels instead of synchronising the device,
// and freed blocks stay cached in the pool (release threshold = max), so the scratch churn of the
// sort/merge pipeline costs microseconds, not cudaMalloc/cudaFree round trips.
void DevBuf::reserve(size_t bytes, cudaStream_t, bool keep, size_t keep_bytes) {
    if (bytes <= cap) return;
    cudaStream_t st = Context::get().stream;
    size_t want = cap ? cap : 4096;
    while (want < bytes) want += want / 2 + 4096;
    want = (want + 255) / 256 * 256;
    void *np = nullptr;
    SM_CUDA(cudaMallocAsync(&np, want, st));
    if (keep && p && keep_bytes) SM_CUDA(cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, st));
    if (p) SM_CUDA(cudaFreeAsync(p, st));
    p = np;
    cap = want;
}
void DevBuf::release() {
    if (p) {
        Context *c = Context::peek();
        if (c) cudaFreeAsync(p, c->stream); else cudaFree(p);
    }
    p = nullptr;
    cap = 0;
}

static int g_requested_device = -1;
static std::mutex g_ctx_mutex;
static Context *g_ctx = nullptr;

void set_requested_device(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    if (g_ctx && g_ctx->device != dev) throw_internal("device already selected for this process");
    g_requested_device = dev;
}

Context *Context::peek() { return g_ctx; }

Context &Context::get() {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    if (g_ctx) {
        // the library's device must be current on whatever thread calls in
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != g_ctx->device) SM_CUDA(cudaSetDevice(g_ctx->device));
        return *g_ctx;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        throw_internal(std::string("no usable CUDA device (") + (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
                       "); this library has no CPU path");
    }
    int dev = g_requested_device;
    if (dev < 0) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    }
    if (dev >= count) throw_internal("requested CUDA device does not exist");
    SM_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    SM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) throw_internal("this build targets sm_100a (B200); found compute capability " +
                                        std::to_string(prop.major) + "." + std::to_string(prop.minor));
    Context *c = new Context();
    c->device = dev;
    c->sm_count = prop.multiProcessorCount;
    {
        cudaMemPool_t pool;
        SM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep_all = ~0ull;
        SM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
    }
    SM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 6; i++) {
        SM_CUDA(cudaStreamCreateWithFlags(&c->k_streams[i], cudaStreamNonBlocking));
        SM_CUDA(cudaEventCreateWithFlags(&c->k_events[i], cudaEventDisableTiming));
    }
    SM_CUDA(cudaEventCreateWithFlags(&c->prep_event, cudaEventDisableTiming));
    c->own_stream = true;
    {
        // Reserve the pool's working set once (default 4 GiB of the 180 GB, SMB200_POOL_PREWARM_MB to
        // change): getting fresh physical memory into the pool costs milliseconds per call, which
        // would otherwise land inside whichever batch first grows a buffer.
        size_t mb = 4096;
        if (const char *e = getenv("SMB200_POOL_PREWARM_MB")) mb = (size_t)strtoull(e, nullptr, 10);
        size_t free_b = 0, total_b = 0;
        if (mb && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > (mb << 20) * 2) {
            void *warm = nullptr;
            if (cudaMallocAsync(&warm, mb << 20, c->stream) == cudaSuccess) cudaFreeAsync(warm, c->stream);
            (void)cudaGetLastError();
            SM_CUDA(cudaStreamSynchronize(c->stream));
        }
    }
    SM_CUDA(cudaMalloc(&c->d_scalars, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(c->d_scalars, 0, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&c->h_scalars, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&c->h_fetch, 2 * sizeof(unsigned long long)));
    g_ctx = c;
    return *g_ctx;
}

// ---- per-kernel timing ---------------------------------------------------------------------------
namespace {
struct ProfState {
    bool enabled = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[PROF_KINDS];
    std::vector<cudaEvent_t> pool;
    double ms[PROF_KINDS] = {0};
    uint64_t launches[PROF_KINDS] = {0};
} g_prof;
cudaEvent_t prof_event() {
    if (!g_prof.pool.empty()) {
        cudaEvent_t e = g_prof.pool.back();
        g_prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    SM_CUDA(cudaEventCreate(&e));
    return e;
}
}  // namespace
ProfScope::ProfScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_) {
    if (!g_prof.enabled) return;
    e0 = prof_event();
    e1 = prof_event();
    cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
    if (!e0) return;
    cudaEventRecord(e1, st);
    g_prof.pending[kind].emplace_back(e0, e1);
}
void prof_enable(bool on) { g_prof.enabled = on; }
bool prof_enabled() { return g_prof.enabled; }
void prof_read(int kind, double *ms, uint64_t *launches, bool reset) {
    if (kind < 0 || kind >= PROF_KINDS) throw_internal("bad profile kind");
    for (auto &pr : g_prof.pending[kind]) {
        SM_CUDA(cudaEventSynchronize(pr.second));
        float t = 0;
        SM_CUDA(cudaEventElapsedTime(&t, pr.first, pr.second));
        g_prof.ms[kind] += t;
        g_prof.launches[kind]++;
        g_prof.pool.push_back(pr.first);
        g_prof.pool.push_back(pr.second);
    }
    g_prof.pending[kind].clear();
    if (ms) *ms = g_prof.ms[kind];
    if (launches) *launches = g_prof.launches[kind];
    if (reset) { g_prof.ms[kind] = 0; g_prof.launches[kind] = 0; }
}

void Context::set_scalar(int idx, unsigned long long v) {
    h_scalars[idx] = v;  // pinned: the async copy reads it when it executes, so wait for it
    SM_CUDA(cudaMemcpyAsync(d_scalars + idx, h_scalars + idx, sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
}
// Small device -> host reads go through a one-warp kernel that stores into pinned (UVA-mapped) host
// memory instead of a cudaMemcpy: a memcpy queues on the device-to-host copy engine behind whatever
// large read-back is in flight on another stream (the compare matrix), a kernel store does not.
__global__ void scalars_to_host_kernel(const unsigned long long *src, unsigned long long *dst, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__global__ void fetch2_kernel(const unsigned long long *a, const unsigned long long *b, unsigned long long *dst) {
    if (threadIdx.x == 0) dst[0] = *a;
    if (threadIdx.x == 1) dst[1] = *b;
}
void Context::read_scalars() {
    scalars_to_host_kernel<<<1, 32, 0, stream>>>(d_scalars, h_scalars, SC_COUNT);
    SM_LAUNCHED();
    SM_CUDA(cudaStreamSynchronize(stream));
}
void Context::fetch2(const void *a, const void *b, uint64_t out[2]) {
    fetch2_kernel<<<1, 32, 0, stream>>>(static_cast<const unsigned long long *>(a), static_cast<const unsigned long long *>(b),
                                        h_fetch);
    SM_LAUNCHED();
    SM_CUDA(cudaStreamSynchronize(stream));
    out[0] = h_fetch[0];
    out[1] = h_fetch[1];
}

}  // namespace smb200
