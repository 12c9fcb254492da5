// device.cu -- device selection and memory pool (process-wide) and the per-thread launch context:
// the thread's stream, device scalars with a pinned mirror, growable scratch.
#include "device.hpp"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>

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
        // stream-ordered free on the calling thread's stream (whoever used the buffer last has been waited
        // for: Context::adopt); a thread that never used the library frees on its per-thread default stream
        Context *c = Context::peek();
        if (cudaFreeAsync(p, c ? c->stream : cudaStreamPerThread) != cudaSuccess) (void)cudaGetLastError();
    }
    p = nullptr;
    cap = 0;
}

// ---- process-wide state: the device, its memory pool -------------------------------------------------
static std::mutex g_proc_mutex;          // guards the three below and the registry of live contexts
static int g_requested_device = -1;
static int g_device = -1;                // bound by the first context; every thread uses the same device
static int g_sm_count = 148;
static uint64_t g_next_ctx_id = 1;
static std::unordered_map<uint64_t, cudaStream_t> g_live_streams;  // context id -> its stream

void set_requested_device(int dev) {
    std::lock_guard<std::mutex> lk(g_proc_mutex);
    if (g_device >= 0 && g_device != dev) throw_internal("device already selected for this process");
    g_requested_device = dev;
}

// called with g_proc_mutex held, once
static void bind_device() {
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
    cudaMemPool_t pool;
    SM_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep_all = ~0ull;
    SM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all));
    {
        // Reserve the pool's working set once (default 4 GiB of the 180 GB, SMB200_POOL_PREWARM_MB to
        // change): getting fresh physical memory into the pool costs milliseconds per call, which
        // would otherwise land inside whichever batch first grows a buffer.
        size_t mb = 4096;
        if (const char *e2 = getenv("SMB200_POOL_PREWARM_MB")) mb = (size_t)strtoull(e2, nullptr, 10);
        size_t free_b = 0, total_b = 0;
        if (mb && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > (mb << 20) * 2) {
            void *warm = nullptr;
            if (cudaMallocAsync(&warm, mb << 20, cudaStreamPerThread) == cudaSuccess) cudaFreeAsync(warm, cudaStreamPerThread);
            (void)cudaGetLastError();
            SM_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
        }
    }
    g_sm_count = prop.multiProcessorCount;
    g_device = dev;
}

// ---- per-thread context ------------------------------------------------------------------------------
namespace {
struct ContextHolder {
    Context *c = nullptr;
    ~ContextHolder() {
        Context *dead = c;
        c = nullptr;
        delete dead;
    }
};
thread_local ContextHolder t_ctx;
}  // namespace

Context *Context::peek() { return t_ctx.c; }

Context &Context::get() {
    if (Context *c = t_ctx.c) {
        // the library's device must be current on whatever thread calls in
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != c->device) SM_CUDA(cudaSetDevice(c->device));
        return *c;
    }
    {
        std::lock_guard<std::mutex> lk(g_proc_mutex);
        if (g_device < 0) bind_device();
    }
    SM_CUDA(cudaSetDevice(g_device));
    std::unique_ptr<Context> c(new Context());
    c->device = g_device;
    c->sm_count = g_sm_count;
    SM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    SM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 6; i++) {
        SM_CUDA(cudaStreamCreateWithFlags(&c->k_streams[i], cudaStreamNonBlocking));
        SM_CUDA(cudaEventCreateWithFlags(&c->k_events[i], cudaEventDisableTiming));
    }
    SM_CUDA(cudaEventCreateWithFlags(&c->prep_event, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        SM_CUDA(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
        SM_CUDA(cudaEventCreateWithFlags(&c->ev_free[k], cudaEventDisableTiming));
    }
    SM_CUDA(cudaMalloc(&c->d_scalars, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(c->d_scalars, 0, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&c->h_scalars, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&c->h_fetch, 2 * sizeof(unsigned long long)));
    {
        std::lock_guard<std::mutex> lk(g_proc_mutex);
        c->id = g_next_ctx_id++;
        g_live_streams[c->id] = c->stream;
    }
    t_ctx.c = c.release();
    return *t_ctx.c;
}

// Thread exit (or process exit for the main thread): everything this thread queued has to be done
// before its stream goes away, because handles it touched may live on in other threads.
Context::~Context() {
    if (id) {
        std::lock_guard<std::mutex> lk(g_proc_mutex);
        g_live_streams.erase(id);
    }
    if (!stream || cudaStreamSynchronize(stream) != cudaSuccess) {  // runtime already torn down: nothing to give back
        (void)cudaGetLastError();
        ascii.p = offsets.p = packed.p = sort_tmp_k.p = sort_tmp_v.p = scan_tmp.p = find_cmat.p = find_bits.p = find_rows.p = nullptr;
        for (DevBuf &b : misc) b.p = nullptr;
        for (DevBuf &b : join) b.p = nullptr;
        return;
    }
    DevBuf *all[] = {&ascii, &offsets, &packed, &sort_tmp_k, &sort_tmp_v, &scan_tmp, &find_cmat, &find_bits, &find_rows};
    for (DevBuf *b : all) { if (b->p) cudaFreeAsync(b->p, stream); b->p = nullptr; b->cap = 0; }
    for (DevBuf &b : misc) { if (b.p) cudaFreeAsync(b.p, stream); b.p = nullptr; b.cap = 0; }
    for (DevBuf &b : join) { if (b.p) cudaFreeAsync(b.p, stream); b.p = nullptr; b.cap = 0; }
    cudaStreamSynchronize(stream);
    for (cudaEvent_t e : chunk_events) cudaEventDestroy(e);
    for (int i = 0; i < 6; i++) {
        if (k_streams[i]) cudaStreamDestroy(k_streams[i]);
        if (k_events[i]) cudaEventDestroy(k_events[i]);
    }
    if (prep_event) cudaEventDestroy(prep_event);
    for (int k = 0; k < 2; k++) {
        if (ev_done[k]) cudaEventDestroy(ev_done[k]);
        if (ev_free[k]) cudaEventDestroy(ev_free[k]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (d_scalars) cudaFree(d_scalars);
    if (h_scalars) cudaFreeHost(h_scalars);
    if (h_fetch) cudaFreeHost(h_fetch);
    cudaStreamDestroy(stream);
    (void)cudaGetLastError();
}

void Context::adopt(StreamOwner &o) {
    if (o.ctx_id == id) return;
    if (o.ctx_id != 0) {
        // held across the wait so that the other thread's context cannot be destroyed under it
        std::lock_guard<std::mutex> lk(g_proc_mutex);
        auto it = g_live_streams.find(o.ctx_id);
        if (it != g_live_streams.end()) SM_CUDA(cudaStreamSynchronize(it->second));  // (a dead context synchronised on exit)
    }
    o.ctx_id = id;
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
std::mutex g_prof_mutex;  // threads time their launches into the same table
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
    {
        std::lock_guard<std::mutex> lk(g_prof_mutex);
        e0 = prof_event();
        e1 = prof_event();
    }
    cudaEventRecord(e0, st);
}
ProfScope::~ProfScope() {
    if (!e0) return;
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    g_prof.pending[kind].emplace_back(e0, e1);
}
void prof_enable(bool on) { g_prof.enabled = on; }
bool prof_enabled() { return g_prof.enabled; }
void prof_read(int kind, double *ms, uint64_t *launches, bool reset) {
    if (kind < 0 || kind >= PROF_KINDS) throw_internal("bad profile kind");
    std::lock_guard<std::mutex> lk(g_prof_mutex);
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
