// device.cu -- process-wide launch context: device selection, the library's stream, device
// scalars with a pinned mirror, growable scratch.
#include "device.hpp"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace smb200 {

std::atomic<uint64_t> g_launch_count{0};
bool g_debug_sync = getenv("SMB200_DEBUG_SYNC") != nullptr;

void DevBuf::reserve(size_t bytes, cudaStream_t st, bool keep, size_t keep_bytes) {
    if (bytes <= cap) return;
    size_t want = cap ? cap : 4096;
    while (want < bytes) want += want / 2 + 4096;
    want = (want + 255) / 256 * 256;
    void *np = nullptr;
    SM_CUDA(cudaMalloc(&np, want));
    if (keep && p && keep_bytes) {
        SM_CUDA(cudaMemcpyAsync(np, p, keep_bytes, cudaMemcpyDeviceToDevice, st));
        SM_CUDA(cudaStreamSynchronize(st));
    }
    if (p) cudaFree(p);  // cudaFree synchronises the device: no kernel still reads the old block
    p = np;
    cap = want;
}

void PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    SM_CUDA(cudaMallocHost(&p, bytes));
    cap = bytes;
}

static int g_requested_device = -1;
static std::mutex g_ctx_mutex;
static Context *g_ctx = nullptr;

void set_requested_device(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    if (g_ctx && g_ctx->device != dev) throw_internal("device already selected for this process");
    g_requested_device = dev;
}

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
    SM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    c->own_stream = true;
    SM_CUDA(cudaMalloc(&c->d_scalars, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMemset(c->d_scalars, 0, SC_COUNT * sizeof(unsigned long long)));
    SM_CUDA(cudaMallocHost(&c->h_scalars, SC_COUNT * sizeof(unsigned long long)));
    g_ctx = c;
    return *g_ctx;
}

void Context::set_scalar(int idx, unsigned long long v) {
    h_scalars[idx] = v;  // pinned: the async copy reads it when it executes, so wait for it
    SM_CUDA(cudaMemcpyAsync(d_scalars + idx, h_scalars + idx, sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
}
void Context::read_scalars() {
    SM_CUDA(cudaMemcpyAsync(h_scalars, d_scalars, SC_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    SM_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace smb200
