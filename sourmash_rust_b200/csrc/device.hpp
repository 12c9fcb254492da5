// Device plumbing shared by every translation unit: error type, CUDA checks,
// RAII device buffers, the per-process launch context (stream + scratch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace smb200 {

// Error codes of the reference (src/errors.rs:28-50, include/sourmash.h:10-27).
enum ErrorCode : uint32_t {
    ERR_NO_ERROR = 0,
    ERR_PANIC = 1,
    ERR_INTERNAL = 2,
    ERR_MSG = 3,
    ERR_UNKNOWN = 4,
    ERR_MISMATCH_KSIZES = 101,
    ERR_MISMATCH_DNAPROT = 102,
    ERR_MISMATCH_MAXHASH = 103,
    ERR_MISMATCH_SEED = 104,
    ERR_INVALID_DNA = 1101,
    ERR_INVALID_PROT = 1102,
    ERR_IO = 100001,
    ERR_UTF8 = 100002,
    ERR_PARSE_INT = 100003,
    ERR_SERDE = 100004,
};

// Thrown by the host classes; the C ABI turns it into the thread-local last
// error (reference: utils.rs:154-166 landingpad).
struct SourmashError : public std::runtime_error {
    uint32_t code;
    SourmashError(uint32_t c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

[[noreturn]] inline void throw_internal(const std::string &msg) {
    throw SourmashError(ERR_INTERNAL, "internal error: " + msg);
}

#define SM_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::smb200::throw_internal(std::string("CUDA: ") + cudaGetErrorString(_e) + " at " +     \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");  \
        }                                                                                          \
    } while (0)

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
extern std::atomic<uint64_t> g_launch_count;
// SMB200_DEBUG_SYNC=1 in the environment: synchronise after every launch so that a faulting
// kernel is reported at its own launch site
extern bool g_debug_sync;
#define SM_LAUNCHED()                                       \
    do {                                                    \
        ::smb200::g_launch_count.fetch_add(1);              \
        SM_CUDA(cudaGetLastError());                        \
        if (::smb200::g_debug_sync) SM_CUDA(cudaDeviceSynchronize()); \
    } while (0)

// Growable raw device allocation (never shrinks; contents not preserved on grow
// unless keep=true).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release();
    void reserve(size_t bytes, cudaStream_t st = nullptr, bool keep = false, size_t keep_bytes = 0);
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Which host thread's context last queued device work on an object (a sketch, a collection, a bloom
// filter).  The reference's contract is "a handle may be used from any thread, but not concurrently"
// (SURVEY 8(b) Threading): when an object turns up on another thread, that thread first waits for the
// stream the object was last used on (Context::adopt), then carries on with its own.
struct StreamOwner {
    uint64_t ctx_id = 0;
};

// One per HOST THREAD (all on the process's one device): the stream every kernel launched from that
// thread goes to, small device scalars, and reusable scratch.  Distinct handles driven from distinct
// threads therefore share nothing but the device's memory pool -- no lock is held around a call
// (utils.rs:14-16: the reference's only per-thread state is the error slot; here it is also this).
struct Context {
    int device = -1;
    uint64_t id = 0;  // never reused; 0 = none
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    // device scalars: [0] candidate counter, [1] first-invalid base, [2] overflow flag, [3..] misc
    unsigned long long *d_scalars = nullptr;
    unsigned long long *h_scalars = nullptr;  // pinned mirror
    // scratch
    DevBuf ascii, offsets, packed, sort_tmp_k, sort_tmp_v, scan_tmp, misc[8], join[8];
    // find_stream.cu: count matrix + touched-row bitmap that every search leaves all-zero again (the hit pass clears what
    // it reads), so that a search does not start with a memset of half a gigabyte; find_clean: that invariant holds
    DevBuf find_cmat, find_bits, find_rows;
    bool find_clean = false;
    std::vector<cudaEvent_t> chunk_events;

    cudaStream_t copy_stream = nullptr;  // host->device staging, overlapped with `stream`
    cudaStream_t k_streams[6] = {nullptr};  // one per sketch of a batch call: their kernels fill each other's tails
    cudaEvent_t k_events[6] = {nullptr};
    cudaEvent_t prep_event = nullptr;

    static Context &get();      // this thread's context; creates on first use; throws if no CUDA device
    static Context *peek();     // nullptr before this thread's first use
    void adopt(StreamOwner &o); // `o` was last used on another thread: wait for that thread's stream
    ~Context();
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};  // compare_matrix's read-back pipeline
    void sync() { SM_CUDA(cudaStreamSynchronize(stream)); }
    void set_scalar(int idx, unsigned long long v);  // synchronous
    void read_scalars();                              // d_scalars -> h_scalars, synchronous
    void fetch2(const void *a, const void *b, uint64_t out[2]);  // two device u64 -> host, synchronous, no copy engine
    unsigned long long *h_fetch = nullptr;
    unsigned long long *dsc(int idx) { return d_scalars + idx; }
};
// Per-kernel device timing (CUDA events on the launching stream), off unless enabled through
// smgpu_profile_enable: bench.py's roofline numbers come from here.
enum { PROF_SKETCH_K21 = 0, PROF_SKETCH_K31 = 1, PROF_SKETCH_K51 = 2, PROF_SKETCH_OTHER = 3, PROF_COMPARE = 4,
       PROF_SORT = 5, PROF_SKETCH_MULTI = 6, PROF_FIND = 7, PROF_WALK = 8, PROF_PROBE = 9, PROF_FILL = 10, PROF_KINDS = 11 };
struct ProfScope {
    int kind;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(int kind, cudaStream_t st);
    ~ProfScope();
};
void prof_enable(bool on);
bool prof_enabled();
void prof_read(int kind, double *ms, uint64_t *launches, bool reset);  // synchronises the recorded events
// which device the context binds to when it is created (default: the thread's current device)
void set_requested_device(int dev);

enum { SC_FIRST_BAD = 0, SC_NUNIQ = 1, SC_TMAX = 2, SC_CNT = 3, SC_FLAG = 4, SC_LEN = 5, SC_PAIR0 = 6, SC_PAIR1 = 7,
       SC_PAIR2 = 8, SC_THRESH = 9, SC_CAND0 = 10 /* .. SC_CAND0+5: per-handle candidate counters of one batch */,
       SC_TOUCHED = 16 /* find_stream: rows holding a count; zero between searches */,
       SC_SPILL = 17 /* find_stream: hashes written out for the deferred lookup */, SC_COUNT = 24 };

}  // namespace smb200
