// sortops.cu -- what replaces the reference's sorted-Vec bookkeeping inside add_hash
// (binary_search + Vec::insert + pop, src/lib.rs:212-242) for a whole batch of survivor
// hashes at once: LSD radix sort, run-length reduce (distinct hashes + abundance sums),
// prefix scan, plus the small helpers of the num+abundance corner case and the ordered
// replay used for non-standard parameter combinations.
#include <algorithm>

#include "device.hpp"
#include "kernels.cuh"
#include "murmur3.cuh"

namespace smb200 {

// =====================================================================================
// fill / filter / check
// =====================================================================================
__global__ void fill_u64_kernel(uint64_t *p, uint64_t v, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}
static unsigned grid_for(uint64_t n, unsigned per_block, unsigned cap = 148 * 32) {
    uint64_t b = (n + per_block - 1) / per_block;
    if (b == 0) b = 1;
    return (unsigned)(b > cap ? cap : b);
}
void launch_fill_u64(uint64_t *p, uint64_t v, uint64_t n, cudaStream_t st) {
    if (!n) return;
    fill_u64_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, v, n);
    SM_LAUNCHED();
}

// add_hash's state-independent gate (lib.rs:198) and the current threshold, over raw hashes
__global__ void filter_hashes_kernel(const uint64_t *__restrict__ hashes, uint64_t n, uint64_t max_hash,
                                     const uint64_t *__restrict__ thr_ptr, uint64_t *out_hash,
                                     uint64_t *out_pos, unsigned long long *counter) {
    const uint64_t thr = *thr_ptr;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const uint64_t n_round = ((n + 31) / 32) * 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        uint64_t h = 0;
        bool pass = false;
        if (i < n) {
            h = hashes[i];
            pass = (h <= max_hash || max_hash == 0) && h <= thr;
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pass);
        if (bal) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (pass) {
                const unsigned long long idx = base + __popc(bal & ((1u << lane) - 1u));
                out_hash[idx] = h;
                if (out_pos) out_pos[idx] = i;
            }
        }
    }
}
void launch_filter_hashes(const uint64_t *hashes, uint64_t n, uint64_t max_hash, const uint64_t *thr,
                          uint64_t *out_hash, uint64_t *out_pos, unsigned long long *counter, cudaStream_t st) {
    if (!n) return;
    filter_hashes_kernel<<<grid_for(n, 256), 256, 0, st>>>(hashes, n, max_hash, thr, out_hash, out_pos, counter);
    SM_LAUNCHED();
}

__global__ void check_sorted_kernel(const uint64_t *keys, uint64_t n, unsigned long long *flag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += stride)
        if (!(keys[i - 1] < keys[i])) *flag = 1;
}
void launch_check_sorted(const uint64_t *keys, uint64_t n, unsigned long long *flag, cudaStream_t st) {
    if (n < 2) return;
    check_sorted_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, n, flag);
    SM_LAUNCHED();
}

// =====================================================================================
// exclusive scan (u64): 2048 elements per block, recursive over block totals
// =====================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_kernel(const uint64_t *in,
                                                                  uint64_t *out, uint64_t n,
                                                                  uint64_t *__restrict__ block_sums) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        sum += v[j];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t warp_off = 0;
    for (int w = 0; w < warp; w++) warp_off += s_warp[w];
    uint64_t run = warp_off + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (block_sums && threadIdx.x == SCAN_THREADS - 1) block_sums[blockIdx.x] = run;
}
__global__ void scan_add_kernel(uint64_t *out, uint64_t n, const uint64_t *__restrict__ block_offs) {
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
    const uint64_t off = block_offs[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        const uint64_t k = i + (uint64_t)j * SCAN_THREADS;
        if (k < n) out[k] += off;
    }
}
size_t scan_tmp_bytes(uint64_t n) {
    size_t total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += ((n + 1) * sizeof(uint64_t) + 255) / 256 * 256;
    }
    return total + 256;
}
void scan_exclusive_u64(const uint64_t *in, uint64_t *out, uint64_t n, void *tmp, cudaStream_t st) {
    if (n == 0) return;
    const uint64_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (blocks == 1) {
        scan_block_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr);
        SM_LAUNCHED();
        return;
    }
    uint64_t *sums = reinterpret_cast<uint64_t *>(tmp);
    void *next_tmp = reinterpret_cast<char *>(tmp) + ((blocks + 1) * sizeof(uint64_t) + 255) / 256 * 256;
    scan_block_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, out, n, sums);
    SM_LAUNCHED();
    scan_exclusive_u64(sums, sums, blocks, next_tmp, st);
    scan_add_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(out, n, sums);
    SM_LAUNCHED();
}

// =====================================================================================
// LSD radix sort, 8-bit digits.  Per pass: per-block digit histogram -> scan of the
// digit-major count table -> stable scatter (warp-private running counters + match_any).
// =====================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // 4096 keys per block
constexpr int RS_WARP_KEYS = 32 * RS_ROUNDS;        // contiguous keys owned by one warp

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                             int shift, uint64_t *__restrict__ counts,
                                                             uint32_t n_blocks) {
    __shared__ uint32_t s_cnt[256];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_cnt[(keys[i] >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    counts[(uint64_t)threadIdx.x * n_blocks + blockIdx.x] = s_cnt[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t *__restrict__ keys,
                                                                const uint64_t *__restrict__ vals, uint64_t n,
                                                                int shift, const uint64_t *__restrict__ offsets,
                                                                uint32_t n_blocks, uint64_t *__restrict__ out_keys,
                                                                uint64_t *__restrict__ out_vals) {
    __shared__ uint32_t s_cnt[RS_WARPS][256];
    __shared__ uint64_t s_base[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < RS_WARPS * 256; j += RS_THREADS) (&s_cnt[0][0])[j] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)warp * RS_WARP_KEYS;
    uint64_t k[RS_ROUNDS];
    // pass A: per-warp digit counts (keys stay in registers)
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint64_t i = wbase + r * 32 + lane;
        const bool in = i < n;
        k[r] = in ? keys[i] : ~0ull;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 0xFF;
        const unsigned act = __ballot_sync(0xFFFFFFFFu, in);
        if (in) {
            const unsigned peers = __match_any_sync(act, d);
            if (lane == __ffs(peers) - 1) s_cnt[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    // digit d: exclusive prefix over the warps of this block, on top of the global offset
    {
        const int d = threadIdx.x;
        uint64_t run = offsets[(uint64_t)d * n_blocks + blockIdx.x];
        s_base[d] = run;
        uint32_t acc = 0;
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t t = s_cnt[w][d];
            s_cnt[w][d] = acc;
            acc += t;
        }
    }
    __syncthreads();
    // pass B: stable ranks, scatter
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const uint64_t i = wbase + r * 32 + lane;
        const bool in = i < n;
        const uint32_t d = (uint32_t)(k[r] >> shift) & 0xFF;
        const unsigned act = __ballot_sync(0xFFFFFFFFu, in);
        if (in) {
            const unsigned peers = __match_any_sync(act, d);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            const uint64_t dst = s_base[d] + s_cnt[warp][d] + rank;
            out_keys[dst] = k[r];
            if (vals) out_vals[dst] = vals[i];
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) s_cnt[warp][d] += __popc(peers);
        }
        __syncwarp();
    }
}

size_t radix_sort_scan_bytes(uint64_t n) {
    const uint64_t blocks = (n + RS_TILE - 1) / RS_TILE;
    const uint64_t table = 256 * (blocks ? blocks : 1);
    return ((table * sizeof(uint64_t) + 255) / 256 * 256) + scan_tmp_bytes(table);
}

void radix_sort_pairs(uint64_t *keys, uint64_t *vals, uint64_t n, uint64_t *tmp_keys, uint64_t *tmp_vals,
                      int end_bit, void *scan_tmp, size_t scan_tmp_bytes_, cudaStream_t st) {
    if (n < 2) return;
    (void)scan_tmp_bytes_;
    const uint64_t blocks64 = (n + RS_TILE - 1) / RS_TILE;
    if (blocks64 > 0x7FFFFFFFull) throw_internal("sort too large");
    const uint32_t blocks = (uint32_t)blocks64;
    const uint64_t table = 256ull * blocks;
    uint64_t *counts = reinterpret_cast<uint64_t *>(scan_tmp);
    void *stmp = reinterpret_cast<char *>(scan_tmp) + ((table * sizeof(uint64_t) + 255) / 256 * 256);
    int passes = (end_bit + 7) / 8;
    if (passes < 1) passes = 1;
    if (passes > 8) passes = 8;
    const bool has_vals = vals != nullptr;
    uint64_t *src_k = keys, *src_v = vals, *dst_k = tmp_keys, *dst_v = has_vals ? tmp_vals : nullptr;
    for (int p = 0; p < passes; p++) {
        const int shift = 8 * p;
        rs_hist_kernel<<<blocks, RS_THREADS, 0, st>>>(src_k, n, shift, counts, blocks);
        SM_LAUNCHED();
        scan_exclusive_u64(counts, counts, table, stmp, st);
        rs_scatter_kernel<<<blocks, RS_THREADS, 0, st>>>(src_k, src_v, n, shift, counts, blocks, dst_k, dst_v);
        SM_LAUNCHED();
        uint64_t *t = src_k; src_k = dst_k; dst_k = t;
        t = src_v; src_v = dst_v; dst_v = t;
    }
    if (src_k != keys) {  // odd number of passes: the sorted data sits in the temporaries
        SM_CUDA(cudaMemcpyAsync(keys, src_k, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        if (has_vals) SM_CUDA(cudaMemcpyAsync(vals, src_v, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    }
}

// =====================================================================================
// reduce-by-key over a sorted key array
// =====================================================================================
__global__ void heads_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint64_t *__restrict__ flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
void launch_heads(const uint64_t *keys, uint64_t n, uint64_t *flags, cudaStream_t st) {
    if (!n) return;
    heads_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, n, flags);
    SM_LAUNCHED();
}
// idx[i] = exclusive scan of head flags; run id of element i = idx[i] + head(i) - 1
__global__ void rbk_scatter_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals,
                                   uint64_t n, uint64_t *__restrict__ idx, uint64_t *__restrict__ ukeys,
                                   uint64_t *__restrict__ usums, unsigned long long *n_unique) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool head = (i == 0 || keys[i] != keys[i - 1]);
        const uint64_t run = idx[i] + (head ? 1 : 0) - 1;
        idx[i] = run;  // now the run id (min_by_key reuses it)
        if (head) ukeys[run] = keys[i];
        atomicAdd(reinterpret_cast<unsigned long long *>(&usums[run]),
                  (unsigned long long)(vals ? vals[i] : 1ull));
        if (i == n - 1) *n_unique = run + 1;
    }
}
void reduce_by_key(const uint64_t *keys, const uint64_t *vals, uint64_t n, uint64_t *ukeys, uint64_t *usums,
                   unsigned long long *n_unique, uint64_t *idx_tmp, void *scan_tmp, cudaStream_t st) {
    if (n == 0) {
        SM_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(unsigned long long), st));
        return;
    }
    heads_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, n, idx_tmp);
    SM_LAUNCHED();
    scan_exclusive_u64(idx_tmp, idx_tmp, n, scan_tmp, st);
    SM_CUDA(cudaMemsetAsync(usums, 0, n * sizeof(uint64_t), st));
    rbk_scatter_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, vals, n, idx_tmp, ukeys, usums, n_unique);
    SM_LAUNCHED();
}
__global__ void min_by_key_kernel(const uint64_t *__restrict__ vals, uint64_t n, const uint64_t *__restrict__ idx,
                                  uint64_t *umins) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicMin(reinterpret_cast<unsigned long long *>(&umins[idx[i]]), (unsigned long long)vals[i]);
}
void min_by_key(const uint64_t *keys, const uint64_t *vals, uint64_t n, const uint64_t *idx, uint64_t *umins,
                cudaStream_t st) {
    (void)keys;
    if (!n) return;
    min_by_key_kernel<<<grid_for(n, 256), 256, 0, st>>>(vals, n, idx, umins);
    SM_LAUNCHED();
}

// =====================================================================================
// num + track_abundance corner (lib.rs:206-208): when the sketch is full, occurrences of
// its largest element X are counted only up to the moment T the sketch became full.
// =====================================================================================
__device__ __forceinline__ bool contains_sorted(const uint64_t *a, uint64_t n, uint64_t x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint64_t v = a[mid];
        if (v == x) return true;
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return false;
}
// T+1 = max over new distinct hashes u <= x that are not in the old sketch of (first position + 1)
__global__ void first_new_max_kernel(const uint64_t *__restrict__ ukeys, const uint64_t *__restrict__ ufirst,
                                     uint64_t nu, const uint64_t *__restrict__ old_keys, uint64_t n_old,
                                     const uint64_t *__restrict__ x_ptr, unsigned long long *t_max_plus1) {
    const uint64_t x = *x_ptr;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nu; i += stride) {
        const uint64_t u = ukeys[i];
        if (u <= x && !contains_sorted(old_keys, n_old, u)) atomicMax(t_max_plus1, (unsigned long long)(ufirst[i] + 1));
    }
}
void launch_first_new_max(const uint64_t *ukeys, const uint64_t *ufirst, uint64_t nu, const uint64_t *old_keys,
                          uint64_t n_old, const uint64_t *x, unsigned long long *t_max_plus1, cudaStream_t st) {
    if (!nu) return;
    first_new_max_kernel<<<grid_for(nu, 256), 256, 0, st>>>(ukeys, ufirst, nu, old_keys, n_old, x, t_max_plus1);
    SM_LAUNCHED();
}
// number of events (key == x) with position < T+1
__global__ void count_key_upto_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ pos,
                                      uint64_t n, const uint64_t *__restrict__ x_ptr,
                                      const unsigned long long *t_max_plus1, unsigned long long *count) {
    const uint64_t x = *x_ptr;
    const unsigned long long lim = *t_max_plus1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (keys[i] == x && pos[i] < lim) local++;
    if (local) atomicAdd(count, local);
}
void launch_count_key_upto(const uint64_t *keys, const uint64_t *pos, uint64_t n, const uint64_t *x,
                           const unsigned long long *t_max_plus1, unsigned long long *count, cudaStream_t st) {
    if (!n) return;
    count_key_upto_kernel<<<grid_for(n, 256), 256, 0, st>>>(keys, pos, n, x, t_max_plus1, count);
    SM_LAUNCHED();
}

// abund(X) currently holds old + (all new occurrences of X); replace the new part by the
// occurrences up to T:  *abund_x = *abund_x - total_new(X) + *count_upto
__global__ void fix_max_abund_kernel(uint64_t *abund_x, const uint64_t *__restrict__ ukeys,
                                     const uint64_t *__restrict__ ucounts, uint64_t nu,
                                     const uint64_t *__restrict__ x_ptr, const unsigned long long *count_upto) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint64_t x = *x_ptr;
    uint64_t lo = 0, hi = nu;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (ukeys[mid] < x) lo = mid + 1; else hi = mid;
    }
    const uint64_t total = (lo < nu && ukeys[lo] == x) ? ucounts[lo] : 0;
    *abund_x = *abund_x - total + *count_upto;
}
void launch_fix_max_abund(uint64_t *abund_x, const uint64_t *ukeys, const uint64_t *ucounts, uint64_t nu,
                          const uint64_t *x, const unsigned long long *count_upto, cudaStream_t st) {
    fix_max_abund_kernel<<<1, 32, 0, st>>>(abund_x, ukeys, ucounts, nu, x, count_upto);
    SM_LAUNCHED();
}

// =====================================================================================
// Ordered replay of add_hash (lib.rs:192-245), one CTA, events in stream order.  Exact for
// every (num, max_hash) combination, including the ones whose result depends on arrival
// order (num>0 && max_hash>0; num==0 && max_hash==0).  mins/abunds must have room for
// len + n_events entries.
// =====================================================================================
constexpr int RP_THREADS = 1024;
__global__ void __launch_bounds__(RP_THREADS) replay_add_hash_kernel(const uint64_t *__restrict__ events,
                                                                     uint64_t n_events, uint32_t num,
                                                                     uint64_t max_hash, uint64_t *mins,
                                                                     uint64_t *abunds,
                                                                     unsigned long long *len_io) {
    __shared__ unsigned long long s_cnt;
    __shared__ int s_found;
    uint64_t len = *len_io;
    const int tid = threadIdx.x;
    for (uint64_t e = 0; e < n_events; e++) {
        const uint64_t h = events[e];
        const uint64_t current_max = len ? mins[len - 1] : ~0ull;
        if (!(h <= max_hash || max_hash == 0)) continue;
        if (len == 0) {
            if (tid == 0) { mins[0] = h; if (abunds) abunds[0] = 1; }
            len = 1;
            __syncthreads();
            continue;
        }
        if (!(h <= max_hash || current_max > h || (uint32_t)len < num)) continue;
        // pos = number of elements < h ; found = h present
        if (tid == 0) { s_cnt = 0; s_found = 0; }
        __syncthreads();
        unsigned long long local = 0;
        int found = 0;
        for (uint64_t i = tid; i < len; i += RP_THREADS) {
            const uint64_t v = mins[i];
            local += (v < h);
            found |= (v == h);
        }
        if (local) atomicAdd(&s_cnt, local);
        if (found) s_found = 1;
        __syncthreads();
        const uint64_t pos = s_cnt;
        const bool present = s_found != 0;
        __syncthreads();
        if (pos == len) {
            if (tid == 0) { mins[len] = h; if (abunds) abunds[len] = 1; }
            len++;
        } else if (!present) {
            // shift [pos, len) up by one, highest block first
            for (uint64_t hi = len; hi > pos;) {
                const uint64_t lo = (hi - pos > RP_THREADS) ? hi - RP_THREADS : pos;
                const uint64_t i = lo + tid;
                uint64_t v = 0, a = 0;
                const bool mine = i < hi;
                if (mine) { v = mins[i]; if (abunds) a = abunds[i]; }
                __syncthreads();
                if (mine) { mins[i + 1] = v; if (abunds) abunds[i + 1] = a; }
                __syncthreads();
                hi = lo;
            }
            if (tid == 0) { mins[pos] = h; if (abunds) abunds[pos] = 1; }
            len++;
            if (num != 0 && len > (uint64_t)num) len--;
        } else if (abunds) {
            if (tid == 0) abunds[pos] += 1;
        }
        __syncthreads();
    }
    if (tid == 0) *len_io = len;
}
void launch_replay_add_hash(const uint64_t *events, uint64_t n_events, uint32_t num, uint64_t max_hash,
                            uint64_t *mins, uint64_t *abunds, unsigned long long *len_io, cudaStream_t st) {
    if (!n_events) return;
    replay_add_hash_kernel<<<1, RP_THREADS, 0, st>>>(events, n_events, num, max_hash, mins, abunds, len_io);
    SM_LAUNCHED();
}

// =====================================================================================
// Fast fold of scaled-sketch candidates into a sorted state (minhash.cu: ingest).
// Once a sketch has seen a few batches of a sample, nearly every candidate hash is already in the
// state: the fold is then "look each candidate up, count it" and no sort at all.
// =====================================================================================
// found: abunds[pos] += 1 (when tracked); not found: appended to news[] through *n_news
__global__ void __launch_bounds__(256) fold_match_kernel(const uint64_t *__restrict__ cand, uint64_t nc,
                                                         const uint64_t *__restrict__ mins, uint64_t na,
                                                         unsigned long long *abunds, uint64_t *news,
                                                         unsigned long long *n_news) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (nc + 31) / 32 * 32;
    const int lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool is_new = false;
        uint64_t h = 0;
        if (i < nc) {
            h = cand[i];
            uint64_t lo = 0, hi = na;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (mins[mid] < h) lo = mid + 1; else hi = mid;
            }
            if (lo < na && mins[lo] == h) {
                if (abunds) atomicAdd(&abunds[lo], 1ull);
            } else {
                is_new = true;
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, is_new);
        if (bal) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(n_news, (unsigned long long)__popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (is_new) news[base + __popc(bal & ((1u << lane) - 1u))] = h;
        }
    }
}
void launch_fold_match(const uint64_t *cand, uint64_t nc, const uint64_t *mins, uint64_t na, uint64_t *abunds, uint64_t *news,
                       unsigned long long *n_news, cudaStream_t st) {
    if (!nc) return;
    const uint64_t blocks = std::min<uint64_t>((nc + 255) / 256, 148 * 16);
    fold_match_kernel<<<(unsigned)blocks, 256, 0, st>>>(cand, nc, mins, na, reinterpret_cast<unsigned long long *>(abunds), news, n_news);
    SM_LAUNCHED();
}
// up to FOLD_SMALL new hashes: sorted and run-length reduced by ONE CTA in shared memory (bitonic),
// ukeys / ucnt / *n_unique out
constexpr int FOLD_SMALL = 2048;
__global__ void __launch_bounds__(1024) fold_small_sort_kernel(const uint64_t *__restrict__ news, uint32_t n, uint64_t *ukeys,
                                                               uint64_t *ucnt, unsigned long long *n_unique) {
    __shared__ uint64_t s_k[FOLD_SMALL];
    __shared__ uint32_t s_head[FOLD_SMALL];
    __shared__ uint32_t s_cnt[FOLD_SMALL];
    const int tid = threadIdx.x;
    for (int i = tid; i < FOLD_SMALL; i += blockDim.x) { s_k[i] = i < (int)n ? news[i] : ~0ull; s_cnt[i] = 0; }
    __syncthreads();
    for (int k = 2; k <= FOLD_SMALL; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < FOLD_SMALL; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = s_k[i], b = s_k[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_k[i] = b; s_k[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    // heads of runs among the first n sorted entries -> inclusive scan (serial over 2048 by warp 0 is enough: rare path)
    for (int i = tid; i < FOLD_SMALL; i += blockDim.x) s_head[i] = (i < (int)n && (i == 0 || s_k[i] != s_k[i - 1])) ? 1u : 0u;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < n; i++) { run += s_head[i]; s_head[i] = run; }  // s_head[i] = 1-based run id
        *n_unique = run;
    }
    __syncthreads();
    for (int i = tid; i < (int)n; i += blockDim.x) {
        const uint32_t r = s_head[i] - 1;
        atomicAdd(&s_cnt[r], 1u);
        if (i == 0 || s_k[i] != s_k[i - 1]) ukeys[r] = s_k[i];
    }
    __syncthreads();
    const uint32_t nu = n ? s_head[n - 1] : 0;
    for (int i = tid; i < (int)nu; i += blockDim.x) ucnt[i] = s_cnt[i];
}
// two sorted, disjoint key lists (a: state with values av or null; b: new keys with counts bv, *nb_dev entries)
// -> merged into out_k / out_v by rank
__global__ void __launch_bounds__(256) fold_merge_kernel(const uint64_t *__restrict__ a, const uint64_t *__restrict__ av, uint64_t na,
                                                         const uint64_t *__restrict__ b, const uint64_t *__restrict__ bv,
                                                         const unsigned long long *nb_dev, uint64_t *out_k, uint64_t *out_v) {
    const uint64_t nb = *nb_dev;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < na + nb; t += stride) {
        const bool from_a = t < na;
        const uint64_t idx = from_a ? t : t - na;
        const uint64_t key = from_a ? a[idx] : b[idx];
        const uint64_t *o = from_a ? b : a;
        uint64_t lo = 0, hi = from_a ? nb : na;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (o[mid] < key) lo = mid + 1; else hi = mid;
        }
        out_k[idx + lo] = key;
        if (out_v) out_v[idx + lo] = from_a ? (av ? av[idx] : 0) : bv[idx];
    }
}
void launch_fold_small(const uint64_t *news, uint32_t n_news, const uint64_t *mins, const uint64_t *abunds, uint64_t na,
                       uint64_t *ukeys, uint64_t *ucnt, unsigned long long *n_unique, uint64_t *out_k, uint64_t *out_v,
                       cudaStream_t st) {
    if (n_news > (uint32_t)FOLD_SMALL) throw_internal("fold_small: too many new hashes");
    fold_small_sort_kernel<<<1, 1024, 0, st>>>(news, n_news, ukeys, ucnt, n_unique);
    SM_LAUNCHED();
    const uint64_t blocks = std::min<uint64_t>((na + n_news + 255) / 256, 148 * 8);
    fold_merge_kernel<<<(unsigned)blocks, 256, 0, st>>>(mins, abunds, na, ukeys, ucnt, n_unique, out_k, out_v);
    SM_LAUNCHED();
}
int fold_small_limit() { return FOLD_SMALL; }

// =====================================================================================
// INT32 issue-rate microbenchmark (roofline denominator for sketch.cu, measured live)
// mode 0: IMAD chain (fma pipe), mode 1: LOP3/SHF chain (alu pipe), mode 2: interleaved
// =====================================================================================
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters, int mode) {
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const uint32_t m = blockIdx.x | 3u, c = 0x9E3779B9u + blockIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (mode == 0 || mode == 2) {
                a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            }
            if (mode == 1 || mode == 2) {
                a4 = __funnelshift_l(a4, a5, 7) ^ c; a5 = (a5 & a6) ^ c; a6 = __funnelshift_l(a6, a7, 13) ^ m;
                a7 = (a7 | a4) ^ m;
            }
            if (mode == 0) { a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c; }
            if (mode == 1) {
                a0 = __funnelshift_l(a0, a1, 5) ^ c; a1 = (a1 & a2) ^ m; a2 = __funnelshift_l(a2, a3, 11) ^ c;
                a3 = (a3 | a0) ^ m;
            }
        }
    }
    const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678u) out[0] = r;  // keep the chain alive
}
// Modes 3..9: eight INDEPENDENT chains per pipe, one SASS instruction per chain step (checked with
// cuobjdump), so that the measured rate is the pipes' and not a dependency chain's:
//   3: 8 LOP3           4: 8 SHF            5: 8 IMAD + 8 LOP3        6: 8 IMAD.WIDE + 8 LOP3
//   7: 4 IMAD.WIDE + 8 IMAD + 12 LOP3/SHF  (the MurmurHash3 mix)      8: 8 IMAD.HI      9: 8 IMAD + 8 SHF
template <int MODE>
__global__ void __launch_bounds__(256) int_peak2_kernel(uint32_t *out, int iters, uint32_t m, uint32_t c) {
    uint32_t a[8], b[8];
    uint64_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = threadIdx.x + j; b[j] = threadIdx.x * 3 + j; w[j] = threadIdx.x + 5 * j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
#define IP_LOP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(m), "r"(c))
#define IP_SHF(x) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x) : "r"(m))
#define IP_MAD(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(c))
#define IP_WIDE(x, y) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(y), "r"(m))
#define IP_HI(x) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(c))
                if (MODE == 3) IP_LOP(a[j]);
                if (MODE == 4) IP_SHF(a[j]);
                if (MODE == 5) { IP_MAD(a[j]); IP_LOP(b[j]); }
                if (MODE == 6) {  // WIDE + LOP3 (both halves of the product are consumed)
                    uint32_t lo, hi;
                    asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t;}" : "=r"(lo), "=r"(hi) : "r"(a[j]), "r"(m));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[j]) : "r"(lo), "r"(hi), "r"(c));
                }
                if (MODE == 7) {
                    if (j < 4) IP_WIDE(w[j], b[j]);
                    IP_MAD(a[j]);
                    if (j < 4) IP_LOP(b[j]); else IP_SHF(b[j]);
                    if (j < 4) IP_LOP(b[j + 4]);
                }
                if (MODE == 8) IP_HI(a[j]);
                if (MODE == 9) { IP_MAD(a[j]); IP_SHF(b[j]); }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r ^= a[j] ^ b[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
    if (r == 0x12345678u) out[0] = r;  // keep the chains alive
}
// Mode 10/11/12: MurmurHash3 x64_128 of a K = 21 / 31 / 51 byte k-mer held in registers, nothing else
// (no staging, no strand choice, no shared memory): the hash-only ceiling of the sketch kernel.
// One "instruction" in the returned count = one hash (iters * 64 hashes per thread).
template <int K>
__global__ void __launch_bounds__(256, 8) hash_peak_kernel(uint32_t *out, int iters, uint32_t m, uint64_t seed) {
    constexpr int NW = (K + 3) / 4;
    uint32_t w[NW];
#pragma unroll
    for (int j = 0; j < NW; j++) w[j] = (threadIdx.x + blockIdx.x * 256u) * 0x9E3779B9u + j * m;
    w[NW - 1] &= (K % 4) ? ((1u << ((K % 4) * 8)) - 1u) : 0xFFFFFFFFu;
    uint64_t acc = 0;
    for (int i = 0; i < iters * 64; i++) {
        const uint64_t h = murmur3_h1_words<K>(w, seed);
        acc ^= h;
#pragma unroll
        for (int j = 0; j < NW - 1; j++) w[j] += m + j;  // next k-mer: every word changes (NW - 1 extra adds per hash)
        w[NW - 1] = (w[NW - 1] + m) & ((K % 4) ? ((1u << ((K % 4) * 8)) - 1u) : 0xFFFFFFFFu);
    }
    if (acc == 0x12345678u) out[0] = (uint32_t)acc;
}
void launch_int_peak(uint32_t *out, int iters, int blocks, int mode, cudaStream_t st) {
    if (mode >= 10 && mode <= 12) {
        const uint32_t mm = 0x01000193u | (uint32_t)blocks;
        if (mode == 10) hash_peak_kernel<21><<<blocks, 256, 0, st>>>(out, iters, mm, 42);
        if (mode == 11) hash_peak_kernel<31><<<blocks, 256, 0, st>>>(out, iters, mm, 42);
        if (mode == 12) hash_peak_kernel<51><<<blocks, 256, 0, st>>>(out, iters, mm, 42);
        SM_LAUNCHED();
        return;
    }
    const uint32_t m = 0x01000193u | (uint32_t)blocks, c = 0x9E3779B9u;  // opaque to the compiler
    switch (mode) {
    case 3: int_peak2_kernel<3><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 4: int_peak2_kernel<4><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 5: int_peak2_kernel<5><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 6: int_peak2_kernel<6><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 7: int_peak2_kernel<7><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 8: int_peak2_kernel<8><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    case 9: int_peak2_kernel<9><<<blocks, 256, 0, st>>>(out, iters, m, c); break;
    default: int_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, mode);
    }
    SM_LAUNCHED();
}

}  // namespace smb200
