// compare.cu -- sorted-set comparison of sketches.
//
// Replaces the reference's Intersection iterator (src/lib.rs:515-544) and its users:
// count_common (lib.rs:428-436), intersection_size (lib.rs:470-499: two merges, two
// intersection passes, three allocations per pair) and compare (lib.rs:501-508), plus
// Leaf<Signature>::containment (src/index.rs:146-160).
//
// Per pair the arithmetic is one merge walk over the two sorted lists that stops after
// `num` union elements (the bottom-num-of-union rule of intersection_size) and counts the
// equal elements met on the way -- identical integers to the reference, no allocation.
//   matrix kernels : one thread per (row, column) pair; the 32 column sketches of a tile are
//                    staged lane-interleaved in shared memory (lane l owns banks 2l,2l+1, so
//                    the per-lane column pointer never conflicts), the 8 row sketches of the
//                    CTA are read through L1.  Bound: shared-memory/L1 wavefronts, not HBM.
//   pair kernel    : one CTA per pair for the per-object C ABI (rank formulation, below).
#include <algorithm>

#include "device.hpp"
#include "kernels.cuh"

namespace smb200 {

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t *__restrict__ a, uint64_t n, uint64_t x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// -------------------------------------------------------------------------------------
// One pair, rank formulation.  x = a[i] also occurs at b[j]; c = number of common elements
// smaller than x.  Its 1-based rank in the sorted union is (i+1)+(j+1)-(c+1), so it lies in
// bottom_num(A∪B) iff i + j + 1 - c <= num.
// -------------------------------------------------------------------------------------
constexpr int PS_THREADS = 256;
__global__ void __launch_bounds__(PS_THREADS) pair_stats_kernel(const uint64_t *__restrict__ a, uint64_t na,
                                                                const uint64_t *__restrict__ b, uint64_t nb,
                                                                uint32_t num, unsigned long long *out3) {
    __shared__ uint32_t s_warp[PS_THREADS / 32];
    __shared__ unsigned long long s_trunc;
    if (threadIdx.x == 0) s_trunc = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t running = 0;  // common elements before this block of A (uniform)
    unsigned long long local_trunc = 0;
    for (uint64_t base = 0; base < na; base += PS_THREADS) {
        const uint64_t i = base + threadIdx.x;
        bool eq = false;
        uint64_t j = 0;
        if (i < na) {
            const uint64_t x = a[i];
            j = lower_bound_u64(b, nb, x);
            eq = (j < nb) && (b[j] == x);
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, eq);
        __syncthreads();  // previous round's s_warp fully consumed
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint64_t before = running, total = 0;
        for (int w = 0; w < PS_THREADS / 32; w++) {
            if (w < warp) before += s_warp[w];
            total += s_warp[w];
        }
        before += __popc(bal & ((1u << lane) - 1u));
        if (eq && (num == 0 || i + j + 1 - before <= (uint64_t)num)) local_trunc++;
        running += total;
    }
    if (local_trunc) atomicAdd(&s_trunc, local_trunc);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t uni = na + nb - running;
        out3[0] = running;
        out3[1] = s_trunc;
        out3[2] = (num != 0 && uni >= (uint64_t)num) ? (uint64_t)num : uni;  // lib.rs:391-401
    }
}
void launch_pair_stats(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint32_t num,
                       unsigned long long *out3, cudaStream_t st) {
    pair_stats_kernel<<<1, PS_THREADS, 0, st>>>(a, na, b, nb, num, out3);
    SM_LAUNCHED();
}

// flags[i] = 1 if a[i] occurs in b (both sorted, distinct)
__global__ void mark_common_kernel(const uint64_t *__restrict__ a, uint64_t na, const uint64_t *__restrict__ b,
                                   uint64_t nb, uint64_t *__restrict__ flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += stride) {
        const uint64_t x = a[i];
        const uint64_t j = lower_bound_u64(b, nb, x);
        flags[i] = (j < nb && b[j] == x) ? 1 : 0;
    }
}
void launch_mark_common(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint64_t *flags,
                        cudaStream_t st) {
    if (!na) return;
    uint64_t blocks = (na + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    mark_common_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, na, b, nb, flags);
    SM_LAUNCHED();
}
// out[i - pre[i]] = vals[i] for every i with flag 0 (pre = exclusive scan of flags)
__global__ void compact_unflagged_kernel(const uint64_t *__restrict__ vals, const uint64_t *__restrict__ flags,
                                         const uint64_t *__restrict__ pre, uint64_t n, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (!flags[i]) out[i - pre[i]] = vals[i];
}
void launch_compact_unflagged(const uint64_t *vals, const uint64_t *flags, const uint64_t *pre, uint64_t n,
                              uint64_t *out, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    compact_unflagged_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, flags, pre, n, out);
    SM_LAUNCHED();
}
// out[pre[i]] = vals[i] for every i with flag 1, as long as pre[i] < limit
__global__ void compact_flagged_kernel(const uint64_t *__restrict__ vals, const uint64_t *__restrict__ flags,
                                       const uint64_t *__restrict__ pre, uint64_t n, uint64_t limit, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (flags[i] && pre[i] < limit) out[pre[i]] = vals[i];
}
void launch_compact_flagged(const uint64_t *vals, const uint64_t *flags, const uint64_t *pre, uint64_t n, uint64_t limit,
                            uint64_t *out, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    compact_flagged_kernel<<<(unsigned)blocks, 256, 0, st>>>(vals, flags, pre, n, limit, out);
    SM_LAUNCHED();
}

// -------------------------------------------------------------------------------------
// Matrix kernels.  CTA = 8 warps = 8 rows x 32 columns.
// -------------------------------------------------------------------------------------
constexpr int CM_ROWS = 8;
constexpr int CM_COLS = 32;
constexpr int CM_THREADS = CM_ROWS * CM_COLS;

struct PairOut {
    uint32_t common, size;
};

// merge walk; B element e of this lane's column is at bcol[e * BSTRIDE]
template <int BSTRIDE>
__device__ __forceinline__ PairOut merge_walk(const uint64_t *__restrict__ arow, uint32_t na,
                                              const uint64_t *bcol, uint32_t nb, uint32_t num, int mode) {
    uint32_t i = 0, j = 0, c = 0, u = 0;
    PairOut o;
    if (mode == 0) {
        const uint32_t limit = num ? num : 0xFFFFFFFFu;
        while (i < na && j < nb && u < limit) {
            const uint64_t x = __ldg(arow + i), y = bcol[(size_t)j * BSTRIDE];
            c += (x == y);
            i += (x <= y);
            j += (y <= x);
            u++;
        }
        const uint64_t uni = (uint64_t)u + (na - i) + (nb - j);
        o.common = c;
        o.size = (uint32_t)((num != 0 && uni >= num) ? num : uni);
    } else {
        while (i < na && j < nb) {
            const uint64_t x = __ldg(arow + i), y = bcol[(size_t)j * BSTRIDE];
            c += (x == y);
            i += (x <= y);
            j += (y <= x);
        }
        o.common = c;
        o.size = na;  // containment denominator: the ROW (index/node) sketch, index.rs:152-154
    }
    return o;
}

__device__ __forceinline__ void store_pair(PairOut o, int mode, uint32_t *common, uint32_t *size, double *ratio,
                                           size_t at) {
    if (common) common[at] = o.common;
    if (size) size[at] = o.size;
    if (ratio) {
        const double den = (mode == 0) ? (double)(o.size > 1 ? o.size : 1) : (double)o.size;
        ratio[at] = (double)o.common / den;  // IEEE division of two exact integers: same bits as the CPU
    }
}

template <bool SMEM_COLS>
__global__ void __launch_bounds__(CM_THREADS)
compare_cross_kernel(const uint64_t *__restrict__ rh, const uint64_t *__restrict__ ro,
                     const uint32_t *__restrict__ rnum, uint64_t r0, uint64_t nr, const uint64_t *__restrict__ ch,
                     const uint64_t *__restrict__ co, uint64_t c0, uint64_t nc, int mode, uint32_t *common,
                     uint32_t *size, double *ratio, uint64_t ld, uint32_t col_cap) {
    extern __shared__ __align__(16) uint64_t s_cols[];  // [col_cap][32] lane-interleaved
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_ctile = (nc + CM_COLS - 1) / CM_COLS;
    const uint64_t n_rtile = (nr + CM_ROWS - 1) / CM_ROWS;
    for (uint64_t ct = blockIdx.x; ct < n_ctile; ct += gridDim.x) {
        const uint64_t col = c0 + ct * CM_COLS + lane;
        const bool col_ok = (ct * CM_COLS + lane) < nc;
        uint64_t cb = 0;
        uint32_t nb = 0;
        if (col_ok) { cb = co[col]; nb = (uint32_t)(co[col + 1] - cb); }
        if (SMEM_COLS) {
            __syncthreads();
            // stage: warp w copies column w, w+8, ... (coalesced 256 B reads, strided smem writes)
            for (int cc = warp; cc < CM_COLS; cc += CM_ROWS) {
                const uint64_t gcol = c0 + ct * CM_COLS + cc;
                if (ct * CM_COLS + cc < nc) {
                    const uint64_t b = co[gcol];
                    const uint32_t len = (uint32_t)(co[gcol + 1] - b);
                    for (uint32_t e = lane; e < len; e += 32) s_cols[(size_t)e * CM_COLS + cc] = __ldg(ch + b + e);
                }
            }
            __syncthreads();
        }
        for (uint64_t rt = blockIdx.y; rt < n_rtile; rt += gridDim.y) {
            const uint64_t rr = rt * CM_ROWS + warp;
            if (rr < nr && col_ok) {
                const uint64_t row = r0 + rr;
                const uint64_t ab = ro[row];
                const uint32_t na = (uint32_t)(ro[row + 1] - ab);
                const uint32_t num = rnum ? rnum[row] : 0;
                PairOut o;
                if (SMEM_COLS) o = merge_walk<CM_COLS>(rh + ab, na, s_cols + lane, nb, num, mode);
                else o = merge_walk<1>(rh + ab, na, ch + cb, nb, num, mode);
                store_pair(o, mode, common, size, ratio, (size_t)rr * ld + (ct * CM_COLS + lane));
            }
        }
    }
    (void)col_cap;
}

void launch_compare_cross(const uint64_t *row_hashes, const uint64_t *row_offsets, const uint32_t *row_nums,
                          uint64_t r0, uint64_t nr, const uint64_t *col_hashes, const uint64_t *col_offsets,
                          uint64_t c0, uint64_t nc, int mode, uint32_t *common, uint32_t *size, double *ratio,
                          uint64_t ld, uint32_t max_col_len, int sm_count, cudaStream_t st) {
    if (nr == 0 || nc == 0) return;
    const uint64_t n_ctile = (nc + CM_COLS - 1) / CM_COLS;
    const uint64_t n_rtile = (nr + CM_ROWS - 1) / CM_ROWS;
    const size_t smem = (size_t)max_col_len * CM_COLS * sizeof(uint64_t);
    const bool use_smem = max_col_len > 0 && smem <= 200 * 1024;
    // x: column tiles, y: row-tile groups; enough CTAs for a few waves over the SMs
    dim3 grid;
    grid.x = (unsigned)(n_ctile > 65535 ? 65535 : n_ctile);
    uint64_t want_y = ((uint64_t)sm_count * 8 + grid.x - 1) / grid.x;
    if (want_y > n_rtile) want_y = n_rtile;
    if (want_y < 1) want_y = 1;
    if (want_y > 65535) want_y = 65535;
    grid.y = (unsigned)want_y;
    ProfScope prof(PROF_COMPARE, st);
    if (use_smem) {
        SM_CUDA(cudaFuncSetAttribute(compare_cross_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        compare_cross_kernel<true><<<grid, CM_THREADS, smem, st>>>(row_hashes, row_offsets, row_nums, r0, nr, col_hashes,
                                                                 col_offsets, c0, nc, mode, common, size, ratio, ld,
                                                                 max_col_len);
    } else {
        compare_cross_kernel<false><<<grid, CM_THREADS, 0, st>>>(row_hashes, row_offsets, row_nums, r0, nr, col_hashes,
                                                                col_offsets, c0, nc, mode, common, size, ratio, ld,
                                                                max_col_len);
    }
    SM_LAUNCHED();
}

// hits of a linear search, transposed: flags[j * nr + i] = ratio[i * nq + j] > threshold (strict,
// src/index/search.rs:3-9; NaN from 0/0 is never a hit)
__global__ void threshold_flags_t_kernel(const double *__restrict__ ratio, uint64_t nr, uint64_t nq, double thr,
                                         uint64_t *__restrict__ flags) {
    const uint64_t n = nr * nq;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const uint64_t i = t / nq, j = t - i * nq;  // coalesced read, strided write (hits are sparse work anyway)
        flags[j * nr + i] = (ratio[t] > thr) ? 1 : 0;
    }
}
void launch_threshold_flags_t(const double *ratio, uint64_t nr, uint64_t nq, double threshold, uint64_t *flags,
                              cudaStream_t st) {
    const uint64_t n = nr * nq;
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    threshold_flags_t_kernel<<<(unsigned)blocks, 256, 0, st>>>(ratio, nr, nq, threshold, flags);
    SM_LAUNCHED();
}
// one warp per row: every row strictly ascending?
__global__ void csr_check_sorted_kernel(const uint64_t *__restrict__ hashes, const uint64_t *__restrict__ offsets,
                                        uint64_t n_rows, unsigned long long *flag) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        for (uint64_t i = b + 1 + lane; i < e; i += 32)
            if (!(hashes[i - 1] < hashes[i])) *flag = 1;
    }
}
void launch_csr_check_sorted(const uint64_t *hashes, const uint64_t *offsets, uint64_t n_rows, unsigned long long *flag,
                             cudaStream_t st) {
    if (!n_rows) return;
    uint64_t blocks = (n_rows * 32 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    csr_check_sorted_kernel<<<(unsigned)blocks, 256, 0, st>>>(hashes, offsets, n_rows, flag);
    SM_LAUNCHED();
}
// out[pre[i]] = i for flagged i (ascending i: insertion order of LinearIndex::find, linear.rs:34-44)
__global__ void compact_indices_kernel(const uint64_t *__restrict__ flags, const uint64_t *__restrict__ pre,
                                       uint64_t n, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (flags[i]) out[pre[i]] = i;
}
void launch_compact_indices(const uint64_t *flags, const uint64_t *pre, uint64_t n, uint64_t *out, cudaStream_t st) {
    if (!n) return;
    uint64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    compact_indices_kernel<<<(unsigned)blocks, 256, 0, st>>>(flags, pre, n, out);
    SM_LAUNCHED();
}

// linear_find, containment with threshold >= 0: cells of the count matrix (row-major [index row][query], ld = nq) whose
// count is non-zero and whose count / |node| exceeds the threshold (strict '>', search.rs:7-9; index.rs:152-154) are
// appended as query * nr + row through *n_found (entries beyond cap are counted, not stored)
// q_offsets != nullptr: similarity of sketches whose node side has num == 0 (lib.rs:470-508 with self.num == 0:
// common = |A n B|, size = |A u B| = |A| + |B| - common), the denominator max(1, size) of lib.rs:504
__global__ void __launch_bounds__(256) count_hits_kernel(const uint32_t *__restrict__ cmat, uint64_t nr, uint64_t nq,
                                                         const uint64_t *__restrict__ row_offsets, uint64_t r0,
                                                         const uint64_t *__restrict__ q_offsets, double threshold,
                                                         uint64_t *found, uint64_t cap, unsigned long long *n_found) {
    for (uint64_t i = blockIdx.y; i < nr; i += gridDim.y) {
        const uint64_t la = row_offsets[r0 + i + 1] - row_offsets[r0 + i];
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nq; j += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t cm = cmat[i * nq + j];
            if (cm == 0) continue;
            double den = (double)la;
            if (q_offsets) den = (double)(la + (q_offsets[j + 1] - q_offsets[j]) - cm);  // >= 1 when cm >= 1
            if ((double)cm / den > threshold) {
                const unsigned long long at = atomicAdd(n_found, 1ull);
                if (at < cap) found[at] = j * nr + i;
            }
        }
    }
}
void launch_count_hits(const uint32_t *cmat, uint64_t nr, uint64_t nq, const uint64_t *row_offsets, uint64_t r0,
                       const uint64_t *q_offsets, double threshold, uint64_t *found, uint64_t cap, unsigned long long *n_found,
                       cudaStream_t st) {
    if (!nr || !nq) return;
    const unsigned gx = (unsigned)std::min<uint64_t>((nq + 255) / 256, 64);
    const unsigned gy = (unsigned)std::min<uint64_t>(nr, std::max<uint64_t>(1, (148 * 32) / gx));
    count_hits_kernel<<<dim3(gx, gy), 256, 0, st>>>(cmat, nr, nq, row_offsets, r0, q_offsets, threshold, found, cap, n_found);
    SM_LAUNCHED();
}

}  // namespace smb200
