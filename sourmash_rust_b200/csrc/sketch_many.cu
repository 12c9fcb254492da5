// sketch_many.cu -- one fresh sketch PER SEQUENCE of a batch, all in one pass: the sketching half of
// "sketch N genomes, then compare them all" (BASELINE configs 3 and 5).
//
// Equivalent reference loop (src/lib.rs:142-174, 252-274; src/index/linear.rs:47-50):
//     for s in sequences { let mut mh = KmerMinHash::new(num, ksize, false, seed, max_hash, false);
//                          mh.add_sequence(s, /*force=*/true); collection.push(mh) }
// Calling the per-object ABI N times pays a launch, a threshold read-back and a fold per genome -- for a
// 5 Mbp genome ten times the 28 us its k-mers take.  Here the sketch kernel (sketch.cu) runs ONCE over
// the whole batch and reports every surviving hash with its position; positions become sequence
// ids, two stable radix sorts (by hash, then by sequence) group them, and a flag / scan / compact
// pass drops duplicates, applies `num` per sequence and leaves the packed CSR of a SketchCollection.
//   scaled sketches (num = 0, max_hash > 0): the threshold is max_hash, exact by construction;
//   num sketches (num > 0, max_hash = 0): the kernel's threshold is an estimate from the median
//     sequence length (16 x num expected survivors); a sequence that ends up with fewer than `num`
//     distinct survivors under a threshold that was not "everything" is sketched again on its own
//     through the per-object path, so the result is exact for every input.
#include <algorithm>
#include <memory>

#include "collection.hpp"
#include "kernels.cuh"

namespace smb200 {

namespace {

__global__ void __launch_bounds__(256) pos_to_seq_kernel(const uint64_t *pos /* may alias seq */, uint64_t n,
                                                         const uint64_t *__restrict__ offsets, uint64_t n_seqs, uint64_t *seq) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t p = pos[i];
        uint64_t lo = 0, hi = n_seqs;  // largest s with offsets[s] <= p
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (offsets[mid] <= p) lo = mid; else hi = mid;
        }
        seq[i] = lo;
    }
}
// sorted by (seq, hash): flags[i] = first occurrence of its (seq, hash); seg_head[i] = first entry of its sequence
__global__ void __launch_bounds__(256) many_flags_kernel(const uint64_t *__restrict__ seq, const uint64_t *__restrict__ hash,
                                                         uint64_t n, uint64_t *flags) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = (i == 0 || seq[i] != seq[i - 1] || hash[i] != hash[i - 1]) ? 1 : 0;
}
// seg_base[s] = number of distinct entries before the first entry of sequence s
__global__ void __launch_bounds__(256) many_seg_base_kernel(const uint64_t *__restrict__ seq, const uint64_t *__restrict__ pre,
                                                            uint64_t n, uint64_t *seg_base) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (i == 0 || seq[i] != seq[i - 1]) seg_base[seq[i]] = pre[i];
}
// keep[i] = distinct and among the first `num` distinct hashes of its sequence (num = 0: all);
// distinct[s] / kept[s] count per sequence
__global__ void __launch_bounds__(256) many_keep_kernel(const uint64_t *__restrict__ seq, const uint64_t *__restrict__ flags,
                                                        const uint64_t *__restrict__ pre, const uint64_t *__restrict__ seg_base,
                                                        uint64_t n, uint32_t num, uint64_t *keep, unsigned long long *distinct,
                                                        unsigned long long *kept) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = 0;
        if (flags[i]) {
            const uint64_t s = seq[i];
            atomicAdd(&distinct[s], 1ull);
            if (num == 0 || pre[i] - seg_base[s] < num) {
                k = 1;
                atomicAdd(&kept[s], 1ull);
            }
        }
        keep[i] = k;
    }
}
__global__ void __launch_bounds__(256) many_compact_kernel(const uint64_t *__restrict__ hash, const uint64_t *__restrict__ keep,
                                                           const uint64_t *__restrict__ kpre, uint64_t n, uint64_t *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (keep[i]) out[kpre[i]] = hash[i];
}
__global__ void many_set_thr_kernel(unsigned long long *thr, unsigned long long v, unsigned long long *counter, uint32_t *tile_ctr) {
    *thr = v;
    *counter = 0;
    tile_ctr[0] = 0;
    tile_ctr[1] = 0;
}

constexpr uint64_t U64_MAX = ~0ull;
unsigned grid_for(uint64_t n) { return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, 148 * 16)); }
int bits_of(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return std::max(1, b); }

}  // namespace

SketchCollection *sketch_collection(const uint8_t *buf, const uint64_t *offsets, uint64_t n_seqs, uint32_t num, uint32_t ksize,
                                    uint64_t seed, uint64_t max_hash, bool on_device) {
    if ((num == 0) == (max_hash == 0))
        throw_internal("sketch_collection takes either num > 0 (bottom-num sketches) or max_hash > 0 (scaled sketches)");
    if (ksize == 0) throw_internal("ksize 0");
    Context &ctx = Context::get();
    cudaStream_t st = ctx.stream;
    std::unique_ptr<SketchCollection> c(new SketchCollection());
    c->have_params = true;
    c->ksize = ksize; c->seed = seed; c->max_hash = max_hash; c->is_protein = false;
    c->n_rows = n_seqs;
    c->h_nums.assign(n_seqs, num);
    c->h_offsets.assign(n_seqs + 1, 0);
    if (n_seqs == 0) { c->dirty = false; return c.release(); }

    // ---- the batch and its offsets on the device -----------------------------------------------------
    std::vector<uint64_t> h_off(n_seqs + 1);
    if (on_device) {
        SM_CUDA(cudaMemcpyAsync(h_off.data(), offsets, (n_seqs + 1) * 8, cudaMemcpyDeviceToHost, st));
        ctx.sync();
    } else {
        std::copy(offsets, offsets + n_seqs + 1, h_off.begin());
    }
    if (h_off[0] != 0) throw_internal("offsets[0] must be 0");
    for (uint64_t s = 0; s < n_seqs; s++)
        if (h_off[s + 1] < h_off[s]) throw_internal("offsets must be non-decreasing");
    const uint64_t n = h_off[n_seqs];
    const uint8_t *d_buf = buf;
    const uint64_t *d_off = offsets;
    if (!on_device) {
        ctx.ascii.reserve(((n + 15) & ~15ull) + 256);
        ctx.offsets.reserve((n_seqs + 1) * 8);
        if (n) SM_CUDA(cudaMemcpyAsync(ctx.ascii.p, buf, n, cudaMemcpyHostToDevice, st));
        SM_CUDA(cudaMemcpyAsync(ctx.offsets.p, h_off.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
        d_buf = ctx.ascii.as<uint8_t>();
        d_off = ctx.offsets.as<uint64_t>();
    } else if ((reinterpret_cast<uintptr_t>(buf) & 15) != 0) {
        throw_internal("device sequence buffer must be 16-byte aligned");
    }

    // ---- threshold ----------------------------------------------------------------------------------
    double frac = 1.0;
    uint64_t thr = U64_MAX;
    if (max_hash) {
        thr = max_hash;
        frac = (double)max_hash / 18446744073709551616.0;
    } else {
        std::vector<uint64_t> lens(n_seqs);
        for (uint64_t s = 0; s < n_seqs; s++) lens[s] = h_off[s + 1] - h_off[s];
        std::nth_element(lens.begin(), lens.begin() + n_seqs / 2, lens.end());
        const double median = (double)std::max<uint64_t>(1, lens[n_seqs / 2]);
        frac = std::min(1.0, 16.0 * (double)num / median);
        if (frac < 1.0) thr = (uint64_t)(frac * 18446744073709551616.0);
    }
    const bool thr_is_everything = (thr == U64_MAX);

    // ---- sketch kernel over the whole batch: survivors (hash, position) -----------------------------------
    // control words in the context's device scalars: threshold, survivor counter, tile counter pair
    unsigned long long *d_thr = ctx.dsc(SC_THRESH), *d_count = ctx.dsc(SC_CNT), *d_tiles = ctx.dsc(SC_PAIR2);
    uint64_t cap = (uint64_t)(frac * 1.25 * (double)n) + 4096;
    uint64_t count = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt > 8) throw_internal("sketch_collection: survivor buffer did not converge");
        ctx.join[0].reserve((cap + 1) * 8);
        ctx.join[1].reserve((cap + 1) * 8);
        many_set_thr_kernel<<<1, 1, 0, st>>>(d_thr, thr, d_count, reinterpret_cast<uint32_t *>(d_tiles));
        SM_LAUNCHED();
        SketchBatch sb;
        sb.buf = d_buf; sb.n = n; sb.n_limit = n; sb.offsets = d_off; sb.n_seqs = n_seqs; sb.read_len = 0; sb.tile_lo = 0;
        sb.seed = seed; sb.pos_base = 0; sb.first_bad = nullptr;  // force = true: invalid k-mers are skipped
        sb.tile_ctr = reinterpret_cast<uint32_t *>(d_tiles);
        SketchOut out;
        out.thr = reinterpret_cast<const uint64_t *>(d_thr);
        out.hash = ctx.join[0].as<uint64_t>();
        out.pos = ctx.join[1].as<uint64_t>();
        out.cap = cap;
        out.counter = d_count;
        if (n) launch_sketch(ksize, sb, out, sketch_tile_count(n, n), ctx.sm_count, st);
        uint64_t two[2];
        ctx.fetch2(d_count, d_count, two);
        count = two[0];
        if (count <= cap) break;
        cap = count + 4096;  // entries beyond cap were dropped but counted: run again with room for all
    }

    std::vector<uint64_t> h_distinct(n_seqs, 0), h_kept(n_seqs, 0);
    if (count) {
        // ---- positions -> sequence ids; group by (sequence, hash) ---------------------------------------
        uint64_t *hash = ctx.join[0].as<uint64_t>(), *seq = ctx.join[1].as<uint64_t>();
        pos_to_seq_kernel<<<grid_for(count), 256, 0, st>>>(seq, count, d_off, n_seqs, seq);  // in place: pos -> seq
        SM_LAUNCHED();
        ctx.sort_tmp_k.reserve((count + 1) * 8);
        ctx.sort_tmp_v.reserve((count + 1) * 8);
        ctx.scan_tmp.reserve(std::max(radix_sort_scan_bytes(count), scan_tmp_bytes(std::max<uint64_t>(count, n_seqs + 1))) + 256);
        radix_sort_pairs(hash, seq, count, ctx.sort_tmp_k.as<uint64_t>(), ctx.sort_tmp_v.as<uint64_t>(), bits_of(thr), ctx.scan_tmp.p,
                         ctx.scan_tmp.cap, st);
        radix_sort_pairs(seq, hash, count, ctx.sort_tmp_k.as<uint64_t>(), ctx.sort_tmp_v.as<uint64_t>(), bits_of(n_seqs), ctx.scan_tmp.p,
                         ctx.scan_tmp.cap, st);
        // ---- distinct, first `num` per sequence, compact ---------------------------------------------------
        for (int i = 2; i < 6; i++) ctx.join[i].reserve((std::max(count, n_seqs) + 2) * 8);
        ctx.misc[0].reserve((n_seqs + 2) * 8);
        ctx.misc[1].reserve((n_seqs + 2) * 8);
        uint64_t *flags = ctx.join[2].as<uint64_t>(), *pre = ctx.join[3].as<uint64_t>(), *keep = ctx.join[4].as<uint64_t>();
        uint64_t *seg_base = ctx.join[5].as<uint64_t>();
        unsigned long long *distinct = ctx.misc[0].as<unsigned long long>(), *kept = ctx.misc[1].as<unsigned long long>();
        SM_CUDA(cudaMemsetAsync(seg_base, 0, (n_seqs + 1) * 8, st));
        SM_CUDA(cudaMemsetAsync(distinct, 0, (n_seqs + 1) * 8, st));
        SM_CUDA(cudaMemsetAsync(kept, 0, (n_seqs + 1) * 8, st));
        many_flags_kernel<<<grid_for(count), 256, 0, st>>>(seq, hash, count, flags);
        SM_LAUNCHED();
        scan_exclusive_u64(flags, pre, count, ctx.scan_tmp.p, st);
        many_seg_base_kernel<<<grid_for(count), 256, 0, st>>>(seq, pre, count, seg_base);
        SM_LAUNCHED();
        many_keep_kernel<<<grid_for(count), 256, 0, st>>>(seq, flags, pre, seg_base, count, num, keep, distinct, kept);
        SM_LAUNCHED();
        scan_exclusive_u64(keep, pre, count, ctx.scan_tmp.p, st);  // pre now: output slot of every kept entry
        SM_CUDA(cudaMemcpyAsync(h_distinct.data(), distinct, n_seqs * 8, cudaMemcpyDeviceToHost, st));
        SM_CUDA(cudaMemcpyAsync(h_kept.data(), kept, n_seqs * 8, cudaMemcpyDeviceToHost, st));
        ctx.sync();
        uint64_t total = 0;
        for (uint64_t s = 0; s < n_seqs; s++) { c->h_offsets[s] = total; total += h_kept[s]; }
        c->h_offsets[n_seqs] = total;
        c->d_hashes.reserve((total + 4) * 8);
        many_compact_kernel<<<grid_for(count), 256, 0, st>>>(hash, keep, pre, count, c->d_hashes.as<uint64_t>());
        SM_LAUNCHED();
        c->n_hashes = total;
    }

    // ---- num sketches: sequences the estimated threshold cut short are sketched again on their own ----------
    std::vector<uint64_t> redo;
    if (num && !thr_is_everything)
        for (uint64_t s = 0; s < n_seqs; s++)
            if (h_distinct[s] < num) redo.push_back(s);
    if (!redo.empty()) {
        // patch on the host: rare, and those sequences are short by construction
        std::vector<uint64_t> h_hashes(c->n_hashes);
        if (c->n_hashes) SM_CUDA(cudaMemcpyAsync(h_hashes.data(), c->d_hashes.p, c->n_hashes * 8, cudaMemcpyDeviceToHost, st));
        ctx.sync();
        std::vector<std::vector<uint64_t>> fixed(redo.size());
        std::vector<uint8_t> tmp;
        for (size_t r = 0; r < redo.size(); r++) {
            const uint64_t s = redo[r], len = h_off[s + 1] - h_off[s];
            KmerMinHash mh(num, ksize, false, seed, 0, false);
            if (len) {
                tmp.resize(len);
                if (on_device) {
                    SM_CUDA(cudaMemcpyAsync(tmp.data(), buf + h_off[s], len, cudaMemcpyDeviceToHost, st));
                    ctx.sync();
                } else {
                    std::copy(buf + h_off[s], buf + h_off[s] + len, tmp.begin());
                }
                mh.add_sequence(tmp.data(), len, true);
            }
            fixed[r] = mh.mins();
        }
        std::vector<uint64_t> out_h, out_o(n_seqs + 1, 0);
        size_t r = 0;
        for (uint64_t s = 0; s < n_seqs; s++) {
            out_o[s] = out_h.size();
            if (r < redo.size() && redo[r] == s) {
                out_h.insert(out_h.end(), fixed[r].begin(), fixed[r].end());
                r++;
            } else {
                out_h.insert(out_h.end(), h_hashes.begin() + c->h_offsets[s], h_hashes.begin() + c->h_offsets[s + 1]);
            }
        }
        out_o[n_seqs] = out_h.size();
        c->h_offsets = out_o;
        c->n_hashes = out_h.size();
        c->d_hashes.reserve((c->n_hashes + 4) * 8);
        if (c->n_hashes) SM_CUDA(cudaMemcpyAsync(c->d_hashes.p, out_h.data(), c->n_hashes * 8, cudaMemcpyHostToDevice, st));
        ctx.sync();
    }

    c->d_offsets.reserve((n_seqs + 1) * 8);
    c->d_nums.reserve((n_seqs + 1) * 4);
    SM_CUDA(cudaMemcpyAsync(c->d_offsets.p, c->h_offsets.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice, st));
    SM_CUDA(cudaMemcpyAsync(c->d_nums.p, c->h_nums.data(), n_seqs * 4, cudaMemcpyHostToDevice, st));
    c->max_len = 0;
    for (uint64_t s = 0; s < n_seqs; s++) c->max_len = std::max<uint32_t>(c->max_len, (uint32_t)(c->h_offsets[s + 1] - c->h_offsets[s]));
    ctx.sync();
    c->dirty = false;
    return c.release();
}

}  // namespace smb200
