// GPU unit test of the sort / scan / reduce kernels (sourmash_rust_b200/csrc/sortops.cu)
// against std::sort / std::map on the host.  Built and run by tests/test_gpu_units.py.
#include <algorithm>
#include <cstdio>
#include <map>
#include <vector>

#include "../../sourmash_rust_b200/csrc/device.hpp"
#include "../../sourmash_rust_b200/csrc/kernels.cuh"

using namespace smb200;

static uint64_t rng_state = 99;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

int main() {
    int fails = 0;
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const uint64_t sizes[] = {1, 2, 3, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 10000, 100000, 1000003};
    for (uint64_t n : sizes) {
        for (int variant = 0; variant < 3; variant++) {
            const int bits = variant == 0 ? 64 : (variant == 1 ? 20 : 54);
            std::vector<uint64_t> k(n), v(n);
            for (uint64_t i = 0; i < n; i++) {
                k[i] = rnd();
                if (bits < 64) k[i] &= (1ull << bits) - 1;
                if (variant == 1) k[i] %= 50;  // heavy duplicates
                v[i] = i;
            }
            uint64_t *dk, *dv, *tk, *tv, *uk, *us, *idx;
            void *scan;
            unsigned long long *dn;
            const size_t sb = std::max(radix_sort_scan_bytes(n), scan_tmp_bytes(n)) + 256;
            CK(cudaMalloc(&dk, (n + 1) * 8)); CK(cudaMalloc(&dv, (n + 1) * 8)); CK(cudaMalloc(&tk, (n + 1) * 8));
            CK(cudaMalloc(&tv, (n + 1) * 8)); CK(cudaMalloc(&uk, (n + 1) * 8)); CK(cudaMalloc(&us, (n + 1) * 8));
            CK(cudaMalloc(&idx, (n + 1) * 8)); CK(cudaMalloc(&scan, sb)); CK(cudaMalloc(&dn, 8));
            CK(cudaMemcpy(dk, k.data(), n * 8, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dv, v.data(), n * 8, cudaMemcpyHostToDevice));
            try {
                radix_sort_pairs(dk, dv, n, tk, tv, bits, scan, sb, st);
                reduce_by_key(dk, nullptr, n, uk, us, dn, idx, scan, st);
            } catch (const SourmashError &e) { printf("n=%llu: %s\n", (unsigned long long)n, e.what()); return 1; }
            CK(cudaStreamSynchronize(st));
            std::vector<uint64_t> gk(n), gv(n);
            CK(cudaMemcpy(gk.data(), dk, n * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(gv.data(), dv, n * 8, cudaMemcpyDeviceToHost));
            std::vector<std::pair<uint64_t, uint64_t>> ref(n);
            for (uint64_t i = 0; i < n; i++) ref[i] = {k[i], v[i]};
            std::stable_sort(ref.begin(), ref.end(), [](auto &a, auto &b) { return a.first < b.first; });
            for (uint64_t i = 0; i < n; i++)
                if (gk[i] != ref[i].first || gv[i] != ref[i].second) { if (fails++ < 5) printf("sort mismatch n=%llu bits=%d at %llu\n", (unsigned long long)n, bits, (unsigned long long)i); break; }
            std::map<uint64_t, uint64_t> cnt;
            for (uint64_t x : k) cnt[x]++;
            unsigned long long nu = 0;
            CK(cudaMemcpy(&nu, dn, 8, cudaMemcpyDeviceToHost));
            if (nu != cnt.size()) { if (fails++ < 5) printf("unique count mismatch n=%llu: %llu vs %zu\n", (unsigned long long)n, nu, cnt.size()); }
            else {
                std::vector<uint64_t> hk(nu), hs(nu);
                CK(cudaMemcpy(hk.data(), uk, nu * 8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hs.data(), us, nu * 8, cudaMemcpyDeviceToHost));
                uint64_t j = 0;
                for (auto &kv : cnt) { if (hk[j] != kv.first || hs[j] != kv.second) { if (fails++ < 5) printf("rle mismatch n=%llu at %llu\n", (unsigned long long)n, (unsigned long long)j); break; } j++; }
            }
            // keys-only sort
            CK(cudaMemcpy(dk, k.data(), n * 8, cudaMemcpyHostToDevice));
            radix_sort_pairs(dk, nullptr, n, tk, tv, bits, scan, sb, st);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(gk.data(), dk, n * 8, cudaMemcpyDeviceToHost));
            for (uint64_t i = 0; i < n; i++)
                if (gk[i] != ref[i].first) { if (fails++ < 5) printf("keys-only sort mismatch n=%llu bits=%d\n", (unsigned long long)n, bits); break; }
            // exclusive scan
            CK(cudaMemcpy(dv, v.data(), n * 8, cudaMemcpyHostToDevice));
            scan_exclusive_u64(dv, tv, n, scan, st);
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(gv.data(), tv, n * 8, cudaMemcpyDeviceToHost));
            uint64_t run = 0;
            for (uint64_t i = 0; i < n; i++) { if (gv[i] != run) { if (fails++ < 5) printf("scan mismatch n=%llu at %llu\n", (unsigned long long)n, (unsigned long long)i); break; } run += v[i]; }
            cudaFree(dk); cudaFree(dv); cudaFree(tk); cudaFree(tv); cudaFree(uk); cudaFree(us); cudaFree(idx); cudaFree(scan); cudaFree(dn);
        }
    }
    printf(fails ? "FAILED %d\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
