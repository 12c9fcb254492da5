// CPU check of the host+device bit helpers the sketch kernel is built from
// (sourmash_rust_b200/csrc/kmer_bits.cuh, murmur3.cuh): emulates one tile the way
// sketch.cu:build_views lays it out and compares every window with a byte-wise restatement of
// the reference loop body (canonical = min(kmer, revcomp), src/lib.rs:260-267).
// Built and run by tests/test_host_bits.py (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../sourmash_rust_b200/csrc/kmer_bits.cuh"
#include "../../sourmash_rust_b200/csrc/murmur3.cuh"

using namespace smb200;

static uint64_t rng_state = 0x1234567;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static bool is_dna(uint8_t c) {
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return true; }
    return false;
}
static uint8_t comp(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; }
    return c;
}

template <int K>
static int run(int TILE, int trials) {
    const int B = TILE + ((K - 1 + 15) / 16) * 16;
    int fails = 0;
    for (int t = 0; t < trials; t++) {
        std::vector<uint8_t> raw(B);
        for (int j = 0; j < B; j++) {
            const uint64_t r = rnd();
            uint8_t c = "ACGT"[r & 3];
            if (((r >> 8) & 7) == 0) c = (uint8_t)(c | 0x20);            // lower case
            if (t % 3 == 1 && ((r >> 16) % 97) == 0) c = (uint8_t)(r >> 24);  // arbitrary byte
            if (t % 3 == 2 && ((r >> 16) % 61) == 0) c = 'N';
            raw[j] = c;
        }
        // ---- views, as sketch.cu:build_views ----
        std::vector<uint32_t> fA(B / 4 + 2, 0), rA(B / 4 + 2, 0), f2(B / 16 + 2, 0), r2(B / 16 + 2, 0),
            bad((B + 31) / 32 + 3, 0xFFFFFFFFu);
        uint16_t *f2h = (uint16_t *)f2.data(), *r2h = (uint16_t *)r2.data();
        uint8_t *badb = (uint8_t *)bad.data();
        const uint32_t *raw32 = (const uint32_t *)raw.data();
        const int groups = B / 8;
        for (int g = 0; g < groups; g++) {
            const Oct o = classify8(raw32[2 * g], raw32[2 * g + 1]);
            fA[2 * g] = o.fA0; fA[2 * g + 1] = o.fA1;
            const int rg = groups - 1 - g;
            rA[2 * rg] = o.rA0; rA[2 * rg + 1] = o.rA1;
            f2h[g] = (uint16_t)o.f2; r2h[rg] = (uint16_t)o.r2;
            badb[g] = (uint8_t)o.bad8;
        }
        if (B % 32) ((uint16_t *)bad.data())[B / 16] = 0xFFFF;
        // base validity
        for (int j = 0; j < B; j++) {
            const bool b = (bad[j >> 5] >> (j & 31)) & 1;
            if (b == is_dna(raw[j])) { if (fails++ < 5) printf("K=%d bad-bit mismatch at %d (byte %02x)\n", K, j, raw[j]); }
        }
        // window-start dilation
        for (int w = 0; w < TILE / 32; w++) {
            const uint32_t d = dilate_word(bad[w], bad[w + 1], bad[w + 2], K);
            for (int p = 0; p < 32; p++) {
                const int q = w * 32 + p;
                bool any = false;
                for (int j = q; j < q + K; j++) any |= !is_dna(raw[j]);
                if ((bool)((d >> p) & 1) != any) { if (fails++ < 5) printf("K=%d dilate mismatch at %d\n", K, q); }
            }
        }
        // per window
        for (int i = 0; i < TILE; i++) {
            bool ok = true;
            for (int j = i; j < i + K; j++) ok &= is_dna(raw[j]);
            if (!ok) continue;
            uint8_t fw[K], rc[K];
            for (int j = 0; j < K; j++) fw[j] = (uint8_t)(raw[i + j] & 0xDF);
            for (int j = 0; j < K; j++) rc[j] = comp(fw[K - 1 - j]);
            const bool fw_lt = memcmp(fw, rc, K) < 0;
            const uint64_t want = murmur3_h1_bytes(fw_lt ? fw : rc, K, 42);

            uint32_t ef[KmerGeom<K>::NE], er[KmerGeom<K>::NE];
            extract2<K>(f2.data(), i, ef);
            const int ri = B - K - i;
            extract2<K>(r2.data(), ri, er);
            const bool use_fw = canonical_is_fw<K>(ef, er);
            uint32_t kw[KmerGeom<K>::NW];
            extractA<K>(use_fw ? fA.data() : rA.data(), use_fw ? i : ri, kw);
            const uint64_t got = murmur3_h1_words<K>(kw, 42);
            if (use_fw != fw_lt && memcmp(fw, rc, K) != 0) { if (fails++ < 5) printf("K=%d strand mismatch at %d\n", K, i); }
            if (got != want) { if (fails++ < 5) printf("K=%d hash mismatch at %d: %llu vs %llu\n", K, i, (unsigned long long)got, (unsigned long long)want); }
        }
    }
    return fails;
}

// dilate_word for every span 0..64 against the bit-by-bit definition
static int dilate_spans() {
    int fails = 0;
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (int trial = 0; trial < 400; trial++) {
        uint32_t in[3];
        for (int j = 0; j < 3; j++) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            in[j] = (uint32_t)x & ((trial % 3) ? (uint32_t)(x >> 32) : 0xFFFFFFFFu) & ((trial % 5 == 0) ? (uint32_t)(x >> 20) : 0xFFFFFFFFu);
        }
        for (int span = 0; span <= 64; span++) {
            const uint32_t d = dilate_word(in[0], in[1], in[2], span);
            for (int p = 0; p < 32; p++) {
                bool any = false;
                for (int j = p; j < p + span; j++) any |= (in[j >> 5] >> (j & 31)) & 1u;
                if ((bool)((d >> p) & 1) != any) { if (fails++ < 5) printf("dilate span %d mismatch at bit %d\n", span, p); }
            }
        }
    }
    return fails;
}

int main() {
    // known answers: tests/test.rs:5 of the reference, and the oracle-derived k=21/31/51 vectors
    int fails = 0;
    if (murmur3_h1_bytes((const uint8_t *)"ACG", 3, 42) != 1731421407650554201ULL) { printf("KAT ACG failed\n"); fails++; }
    if (murmur3_h1_bytes((const uint8_t *)"GTCACCCGGTGCTGGGCGGCA", 21, 42) != 529147935188082428ULL) { printf("KAT k21 failed\n"); fails++; }
    if (murmur3_h1_bytes((const uint8_t *)"GCTCAACCTAGTCACCCGGTGCTGGGCGGCA", 31, 42) != 13824550005532878703ULL) { printf("KAT k31 failed\n"); fails++; }
    if (murmur3_h1_bytes((const uint8_t *)"CTCATTGCAGGTTAATCATGGCTCAACCTAGTCACCCGGTGCTGGGCGGCA", 51, 42) != 14476355676789784531ULL) { printf("KAT k51 failed\n"); fails++; }
    fails += dilate_spans();
    fails += run<21>(2048, 6);
    fails += run<31>(2048, 6);
    fails += run<51>(2048, 6);
    fails += run<4>(256, 3);
    fails += run<10>(256, 3);
    fails += run<32>(256, 3);
    fails += run<33>(256, 3);
    fails += run<64>(256, 3);
    printf(fails ? "FAILED %d\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
