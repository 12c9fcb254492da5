// ThreadSanitizer harness for the host side of the C ABI (no GPU needed): N threads drive the calls that never reach the
// device -- hashing, handle lifecycle and getters, the thread-local error slot, signature JSON in and out -- plus the
// failure path of a call that does need the device (every thread then finds "no usable CUDA device" at once), on
// private handles and on handles shared by all threads (the per-handle lock of csrc/ffi.cpp must make that safe).
// Built and run by tests/test_tsan_cpu.py against a -fsanitize=thread build of the library; exit status 0 = every
// result was what one thread alone computed (ThreadSanitizer reports races on stderr and sets its own exit code).
//   usage: tsan_host_calls <signatures.json> [threads=8] [rounds=40]
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "sourmash.h"

static std::atomic<int> g_failures{0};
#define CHECK(cond)                                                                   \
    do {                                                                              \
        if (!(cond)) {                                                                \
            fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #cond);        \
            g_failures.fetch_add(1);                                                  \
        }                                                                             \
    } while (0)

static std::string take(SourmashStr s) {
    std::string out(s.data ? s.data : "", s.len);
    sourmash_str_free(&s);
    return out;
}

// every signature of the file -> its JSON, concatenated (what one thread alone gets is the expected value)
static std::string reload_and_save(const std::string &json) {
    uintptr_t n = 0;
    Signature **sigs = signatures_load_buffer(json.data(), json.size(), false, 0, nullptr, &n);
    if (!sigs) return "<load failed>";
    std::string out = take(signatures_save_buffer(sigs, n));
    for (uintptr_t i = 0; i < n; i++) {
        out += "|" + take(signature_get_name(sigs[i])) + "|" + take(signature_save_json(sigs[i]));
        uintptr_t m = 0;
        KmerMinHash **mhs = signature_get_mhs(sigs[i], &m);
        for (uintptr_t j = 0; mhs && j < m; j++) {
            out += "," + std::to_string(kmerminhash_ksize(mhs[j])) + ":" + std::to_string(kmerminhash_num(mhs[j]));
            kmerminhash_free(mhs[j]);
        }
        free(mhs);
        signature_free(sigs[i]);
    }
    free(sigs);
    return out;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s signatures.json [threads] [rounds]\n", argv[0]);
        return 2;
    }
    const int n_threads = argc > 2 ? atoi(argv[2]) : 8, rounds = argc > 3 ? atoi(argv[3]) : 40;
    std::string json;
    {
        FILE *f = fopen(argv[1], "rb");
        if (!f) { perror(argv[1]); return 2; }
        char buf[65536];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, f)) > 0) json.append(buf, k);
        fclose(f);
    }
    sourmash_init();
    const char *kmer = "ACGTTGCAACGTTGCAACGTTGCAACGTTGC";
    const uint64_t want_hash = hash_murmur(kmer, 42);
    const std::string want_json = reload_and_save(json);
    CHECK(want_json != "<load failed>" && want_json.size() > 100);

    // handles every thread uses at once
    uintptr_t n_shared = 0;
    Signature **shared_sigs = signatures_load_buffer(json.data(), json.size(), false, 0, nullptr, &n_shared);
    CHECK(shared_sigs && n_shared > 0);
    const std::string want_shared = n_shared ? take(signature_save_json(shared_sigs[0])) : "";
    KmerMinHash *shared_mh = kmerminhash_new(500, 31, false, 42, 0, true);
    CHECK(shared_mh != nullptr);

    std::vector<std::thread> threads;
    for (int t = 0; t < n_threads; t++) {
        threads.emplace_back([&, t]() {
            for (int r = 0; r < rounds; r++) {
                CHECK(hash_murmur(kmer, 42) == want_hash);
                // private handle: lifecycle and getters
                KmerMinHash *mh = kmerminhash_new(100 + t, 21, (t & 1) != 0, 42 + r, 0, (r & 1) != 0);
                CHECK(mh && kmerminhash_num(mh) == (uint32_t)(100 + t) && kmerminhash_ksize(mh) == 21);
                CHECK(kmerminhash_is_protein(mh) == ((t & 1) != 0) && kmerminhash_seed(mh) == (uint64_t)(42 + r));
                CHECK(kmerminhash_track_abundance(mh) == ((r & 1) != 0) && kmerminhash_max_hash(mh) == 0);
                // the error slot is this thread's own: even threads raise and read an error, odd ones must never see one
                if ((t & 1) == 0) {
                    CHECK(kmerminhash_get_mins_size(nullptr) == 0);
                    CHECK(sourmash_err_get_last_code() == SOURMASH_ERROR_CODE_PANIC);
                    CHECK(!take(sourmash_err_get_last_message()).empty());
                    sourmash_err_clear();
                }
                CHECK(sourmash_err_get_last_code() == 0);
                // a call that needs the device: without one every thread fails the same way, at the same time
                // (with one it simply succeeds); either way the slot is read and cleared by the thread that raised it
                kmerminhash_add_sequence(mh, "ACGTACGTTGCATGCAACGTTTGACCATGACCA", true);
                const SourmashErrorCode code = sourmash_err_get_last_code();
                if (code != 0) {
                    CHECK(take(sourmash_err_get_last_message()).find("CUDA") != std::string::npos);
                    sourmash_err_clear();
                }
                kmerminhash_free(mh);
                // signature JSON in and out, private objects
                if (r % 4 == 0) CHECK(reload_and_save(json) == want_json);
                // shared handles: getters and the writer, all threads at once
                CHECK(kmerminhash_num(shared_mh) == 500 && kmerminhash_track_abundance(shared_mh));
                if (n_shared) {
                    CHECK(take(signature_save_json(shared_sigs[0])) == want_shared);
                    KmerMinHash *first = signature_first_mh(shared_sigs[r % n_shared]);
                    CHECK(first != nullptr);
                    kmerminhash_free(first);
                }
            }
        });
    }
    for (auto &th : threads) th.join();
    kmerminhash_free(shared_mh);
    for (uintptr_t i = 0; i < n_shared; i++) signature_free(shared_sigs[i]);
    free(shared_sigs);
    const int bad = g_failures.load();
    printf("tsan_host_calls: %d threads x %d rounds, %d failed checks\n", n_threads, rounds, bad);
    return bad ? 1 : 0;
}
