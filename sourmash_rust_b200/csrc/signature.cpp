// signature.cpp -- Signature JSON output (byte-compatible with serde_json's compact writer for
// the field order of src/lib.rs:79-100 and 546-565) and input (Signature::load_signatures,
// src/lib.rs:593-645).
#include "signature.hpp"

#include "md5.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>
#include <sstream>
#include <system_error>
#include <thread>

#include <zlib.h>

namespace smb200 {

// ---------------------------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------------------------
static void put_str(std::string &o, const std::string &s) {  // serde_json string escaping
    static const char *hex = "0123456789abcdef";
    o.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\b': o += "\\b"; break;
        case '\f': o += "\\f"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        default:
            if (c < 0x20) { o += "\\u00"; o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
            else o.push_back((char)c);
        }
    }
    o.push_back('"');
}
// Decimal digits of v into buf, returns the count (1..20).  buf must have 20 writable bytes: all 20 are
// stored, the first `count` are the number.  The value is cut into 4 + 8 + 8 digits so that the three
// conversions are independent 32-bit chains rather than one 64-bit divide chain.
static const char kDigitPairs[] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
static inline void put_8_digits(uint32_t x, char *out) {  // x < 10^8, zero padded
    const uint32_t hi = x / 10000, lo = x % 10000;
    memcpy(out + 0, kDigitPairs + 2 * (hi / 100), 2);
    memcpy(out + 2, kDigitPairs + 2 * (hi % 100), 2);
    memcpy(out + 4, kDigitPairs + 2 * (lo / 100), 2);
    memcpy(out + 6, kDigitPairs + 2 * (lo % 100), 2);
}
static inline int u64_digits(uint64_t v, char *buf) {
    char tmp[40];
    const uint64_t top = v / 10000000000000000ull, rest = v % 10000000000000000ull;  // top < 1845
    const uint32_t mid = (uint32_t)(rest / 100000000ull), low = (uint32_t)(rest % 100000000ull);
    memcpy(tmp + 0, kDigitPairs + 2 * (top / 100), 2);
    memcpy(tmp + 2, kDigitPairs + 2 * (top % 100), 2);
    put_8_digits(mid, tmp + 4);
    put_8_digits(low, tmp + 12);
    int lz = 0;
    while (lz < 19 && tmp[lz] == '0') lz++;
    memcpy(buf, tmp + lz, 20);
    return 20 - lz;
}
static void put_u64(std::string &o, uint64_t v) {
    char buf[20];
    o.append(buf, (size_t)u64_digits(v, buf));
}
// "[a,b,c]"; when `md5` is given it also receives every element's digits (no separators), which is what the
// md5sum field is taken over (lib.rs:72-77)
static void put_u64_array(std::string &o, const std::vector<uint64_t> &v, Md5 *md5 = nullptr) {
    const size_t base = o.size();
    o.resize(base + 2 + v.size() * 21);
    char *w = &o[base];
    char plain[4096 + 20];
    size_t np = 0;
    *w++ = '[';
    for (size_t i = 0; i < v.size(); i++) {
        if (i) *w++ = ',';
        const int n = u64_digits(v[i], w);
        if (md5) {
            memcpy(plain + np, w, (size_t)n);
            np += (size_t)n;
            if (np >= 4096) { md5->update(reinterpret_cast<const uint8_t *>(plain), np); np = 0; }
        }
        w += n;
    }
    *w++ = ']';
    if (md5 && np) md5->update(reinterpret_cast<const uint8_t *>(plain), np);
    o.resize((size_t)(w - o.data()));
}
// shortest decimal that round-trips, in the shape serde_json (ryu) prints: "0.4", "1.0", "1e21"
// f64 the way serde_json writes it (the ryu crate's `pretty` layout): the shortest digit string that
// round-trips, then with kk = position of the decimal point relative to those digits
//   digits then zeros then ".0" when the value is an integer below 1e16, "12.34", "0.001234" down to 1e-5,
//   and d[.ddd]e[-]x otherwise.
static void put_f64(std::string &o, double x) {
    if (!std::isfinite(x)) { o += "null"; return; }
    if (x == 0) { o += std::signbit(x) ? "-0.0" : "0.0"; return; }
    char buf[48];
    for (int prec = 0; prec <= 16; prec++) {
        snprintf(buf, sizeof buf, "%.*e", prec, x);
        if (strtod(buf, nullptr) == x) break;
    }
    const char *p = buf;
    if (*p == '-') { o.push_back('-'); p++; }
    std::string digits;
    for (; *p && *p != 'e'; p++)
        if (*p != '.') digits.push_back(*p);
    const int e10 = atoi(p + 1);                    // value = d.ddd x 10^e10
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int len = (int)digits.size();
    const int k = e10 - (len - 1);                  // value = digits x 10^k
    const int kk = len + k;
    if (k >= 0 && kk <= 16) {
        o += digits;
        o.append((size_t)k, '0');
        o += ".0";
    } else if (kk > 0 && kk <= 16) {
        o.append(digits, 0, (size_t)kk);
        o.push_back('.');
        o.append(digits, (size_t)kk, std::string::npos);
    } else if (kk > -5 && kk <= 0) {
        o += "0.";
        o.append((size_t)(-kk), '0');
        o += digits;
    } else {
        o.push_back(digits[0]);
        if (len > 1) {
            o.push_back('.');
            o.append(digits, 1, std::string::npos);
        }
        o.push_back('e');
        o += std::to_string(kk - 1);
    }
}
static void put_minhash(std::string &o, KmerMinHash &mh) {  // lib.rs:62-102
    o += "{\"num\":"; put_u64(o, mh.num);
    o += ",\"ksize\":"; put_u64(o, mh.ksize);
    o += ",\"seed\":"; put_u64(o, mh.seed);
    o += ",\"max_hash\":"; put_u64(o, mh.max_hash);
    Md5 md5;                                     // over ksize and every min in decimal, lib.rs:72-77
    md5.update(std::to_string(mh.ksize));
    o += ",\"mins\":"; put_u64_array(o, mh.mins(), &md5);
    o += ",\"md5sum\":\""; o += md5.hex(); o.push_back('"');
    if (mh.track_abundance()) { o += ",\"abundances\":"; put_u64_array(o, mh.abunds()); }
    o += ",\"molecule\":"; o += mh.is_protein ? "\"protein\"" : "\"DNA\"";
    o.push_back('}');
}

void Signature::to_json(std::string &o) {
    o += "{\"class\":"; put_str(o, class_);
    o += ",\"email\":"; put_str(o, email);
    o += ",\"hash_function\":"; put_str(o, hash_function);
    o += ",\"filename\":"; if (has_filename) put_str(o, filename); else o += "null";
    o += ",\"name\":"; if (has_name) put_str(o, name); else o += "null";
    o += ",\"license\":"; put_str(o, license);
    o += ",\"signatures\":[";
    for (size_t i = 0; i < signatures.size(); i++) {
        if (i) o.push_back(',');
        put_minhash(o, *signatures[i]);
    }
    o += "],\"version\":"; put_f64(o, version);
    o.push_back('}');
}

// host threads for reading / writing large signature files (SMB200_JSON_THREADS; default min(8, cores))
static unsigned json_threads() {
    static const unsigned n = [] {
        const char *e = getenv("SMB200_JSON_THREADS");
        if (e && *e) return (unsigned)std::max(1, atoi(e));
        return std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
    }();
    return n;
}

// fn(0) .. fn(T-1), each on its own thread (fn(0) on the caller's); a piece whose thread cannot be started runs
// on the caller's thread instead.  fn must not throw.
template <typename F>
static void run_pieces(unsigned T, F fn) {
    std::vector<std::thread> workers;
    std::vector<unsigned> inline_pieces{0};
    for (unsigned t = 1; t < T; t++) {
        try {
            workers.emplace_back(fn, t);
        } catch (const std::system_error &) {
            inline_pieces.push_back(t);
        }
    }
    for (unsigned t : inline_pieces) fn(t);
    for (auto &w : workers) w.join();
}

// Output goes to a malloc'd buffer the C ABI hands to the caller as is (SourmashStr, freed by sourmash_str_free).
// Large outputs are written in contiguous runs of signatures on several host threads; each thread then copies
// its run into place.  The host copies of the sketches are fetched first, on the calling thread (a device
// read-back is not thread-safe).
char *signatures_to_json(Signature *const *sigs, size_t n, size_t *len) {
    for (size_t i = 0; i < n; i++)
        if (!sigs[i]) throw SourmashError(ERR_PANIC, "sourmash panicked: null signature pointer");
    size_t hashes = 0;
    for (size_t i = 0; i < n; i++)
        for (auto &mh : sigs[i]->signatures) hashes += mh->mins().size();
    const unsigned T = std::max(1u, (unsigned)std::min<size_t>({(size_t)json_threads(), n, hashes / 50000}));
    std::vector<std::string> parts(T);
    std::vector<std::exception_ptr> errs(T);
    auto write_run = [&](unsigned t) {
        try {
            const size_t lo = n * t / T, hi = n * (t + 1) / T;
            for (size_t i = lo; i < hi; i++) {
                if (i > lo) parts[t].push_back(',');
                sigs[i]->to_json(parts[t]);
            }
        } catch (...) {
            errs[t] = std::current_exception();
        }
    };
    run_pieces(T, write_run);
    for (auto &e : errs)
        if (e) std::rethrow_exception(e);
    std::vector<size_t> at(T);
    size_t total = 1;  // '['
    for (unsigned t = 0; t < T; t++) {
        if (t) total++;  // ',' between runs
        at[t] = total;
        total += parts[t].size();
    }
    total++;  // ']'
    char *out = static_cast<char *>(malloc(total));
    if (!out) throw std::bad_alloc();
    out[0] = '[';
    out[total - 1] = ']';
    run_pieces(T, [&](unsigned t) {
        if (t) out[at[t] - 1] = ',';
        memcpy(out + at[t], parts[t].data(), parts[t].size());
        std::string().swap(parts[t]);
    });
    *len = total;
    return out;
}

Signature *Signature::clone_meta() const {
    Signature *s = new Signature();
    s->class_ = class_; s->email = email; s->hash_function = hash_function;
    s->has_filename = has_filename; s->filename = filename;
    s->has_name = has_name; s->name = name;
    s->license = license; s->version = version;
    return s;
}

bool Signature::equals(Signature &o) {
    const bool meta = class_ == o.class_ && email == o.email && hash_function == o.hash_function &&
                      has_filename == o.has_filename && (!has_filename || filename == o.filename) &&
                      has_name == o.has_name && (!has_name || name == o.name);
    if (signatures.empty() || o.signatures.empty())  // the reference indexes [0] and panics
        throw SourmashError(ERR_PANIC, "sourmash panicked: index out of bounds: the len is 0 but the index is 0");
    return meta && signatures[0]->equals(*o.signatures[0]);
}

// ---------------------------------------------------------------------------------------------
// reader: a typed, single-pass JSON reader with the acceptance rules of serde_json + the serde derive
// of the two structs (lib.rs:104-139 TempSig, lib.rs:546-565 Signature): strict RFC 8259 grammar,
// unknown fields skipped, duplicate known fields rejected, integers must fit the field's type, both
// the map and the sequence form of a struct are read.  `mins` / `abundances` go straight into u64
// vectors -- no value tree is built.
// ---------------------------------------------------------------------------------------------
namespace {
[[noreturn]] void jfail(const std::string &m) { throw SourmashError(ERR_UNKNOWN, "JSON error: " + m); }

// strict UTF-8 (what Rust's str::from_utf8 accepts): no overlong forms, no surrogates, nothing above U+10FFFF
bool utf8_ok(const std::string &s) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(s.data()), *e = p + s.size();
    while (p < e) {
        const unsigned c = *p;
        if (c < 0x80) { p++; continue; }
        int n;
        unsigned lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) n = 1;
        else if (c == 0xE0) { n = 2; lo = 0xA0; }
        else if (c == 0xED) { n = 2; hi = 0x9F; }
        else if (c >= 0xE1 && c <= 0xEF) n = 2;
        else if (c == 0xF0) { n = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) n = 3;
        else if (c == 0xF4) { n = 3; hi = 0x8F; }
        else return false;
        if (e - p <= n) return false;
        if (p[1] < lo || p[1] > hi) return false;
        for (int i = 2; i <= n; i++)
            if (p[i] < 0x80 || p[i] > 0xBF) return false;
        p += n + 1;
    }
    return true;
}

struct JReader {
    const char *p, *e;
    JReader(const char *d, size_t n) : p(d), e(d + n) {}

    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    char peek() {
        ws();
        if (p >= e) jfail("EOF while parsing a value");
        return *p;
    }
    bool eat(char c) {
        ws();
        if (p < e && *p == c) { p++; return true; }
        return false;
    }
    void literal(const char *word) {
        const size_t n = strlen(word);
        if ((size_t)(e - p) < n || strncmp(p, word, n) != 0) jfail("expected ident");
        p += n;
    }
    bool null() {  // consumes `null` if that is the next value
        if (peek() != 'n') return false;
        literal("null");
        return true;
    }

    static void put_utf8(std::string &s, uint32_t cp) {
        if (cp < 0x80) s.push_back((char)cp);
        else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 63))); }
        else if (cp < 0x10000) { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
        else { s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 63))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
    }
    uint32_t hex4() {
        if (e - p < 4) jfail("EOF while parsing a string");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else jfail("invalid escape");
        }
        return v;
    }
    // A string whose opening quote is at *p.  out == nullptr: the value is being skipped (same grammar checks,
    // nothing kept).  Kept strings must be valid UTF-8.
    void string(std::string *out) {
        if (peek() != '"') jfail("invalid type: expected a string");
        p++;
        if (out) out->clear();
        while (true) {
            const char *run = p;
            while (p < e && *p != '"' && *p != '\\' && (unsigned char)*p >= 0x20) p++;
            if (out) out->append(run, p);
            if (p >= e) jfail("EOF while parsing a string");
            const char c = *p++;
            if (c == '"') break;
            if (c != '\\') jfail("control character (\\u0000-\\u001F) found while parsing a string");
            if (p >= e) jfail("EOF while parsing a string");
            const char x = *p++;
            uint32_t cp;
            switch (x) {
            case '"': cp = '"'; break;
            case '\\': cp = '\\'; break;
            case '/': cp = '/'; break;
            case 'b': cp = '\b'; break;
            case 'f': cp = '\f'; break;
            case 'n': cp = '\n'; break;
            case 'r': cp = '\r'; break;
            case 't': cp = '\t'; break;
            case 'u':
                cp = hex4();
                if (cp >= 0xDC00 && cp <= 0xDFFF) jfail("lone leading surrogate in hex escape");
                if (cp >= 0xD800 && cp <= 0xDBFF) {
                    if (e - p < 2 || p[0] != '\\' || p[1] != 'u') jfail("unexpected end of hex escape");
                    p += 2;
                    const uint32_t lo = hex4();
                    if (lo < 0xDC00 || lo > 0xDFFF) jfail("lone leading surrogate in hex escape");
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                break;
            default: jfail("invalid escape");
            }
            if (out) put_utf8(*out, cp);
        }
        if (out && !utf8_ok(*out)) jfail("invalid unicode code point");
    }

    // Number token at *p, checked against -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?.  Returns its extent;
    // *plain_uint says it was digits only (no sign, fraction or exponent).
    const char *number(bool *plain_uint) {
        const char *s0 = p;
        bool plain = true;
        if (p < e && *p == '-') { plain = false; p++; }
        if (p >= e || *p < '0' || *p > '9') jfail("invalid number");
        if (*p == '0') {
            p++;
            if (p < e && *p >= '0' && *p <= '9') jfail("invalid number");
        } else {
            while (p < e && *p >= '0' && *p <= '9') p++;
        }
        if (p < e && *p == '.') {
            plain = false;
            p++;
            if (p >= e || *p < '0' || *p > '9') jfail("invalid number");
            while (p < e && *p >= '0' && *p <= '9') p++;
        }
        if (p < e && (*p == 'e' || *p == 'E')) {
            plain = false;
            p++;
            if (p < e && (*p == '+' || *p == '-')) p++;
            if (p >= e || *p < '0' || *p > '9') jfail("invalid number");
            while (p < e && *p >= '0' && *p <= '9') p++;
        }
        *plain_uint = plain;
        return s0;
    }
    // an unsigned integer field of at most `maxv` (u32 / u64 visitors: anything negative, fractional, with an
    // exponent, or out of range is an error)
    uint64_t uint_field(uint64_t maxv, const char *what) {
        const char c = peek();
        if (c != '-' && (c < '0' || c > '9')) jfail(std::string("invalid type for ") + what + ": expected an unsigned integer");
        bool plain;
        const char *s0 = number(&plain);
        if (!plain) jfail(std::string("invalid type or value for ") + what + ": expected an unsigned integer");
        uint64_t u = 0;
        for (const char *q = s0; q < p; q++) {
            const uint64_t d = (uint64_t)(*q - '0');
            if (u > (~0ull - d) / 10) jfail(std::string("invalid value for ") + what + ": integer out of range");
            u = u * 10 + d;
        }
        if (u > maxv) jfail(std::string("invalid value for ") + what + ": integer out of range");
        return u;
    }
    double f64_field(const char *what) {
        const char c = peek();
        if (c != '-' && (c < '0' || c > '9')) jfail(std::string("invalid type for ") + what + ": expected a number");
        bool plain;
        const char *s0 = number(&plain);
        const std::string tok(s0, p);
        const double d = strtod(tok.c_str(), nullptr);
        if (!std::isfinite(d)) jfail("number out of range");
        return d;
    }
    // Vec<u64>
    void u64_array(std::vector<uint64_t> &out, const char *what) {
        if (peek() != '[') jfail(std::string("invalid type for ") + what + ": expected a sequence");
        p++;
        out.clear();
        if (eat(']')) return;
        while (true) {
            ws();
            // fast path: a run of digits not starting with 0
            if (p < e && *p >= '1' && *p <= '9') {
                const char *s0 = p;
                uint64_t u = 0;
                while (p < e && *p >= '0' && *p <= '9') { u = u * 10 + (uint64_t)(*p - '0'); p++; }
                const char nx = p < e ? *p : '\0';
                if (p - s0 <= 19 && nx != '.' && nx != 'e' && nx != 'E') {
                    out.push_back(u);
                } else {
                    p = s0;
                    out.push_back(uint_field(~0ull, what));
                }
            } else {
                out.push_back(uint_field(~0ull, what));
            }
            if (eat(',')) continue;
            if (eat(']')) break;
            if (p >= e) jfail("EOF while parsing a list");
            jfail("expected `,` or `]`");
        }
    }

    // any value, discarded (unknown fields).  Iterative: nesting depth costs heap, not stack.
    void skip_value() {
        std::vector<char> open;  // '[' or '{' per enclosing container
        while (true) {
            const char c = peek();
            bool container = false;
            if (c == '"') string(nullptr);
            else if (c == 't') literal("true");
            else if (c == 'f') literal("false");
            else if (c == 'n') literal("null");
            else if (c == '-' || (c >= '0' && c <= '9')) { bool plain; number(&plain); }
            else if (c == '[') {
                p++;
                if (!eat(']')) { open.push_back('['); container = true; }
            } else if (c == '{') {
                p++;
                if (!eat('}')) { open.push_back('{'); container = true; }
            } else jfail("expected value");
            if (container) {
                if (open.back() == '{') skip_key();
                continue;
            }
            // a value just ended: close / continue the enclosing containers
            while (true) {
                if (open.empty()) return;
                if (eat(',')) {
                    if (open.back() == '{') skip_key();
                    break;
                }
                if (eat(open.back() == '[' ? ']' : '}')) { open.pop_back(); continue; }
                if (p >= e) jfail(open.back() == '[' ? "EOF while parsing a list" : "EOF while parsing an object");
                jfail(open.back() == '[' ? "expected `,` or `]`" : "expected `,` or `}`");
            }
        }
    }
    void skip_key() {
        if (peek() != '"') jfail("key must be a string");
        std::string k;
        string(&k);
        if (!eat(':')) jfail("expected `:`");
    }

    // Drive a struct visitor over the map form {"name": value, ...} or the sequence form [value, ...].
    // names[0..n): the struct's fields in declaration order; field(i) must consume field i's value.
    template <typename F>
    void read_struct(const char *const *names, int n, uint32_t *seen, const char *what, F field) {
        *seen = 0;
        const char c = peek();
        if (c == '{') {
            p++;
            if (eat('}')) return;
            std::string key;
            while (true) {
                if (peek() != '"') jfail("key must be a string");
                string(&key);
                if (!eat(':')) jfail("expected `:`");
                int idx = -1;
                for (int i = 0; i < n; i++)
                    if (key == names[i]) { idx = i; break; }
                if (idx < 0) {
                    skip_value();
                } else {
                    if (*seen & (1u << idx)) jfail(std::string("duplicate field `") + names[idx] + "`");
                    *seen |= 1u << idx;
                    field(idx);
                }
                if (eat(',')) continue;
                if (eat('}')) break;
                if (p >= e) jfail("EOF while parsing an object");
                jfail("expected `,` or `}`");
            }
        } else if (c == '[') {
            p++;
            int idx = 0;
            if (!eat(']')) {
                while (true) {
                    if (idx >= n) jfail("trailing characters");  // more elements than fields
                    *seen |= 1u << idx;
                    field(idx++);
                    if (eat(',')) continue;
                    if (eat(']')) break;
                    if (p >= e) jfail("EOF while parsing a list");
                    jfail("expected `,` or `]`");
                }
            }
            // a sequence is positional: what it does not reach is filled from #[serde(default)] or fails, in order
            *seen |= 0x80000000u;
        } else {
            jfail(std::string("invalid type: expected struct ") + what);
        }
    }
};

struct SketchFields {  // TempSig, lib.rs:109-119
    uint32_t num = 0, ksize = 0;
    uint64_t seed = 0, max_hash = 0;
    std::string md5sum, molecule;
    std::vector<uint64_t> mins, abunds;
    bool has_abunds = false;
};
}  // namespace

// KmerMinHash Deserialize, lib.rs:104-139
static std::unique_ptr<KmerMinHash> read_minhash(JReader &r, SketchFields &f) {
    static const char *const names[8] = {"num", "ksize", "seed", "max_hash", "md5sum", "mins", "abundances", "molecule"};
    f.has_abunds = false;
    uint32_t seen;
    r.read_struct(names, 8, &seen, "TempSig", [&](int i) {
        switch (i) {
        case 0: f.num = (uint32_t)r.uint_field(0xFFFFFFFFull, "`num`"); break;
        case 1: f.ksize = (uint32_t)r.uint_field(0xFFFFFFFFull, "`ksize`"); break;
        case 2: f.seed = r.uint_field(~0ull, "`seed`"); break;
        case 3: f.max_hash = r.uint_field(~0ull, "`max_hash`"); break;
        case 4: r.string(&f.md5sum); break;
        case 5: r.u64_array(f.mins, "`mins`"); break;
        case 6:
            f.has_abunds = !r.null();
            if (f.has_abunds) r.u64_array(f.abunds, "`abundances`");
            break;
        default: r.string(&f.molecule); break;
        }
    });
    const bool positional = (seen & 0x80000000u) != 0;
    for (int i = 0; i < 8; i++) {
        if (seen & (1u << i)) continue;
        if (i == 6 && !positional) continue;  // Option<_>: absent from a map means None
        if (positional) jfail("invalid length " + std::to_string(i) + ", expected struct TempSig with 8 elements");
        jfail(std::string("missing field `") + names[i] + "`");
    }
    const uint32_t num = f.max_hash != 0 ? 0 : f.num;  // lib.rs:124
    const bool prot = f.molecule == "protein";          // lib.rs:132-136 (anything else: DNA)
    std::unique_ptr<KmerMinHash> mh(new KmerMinHash(num, f.ksize, prot, f.seed, f.max_hash, f.has_abunds));
    mh->set_from_host(f.mins.data(), f.mins.size(), f.has_abunds ? f.abunds.data() : nullptr, f.abunds.size());
    return mh;
}

static bool ieq(const char *a, const char *b) {
    for (; *a && *b; a++, b++) {
        char x = *a, y = *b;
        if (x >= 'A' && x <= 'Z') x = (char)(x + 32);
        if (y >= 'A' && y <= 'Z') y = (char)(y + 32);
        if (x != y) return false;
    }
    return *a == *b;
}

namespace {
typedef std::vector<std::unique_ptr<Signature>> SigList;
struct LoadFilter {
    size_t ksize;         // 0: any
    const char *moltype;  // nullptr: any
};

// One element of the top-level array: a Signature, flattened to one Signature per sketch and filtered
// (lib.rs:603-644), appended to `out`.
void read_signature(JReader &r, SketchFields &scratch, const LoadFilter &flt, SigList &out) {
    static const char *const names[8] = {"class", "email", "hash_function", "filename", "name", "license", "signatures", "version"};
    Signature meta;
    std::vector<std::unique_ptr<KmerMinHash>> sketches;
    uint32_t seen;
    r.read_struct(names, 8, &seen, "Signature", [&](int i) {
        switch (i) {
        case 0: r.string(&meta.class_); break;
        case 1: r.string(&meta.email); break;
        case 2: r.string(&meta.hash_function); break;
        case 3:
            meta.has_filename = !r.null();
            if (meta.has_filename) r.string(&meta.filename);
            break;
        case 4:
            meta.has_name = !r.null();
            if (meta.has_name) r.string(&meta.name);
            break;
        case 5: r.string(&meta.license); break;
        case 6:
            if (r.peek() != '[') jfail("invalid type for `signatures`: expected a sequence");
            r.p++;
            if (!r.eat(']')) {
                while (true) {
                    sketches.push_back(read_minhash(r, scratch));
                    if (r.eat(',')) continue;
                    if (r.eat(']')) break;
                    if (r.p >= r.e) jfail("EOF while parsing a list");
                    jfail("expected `,` or `]`");
                }
            }
            break;
        default: meta.version = r.f64_field("`version`"); break;
        }
    });
    const bool positional = (seen & 0x80000000u) != 0;
    for (int i = 0; i < 8; i++) {
        if (seen & (1u << i)) continue;
        const bool has_default = i == 0 || i == 1 || i == 5 || i == 7;       // #[serde(default ...)], lib.rs:548-564
        if (has_default) continue;                                           // Signature() already holds them
        if ((i == 3 || i == 4) && !positional) continue;                     // Option<_>: absent from a map means None
        if (positional) jfail("invalid length " + std::to_string(i) + ", expected struct Signature with 8 elements");
        jfail(std::string("missing field `") + names[i] + "`");
    }
    for (auto &mh : sketches) {
        bool keep = false;
        if (flt.ksize == 0 || flt.ksize == (size_t)mh->ksize) {
            if (!flt.moltype) keep = true;
            else if (ieq(flt.moltype, "dna") && !mh->is_protein) keep = true;
            else if (ieq(flt.moltype, "protein") && mh->is_protein) keep = true;
        }
        if (!keep) continue;
        std::unique_ptr<Signature> s(meta.clone_meta());
        s->signatures.push_back(std::move(mh));
        out.push_back(std::move(s));
    }
}

SigList load_sequential(const char *data, size_t len, const LoadFilter &flt) {
    JReader r(data, len);
    SigList out;
    SketchFields scratch;
    if (r.peek() != '[') jfail("invalid type: expected a sequence of signatures");
    r.p++;
    if (!r.eat(']')) {
        while (true) {
            read_signature(r, scratch, flt, out);
            if (r.eat(',')) continue;
            if (r.eat(']')) break;
            if (r.p >= r.e) jfail("EOF while parsing a list");
            jfail("expected `,` or `]`");
        }
    }
    r.ws();
    if (r.p != r.e) jfail("trailing characters");
    return out;
}

// Large files: the top-level array is cut into pieces that are parsed on several host threads.  The cuts are
// guesses -- a `{` that follows `} ,` and is followed by the key "class" (first in the output of both the Rust
// and the Python writer) -- and every guess is verified: piece t counts only if piece t-1, which by induction
// started on a true element boundary, ends exactly on piece t's first byte.  Any failed guess, and any error
// inside any piece, discards the attempt; the caller then parses sequentially, so the result and the error
// reported are always those of the single-pass reader.
bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }
const char *guess_element_start(const char *data, const char *from, const char *to) {
    static const char key[] = "\"class\"";
    for (const char *p = from; p < to; p++) {
        if (*p != '{') continue;
        const char *b = p - 1;
        while (b > data && is_ws(*b)) b--;
        if (b <= data || *b != ',') continue;
        b--;
        while (b > data && is_ws(*b)) b--;
        if (*b != '}') continue;
        const char *f = p + 1;
        while (f < to && is_ws(*f)) f++;
        if ((size_t)(to - f) >= sizeof(key) - 1 && memcmp(f, key, sizeof(key) - 1) == 0) return p;
    }
    return nullptr;
}

bool load_parallel(const char *data, size_t len, const LoadFilter &flt, SigList &out) {
    const size_t piece_min = size_t(1) << 20;
    unsigned T = (unsigned)std::min<size_t>(json_threads(), len / piece_min);
    if (T < 2) return false;
    const char *end = data + len;
    JReader head(data, len);
    head.ws();
    if (head.p >= end || *head.p != '[') return false;
    head.p++;
    head.ws();
    if (head.p >= end || *head.p != '{') return false;
    std::vector<const char *> start{head.p};
    for (unsigned t = 1; t < T; t++) {
        const char *from = std::max(data + len / T * t, start.back() + 1);
        const char *g = guess_element_start(data, from, std::min(end, from + len / T));
        if (g) start.push_back(g);
    }
    T = (unsigned)start.size();
    if (T < 2) return false;
    std::vector<SigList> parts(T);
    std::vector<char> ok(T, 0);
    auto piece = [&](unsigned t) {
        try {
            JReader r(start[t], (size_t)(end - start[t]));
            const char *stop = t + 1 < T ? start[t + 1] : nullptr;
            SketchFields scratch;
            while (true) {
                read_signature(r, scratch, flt, parts[t]);
                if (stop) {
                    if (!r.eat(',')) return;
                    r.ws();
                    if (r.p == stop) break;
                    if (r.p > stop) return;
                } else {
                    if (r.eat(',')) continue;
                    if (!r.eat(']')) return;
                    r.ws();
                    if (r.p != r.e) return;
                    break;
                }
            }
            ok[t] = 1;
        } catch (...) {
        }
    };
    run_pieces(T, piece);
    for (unsigned t = 0; t < T; t++)
        if (!ok[t]) return false;
    size_t total = 0;
    for (auto &p : parts) total += p.size();
    out.reserve(total);
    for (auto &p : parts)
        for (auto &s : p) out.push_back(std::move(s));
    return true;
}
}  // namespace

std::vector<std::unique_ptr<Signature>> load_signatures(const char *data, size_t len, size_t ksize, const char *moltype) {
    const LoadFilter flt{ksize, moltype};
    static const bool trace = getenv("SMB200_JSON_TRACE") != nullptr;  // which reader ran, for the tests
    SigList out;
    if (load_parallel(data, len, flt, out)) {
        if (trace) fprintf(stderr, "smb200 json: %zu bytes read in pieces\n", len);
        return out;
    }
    if (trace) fprintf(stderr, "smb200 json: %zu bytes read in one pass\n", len);
    return load_sequential(data, len, flt);
}

// First gzip member of `data`, inflated (what flate2::read::GzDecoder yields; src/file.rs:62-66).
static std::string gunzip_first_member(const std::string &data) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) throw_internal("zlib: inflateInit2 failed");
    std::string out;
    std::vector<unsigned char> chunk(1 << 16);
    size_t consumed = 0;
    int rc = Z_OK;
    while (rc != Z_STREAM_END) {
        if (zs.avail_in == 0) {
            const size_t left = data.size() - consumed;
            if (left == 0) break;
            const size_t take = left < (size_t(1) << 30) ? left : (size_t(1) << 30);
            zs.next_in = (Bytef *)(data.data() + consumed);
            zs.avail_in = (uInt)take;
            consumed += take;
        }
        zs.next_out = chunk.data();
        zs.avail_out = (uInt)chunk.size();
        rc = inflate(&zs, Z_NO_FLUSH);
        if (rc != Z_OK && rc != Z_STREAM_END && rc != Z_BUF_ERROR) {
            inflateEnd(&zs);
            throw SourmashError(ERR_UNKNOWN, "corrupt gzip stream");
        }
        out.append((const char *)chunk.data(), chunk.size() - zs.avail_out);
        if (rc == Z_BUF_ERROR && zs.avail_in == 0 && consumed == data.size()) break;
    }
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) throw SourmashError(ERR_UNKNOWN, "unexpected end of gzip stream");
    return out;
}

// Signature::from_path by way of file.rs:47-77: "-" is stdin, otherwise the file; the first bytes pick the
// decoder (0x1F8B gzip, 0x425A bzip2, 0xFD377A585A xz).  gzip is inflated here; this image carries no bzip2 or
// xz development files, so those two report an error instead of being decoded.
std::vector<std::unique_ptr<Signature>> load_signatures_path(const char *path, size_t ksize, const char *moltype) {
    std::stringstream ss;
    if (strcmp(path, "-") == 0) {
        ss << std::cin.rdbuf();
    } else {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw SourmashError(ERR_PANIC, std::string("Can't open input file ") + path);  // file.rs:83 expect()
        ss << f.rdbuf();
    }
    std::string data = ss.str();
    auto starts = [&](std::initializer_list<unsigned char> magic) {
        if (data.size() < magic.size()) return false;
        size_t i = 0;
        for (unsigned char m : magic)
            if ((unsigned char)data[i++] != m) return false;
        return true;
    };
    if (starts({0xFD, 0x37, 0x7A, 0x58, 0x5A})) throw_internal("xz-compressed signature files are not supported by this build");
    if (starts({0x42, 0x5A})) throw_internal("bzip2-compressed signature files are not supported by this build");
    if (starts({0x1F, 0x8B})) data = gunzip_first_member(data);
    return load_signatures(data.data(), data.size(), ksize, moltype);
}

}  // namespace smb200
