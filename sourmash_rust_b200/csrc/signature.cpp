// signature.cpp -- Signature JSON output (byte-compatible with serde_json's compact writer for
// the field order of src/lib.rs:79-100 and 546-565) and input (Signature::load_signatures,
// src/lib.rs:593-645).
#include "signature.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

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
static void put_u64(std::string &o, uint64_t v) {
    char buf[24];
    int n = 0;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) o.push_back(buf[--n]);
}
static void put_u64_array(std::string &o, const std::vector<uint64_t> &v) {
    o.push_back('[');
    for (size_t i = 0; i < v.size(); i++) {
        if (i) o.push_back(',');
        put_u64(o, v[i]);
    }
    o.push_back(']');
}
// shortest decimal that round-trips, in the shape serde_json (ryu) prints: "0.4", "1.0", "1e21"
static void put_f64(std::string &o, double x) {
    if (!std::isfinite(x)) { o += "null"; return; }
    char buf[40];
    for (int prec = 1; prec <= 17; prec++) {
        snprintf(buf, sizeof buf, "%.*g", prec, x);
        if (strtod(buf, nullptr) == x) break;
    }
    std::string s(buf);
    const size_t e = s.find('e');
    std::string mant = (e == std::string::npos) ? s : s.substr(0, e);
    if (e == std::string::npos) {
        if (mant.find('.') == std::string::npos) mant += ".0";
        o += mant;
    } else {
        int ex = atoi(s.c_str() + e + 1);
        o += mant;
        o.push_back('e');
        o += std::to_string(ex);
    }
}
static void put_minhash(std::string &o, KmerMinHash &mh) {  // lib.rs:62-102
    o += "{\"num\":"; put_u64(o, mh.num);
    o += ",\"ksize\":"; put_u64(o, mh.ksize);
    o += ",\"seed\":"; put_u64(o, mh.seed);
    o += ",\"max_hash\":"; put_u64(o, mh.max_hash);
    o += ",\"mins\":"; put_u64_array(o, mh.mins());
    o += ",\"md5sum\":"; put_str(o, mh.md5sum());
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

std::string signatures_to_json(Signature *const *sigs, size_t n) {
    std::string o = "[";
    for (size_t i = 0; i < n; i++) {
        if (i) o.push_back(',');
        if (!sigs[i]) throw SourmashError(ERR_PANIC, "sourmash panicked: null signature pointer");
        sigs[i]->to_json(o);
    }
    o.push_back(']');
    return o;
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
// reader: a small recursive-descent JSON parser producing a tree of values
// ---------------------------------------------------------------------------------------------
namespace {
struct JVal {
    enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
    bool b = false;
    bool is_u64 = false;
    uint64_t u = 0;
    double d = 0;
    std::string s;
    std::vector<JVal> a;
    std::vector<std::pair<std::string, JVal>> o;
    const JVal *get(const char *k) const {
        for (auto &kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
[[noreturn]] void jfail(const std::string &m) { throw SourmashError(ERR_UNKNOWN, "JSON error: " + m); }
struct JParser {
    const char *p, *e;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    static void utf8(std::string &s, uint32_t cp) {
        if (cp < 0x80) s.push_back((char)cp);
        else if (cp < 0x800) { s.push_back((char)(0xC0 | (cp >> 6))); s.push_back((char)(0x80 | (cp & 63))); }
        else if (cp < 0x10000) { s.push_back((char)(0xE0 | (cp >> 12))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
        else { s.push_back((char)(0xF0 | (cp >> 18))); s.push_back((char)(0x80 | ((cp >> 12) & 63))); s.push_back((char)(0x80 | ((cp >> 6) & 63))); s.push_back((char)(0x80 | (cp & 63))); }
    }
    uint32_t hex4() {
        if (e - p < 4) jfail("truncated \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else jfail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        std::string s;
        p++;  // opening quote
        while (true) {
            if (p >= e) jfail("unterminated string");
            const char c = *p++;
            if (c == '"') break;
            if (c != '\\') { s.push_back(c); continue; }
            if (p >= e) jfail("unterminated escape");
            const char x = *p++;
            switch (x) {
            case '"': s.push_back('"'); break;
            case '\\': s.push_back('\\'); break;
            case '/': s.push_back('/'); break;
            case 'b': s.push_back('\b'); break;
            case 'f': s.push_back('\f'); break;
            case 'n': s.push_back('\n'); break;
            case 'r': s.push_back('\r'); break;
            case 't': s.push_back('\t'); break;
            case 'u': {
                uint32_t cp = hex4();
                if (cp >= 0xD800 && cp < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                    p += 2;
                    const uint32_t lo = hex4();
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                }
                utf8(s, cp);
                break;
            }
            default: jfail("bad escape");
            }
        }
        return s;
    }
    JVal val() {
        ws();
        if (p >= e) jfail("unexpected end of input");
        JVal v;
        const char c = *p;
        if (c == '{') {
            v.t = JVal::OBJ; p++; ws();
            if (p < e && *p == '}') { p++; return v; }
            while (true) {
                ws();
                if (p >= e || *p != '"') jfail("expected object key");
                std::string k = str();
                ws();
                if (p >= e || *p != ':') jfail("expected ':'");
                p++;
                v.o.emplace_back(std::move(k), val());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; break; }
                jfail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.t = JVal::ARR; p++; ws();
            if (p < e && *p == ']') { p++; return v; }
            while (true) {
                v.a.push_back(val());
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; break; }
                jfail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.t = JVal::STR; v.s = str();
        } else if (c == 't' && e - p >= 4 && !strncmp(p, "true", 4)) { v.t = JVal::BOOL; v.b = true; p += 4; }
        else if (c == 'f' && e - p >= 5 && !strncmp(p, "false", 5)) { v.t = JVal::BOOL; v.b = false; p += 5; }
        else if (c == 'n' && e - p >= 4 && !strncmp(p, "null", 4)) { v.t = JVal::NUL; p += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const char *s0 = p;
            if (*p == '-') p++;
            bool integral = true;
            while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) {
                if (*p == '.' || *p == 'e' || *p == 'E') integral = false;
                p++;
            }
            const std::string tok(s0, p);
            v.t = JVal::NUM;
            v.d = strtod(tok.c_str(), nullptr);
            if (integral && tok[0] != '-' && tok.size() <= 20) {
                uint64_t u = 0;
                bool ok = true;
                for (char ch : tok) {
                    const uint64_t nu = u * 10 + (uint64_t)(ch - '0');
                    if (u > (~0ull) / 10 || nu < u * 10) { ok = false; break; }
                    u = nu;
                }
                v.is_u64 = ok;
                v.u = u;
            }
        } else jfail(std::string("unexpected character '") + c + "'");
        return v;
    }
};

uint64_t need_u64(const JVal &o, const char *k, uint64_t maxv) {
    const JVal *v = o.get(k);
    if (!v) jfail(std::string("missing field `") + k + "`");
    if (v->t != JVal::NUM || !v->is_u64 || v->u > maxv) jfail(std::string("invalid value for `") + k + "`");
    return v->u;
}
std::vector<uint64_t> u64_array(const JVal &v, const char *k) {
    if (v.t != JVal::ARR) jfail(std::string("`") + k + "` is not an array");
    std::vector<uint64_t> out;
    out.reserve(v.a.size());
    for (auto &x : v.a) {
        if (x.t != JVal::NUM || !x.is_u64) jfail(std::string("non-u64 entry in `") + k + "`");
        out.push_back(x.u);
    }
    return out;
}
std::string opt_str(const JVal &o, const char *k, const std::string &dflt, bool *present = nullptr) {
    const JVal *v = o.get(k);
    if (present) *present = false;
    if (!v || v->t == JVal::NUL) return dflt;
    if (v->t != JVal::STR) jfail(std::string("`") + k + "` is not a string");
    if (present) *present = true;
    return v->s;
}
}  // namespace

// KmerMinHash Deserialize, lib.rs:104-139
static std::unique_ptr<KmerMinHash> minhash_from_json(const JVal &o) {
    if (o.t != JVal::OBJ) jfail("sketch is not an object");
    const uint32_t num_in = (uint32_t)need_u64(o, "num", 0xFFFFFFFFull);
    const uint32_t ksize = (uint32_t)need_u64(o, "ksize", 0xFFFFFFFFull);
    const uint64_t seed = need_u64(o, "seed", ~0ull);
    const uint64_t max_hash = need_u64(o, "max_hash", ~0ull);
    if (!o.get("md5sum") || o.get("md5sum")->t != JVal::STR) jfail("missing field `md5sum`");
    if (!o.get("mins")) jfail("missing field `mins`");
    if (!o.get("molecule") || o.get("molecule")->t != JVal::STR) jfail("missing field `molecule`");
    const std::vector<uint64_t> mins = u64_array(*o.get("mins"), "mins");
    const JVal *ab = o.get("abundances");
    const bool has_ab = ab && ab->t != JVal::NUL;
    std::vector<uint64_t> abunds;
    if (has_ab) abunds = u64_array(*ab, "abundances");
    const uint32_t num = max_hash != 0 ? 0 : num_in;            // lib.rs:124
    const bool prot = o.get("molecule")->s == "protein";        // lib.rs:132-136 (anything else: DNA)
    std::unique_ptr<KmerMinHash> mh(new KmerMinHash(num, ksize, prot, seed, max_hash, has_ab));
    mh->set_from_host(mins.data(), mins.size(), has_ab ? abunds.data() : nullptr, abunds.size());
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

std::vector<std::unique_ptr<Signature>> load_signatures(const char *data, size_t len, size_t ksize, const char *moltype) {
    JParser jp{data, data + len};
    const JVal root = jp.val();
    jp.ws();
    if (jp.p != jp.e) jfail("trailing characters");
    if (root.t != JVal::ARR) jfail("expected an array of signatures");
    std::vector<std::unique_ptr<Signature>> out;
    for (const JVal &js : root.a) {
        if (js.t != JVal::OBJ) jfail("signature is not an object");
        Signature meta;
        meta.class_ = opt_str(js, "class", "sourmash_signature");
        meta.email = opt_str(js, "email", "");
        if (!js.get("hash_function") || js.get("hash_function")->t != JVal::STR) jfail("missing field `hash_function`");
        meta.hash_function = js.get("hash_function")->s;
        meta.filename = opt_str(js, "filename", "", &meta.has_filename);
        meta.name = opt_str(js, "name", "", &meta.has_name);
        meta.license = opt_str(js, "license", "CC0");
        const JVal *ver = js.get("version");
        if (ver && ver->t == JVal::NUM) meta.version = ver->d;
        else if (ver && ver->t != JVal::NUL) jfail("`version` is not a number");
        const JVal *sk = js.get("signatures");
        if (!sk || sk->t != JVal::ARR) jfail("missing field `signatures`");
        // flatten: one Signature per sketch, then filter (lib.rs:603-644)
        for (const JVal &jm : sk->a) {
            std::unique_ptr<KmerMinHash> mh = minhash_from_json(jm);
            bool keep = false;
            if (ksize == 0 || ksize == (size_t)mh->ksize) {
                if (!moltype) keep = true;
                else if (ieq(moltype, "dna") && !mh->is_protein) keep = true;
                else if (ieq(moltype, "protein") && mh->is_protein) keep = true;
            }
            if (!keep) continue;
            std::unique_ptr<Signature> s(meta.clone_meta());
            s->signatures.push_back(std::move(mh));
            out.push_back(std::move(s));
        }
    }
    return out;
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
            throw SourmashError(ERR_SERDE, "corrupt gzip stream");
        }
        out.append((const char *)chunk.data(), chunk.size() - zs.avail_out);
        if (rc == Z_BUF_ERROR && zs.avail_in == 0 && consumed == data.size()) break;
    }
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) throw SourmashError(ERR_SERDE, "unexpected end of gzip stream");
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
