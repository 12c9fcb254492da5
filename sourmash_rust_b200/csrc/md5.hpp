// md5.hpp -- RFC 1321 message digest, used only for the `md5sum` field of the Signature JSON
// (reference src/lib.rs:72-77,86: md5 over ksize and every min rendered in decimal; the
// reference uses the third-party `md5` crate, Cargo.toml:48).
#pragma once
#include <stdint.h>

#include <cstring>
#include <string>

namespace smb200 {

class Md5 {
  public:
    Md5() {
        st_[0] = 0x67452301u; st_[1] = 0xefcdab89u; st_[2] = 0x98badcfeu; st_[3] = 0x10325476u;
    }
    void update(const std::string &s) { update(reinterpret_cast<const uint8_t *>(s.data()), s.size()); }
    void update(const uint8_t *data, size_t n) {
        total_ += n;
        while (n) {
            const size_t take = (64 - fill_ < n) ? 64 - fill_ : n;
            std::memcpy(buf_ + fill_, data, take);
            fill_ += take; data += take; n -= take;
            if (fill_ == 64) { block(buf_); fill_ = 0; }
        }
    }
    std::string hex() {
        const uint64_t bits = total_ * 8;
        static const uint8_t pad[64] = {0x80};
        const size_t padlen = (fill_ < 56) ? 56 - fill_ : 120 - fill_;
        update(pad, padlen);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (8 * i));
        update(len, 8);
        static const char *digits = "0123456789abcdef";
        std::string out(32, '0');
        for (int i = 0; i < 16; i++) {
            const uint8_t b = (uint8_t)(st_[i / 4] >> (8 * (i % 4)));
            out[2 * i] = digits[b >> 4];
            out[2 * i + 1] = digits[b & 15];
        }
        return out;
    }

  private:
    uint32_t st_[4];
    uint8_t buf_[64];
    size_t fill_ = 0;
    uint64_t total_ = 0;
    static uint32_t rol(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
    void block(const uint8_t *p) {
        static const uint32_t T[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9,
                                  14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
        uint32_t w[16];
        for (int i = 0; i < 16; i++)
            w[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
        uint32_t a = st_[0], b = st_[1], c = st_[2], d = st_[3];
        for (int i = 0; i < 64; i++) {
            uint32_t f;
            int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            const uint32_t tmp = d;
            d = c; c = b;
            b = b + rol(a + f + T[i] + w[g], S[i]);
            a = tmp;
        }
        st_[0] += a; st_[1] += b; st_[2] += c; st_[3] += d;
    }
};

}  // namespace smb200
