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
        if (fill_ == 0)
            for (; n >= 64; data += 64, n -= 64) block(data);
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
    // the 64 steps of RFC 1321 section 3.4, written out so that the register roles rotate at compile time
#define SMB_MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define SMB_MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define SMB_MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define SMB_MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define SMB_MD5_STEP(f, a, b, c, d, x, t, s) (a) = (b) + rol((a) + SMB_MD5_##f((b), (c), (d)) + (x) + (t), (s))
    void block(const uint8_t *p) {
        uint32_t w[16];
        std::memcpy(w, p, 64);  // little-endian host (x86-64 / aarch64 Linux)
        uint32_t a = st_[0], b = st_[1], c = st_[2], d = st_[3];
        SMB_MD5_STEP(F, a, b, c, d, w[0], 0xd76aa478u, 7);
        SMB_MD5_STEP(F, d, a, b, c, w[1], 0xe8c7b756u, 12);
        SMB_MD5_STEP(F, c, d, a, b, w[2], 0x242070dbu, 17);
        SMB_MD5_STEP(F, b, c, d, a, w[3], 0xc1bdceeeu, 22);
        SMB_MD5_STEP(F, a, b, c, d, w[4], 0xf57c0fafu, 7);
        SMB_MD5_STEP(F, d, a, b, c, w[5], 0x4787c62au, 12);
        SMB_MD5_STEP(F, c, d, a, b, w[6], 0xa8304613u, 17);
        SMB_MD5_STEP(F, b, c, d, a, w[7], 0xfd469501u, 22);
        SMB_MD5_STEP(F, a, b, c, d, w[8], 0x698098d8u, 7);
        SMB_MD5_STEP(F, d, a, b, c, w[9], 0x8b44f7afu, 12);
        SMB_MD5_STEP(F, c, d, a, b, w[10], 0xffff5bb1u, 17);
        SMB_MD5_STEP(F, b, c, d, a, w[11], 0x895cd7beu, 22);
        SMB_MD5_STEP(F, a, b, c, d, w[12], 0x6b901122u, 7);
        SMB_MD5_STEP(F, d, a, b, c, w[13], 0xfd987193u, 12);
        SMB_MD5_STEP(F, c, d, a, b, w[14], 0xa679438eu, 17);
        SMB_MD5_STEP(F, b, c, d, a, w[15], 0x49b40821u, 22);
        SMB_MD5_STEP(G, a, b, c, d, w[1], 0xf61e2562u, 5);
        SMB_MD5_STEP(G, d, a, b, c, w[6], 0xc040b340u, 9);
        SMB_MD5_STEP(G, c, d, a, b, w[11], 0x265e5a51u, 14);
        SMB_MD5_STEP(G, b, c, d, a, w[0], 0xe9b6c7aau, 20);
        SMB_MD5_STEP(G, a, b, c, d, w[5], 0xd62f105du, 5);
        SMB_MD5_STEP(G, d, a, b, c, w[10], 0x02441453u, 9);
        SMB_MD5_STEP(G, c, d, a, b, w[15], 0xd8a1e681u, 14);
        SMB_MD5_STEP(G, b, c, d, a, w[4], 0xe7d3fbc8u, 20);
        SMB_MD5_STEP(G, a, b, c, d, w[9], 0x21e1cde6u, 5);
        SMB_MD5_STEP(G, d, a, b, c, w[14], 0xc33707d6u, 9);
        SMB_MD5_STEP(G, c, d, a, b, w[3], 0xf4d50d87u, 14);
        SMB_MD5_STEP(G, b, c, d, a, w[8], 0x455a14edu, 20);
        SMB_MD5_STEP(G, a, b, c, d, w[13], 0xa9e3e905u, 5);
        SMB_MD5_STEP(G, d, a, b, c, w[2], 0xfcefa3f8u, 9);
        SMB_MD5_STEP(G, c, d, a, b, w[7], 0x676f02d9u, 14);
        SMB_MD5_STEP(G, b, c, d, a, w[12], 0x8d2a4c8au, 20);
        SMB_MD5_STEP(H, a, b, c, d, w[5], 0xfffa3942u, 4);
        SMB_MD5_STEP(H, d, a, b, c, w[8], 0x8771f681u, 11);
        SMB_MD5_STEP(H, c, d, a, b, w[11], 0x6d9d6122u, 16);
        SMB_MD5_STEP(H, b, c, d, a, w[14], 0xfde5380cu, 23);
        SMB_MD5_STEP(H, a, b, c, d, w[1], 0xa4beea44u, 4);
        SMB_MD5_STEP(H, d, a, b, c, w[4], 0x4bdecfa9u, 11);
        SMB_MD5_STEP(H, c, d, a, b, w[7], 0xf6bb4b60u, 16);
        SMB_MD5_STEP(H, b, c, d, a, w[10], 0xbebfbc70u, 23);
        SMB_MD5_STEP(H, a, b, c, d, w[13], 0x289b7ec6u, 4);
        SMB_MD5_STEP(H, d, a, b, c, w[0], 0xeaa127fau, 11);
        SMB_MD5_STEP(H, c, d, a, b, w[3], 0xd4ef3085u, 16);
        SMB_MD5_STEP(H, b, c, d, a, w[6], 0x04881d05u, 23);
        SMB_MD5_STEP(H, a, b, c, d, w[9], 0xd9d4d039u, 4);
        SMB_MD5_STEP(H, d, a, b, c, w[12], 0xe6db99e5u, 11);
        SMB_MD5_STEP(H, c, d, a, b, w[15], 0x1fa27cf8u, 16);
        SMB_MD5_STEP(H, b, c, d, a, w[2], 0xc4ac5665u, 23);
        SMB_MD5_STEP(I, a, b, c, d, w[0], 0xf4292244u, 6);
        SMB_MD5_STEP(I, d, a, b, c, w[7], 0x432aff97u, 10);
        SMB_MD5_STEP(I, c, d, a, b, w[14], 0xab9423a7u, 15);
        SMB_MD5_STEP(I, b, c, d, a, w[5], 0xfc93a039u, 21);
        SMB_MD5_STEP(I, a, b, c, d, w[12], 0x655b59c3u, 6);
        SMB_MD5_STEP(I, d, a, b, c, w[3], 0x8f0ccc92u, 10);
        SMB_MD5_STEP(I, c, d, a, b, w[10], 0xffeff47du, 15);
        SMB_MD5_STEP(I, b, c, d, a, w[1], 0x85845dd1u, 21);
        SMB_MD5_STEP(I, a, b, c, d, w[8], 0x6fa87e4fu, 6);
        SMB_MD5_STEP(I, d, a, b, c, w[15], 0xfe2ce6e0u, 10);
        SMB_MD5_STEP(I, c, d, a, b, w[6], 0xa3014314u, 15);
        SMB_MD5_STEP(I, b, c, d, a, w[13], 0x4e0811a1u, 21);
        SMB_MD5_STEP(I, a, b, c, d, w[4], 0xf7537e82u, 6);
        SMB_MD5_STEP(I, d, a, b, c, w[11], 0xbd3af235u, 10);
        SMB_MD5_STEP(I, c, d, a, b, w[2], 0x2ad7d2bbu, 15);
        SMB_MD5_STEP(I, b, c, d, a, w[9], 0xeb86d391u, 21);
        st_[0] += a; st_[1] += b; st_[2] += c; st_[3] += d;
    }
#undef SMB_MD5_STEP
#undef SMB_MD5_I
#undef SMB_MD5_H
#undef SMB_MD5_G
#undef SMB_MD5_F
};

}  // namespace smb200
