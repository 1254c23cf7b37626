// Blake2b-512 (RFC 7693), unkeyed — host side of setup_utils::calculate_hash
// (reference src/utils.rs:618-623: the 64-byte hash chained through challenge/response files).
// The hash is inherently sequential; it runs on the host, overlapped with the GPU work.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>

namespace sso {

struct Blake2b {
  uint64_t h[8];
  uint64_t t0 = 0, t1 = 0;
  uint8_t buf[128];
  size_t buflen = 0;

  static inline uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
  static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }   // little-endian hosts

  static const uint64_t* iv() {
    static const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                   0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    return IV;
  }

  explicit Blake2b(size_t outlen = 64) {
    for (int i = 0; i < 8; i++) h[i] = iv()[i];
    h[0] ^= 0x01010000ULL ^ (uint64_t)outlen;
  }

  void compress(const uint8_t* block, bool last) {
    static const uint8_t S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) m[i] = load64(block + 8 * i);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = iv()[i]; }
    v[12] ^= t0; v[13] ^= t1;
    if (last) v[14] = ~v[14];
#define SSO_B2G(a, b, c, d, x, y) \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
      const uint8_t* s = S[r];
      SSO_B2G(0, 4, 8, 12, m[s[0]], m[s[1]])  SSO_B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
      SSO_B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) SSO_B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
      SSO_B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) SSO_B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
      SSO_B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) SSO_B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef SSO_B2G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
  }

  void update(const uint8_t* in, size_t len) {
    while (len > 0) {
      if (buflen == 128) {                      // buffer full and more input follows: not the last block
        t0 += 128; if (t0 < 128) t1++;
        compress(buf, false);
        buflen = 0;
      }
      size_t take = 128 - buflen;
      if (take > len) take = len;
      memcpy(buf + buflen, in, take);
      buflen += take; in += take; len -= take;
    }
  }

  void final(uint8_t* out, size_t outlen = 64) {
    t0 += buflen; if (t0 < buflen) t1++;
    memset(buf + buflen, 0, 128 - buflen);
    compress(buf, true);
    uint8_t full[64];
    memcpy(full, h, 64);
    memcpy(out, full, outlen);
  }
};

inline void blake2b_512(const uint8_t* data, size_t len, uint8_t out[64]) {
  Blake2b b(64);
  b.update(data, len);
  b.final(out, 64);
}

}  // namespace sso
