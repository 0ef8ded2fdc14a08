// FLAC decoder (host side) behind the C ABI: aad_flac_info / aad_flac_decode (include/aad.h).
//
// The reference reads its corpus -- ASVspoof 2019/2021 ships as 16-bit mono FLAC -- through libsndfile:
// soundfile.info for the chunk index (ASV_dl_func.py:280) and librosa.load for every extractor call
// (ASV_dl_func.py:406,425,524).  This is the native replacement on the input side of the path: the whole file is
// decoded once into integer PCM (int16 PCM goes to the GPU as it is).  Written from the format
// specification (RFC 9639): frame header with CRC-8, subframes CONSTANT / VERBATIM / FIXED / LPC with wasted bits,
// Rice / Rice2 partitioned residuals with escape partitions, the three stereo decorrelation modes, CRC-16 footer.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/aad.h"

namespace {

struct BitReader {
  // 64-bit left-aligned cache refilled a byte at a time: bits(k) and the unary code of the Rice residuals (a count of
  // leading zeros) work on whole words instead of walking bytes
  const uint8_t* p;
  size_t n, next = 0;   // next byte to load
  uint64_t cache = 0;   // the next `have` bits of the stream, most significant first; zero below them
  int have = 0;
  bool bad = false;
  BitReader(const uint8_t* d, size_t size) : p(d), n(size) {}
  inline void refill() {
    while (have <= 56 && next < n) {
      cache |= (uint64_t)p[next++] << (56 - have);
      have += 8;
    }
  }
  inline size_t bitpos() const { return next * 8 - (size_t)have; }  // bits consumed so far
  inline uint64_t bits(int k) {  // k <= 57
    if (k == 0) return 0;
    if (have < k) {
      refill();
      if (have < k) {
        bad = true;
        have = 0;
        cache = 0;
        return 0;
      }
    }
    const uint64_t v = cache >> (64 - k);
    cache <<= k;
    have -= k;
    return v;
  }
  inline uint32_t bit() { return (uint32_t)bits(1); }
  inline int64_t sbits(int k) {
    if (k == 0) return 0;
    uint64_t v = bits(k);
    const uint64_t sign = 1ull << (k - 1);
    return (int64_t)((v ^ sign) - sign);
  }
  inline uint32_t unary() {  // number of 0 bits before the next 1 bit
    uint32_t q = 0;
    for (;;) {
      if (cache != 0) {
        const int lead = __builtin_clzll(cache);  // < have: the cache is zero below its valid bits
        q += (uint32_t)lead;
        cache <<= lead;
        cache <<= 1;
        have -= lead + 1;
        return q;
      }
      q += (uint32_t)have;
      have = 0;
      refill();
      if (have == 0) {
        bad = true;
        return q;
      }
    }
  }
  inline void align() {
    const int drop = have & 7;  // next * 8 is byte-aligned, so the bits in front of the next boundary are have mod 8
    cache <<= drop;
    have -= drop;
  }
};


uint8_t crc8(const uint8_t* d, size_t n) {
  uint8_t c = 0;
  for (size_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int k = 0; k < 8; ++k) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1));
  }
  return c;
}
struct Crc16Table {
  // slicing-by-8: t[k][i] = CRC of byte i followed by k zero bytes.  The byte-at-a-time form is a chain of dependent
  // table look-ups (6 ns per byte on the test machine: three quarters of the decode time of a 16-bit stream).
  uint16_t t[8][256];
  Crc16Table() {
    for (int i = 0; i < 256; ++i) {
      uint16_t c = (uint16_t)(i << 8);
      for (int k = 0; k < 8; ++k) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
      t[0][i] = c;
    }
    for (int k = 1; k < 8; ++k)
      for (int i = 0; i < 256; ++i) t[k][i] = (uint16_t)((t[k - 1][i] << 8) ^ t[0][t[k - 1][i] >> 8]);
  }
};
uint16_t crc16(const uint8_t* d, size_t n) {
  static const Crc16Table table;  // thread-safe initialisation
  uint16_t c = 0;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const uint8_t* q = d + i;
    const unsigned hi = (unsigned)(c >> 8) ^ q[0], lo = (unsigned)(c & 0xff) ^ q[1];
    c = (uint16_t)(table.t[7][hi] ^ table.t[6][lo] ^ table.t[5][q[2]] ^ table.t[4][q[3]] ^ table.t[3][q[4]] ^
                   table.t[2][q[5]] ^ table.t[1][q[6]] ^ table.t[0][q[7]]);
  }
  for (; i < n; ++i) c = (uint16_t)((c << 8) ^ table.t[0][((c >> 8) ^ d[i]) & 0xff]);
  return c;
}

struct StreamInfo {
  int sample_rate = 0, channels = 0, bps = 0, max_block = 0;
  int64_t total = 0;
  uint8_t md5[16];
  size_t audio_start = 0;
};

int parse_metadata(const uint8_t* d, size_t n, StreamInfo& si) {
  if (n < 42 || memcmp(d, "fLaC", 4) != 0) return AAD_ERR_FORMAT;
  size_t pos = 4;
  bool have = false;
  for (;;) {
    if (pos + 4 > n) return AAD_ERR_FORMAT;
    const bool last = (d[pos] & 0x80) != 0;
    const int type = d[pos] & 0x7f;
    const size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
    pos += 4;
    if (pos + len > n) return AAD_ERR_FORMAT;
    if (type == 0) {
      if (len < 34) return AAD_ERR_FORMAT;
      const uint8_t* s = d + pos;
      si.max_block = (s[2] << 8) | s[3];
      si.sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      si.channels = ((s[12] >> 1) & 7) + 1;
      si.bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      si.total = ((int64_t)(s[13] & 0xf) << 32) | ((int64_t)s[14] << 24) | ((int64_t)s[15] << 16) | ((int64_t)s[16] << 8) | s[17];
      memcpy(si.md5, s + 18, 16);
      have = true;
    }
    pos += len;
    if (last) break;
  }
  if (!have || si.sample_rate <= 0 || si.bps < 4 || si.bps > 32) return AAD_ERR_FORMAT;
  si.audio_start = pos;
  return AAD_OK;
}

bool read_residual(BitReader& br_, int order, int blocksize, int32_t* res) {
  BitReader br = br_;  // a local whose address does not escape: the stores to res[] cannot alias its state
  const int method = (int)br.bits(2);
  if (method > 1) return false;
  const int pbits = method == 0 ? 4 : 5, esc = method == 0 ? 15 : 31;
  const int porder = (int)br.bits(4);
  const int parts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) return false;
  int idx = order;  // samples produced so far (warm-up included)
  for (int p = 0; p < parts; ++p) {
    int count = porder == 0 ? blocksize - order : (p == 0 ? (blocksize >> porder) - order : (blocksize >> porder));
    if (count < 0 || idx + count > blocksize) return false;
    const int k = (int)br.bits(pbits);
    if (k == esc) {
      const int nb = (int)br.bits(5);
      for (int i = 0; i < count; ++i) res[idx++] = (int32_t)br.sbits(nb);
    } else {
      for (int i = 0; i < count; ++i) {
        const uint32_t q = br.unary();
        const uint32_t u = (q << k) | (uint32_t)br.bits(k);
        res[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
      }
    }
    if (br.bad) return false;
  }
  br_ = br;
  return idx == blocksize;
}

// out[i] += (sum_j coef[j] * out[i - 1 - j]) >> shift for i >= ORDER; ACC = int32_t when the sum provably fits
// (bits per sample + coefficient precision + log2(order) <= 32, as in libFLAC), else int64_t
template <int ORDER, typename ACC>
void lpc_restore_fixed(const int32_t* coef, int shift, int blocksize, int32_t* out) {
  for (int i = ORDER; i < blocksize; ++i) {
    ACC acc = 0;
#pragma GCC unroll 32
    for (int j = 0; j < ORDER; ++j) acc += (ACC)coef[j] * (ACC)out[i - 1 - j];
    out[i] = (int32_t)(out[i] + (int32_t)(acc >> shift));
  }
}
template <typename ACC>
void lpc_restore(const int32_t* coef, int order, int shift, int blocksize, int32_t* out) {
  switch (order) {
#define AAD_LPC_CASE(N) case N: lpc_restore_fixed<N, ACC>(coef, shift, blocksize, out); return;
    AAD_LPC_CASE(1) AAD_LPC_CASE(2) AAD_LPC_CASE(3) AAD_LPC_CASE(4) AAD_LPC_CASE(5) AAD_LPC_CASE(6) AAD_LPC_CASE(7) AAD_LPC_CASE(8)
    AAD_LPC_CASE(9) AAD_LPC_CASE(10) AAD_LPC_CASE(11) AAD_LPC_CASE(12)
#undef AAD_LPC_CASE
    default: break;
  }
  for (int i = order; i < blocksize; ++i) {
    ACC acc = 0;
    for (int j = 0; j < order; ++j) acc += (ACC)coef[j] * (ACC)out[i - 1 - j];
    out[i] = (int32_t)(out[i] + (int32_t)(acc >> shift));
  }
}

bool read_subframe(BitReader& br, int bps, int blocksize, int32_t* out) {
  if (br.bit() != 0) return false;
  const int type = (int)br.bits(6);
  int wasted = 0;
  if (br.bit()) wasted = (int)br.unary() + 1;
  if (wasted >= bps) return false;
  bps -= wasted;
  if (type == 0) {  // CONSTANT
    const int32_t v = (int32_t)br.sbits(bps);
    for (int i = 0; i < blocksize; ++i) out[i] = v;
  } else if (type == 1) {  // VERBATIM
    BitReader r = br;
    for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)r.sbits(bps);
    br = r;
  } else if (type >= 8 && type <= 12) {  // FIXED
    const int order = type - 8;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = (int32_t)br.sbits(bps);
    if (!read_residual(br, order, blocksize, out)) return false;
    for (int i = order; i < blocksize; ++i) {
      int64_t pred = 0;
      switch (order) {
        case 1: pred = out[i - 1]; break;
        case 2: pred = 2 * (int64_t)out[i - 1] - out[i - 2]; break;
        case 3: pred = 3 * (int64_t)out[i - 1] - 3 * (int64_t)out[i - 2] + out[i - 3]; break;
        case 4: pred = 4 * (int64_t)out[i - 1] - 6 * (int64_t)out[i - 2] + 4 * (int64_t)out[i - 3] - out[i - 4]; break;
        default: break;
      }
      out[i] = (int32_t)(out[i] + pred);
    }
  } else if (type >= 32) {  // LPC
    const int order = (type & 31) + 1;
    if (order > blocksize) return false;
    for (int i = 0; i < order; ++i) out[i] = (int32_t)br.sbits(bps);
    const int prec = (int)br.bits(4) + 1;
    if (prec == 16) return false;
    const int shift = (int)br.sbits(5);
    if (shift < 0) return false;
    int32_t coef[32];
    for (int j = 0; j < order; ++j) coef[j] = (int32_t)br.sbits(prec);
    if (!read_residual(br, order, blocksize, out)) return false;
    int lg = 0;
    while ((1 << lg) < order) ++lg;
    if (bps + prec + lg <= 32) lpc_restore<int32_t>(coef, order, shift, blocksize, out);
    else lpc_restore<int64_t>(coef, order, shift, blocksize, out);
  } else {
    return false;  // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
  return !br.bad;
}


// MD5 (RFC 1321) of the decoded PCM, as the encoder stored it in STREAMINFO: the mono 16-bit path checks it here so
// that a file costs one library call (no Python-side copies; the call runs without the interpreter lock)
struct Md5 {
  uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  uint64_t len = 0;
  uint8_t buf[64];
  size_t fill = 0;
  static inline uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
  void block(const uint8_t* p) {
    uint32_t m[16];
    memcpy(m, p, 64);  // little-endian host (x86-64 / aarch64 as built here)
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
#define AAD_MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define AAD_MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define AAD_MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define AAD_MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define AAD_MD5_STEP(f, a, b, c, d, k, s, t) a = b + rol(a + f(b, c, d) + m[k] + t, s)
    AAD_MD5_STEP(AAD_MD5_F, a, b, c, d, 0, 7, 0xd76aa478u);  AAD_MD5_STEP(AAD_MD5_F, d, a, b, c, 1, 12, 0xe8c7b756u);
    AAD_MD5_STEP(AAD_MD5_F, c, d, a, b, 2, 17, 0x242070dbu); AAD_MD5_STEP(AAD_MD5_F, b, c, d, a, 3, 22, 0xc1bdceeeu);
    AAD_MD5_STEP(AAD_MD5_F, a, b, c, d, 4, 7, 0xf57c0fafu);  AAD_MD5_STEP(AAD_MD5_F, d, a, b, c, 5, 12, 0x4787c62au);
    AAD_MD5_STEP(AAD_MD5_F, c, d, a, b, 6, 17, 0xa8304613u); AAD_MD5_STEP(AAD_MD5_F, b, c, d, a, 7, 22, 0xfd469501u);
    AAD_MD5_STEP(AAD_MD5_F, a, b, c, d, 8, 7, 0x698098d8u);  AAD_MD5_STEP(AAD_MD5_F, d, a, b, c, 9, 12, 0x8b44f7afu);
    AAD_MD5_STEP(AAD_MD5_F, c, d, a, b, 10, 17, 0xffff5bb1u); AAD_MD5_STEP(AAD_MD5_F, b, c, d, a, 11, 22, 0x895cd7beu);
    AAD_MD5_STEP(AAD_MD5_F, a, b, c, d, 12, 7, 0x6b901122u); AAD_MD5_STEP(AAD_MD5_F, d, a, b, c, 13, 12, 0xfd987193u);
    AAD_MD5_STEP(AAD_MD5_F, c, d, a, b, 14, 17, 0xa679438eu); AAD_MD5_STEP(AAD_MD5_F, b, c, d, a, 15, 22, 0x49b40821u);
    AAD_MD5_STEP(AAD_MD5_G, a, b, c, d, 1, 5, 0xf61e2562u);  AAD_MD5_STEP(AAD_MD5_G, d, a, b, c, 6, 9, 0xc040b340u);
    AAD_MD5_STEP(AAD_MD5_G, c, d, a, b, 11, 14, 0x265e5a51u); AAD_MD5_STEP(AAD_MD5_G, b, c, d, a, 0, 20, 0xe9b6c7aau);
    AAD_MD5_STEP(AAD_MD5_G, a, b, c, d, 5, 5, 0xd62f105du);  AAD_MD5_STEP(AAD_MD5_G, d, a, b, c, 10, 9, 0x02441453u);
    AAD_MD5_STEP(AAD_MD5_G, c, d, a, b, 15, 14, 0xd8a1e681u); AAD_MD5_STEP(AAD_MD5_G, b, c, d, a, 4, 20, 0xe7d3fbc8u);
    AAD_MD5_STEP(AAD_MD5_G, a, b, c, d, 9, 5, 0x21e1cde6u);  AAD_MD5_STEP(AAD_MD5_G, d, a, b, c, 14, 9, 0xc33707d6u);
    AAD_MD5_STEP(AAD_MD5_G, c, d, a, b, 3, 14, 0xf4d50d87u); AAD_MD5_STEP(AAD_MD5_G, b, c, d, a, 8, 20, 0x455a14edu);
    AAD_MD5_STEP(AAD_MD5_G, a, b, c, d, 13, 5, 0xa9e3e905u); AAD_MD5_STEP(AAD_MD5_G, d, a, b, c, 2, 9, 0xfcefa3f8u);
    AAD_MD5_STEP(AAD_MD5_G, c, d, a, b, 7, 14, 0x676f02d9u); AAD_MD5_STEP(AAD_MD5_G, b, c, d, a, 12, 20, 0x8d2a4c8au);
    AAD_MD5_STEP(AAD_MD5_H, a, b, c, d, 5, 4, 0xfffa3942u);  AAD_MD5_STEP(AAD_MD5_H, d, a, b, c, 8, 11, 0x8771f681u);
    AAD_MD5_STEP(AAD_MD5_H, c, d, a, b, 11, 16, 0x6d9d6122u); AAD_MD5_STEP(AAD_MD5_H, b, c, d, a, 14, 23, 0xfde5380cu);
    AAD_MD5_STEP(AAD_MD5_H, a, b, c, d, 1, 4, 0xa4beea44u);  AAD_MD5_STEP(AAD_MD5_H, d, a, b, c, 4, 11, 0x4bdecfa9u);
    AAD_MD5_STEP(AAD_MD5_H, c, d, a, b, 7, 16, 0xf6bb4b60u); AAD_MD5_STEP(AAD_MD5_H, b, c, d, a, 10, 23, 0xbebfbc70u);
    AAD_MD5_STEP(AAD_MD5_H, a, b, c, d, 13, 4, 0x289b7ec6u); AAD_MD5_STEP(AAD_MD5_H, d, a, b, c, 0, 11, 0xeaa127fau);
    AAD_MD5_STEP(AAD_MD5_H, c, d, a, b, 3, 16, 0xd4ef3085u); AAD_MD5_STEP(AAD_MD5_H, b, c, d, a, 6, 23, 0x04881d05u);
    AAD_MD5_STEP(AAD_MD5_H, a, b, c, d, 9, 4, 0xd9d4d039u);  AAD_MD5_STEP(AAD_MD5_H, d, a, b, c, 12, 11, 0xe6db99e5u);
    AAD_MD5_STEP(AAD_MD5_H, c, d, a, b, 15, 16, 0x1fa27cf8u); AAD_MD5_STEP(AAD_MD5_H, b, c, d, a, 2, 23, 0xc4ac5665u);
    AAD_MD5_STEP(AAD_MD5_I, a, b, c, d, 0, 6, 0xf4292244u);  AAD_MD5_STEP(AAD_MD5_I, d, a, b, c, 7, 10, 0x432aff97u);
    AAD_MD5_STEP(AAD_MD5_I, c, d, a, b, 14, 15, 0xab9423a7u); AAD_MD5_STEP(AAD_MD5_I, b, c, d, a, 5, 21, 0xfc93a039u);
    AAD_MD5_STEP(AAD_MD5_I, a, b, c, d, 12, 6, 0x655b59c3u); AAD_MD5_STEP(AAD_MD5_I, d, a, b, c, 3, 10, 0x8f0ccc92u);
    AAD_MD5_STEP(AAD_MD5_I, c, d, a, b, 10, 15, 0xffeff47du); AAD_MD5_STEP(AAD_MD5_I, b, c, d, a, 1, 21, 0x85845dd1u);
    AAD_MD5_STEP(AAD_MD5_I, a, b, c, d, 8, 6, 0x6fa87e4fu);  AAD_MD5_STEP(AAD_MD5_I, d, a, b, c, 15, 10, 0xfe2ce6e0u);
    AAD_MD5_STEP(AAD_MD5_I, c, d, a, b, 6, 15, 0xa3014314u); AAD_MD5_STEP(AAD_MD5_I, b, c, d, a, 13, 21, 0x4e0811a1u);
    AAD_MD5_STEP(AAD_MD5_I, a, b, c, d, 4, 6, 0xf7537e82u);  AAD_MD5_STEP(AAD_MD5_I, d, a, b, c, 11, 10, 0xbd3af235u);
    AAD_MD5_STEP(AAD_MD5_I, c, d, a, b, 2, 15, 0x2ad7d2bbu); AAD_MD5_STEP(AAD_MD5_I, b, c, d, a, 9, 21, 0xeb86d391u);
#undef AAD_MD5_STEP
#undef AAD_MD5_F
#undef AAD_MD5_G
#undef AAD_MD5_H
#undef AAD_MD5_I
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
  }
  void update(const uint8_t* p, size_t n) {
    len += n;
    if (fill) {
      const size_t take = n < 64 - fill ? n : 64 - fill;
      memcpy(buf + fill, p, take);
      fill += take; p += take; n -= take;
      if (fill == 64) { block(buf); fill = 0; }
    }
    for (; n >= 64; p += 64, n -= 64) block(p);
    if (n) { memcpy(buf, p, n); fill = n; }
  }
  void finish(uint8_t out[16]) {
    const uint64_t bits = len * 8;
    const uint8_t one = 0x80, zero = 0;
    update(&one, 1);
    while (fill != 56) update(&zero, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; ++i) lb[i] = (uint8_t)(bits >> (8 * i));
    update(lb, 8);
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (8 * j));
  }
};

}  // namespace

extern "C" {

int aad_flac_info(const uint8_t* data, size_t size, aad_flac_info_t* info) {
  if (!data || !info) return AAD_ERR_INVALID_ARG;
  StreamInfo si;
  int rc = parse_metadata(data, size, si);
  if (rc != AAD_OK) return rc;
  info->sample_rate = si.sample_rate;
  info->channels = si.channels;
  info->bits_per_sample = si.bps;
  info->total_samples = si.total;
  memcpy(info->md5, si.md5, 16);
  return AAD_OK;
}

int aad_flac_decode(const uint8_t* data, size_t size, int32_t* out, int64_t capacity_samples, int64_t* n_decoded) {
  if (!data || !out || !n_decoded || capacity_samples < 0) return AAD_ERR_INVALID_ARG;
  StreamInfo si;
  int rc = parse_metadata(data, size, si);
  if (rc != AAD_OK) return rc;
  const int nch = si.channels;
  std::vector<int32_t> buf[8];
  int64_t done = 0;
  size_t pos = si.audio_start;
  while (pos + 2 <= size) {
    if (!(data[pos] == 0xff && (data[pos + 1] & 0xfe) == 0xf8)) return AAD_ERR_FORMAT;  // sync + reserved bit
    BitReader br(data + pos, size - pos);
    br.bits(15);
    br.bit();  // blocking strategy: the frame / sample number is not needed for sequential decoding
    const int bs_code = (int)br.bits(4), sr_code = (int)br.bits(4), ch_code = (int)br.bits(4), ss_code = (int)br.bits(3);
    if (br.bit() != 0) return AAD_ERR_FORMAT;
    {  // UTF-8 style coded number
      const uint32_t first = (uint32_t)br.bits(8);
      int extra = 0;
      if (first >= 0xfe) extra = 6;
      else if (first >= 0xfc) extra = 5;
      else if (first >= 0xf8) extra = 4;
      else if (first >= 0xf0) extra = 3;
      else if (first >= 0xe0) extra = 2;
      else if (first >= 0xc0) extra = 1;
      else if (first >= 0x80) return AAD_ERR_FORMAT;
      for (int i = 0; i < extra; ++i)
        if ((br.bits(8) & 0xc0) != 0x80) return AAD_ERR_FORMAT;
    }
    int blocksize = 0;
    if (bs_code == 1) blocksize = 192;
    else if (bs_code >= 2 && bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) blocksize = (int)br.bits(8) + 1;
    else if (bs_code == 7) blocksize = (int)br.bits(16) + 1;
    else if (bs_code >= 8) blocksize = 256 << (bs_code - 8);
    else return AAD_ERR_FORMAT;
    if (sr_code == 12) br.bits(8);
    else if (sr_code == 13 || sr_code == 14) br.bits(16);
    else if (sr_code == 15) return AAD_ERR_FORMAT;
    static const int ss_table[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    const int bps = ss_code == 0 ? si.bps : ss_table[ss_code];
    if (bps <= 0 || br.bad) return AAD_ERR_FORMAT;
    const size_t hdr_bytes = br.bitpos() >> 3;
    if (pos + hdr_bytes + 1 > size) return AAD_ERR_FORMAT;
    if (crc8(data + pos, hdr_bytes) != data[pos + hdr_bytes]) return AAD_ERR_FORMAT;
    br.bits(8);
    int frame_ch = 0, side = -1;  // side: channel index coded with one extra bit
    if (ch_code < 8) frame_ch = ch_code + 1;
    else if (ch_code <= 10) {
      frame_ch = 2;
      side = ch_code == 9 ? 0 : 1;
    } else return AAD_ERR_FORMAT;
    if (frame_ch != nch) return AAD_ERR_FORMAT;
    for (int c = 0; c < nch; ++c) {
      buf[c].resize(blocksize);
      if (!read_subframe(br, bps + (c == side ? 1 : 0), blocksize, buf[c].data())) return AAD_ERR_FORMAT;
    }
    br.align();
    const size_t body = br.bitpos() >> 3;
    if (pos + body + 2 > size) return AAD_ERR_FORMAT;
    if (crc16(data + pos, body) != (uint16_t)((data[pos + body] << 8) | data[pos + body + 1])) return AAD_ERR_FORMAT;
    pos += body + 2;
    if (ch_code == 8) {         // left, side
      for (int i = 0; i < blocksize; ++i) buf[1][i] = buf[0][i] - buf[1][i];
    } else if (ch_code == 9) {  // side, right
      for (int i = 0; i < blocksize; ++i) buf[0][i] = buf[0][i] + buf[1][i];
    } else if (ch_code == 10) { // mid, side
      for (int i = 0; i < blocksize; ++i) {
        const int32_t s = buf[1][i];
        const int32_t m = (int32_t)(((uint32_t)buf[0][i] << 1) | (uint32_t)(s & 1));
        buf[0][i] = (m + s) >> 1;
        buf[1][i] = (m - s) >> 1;
      }
    }
    if (done + blocksize > capacity_samples) return AAD_ERR_INVALID_ARG;
    for (int i = 0; i < blocksize; ++i)
      for (int c = 0; c < nch; ++c) out[(done + i) * nch + c] = buf[c][i];
    done += blocksize;
    if (si.total > 0 && done >= si.total) break;
  }
  *n_decoded = done;
  return AAD_OK;
}

int aad_flac_decode_pcm16(const uint8_t* data, size_t size, int16_t* out, int64_t capacity_samples, int64_t* n_decoded,
                          int32_t* md5_state) {
  if (!data || !out || !n_decoded || capacity_samples < 0) return AAD_ERR_INVALID_ARG;
  StreamInfo si;
  int rc = parse_metadata(data, size, si);
  if (rc != AAD_OK) return rc;
  if (si.channels != 1 || si.bps != 16) return AAD_ERR_UNSUPPORTED;
  std::vector<int32_t> tmp((size_t)std::max<int64_t>(capacity_samples, 1));
  rc = aad_flac_decode(data, size, tmp.data(), capacity_samples, n_decoded);
  if (rc != AAD_OK) return rc;
  const int64_t n = *n_decoded;
  for (int64_t i = 0; i < n; ++i) out[i] = (int16_t)tmp[i];
  if (md5_state) {
    bool any = false;
    for (int i = 0; i < 16; ++i) any = any || si.md5[i] != 0;
    if (!any) *md5_state = 0;  // the encoder stored no checksum
    else {
      Md5 m;
      m.update(reinterpret_cast<const uint8_t*>(out), (size_t)n * 2);  // little-endian host: the PCM as the encoder hashed it
      uint8_t dg[16];
      m.finish(dg);
      *md5_state = memcmp(dg, si.md5, 16) == 0 ? 1 : -1;
    }
  }
  return AAD_OK;
}

}  // extern "C"
