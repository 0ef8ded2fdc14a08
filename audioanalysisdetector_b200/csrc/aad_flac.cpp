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

#include <vector>

#include "../../include/aad.h"

namespace {

struct BitReader {
  const uint8_t* p;
  size_t n, pos = 0;  // pos in bits
  bool bad = false;
  BitReader(const uint8_t* d, size_t size) : p(d), n(size) {}
  inline uint32_t bit() {
    if ((pos >> 3) >= n) {
      bad = true;
      return 0;
    }
    uint32_t b = (p[pos >> 3] >> (7 - (pos & 7))) & 1u;
    ++pos;
    return b;
  }
  inline uint64_t bits(int k) {  // k <= 57
    uint64_t v = 0;
    while (k > 0) {
      if ((pos >> 3) >= n) {
        bad = true;
        return 0;
      }
      const int avail = 8 - (int)(pos & 7), take = k < avail ? k : avail;
      const uint32_t cur = p[pos >> 3];
      v = (v << take) | ((cur >> (avail - take)) & ((1u << take) - 1u));
      pos += take;
      k -= take;
    }
    return v;
  }
  inline int64_t sbits(int k) {
    if (k == 0) return 0;
    uint64_t v = bits(k);
    const uint64_t sign = 1ull << (k - 1);
    return (int64_t)((v ^ sign) - sign);
  }
  inline uint32_t unary() {  // number of 0 bits before the next 1 bit
    uint32_t q = 0;
    while (!bad) {
      // fast path: look at the rest of the current byte
      const size_t byte = pos >> 3;
      if (byte >= n) {
        bad = true;
        break;
      }
      const int avail = 8 - (int)(pos & 7);
      const uint32_t cur = p[byte] & ((1u << avail) - 1u);
      if (cur == 0) {
        q += avail;
        pos += avail;
      } else {
        const int lead = __builtin_clz(cur) - (32 - avail);
        q += lead;
        pos += lead + 1;
        return q;
      }
    }
    return q;
  }
  inline void align() { pos = (pos + 7) & ~(size_t)7; }
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
  uint16_t t[256];
  Crc16Table() {
    for (int i = 0; i < 256; ++i) {
      uint16_t c = (uint16_t)(i << 8);
      for (int k = 0; k < 8; ++k) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
      t[i] = c;
    }
  }
};
uint16_t crc16(const uint8_t* d, size_t n) {
  static const Crc16Table table;  // thread-safe initialisation
  uint16_t c = 0;
  for (size_t i = 0; i < n; ++i) c = (uint16_t)((c << 8) ^ table.t[((c >> 8) ^ d[i]) & 0xff]);
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

bool read_residual(BitReader& br, int order, int blocksize, int32_t* res) {
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
  return idx == blocksize;
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
    for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)br.sbits(bps);
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
    for (int i = order; i < blocksize; ++i) {
      int64_t acc = 0;
      for (int j = 0; j < order; ++j) acc += (int64_t)coef[j] * out[i - 1 - j];
      out[i] = (int32_t)(out[i] + (acc >> shift));
    }
  } else {
    return false;  // reserved subframe type
  }
  if (wasted)
    for (int i = 0; i < blocksize; ++i) out[i] = (int32_t)((uint32_t)out[i] << wasted);
  return !br.bad;
}

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
    const size_t hdr_bytes = br.pos >> 3;
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
    const size_t body = br.pos >> 3;
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

}  // extern "C"
