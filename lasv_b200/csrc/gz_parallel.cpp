// gz_parallel.cpp -- see gz_parallel.h.  Host code (one of the rows next to the hot path, SURVEY 8f-4: the text
// has to reach the device at the rate the graph is built); nothing of zlib's inflate is used here, only its
// crc32 / crc32_combine for the member checks.
#include "gz_parallel.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#elif defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <thread>

namespace lasv {

// ------------------------------------------------------------------ CRC-32 by carry-less multiplication
// (gzip's CRC-32 folded 64 bytes at a time with PCLMULQDQ: Gopal et al., "Fast CRC computation for generic polynomials
// using PCLMULQDQ instruction", Intel 2009; the constants are x^n mod P for the reflected polynomial 0x1DB710641.
// Checked against zlib's crc32 when first used; zlib's is the fallback.)
#if defined(__x86_64__)
namespace {
__attribute__((target("pclmul,sse4.1")))
uint32_t crc32_clmul_core(uint32_t crc, const unsigned char *buf, size_t len) {  // len >= 64, multiple of 16; raw (not inverted) state
  const __m128i k12 = _mm_set_epi64x(0x00000001c6e41596ll, 0x0000000154442bd4ll);
  const __m128i k34 = _mm_set_epi64x(0x00000000ccaa009ell, 0x00000001751997d0ll);
  const __m128i k5 = _mm_set_epi64x(0, 0x0000000163cd6124ll);
  const __m128i poly = _mm_set_epi64x(0x00000001F7011641ll, 0x00000001DB710641ll);
  const __m128i mask32 = _mm_set_epi32(0, 0, 0, -1);
  __m128i x1 = _mm_loadu_si128((const __m128i *)buf), x2 = _mm_loadu_si128((const __m128i *)(buf + 16));
  __m128i x3 = _mm_loadu_si128((const __m128i *)(buf + 32)), x4 = _mm_loadu_si128((const __m128i *)(buf + 48));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  buf += 64; len -= 64;
  while (len >= 64) {
    __m128i t1 = _mm_clmulepi64_si128(x1, k12, 0x00), t2 = _mm_clmulepi64_si128(x2, k12, 0x00);
    __m128i t3 = _mm_clmulepi64_si128(x3, k12, 0x00), t4 = _mm_clmulepi64_si128(x4, k12, 0x00);
    x1 = _mm_clmulepi64_si128(x1, k12, 0x11); x2 = _mm_clmulepi64_si128(x2, k12, 0x11);
    x3 = _mm_clmulepi64_si128(x3, k12, 0x11); x4 = _mm_clmulepi64_si128(x4, k12, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t1), _mm_loadu_si128((const __m128i *)buf));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, t2), _mm_loadu_si128((const __m128i *)(buf + 16)));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, t3), _mm_loadu_si128((const __m128i *)(buf + 32)));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, t4), _mm_loadu_si128((const __m128i *)(buf + 48)));
    buf += 64; len -= 64;
  }
  __m128i t = _mm_clmulepi64_si128(x1, k34, 0x00); x1 = _mm_clmulepi64_si128(x1, k34, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x2);
  t = _mm_clmulepi64_si128(x1, k34, 0x00); x1 = _mm_clmulepi64_si128(x1, k34, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x3);
  t = _mm_clmulepi64_si128(x1, k34, 0x00); x1 = _mm_clmulepi64_si128(x1, k34, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, t), x4);
  while (len >= 16) {
    t = _mm_clmulepi64_si128(x1, k34, 0x00); x1 = _mm_clmulepi64_si128(x1, k34, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, t), _mm_loadu_si128((const __m128i *)buf));
    buf += 16; len -= 16;
  }
  // 128 -> 64
  x2 = _mm_clmulepi64_si128(x1, k34, 0x10);
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
  // 64 -> 32
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, mask32);
  x1 = _mm_clmulepi64_si128(x1, k5, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  // Barrett
  x2 = x1;
  x1 = _mm_and_si128(x1, mask32);
  x1 = _mm_clmulepi64_si128(x1, poly, 0x10);
  x1 = _mm_and_si128(x1, mask32);
  x1 = _mm_clmulepi64_si128(x1, poly, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}

bool clmul_usable() {
  if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
  unsigned char t[211];
  for (int i = 0; i < 211; i++) t[i] = (unsigned char)(i * 37 + 11);
  for (size_t n : {64u, 80u, 208u}) {
    const uint32_t want = (uint32_t)crc32(0x1234567u, t, (uInt)n);
    if (~crc32_clmul_core(~0x1234567u, t, n) != want) return false;
  }
  return true;
}
}  // namespace
#endif

uint32_t fast_crc32(uint32_t crc, const unsigned char *buf, size_t len) {
#if defined(__x86_64__)
  static const bool usable = clmul_usable();
  if (usable && len >= 64) {
    const size_t n = len & ~(size_t)15;
    crc = ~crc32_clmul_core(~crc, buf, n);
    buf += n;
    len -= n;
  }
#endif
  while (len) {  // (zlib takes 32-bit lengths)
    const size_t n = std::min<size_t>(len, 1u << 30);
    crc = (uint32_t)crc32(crc, buf, (uInt)n);
    buf += n;
    len -= n;
  }
  return crc;
}

namespace {

constexpr uint32_t CTX = 32768;             // deflate's window
constexpr int LIT_PB = 11, DIST_PB = 8;     // bits of the first-level decoding tables
constexpr int LIT_SIZE = (1 << LIT_PB) + 288 * 16, DIST_SIZE = (1 << DIST_PB) + 32 * 128;
constexpr size_t WINDOW = 8u << 20;         // text handed out per call
constexpr size_t PIECE = 1u << 20;          // ... translated in pieces of this size
constexpr size_t OUT_CAP = 64u << 20;       // symbols after which a chunk stops at the next block boundary
constexpr size_t TRIAL_CAP = 16u << 20;     // symbols a block may inflate to while its start is only a guess
constexpr uint64_t PENDING = ~0ull, NONE = ~0ull - 1;

// table entries: bits 0-7 code bits to consume, 8-11 type, 12-15 extra bits (T_SUB: bits of the second level),
// 16-31 value (literal, base length, base distance, offset of the second-level table)
enum : uint32_t { T_LIT = 0, T_LEN = 1, T_EOB = 2, T_SUB = 3, T_BAD = 4 };
enum { IR_OK = 0, IR_ERR = -1, IR_TRUNC = -2, IR_SPACE = 1, IR_CAREFUL = 2 };

inline uint32_t entry(uint32_t type, uint32_t nbits, uint32_t extra, uint32_t value) {
  return nbits | (type << 8) | (extra << 12) | (value << 16);
}
inline uint32_t e_bits(uint32_t e) { return e & 0xffu; }
inline uint32_t e_type(uint32_t e) { return (e >> 8) & 0xfu; }
inline uint32_t e_extra(uint32_t e) { return (e >> 12) & 0xfu; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

// RFC 1951, 3.2.5
const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769,
                                1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t PRECODE_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint64_t load64(const uint8_t *p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;  // (little endian hosts only, like the rest of the library)
}

inline uint32_t bit_reverse(uint32_t c, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) {
    r = (r << 1) | (c & 1);
    c >>= 1;
  }
  return r;
}

// Two-level decoding table of a canonical Huffman code (code lengths <= 15).  -1: the lengths are over-subscribed,
// or incomplete with more than one code (what zlib's inflate_table rejects for these two alphabets).
int build_table(const uint8_t *lens, int n, int pb, uint32_t *table, bool is_dist) {
  int count[16] = {0};
  for (int i = 0; i < n; i++) count[lens[i]]++;
  count[0] = 0;
  int max = 15;
  while (max > 0 && !count[max]) max--;
  const uint32_t bad = entry(T_BAD, 1, 0, 0);
  const int primary = 1 << pb;
  for (int i = 0; i < primary; i++) table[i] = bad;
  if (max == 0) return 0;  // no codes at all: an error only if one is used
  int left = 1;
  for (int l = 1; l <= 15; l++) {
    left <<= 1;
    left -= count[l];
    if (left < 0) return -1;
  }
  if (left > 0 && max != 1) return -1;
  uint32_t next_code[16], code = 0;
  for (int l = 1; l <= 15; l++) {
    code = (code + (uint32_t)count[l - 1]) << 1;
    next_code[l] = code;
  }
  uint16_t rev[288];
  uint8_t submax[1 << LIT_PB];
  if (max > pb) memset(submax, 0, (size_t)primary);
  for (int i = 0; i < n; i++) {
    const int l = lens[i];
    if (!l) continue;
    rev[i] = (uint16_t)bit_reverse(next_code[l]++, l);
    if (l > pb) {
      uint8_t &m = submax[rev[i] & (primary - 1)];
      if (m < l) m = (uint8_t)l;
    }
  }
  if (max > pb) {
    int next = primary;
    for (int p = 0; p < primary; p++)
      if (submax[p]) {
        const int sb = submax[p] - pb;
        table[p] = entry(T_SUB, (uint32_t)pb, (uint32_t)sb, (uint32_t)next);
        for (int j = 0; j < (1 << sb); j++) table[next + j] = bad;
        next += 1 << sb;
      }
  }
  for (int i = 0; i < n; i++) {
    const int l = lens[i];
    if (!l) continue;
    const uint32_t nb = (uint32_t)(l > pb ? l - pb : l);
    uint32_t e;
    // (lengths and distances: bits = code + extra bits, consumed at once)
    if (is_dist) e = i < 30 ? entry(T_LIT, nb + DIST_EXTRA[i], DIST_EXTRA[i], DIST_BASE[i]) : entry(T_BAD, nb, 0, 0);
    else if (i < 256) e = entry(T_LIT, nb, 0, (uint32_t)i);
    else if (i == 256) e = entry(T_EOB, nb, 0, 0);
    else if (i < 286) e = entry(T_LEN, nb + LEN_EXTRA[i - 257], LEN_EXTRA[i - 257], LEN_BASE[i - 257]);
    else e = entry(T_BAD, nb, 0, 0);
    if (l <= pb) {
      for (int j = rev[i]; j < primary; j += 1 << l) table[j] = e;
    } else {
      const uint32_t t = table[rev[i] & (primary - 1)];
      const int off = (int)e_value(t), sb = (int)e_extra(t);
      for (int j = rev[i] >> pb; j < (1 << sb); j += 1 << (l - pb)) table[off + j] = e;
    }
  }
  return 0;
}

// Length of the gzip member header at d[0..n) (RFC 1952): > 0; 0 if this is not a gzip header at all; -1 if it is one
// that zlib rejects (method, reserved flags) or that is cut off.
long gzip_header_len(const uint8_t *d, size_t n) {
  if (n < 2 || d[0] != 0x1f || d[1] != 0x8b) return 0;
  if (n < 10 || d[2] != 8 || (d[3] & 0xe0)) return -1;
  const int flg = d[3];
  size_t pos = 10;
  if (flg & 4) {
    if (pos + 2 > n) return -1;
    pos += 2 + ((size_t)d[pos] | ((size_t)d[pos + 1] << 8));
  }
  for (int f = 8; f <= 16; f <<= 1)
    if (flg & f) {
      while (pos < n && d[pos]) pos++;
      pos++;
    }
  if (flg & 2) pos += 2;
  return pos > n ? -1 : (long)pos;
}

// One deflate decoder: bits from a mapped file, block by block; out come 16-bit symbols (see gz_parallel.h) where the
// text in front is not known, plain bytes where it is (the members of a BGZF file).
template <typename SYM>
struct InflaterT {
  const uint8_t *base = nullptr, *end = nullptr, *p = nullptr;
  uint64_t buf = 0;
  int cnt = 0;        // valid bits in buf
  uint32_t over = 0;  // zero bytes fed behind the end of the file
  SYM *mem = nullptr;  // [CTX symbols in front of the chunk][its output ...]
  size_t cap = 0;
  SYM *out = nullptr, *out_min = nullptr;  // next symbol; oldest symbol a distance may reach
  size_t limit = 0;                             // symbols of output after which decoding fails (0: none)
  uint32_t lit[LIT_SIZE], dist[DIST_SIZE];

  ~InflaterT() { free(mem); }
  InflaterT() {}
  InflaterT(const InflaterT &) = delete;
  InflaterT &operator=(const InflaterT &) = delete;

  bool alloc(size_t symbols) {
    cap = CTX + symbols + 1024;
    mem = (SYM *)malloc(cap * sizeof(SYM));
    out = out_min = mem ? mem + CTX : nullptr;
    return mem != nullptr;
  }
  // a buffer used before (its pages are there already), and the way back
  void adopt(SYM *m, size_t c) {
    mem = m;
    cap = c;
    out = out_min = mem + CTX;
  }
  SYM *release(size_t *c) {
    SYM *m = mem;
    *c = cap;
    mem = nullptr;
    cap = 0;
    return m;
  }
  bool grow() {
    const size_t o = (size_t)(out - mem), m = (size_t)(out_min - mem);
    SYM *n = (SYM *)realloc(mem, cap * 2 * sizeof(SYM));
    if (!n) return false;
    mem = n;
    cap *= 2;
    out = mem + o;
    out_min = mem + m;
    return true;
  }
  size_t produced() const { return (size_t)(out - (mem + CTX)); }

  void start(const uint8_t *b, const uint8_t *e, uint64_t bitpos) {
    base = b;
    end = e;
    p = b + (bitpos >> 3);
    buf = 0;
    cnt = 0;
    over = 0;
    const int skip = (int)(bitpos & 7);
    if (skip) {
      need(8);
      buf >>= skip;
      cnt -= skip;
    }
  }
  uint64_t bitpos() const { return ((uint64_t)(p - base) + over) * 8 - (uint64_t)cnt; }
  bool past_end() const { return bitpos() > (uint64_t)(end - base) * 8; }
  // at least n <= 57 bits in buf (zeros behind the end of the file: past_end() tells)
  void need(int n) {
    while (cnt < n) {
      if (p < end) buf |= (uint64_t)*p++ << cnt;
      else over++;
      cnt += 8;
    }
  }
  uint32_t take(int n) {
    need(n);
    const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1));
    buf >>= n;
    cnt -= n;
    return v;
  }
  void align_byte() {
    const int r = cnt & 7;
    buf >>= r;
    cnt -= r;
  }

  // the symbols of a compressed block, up to and including its end-of-block code
  template <bool CAREFUL>
  int run() {
    uint64_t b = buf;
    int c = cnt;
    const uint8_t *q = p;
    SYM *o = out;
    SYM *const lim = mem + cap - 300;
    const SYM *const omin = out_min;
    constexpr uint32_t WIDE = 32 / sizeof(SYM), NARROW = 8 / sizeof(SYM);  // symbols per 32-byte / 8-byte copy
    constexpr uint32_t LM = (1u << LIT_PB) - 1, DM = (1u << DIST_PB) - 1;
    int rc;
#define LASV_GZ_REFILL()                                  \
  if (!CAREFUL) {                                         \
    if ((size_t)(end - q) < 8) {                          \
      rc = IR_CAREFUL;                                    \
      break;                                              \
    }                                                     \
    b |= load64(q) << c;                                  \
    q += (63 - c) >> 3;                                   \
    c |= 56;                                              \
  } else {                                                \
    while (c <= 56) {                                     \
      if (q < end) b |= (uint64_t)*q++ << c;              \
      else over++;                                        \
      c += 8;                                             \
    }                                                     \
    if (over > 8) {                                       \
      rc = IR_TRUNC;                                      \
      break;                                              \
    }                                                     \
  }
    // e: the entry of the bits at the head of b, looked up ahead (a refill only adds bits above the ones it was
    // looked up with, so it stays valid across one)
    uint32_t e = lit[b & LM];
    for (;;) {
      if (o >= lim) {
        rc = IR_SPACE;
        break;
      }
      if (c < 48) {
        LASV_GZ_REFILL();
        e = lit[b & LM];  // (c may have been 0: nothing looked up yet)
      }
      // up to three literals of the first level on one refill (at most 11 bits each; a match needs 48)
      if ((e & 0xf00u) == (T_LIT << 8)) {
        b >>= e_bits(e);
        c -= (int)e_bits(e);
        *o++ = (SYM)e_value(e);
        e = lit[b & LM];
        if ((e & 0xf00u) == (T_LIT << 8)) {
          b >>= e_bits(e);
          c -= (int)e_bits(e);
          *o++ = (SYM)e_value(e);
          e = lit[b & LM];
          if ((e & 0xf00u) == (T_LIT << 8)) {
            b >>= e_bits(e);
            c -= (int)e_bits(e);
            *o++ = (SYM)e_value(e);
            e = lit[b & LM];
          }
        }
        continue;
      }
      if (e_type(e) == T_SUB) {
        b >>= LIT_PB;
        c -= LIT_PB;
        e = lit[e_value(e) + (uint32_t)(b & ((1u << e_extra(e)) - 1))];
      }
      uint64_t sv = b;
      b >>= e_bits(e);
      c -= (int)e_bits(e);
      const uint32_t t = e_type(e);
      if (t == T_LIT) {  // (a literal with a long code)
        *o++ = (SYM)e_value(e);
        e = lit[b & LM];
        continue;
      }
      if (t == T_EOB) {
        rc = IR_OK;
        break;
      }
      if (t != T_LEN) {
        rc = IR_ERR;
        break;
      }
      // (entries of lengths and distances: bits = code + extra bits)
      const uint32_t len = e_value(e) + (uint32_t)((sv >> (e_bits(e) - e_extra(e))) & ((1u << e_extra(e)) - 1));
      uint32_t d = dist[b & DM];
      if (e_type(d) == T_SUB) {
        b >>= DIST_PB;
        c -= DIST_PB;
        d = dist[e_value(d) + (uint32_t)(b & ((1u << e_extra(d)) - 1))];
      }
      sv = b;
      b >>= e_bits(d);
      c -= (int)e_bits(d);
      if (e_type(d) != T_LIT) {
        rc = IR_ERR;
        break;
      }
      const uint32_t dd = e_value(d) + (uint32_t)((sv >> (e_bits(d) - e_extra(d))) & ((1u << e_extra(d)) - 1));
      e = lit[b & LM];  // the next symbol's entry, on its way while the match is copied
      if ((size_t)(o - omin) < dd) {  // "invalid distance too far back"
        rc = IR_ERR;
        break;
      }
      const SYM *s = o - dd;
      SYM *const oe = o + len;
      if (dd >= WIDE) {  // 32 bytes at once cover most matches (the buffer has room behind the match)
        memcpy(o, s, 32);
        if (len > WIDE) {
          do {
            o += WIDE;
            s += WIDE;
            memcpy(o, s, 32);
          } while (o + WIDE < oe);
        }
      } else if (dd >= NARROW) {
        do {
          memcpy(o, s, 8);
          o += NARROW;
          s += NARROW;
        } while (o < oe);
      } else if (dd == 1) {  // a run of one symbol
        const uint64_t v = (uint64_t)s[0] * (sizeof(SYM) == 2 ? 0x0001000100010001ull : 0x0101010101010101ull);
        do {
          memcpy(o, &v, 8);
          o += NARROW;
        } while (o < oe);
      } else {
        do *o++ = *s++;
        while (o < oe);
      }
      o = oe;
    }
#undef LASV_GZ_REFILL
    buf = b;
    cnt = c;
    p = q;
    out = o;
    return rc;
  }

  int block_body() {
    for (;;) {
      int rc = this->template run<false>();
      if (rc == IR_CAREFUL) rc = this->template run<true>();
      if (rc != IR_SPACE) return rc;
      if (limit && produced() > limit) return IR_ERR;
      if (!grow()) return IR_ERR;
    }
  }

  // header of the next block; then the block's content.  *final_block: BFINAL
  int block(bool *final_block) {
    *final_block = take(1) != 0;
    const uint32_t btype = take(2);
    if (btype == 3) return IR_ERR;
    if (btype == 0) {  // stored
      align_byte();
      const uint32_t len = take(16), nlen = take(16);
      if (past_end()) return IR_TRUNC;
      if ((len ^ 0xffffu) != nlen) return IR_ERR;
      while ((size_t)(mem + cap - out) < (size_t)len + 400)
        if (!grow()) return IR_ERR;
      uint32_t left = len;
      while (left && cnt >= 8) {
        *out++ = (SYM)(buf & 0xff);
        buf >>= 8;
        cnt -= 8;
        left--;
      }
      if (left) {
        buf = 0;  // (cnt == 0 here: the bytes continue at p)
        cnt = 0;
        if ((size_t)(end - p) < left) return IR_TRUNC;
        for (uint32_t i = 0; i < left; i++) out[i] = p[i];
        out += left;
        p += left;
      }
      return IR_OK;
    }
    uint8_t lens[320];
    if (btype == 1) {  // fixed codes (RFC 1951, 3.2.6)
      int i = 0;
      for (; i < 144; i++) lens[i] = 8;
      for (; i < 256; i++) lens[i] = 9;
      for (; i < 280; i++) lens[i] = 7;
      for (; i < 288; i++) lens[i] = 8;
      build_table(lens, 288, LIT_PB, lit, false);
      for (i = 0; i < 32; i++) lens[i] = 5;
      build_table(lens, 32, DIST_PB, dist, true);
    } else {
      const int hlit = (int)take(5) + 257, hdist = (int)take(5) + 1, hclen = (int)take(4) + 4;
      if (hlit > 286 || hdist > 30) return IR_ERR;
      uint8_t pl[19] = {0};
      for (int i = 0; i < hclen; i++) pl[PRECODE_ORDER[i]] = (uint8_t)take(3);
      // the code of the code lengths: must be complete (zlib: inflate_table(CODES, ...))
      int count[8] = {0};
      for (int i = 0; i < 19; i++) count[pl[i]]++;
      count[0] = 0;
      int left = 1;
      for (int l = 1; l <= 7; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return IR_ERR;
      }
      if (left != 0) return IR_ERR;
      uint16_t pre[128];
      uint32_t next_code[8], code = 0;
      for (int l = 1; l <= 7; l++) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
      }
      for (int i = 0; i < 19; i++) {
        const int l = pl[i];
        if (!l) continue;
        const uint32_t r = bit_reverse(next_code[l]++, l);
        for (uint32_t j = r; j < 128; j += 1u << l) pre[j] = (uint16_t)((l << 8) | i);
      }
      const int total = hlit + hdist;
      int i = 0;
      while (i < total) {
        need(14);
        const uint16_t pe = pre[buf & 127];
        buf >>= pe >> 8;
        cnt -= pe >> 8;
        const int sym = pe & 0xff;
        if (sym < 16) {
          lens[i++] = (uint8_t)sym;
          continue;
        }
        int rep;
        uint8_t val = 0;
        if (sym == 16) {
          if (i == 0) return IR_ERR;
          val = lens[i - 1];
          rep = 3 + (int)take(2);
        } else if (sym == 17) {
          rep = 3 + (int)take(3);
        } else {
          rep = 11 + (int)take(7);
        }
        if (i + rep > total) return IR_ERR;
        while (rep--) lens[i++] = val;
      }
      if (past_end()) return IR_TRUNC;
      if (lens[256] == 0) return IR_ERR;  // no end-of-block code
      if (build_table(lens, hlit, LIT_PB, lit, false) != 0) return IR_ERR;
      if (build_table(lens + hlit, hdist, DIST_PB, dist, true) != 0) return IR_ERR;
    }
    return block_body();
  }
};

typedef InflaterT<uint16_t> Inflater;

struct TextTable {
  bool ok[256];
  TextTable() {
    for (int i = 0; i < 256; i++) ok[i] = (i >= 32 && i < 127) || i == '\n' || i == '\r' || i == '\t';
  }
};
const TextTable TEXT;

// symbols -> bytes.  Runs of 16 literals (nearly everything behind the first 32 KB of a chunk) are packed at once.
void translate(const uint16_t *s, unsigned char *d, size_t n, const unsigned char *lut) {
  size_t j = 0;
#if defined(__SSE2__)
  const __m128i zero = _mm_setzero_si128();
  for (; j + 16 <= n; j += 16) {
    const __m128i a = _mm_loadu_si128((const __m128i *)(s + j)), b = _mm_loadu_si128((const __m128i *)(s + j + 8));
    const __m128i hi = _mm_srli_epi16(_mm_or_si128(a, b), 8);
    if (_mm_movemask_epi8(_mm_cmpeq_epi8(hi, zero)) == 0xffff) {
      _mm_storeu_si128((__m128i *)(d + j), _mm_packus_epi16(a, b));
    } else {
      for (size_t k = j; k < j + 16; k++) d[k] = lut[s[k]];
    }
  }
#endif
  for (; j < n; j++) d[j] = lut[s[j]];
}

struct MemberEnd {
  size_t pos;  // symbols of the chunk in front of it
  uint32_t crc, isize;
};

}  // namespace

struct MemberInflater::Impl {
  InflaterT<uint8_t> in;
};

MemberInflater::MemberInflater() : impl_(new Impl()) {}
MemberInflater::~MemberInflater() {}

bool MemberInflater::inflate(const unsigned char *src, size_t n, unsigned char *dst, size_t expect) {
  InflaterT<uint8_t> &in = impl_->in;
  // (inflated into a scratch buffer that stays in the cache and copied: the decoder writes up to 31 bytes behind a
  // match, which must not touch the neighbouring block's text another thread is writing)
  if (!in.mem && !in.alloc(std::max<size_t>(expect, 65536) + 4096)) return false;
  in.start(src, src + n, 0);
  in.out = in.out_min = in.mem + CTX;
  in.limit = expect;
  bool fin = false;
  while (!fin) {
    if (in.block(&fin) != IR_OK || in.produced() > expect) return false;
  }
  if (in.produced() != expect || in.past_end()) return false;
  memcpy(dst, in.mem + CTX, expect);
  return true;
}

struct PgzStream::Chunk {
  Inflater inf;
  size_t search_from = 0, search_to = 0;  // the chunk's compressed bytes
  std::atomic<uint64_t> start_bit{PENDING};
  uint64_t end_bit = 0;
  int synced_with = -1;  // index of the chunk whose start this one stopped on (>= the number of chunks: none)
  int status = 0;        // 0 stopped at a block boundary, 1 the stream has ended, -1 error (err)
  std::string err;
  size_t n_out = 0;
  std::vector<MemberEnd> ends;
  unsigned char lut[256 + CTX];  // symbol -> byte, once the text in front of the chunk is known
};

PgzStream::PgzStream() {}

void PgzStream::recycle(std::unique_ptr<Chunk> &c) {
  if (!c) return;
  size_t cap = 0;
  uint16_t *m = c->inf.release(&cap);
  if (m && pool_.size() < (size_t)threads_ && cap <= (OUT_CAP + CTX) * 2) pool_.push_back({m, cap});
  else free(m);
  c.reset();
}

PgzStream::~PgzStream() {
  queue_.clear();
  for (auto &b : pool_) free(b.first);
  if (data_) munmap((void *)data_, size_);
  if (fd_ >= 0) ::close(fd_);
}

bool PgzStream::open(const char *path, int threads, size_t min_bytes, size_t chunk_bytes) {
  fd_ = ::open(path, O_RDONLY);
  if (fd_ < 0) return false;
  struct stat st;
  if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < std::max<size_t>(min_bytes, 64)) {
    ::close(fd_);
    fd_ = -1;
    return false;
  }
  void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd_, 0);
  if (m == MAP_FAILED) {
    ::close(fd_);
    fd_ = -1;
    return false;
  }
  data_ = (const unsigned char *)m;
  size_ = (size_t)st.st_size;
  if (gzip_header_len(data_, size_) <= 0) {
    munmap(m, size_);
    data_ = nullptr;
    ::close(fd_);
    fd_ = -1;
    return false;
  }
  madvise(m, size_, MADV_SEQUENTIAL);
  threads_ = std::max(1, std::min(threads, 64));
  chunk_bytes_ = std::max<size_t>(chunk_bytes, 4096);
  reset();
  return true;
}

void PgzStream::reset() {
  for (auto &c : queue_) recycle(c);  // (the buffers stay for the next round)
  queue_.clear();
  q_chunk_ = q_off_ = q_end_ = 0;
  cur_bit_ = (uint64_t)gzip_header_len(data_, size_) * 8;
  window_.assign(CTX, 0);
  window_valid_ = 0;
  finished_ = false;
  sequential_ = false;
  failed_rounds_ = 0;
  pending_err_.clear();
  crc_run_ = len_run_ = 0;
}

void PgzStream::rewind() { reset(); }

// Is `bit` the start of a deflate block?  (see gz_parallel.h)  The block is inflated on the way and stays in `in`.
static bool try_start(Inflater &in, const uint8_t *data, size_t size, uint64_t bit) {
  in.start(data, data + size, bit);
  in.out = in.mem + CTX;
  in.out_min = in.mem;
  in.limit = TRIAL_CAP;
  bool fin = false;
  if (in.block(&fin) != IR_OK || fin || in.past_end()) return false;
  const size_t n = in.produced();
  if (n < 64) return false;
  const uint16_t *s = in.mem + CTX;
  for (size_t i = 0; i < n; i++)
    if (s[i] < 256 && !TEXT.ok[s[i]]) return false;
  in.need(3);
  if (((in.buf >> 1) & 3) == 3) return false;  // what follows is not a block header
  return true;
}

// Tables of the search: for 10 bits of the stream, the bit offsets 0..7 at which "BFINAL = 0, BTYPE = 2" stands;
// for four 3-bit code lengths, their share of the Kraft sum in 1/128ths.
struct ScanTables {
  uint8_t cand[1024];
  uint16_t kraft[4096];
  ScanTables() {
    for (int v = 0; v < 1024; v++) {
      cand[v] = 0;
      for (int s = 0; s < 8; s++)
        if (((v >> s) & 7) == 4) cand[v] |= (uint8_t)(1 << s);
    }
    for (int v = 0; v < 4096; v++) {
      int k = 0;
      for (int i = 0; i < 4; i++) {
        const int l = (v >> (3 * i)) & 7;
        if (l) k += 128 >> l;
      }
      kraft[v] = (uint16_t)k;
    }
  }
};
const ScanTables SCAN;

// the first start of a block in data[from, to), or NONE
static uint64_t find_block_start(Inflater &in, size_t from, size_t to, const uint8_t *data, size_t size) {
  const size_t lim = std::min(to, size >= 32 ? size - 32 : 0);
  for (size_t byte = from; byte < lim; byte++) {
    const uint64_t w0 = load64(data + byte);
    uint32_t m = SCAN.cand[w0 & 1023];
    while (m) {
      const int s = __builtin_ctz(m);
      m &= m - 1;
      const uint64_t w = w0 >> s;
      if (((w >> 3) & 31) > 29 || ((w >> 8) & 31) > 29) continue;  // HLIT, HDIST
      // the code of the code lengths must be complete: HCLEN + 4 lengths of 3 bits
      const int n = (int)((w >> 13) & 15) + 4;
      const uint64_t at = (uint64_t)byte * 8 + (uint64_t)s + 17;
      const uint64_t pc = (load64(data + (at >> 3)) >> (at & 7)) & ((1ull << (3 * n)) - 1);
      const int kraft = SCAN.kraft[pc & 4095] + SCAN.kraft[(pc >> 12) & 4095] + SCAN.kraft[(pc >> 24) & 4095] +
                        SCAN.kraft[(pc >> 36) & 4095] + SCAN.kraft[(pc >> 48) & 4095];
      if (kraft != 128) continue;
      if (try_start(in, data, size, (uint64_t)byte * 8 + (uint64_t)s)) return (uint64_t)byte * 8 + (uint64_t)s;
    }
  }
  return NONE;
}

bool PgzStream::run_round(std::string &err) {
  const size_t first_byte = (size_t)(cur_bit_ >> 3);
  int T = sequential_ ? 1 : threads_;
  const size_t span = sequential_ ? chunk_bytes_ * (size_t)threads_ : chunk_bytes_;
  std::vector<std::unique_ptr<Chunk>> chunks;
  for (int t = 0; t < T; t++) {
    const size_t from = first_byte + (size_t)t * span;
    if (t && from >= size_) break;
    std::unique_ptr<Chunk> c(new Chunk());
    c->search_from = from;
    c->search_to = std::min(size_, from + span);
    if (!pool_.empty()) {
      c->inf.adopt(pool_.back().first, pool_.back().second);
      pool_.pop_back();
    } else if (!c->inf.alloc(span * 5)) {
      err = "out of memory";
      finished_ = true;
      return false;
    }
    chunks.push_back(std::move(c));
  }
  T = (int)chunks.size();
  const uint64_t round_end_bit = (uint64_t)chunks.back()->search_to * 8;
  const uint8_t *const data = data_;
  const size_t size = size_;

  auto work = [&](int t) {
    Chunk &c = *chunks[(size_t)t];
    Inflater &in = c.inf;
    if (t == 0) {
      for (uint32_t i = 0; i < CTX; i++) in.mem[i] = window_[i];
      in.start(data, data + size, cur_bit_);
      in.out = in.mem + CTX;
      in.out_min = in.out - window_valid_;
      c.start_bit.store(cur_bit_, std::memory_order_release);
    } else {
      for (uint32_t i = 0; i < CTX; i++) in.mem[i] = (uint16_t)(256 + i);
      const uint64_t sb = find_block_start(in, c.search_from, c.search_to, data, size);
      c.start_bit.store(sb, std::memory_order_release);
      if (sb == NONE) {
        c.status = -1;
        return;
      }
    }
    in.limit = 0;
    int u = t + 1;
    for (;;) {
      const uint64_t pos = in.bitpos();
      bool stop = false;
      while (u < T) {
        Chunk &nx = *chunks[(size_t)u];
        if (pos < (uint64_t)nx.search_from * 8) break;
        uint64_t sb;
        while ((sb = nx.start_bit.load(std::memory_order_acquire)) == PENDING) std::this_thread::yield();
        if (sb == NONE || sb < pos) {
          u++;
          continue;
        }
        if (sb == pos) stop = true;
        break;
      }
      if (!stop && u >= T && pos >= round_end_bit) stop = true;
      if (!stop && in.produced() >= OUT_CAP) {
        stop = true;
        u = T;  // (whatever the later chunks did is dropped)
      }
      if (stop) {
        c.synced_with = u;
        c.status = 0;
        break;
      }
      bool fin = false;
      const int rc = in.block(&fin);
      if (rc != IR_OK || in.past_end()) {
        c.status = -1;
        c.err = (rc == IR_ERR) ? "corrupt gzip file (invalid deflate data)" : "gzip file ends inside the compressed stream";
        break;
      }
      if (!fin) continue;
      // the member's trailer, then another member, the end of the file, or bytes that are not gzip (ignored, as
      // zlib's gzread does)
      in.align_byte();
      MemberEnd me;
      me.pos = in.produced();
      me.crc = in.take(32);
      me.isize = in.take(32);
      if (in.past_end()) {
        c.status = -1;
        c.err = "gzip file ends inside the compressed stream";
        break;
      }
      c.ends.push_back(me);
      in.out_min = in.out;
      const size_t at = (size_t)(in.bitpos() >> 3);
      const long hl = gzip_header_len(data + at, size - at);
      if (hl == 0) {
        c.status = 1;
        break;
      }
      if (hl < 0) {
        c.status = -1;
        c.err = "corrupt gzip file (bad member header)";
        break;
      }
      in.start(data, data + size, (uint64_t)(at + (size_t)hl) * 8);
    }
    c.n_out = in.produced();
    c.end_bit = in.bitpos();
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
  }
  // the chunks whose starts were met, in order
  std::vector<int> chain{0};
  for (;;) {
    const Chunk &c = *chunks[(size_t)chain.back()];
    if (c.status != 0 || c.synced_with >= T) break;
    chain.push_back(c.synced_with);
  }
  confirmed_ += chain.size() - 1;
  for (int t = 1; t < T; t++)
    if (chunks[(size_t)t]->start_bit.load() != NONE && std::find(chain.begin(), chain.end(), t) == chain.end()) dropped_++;
  if (T > 1 && chain.size() == 1 && chunks[0]->status == 0 && ++failed_rounds_ >= 2) sequential_ = true;  // (binary data, stored blocks ...)
  // the text in front of every chunk of the chain: from the chunk before it
  size_t valid = window_valid_;
  unsigned char ctx[CTX];
  memcpy(ctx, window_.data(), CTX);
  queue_.clear();
  for (size_t k = 0; k < chain.size(); k++) {
    std::unique_ptr<Chunk> &c = chunks[(size_t)chain[k]];
    for (int i = 0; i < 256; i++) c->lut[i] = (unsigned char)i;
    memcpy(c->lut + 256, ctx, CTX);
    const uint16_t *sym = c->inf.mem + CTX;
    const size_t n = c->n_out;
    if (n >= CTX) {
      for (size_t j = 0; j < CTX; j++) ctx[j] = c->lut[sym[n - CTX + j]];
    } else {
      memmove(ctx, ctx + n, CTX - n);
      for (size_t j = 0; j < n; j++) ctx[CTX - n + j] = c->lut[sym[j]];
    }
    valid = c->ends.empty() ? valid + n : n - c->ends.back().pos;
    if (valid > CTX) valid = CTX;
  }
  const Chunk &last = *chunks[(size_t)chain.back()];
  if (last.status == 0) {
    cur_bit_ = last.end_bit;
  } else {
    finished_ = true;
    if (last.status < 0) pending_err_ = last.err.empty() ? "corrupt gzip file" : last.err;
  }
  memcpy(window_.data(), ctx, CTX);
  window_valid_ = valid;
  for (int t : chain) queue_.push_back(std::move(chunks[(size_t)t]));
  for (auto &c : chunks) recycle(c);  // (the dropped ones)
  q_chunk_ = q_off_ = q_end_ = 0;
  return true;
}

bool PgzStream::next_impl(std::vector<unsigned char> *outv, unsigned char *dst, size_t dst_cap, size_t *n_out, std::string &err) {
  struct Piece {
    size_t chunk, off, n, at;
    const MemberEnd *end;  // a member ends in front of this (empty) piece
    uint32_t crc;
  };
  const size_t want = outv ? WINDOW : std::min(WINDOW, dst_cap);
  for (;;) {
    if (q_chunk_ >= queue_.size()) {
      if (finished_) {
        if (!pending_err_.empty()) err = pending_err_;
        return false;
      }
      if (!run_round(err)) return false;
      continue;
    }
    std::vector<Piece> pieces;
    size_t total = 0, ch = q_chunk_, off = q_off_, ei = q_end_;
    while (ch < queue_.size() && total < want) {
      const Chunk &c = *queue_[ch];
      if (ei < c.ends.size() && c.ends[ei].pos == off) {
        pieces.push_back(Piece{ch, off, 0, total, &c.ends[ei], 0});
        ei++;
        continue;
      }
      if (off == c.n_out) {
        ch++;
        off = 0;
        ei = 0;
        continue;
      }
      size_t n = (ei < c.ends.size() ? c.ends[ei].pos : c.n_out) - off;
      n = std::min(n, std::min(PIECE, want - total));
      pieces.push_back(Piece{ch, off, n, total, nullptr, 0});
      off += n;
      total += n;
    }
    if (outv) {
      outv->resize(total);
      dst = outv->data();
    }
    std::atomic<size_t> next_piece{0};
    auto work = [&]() {
      for (;;) {
        const size_t i = next_piece.fetch_add(1);
        if (i >= pieces.size()) break;
        Piece &pc = pieces[i];
        if (!pc.n) continue;
        const Chunk &c = *queue_[pc.chunk];
        const uint16_t *s = c.inf.mem + CTX + pc.off;
        unsigned char *d = dst + pc.at;
        translate(s, d, pc.n, c.lut);
        pc.crc = fast_crc32(0, d, pc.n);
      }
    };
    {
      const int n_thr = (int)std::min<size_t>((size_t)threads_, (total + PIECE - 1) / PIECE);
      std::vector<std::thread> pool;
      for (int t = 1; t < n_thr; t++) pool.emplace_back(work);
      work();
      for (auto &th : pool) th.join();
    }
    for (const Piece &pc : pieces) {
      if (pc.end) {
        if (crc_run_ != pc.end->crc || len_run_ != pc.end->isize) {
          err = "corrupt gzip file (CRC or length of a member does not match its text)";
          finished_ = true;
          pending_err_.clear();
          queue_.clear();
          q_chunk_ = 0;
          return false;
        }
        crc_run_ = len_run_ = 0;
      } else {
        crc_run_ = (uint32_t)crc32_combine(crc_run_, pc.crc, (z_off_t)pc.n);
        len_run_ += (uint32_t)pc.n;
      }
    }
    for (size_t k = q_chunk_; k < ch && k < queue_.size(); k++) recycle(queue_[k]);  // handed out in full
    q_chunk_ = ch;
    q_off_ = off;
    q_end_ = ei;
    if (!total) continue;  // (only member ends: an empty member, or the end of the round)
    if (n_out) *n_out = total;
    return true;
  }
}

}  // namespace lasv
