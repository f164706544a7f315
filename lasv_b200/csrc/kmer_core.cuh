// kmer_core.cuh -- host/device arithmetic shared by every kernel of the graph
// build: base classification + 2-bit packing, window validity, O(1) k-mer /
// reverse-complement extraction from a packed tile, canonical key, per-occurrence
// edge byte, and Bob Jenkins lookup3 over 1/2/3-word keys.
//
// Everything here is a pure function (no memory side effects beyond its
// arguments) marked __host__ __device__, so tests/host_emul can run exactly the
// code the kernels run, on the CPU, against the oracle.
//
// Reference semantics restated (file:line under /root/reference):
//   base codes A0 C1 G2 T3 ............ include/basic/event_encoding/base_encoding/event_encoding.h:37-44
//   "bad" base (non-ACGT or qual<=cutoff) src/dB_graph_operations/file_reader.c:120-178,389
//   BinaryKmer word order (word 0 = MSW)  include/basic/binary_kmer.h:41-47, src/basic/binary_kmer.c:115-143
//   reverse complement .................. src/basic/binary_kmer.c:456-523
//   canonical key = min(kmer, revcomp) .. src/dB_graph_operations/element.c:308-336, binary_kmer.c:66-97
//   orientation / edge bits ............. src/dB_graph_operations/element.c:470-560
//   lookup3 hashlittle(key, 8N, 10) ..... src/hash_table/hash_key/bob_jenkins/hash_value.c:114-158,282-449,767-773
//   rehash = add r to the last word ..... src/hash_table/open_hash/hash_table.c:73-78
#pragma once
#include <stdint.h>
#include <vector_types.h>

#if defined(__CUDACC__)
#define LASV_HD __host__ __device__ __forceinline__
#else
#define LASV_HD inline
#endif

namespace lasv {

// ---------------------------------------------------------------- intrinsics
LASV_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
LASV_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __brev(x);
#else
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
#endif
}
LASV_HD uint32_t rot32(uint32_t x, int k) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(x, x, k);
#else
  return (x << k) | (x >> (32 - k));
#endif
}
LASV_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

// ------------------------------------------------------------------- lookup3
struct Hash3 {
  uint32_t a, b, c;  // c is the reference's hashlittle() value
};

#define LASV_MIX(a, b, c)                 \
  {                                       \
    a -= c; a ^= rot32(c, 4);  c += b;    \
    b -= a; b ^= rot32(a, 6);  a += c;    \
    c -= b; c ^= rot32(b, 8);  b += a;    \
    a -= c; a ^= rot32(c, 16); c += b;    \
    b -= a; b ^= rot32(a, 19); a += c;    \
    c -= b; c ^= rot32(b, 4);  b += a;    \
  }
#define LASV_FINAL(a, b, c)               \
  {                                       \
    c ^= b; c -= rot32(b, 14);            \
    a ^= c; a -= rot32(c, 11);            \
    b ^= a; b -= rot32(a, 25);            \
    c ^= b; c -= rot32(b, 16);            \
    a ^= c; a -= rot32(c, 4);             \
    b ^= a; b -= rot32(a, 14);            \
    c ^= b; c -= rot32(b, 24);            \
  }

// hashlittle over the 8N key bytes in memory order (word 0 first, each word
// little-endian), initval 10, with `rehash` added to the last word first.
// Returns the full final (a,b,c) state: c is what the reference uses for the
// bucket; a and b are spare, equally mixed bits (used for the in-bucket start
// slot and the slot tag -- device-layout only, invisible in any output).
template <int N>
LASV_HD Hash3 lookup3_key(const uint64_t (&key)[N], uint32_t rehash) {
  uint32_t w[2 * N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t v = key[i] + ((i == N - 1) ? (uint64_t)rehash : 0ull);
    w[2 * i] = (uint32_t)v;
    w[2 * i + 1] = (uint32_t)(v >> 32);
  }
  uint32_t a, b, c;
  a = b = c = 0xdeadbeefu + (uint32_t)(8 * N) + 10u;
  if (N == 1) {
    a += w[0]; b += w[1];
  } else if (N == 2) {
    a += w[0]; b += w[1]; c += w[2];
    LASV_MIX(a, b, c);
    a += w[3];
  } else {
    a += w[0]; b += w[1]; c += w[2];
    LASV_MIX(a, b, c);
    a += w[3]; b += w[4 % (2 * N)]; c += w[5 % (2 * N)];
  }
  LASV_FINAL(a, b, c);
  Hash3 h;
  h.a = a; h.b = b; h.c = c;
  return h;
}

// ------------------------------------------------- base classification (16 B)
// Classify 16 consecutive bytes of the sequence stream (and the 16 matching
// quality bytes).  Out:
//   code32 : 2-bit codes, FIRST base in bits 31:30 ... 16th base in bits 1:0
//   bad16  : bit 15 = first base ... bit 0 = 16th base; 1 = base breaks k-mers
// A base is bad iff it is not one of ACGTacgt, or has_qual and the RAW quality
// byte, compared as signed char, is <= cutoff (file_reader.c:150,389), or its
// byte lies outside the stream (inrange16 bit clear; bit i <-> byte i).
// Read separators ('\n') are ordinary non-ACGT bytes, hence bad: no window ever
// spans two reads.
LASV_HD void classify16(uint4 sv, uint4 qv, bool has_qual, int cutoff, uint32_t inrange16,
                        uint32_t &code32, uint32_t &bad16) {
  const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
  const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
  uint32_t code = 0, bad = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    uint32_t ch = (sw[i >> 2] >> (8 * (i & 3))) & 0xFFu;
    uint32_t u = ch & 0xDFu;  // fold lower case onto upper case (clears bit 5 only)
    bool isb = (u == 0x41u) | (u == 0x43u) | (u == 0x47u) | (u == 0x54u);
    uint32_t c2 = ((u >> 1) & 3u) ^ ((u >> 2) & 1u);  // A0 C1 G2 T3
    int q = (int)(int8_t)((qw[i >> 2] >> (8 * (i & 3))) & 0xFFu);
    bool lowq = has_qual & (q <= cutoff);
    bool in = (inrange16 >> i) & 1u;
    bool isbad = (!isb) | lowq | (!in);
    code |= (isb ? c2 : 0u) << (30 - 2 * i);
    bad |= (uint32_t)isbad << (15 - i);
  }
  code32 = code;
  bad16 = bad;
}

// Which of the 16 bytes of the chunk at byte offset g (relative to the 16-byte
// aligned base of the stream buffer) belong to the stream [lead, n_total):
// bit i <-> byte g+i.  `lead` (< 16) is the misalignment of the caller's pointer.
LASV_HD uint32_t chunk_inrange16(long long g, uint32_t lead, uint64_t n_total) {
  if (g < 0 || (uint64_t)g >= n_total) return 0u;
  const uint64_t left = n_total - (uint64_t)g;
  uint32_t m = left >= 16 ? 0xFFFFu : ((1u << left) - 1u);
  if ((uint64_t)g < lead) m &= ~((1u << (lead - (uint32_t)g)) - 1u);
  return m;
}

// Reverse-complement of a 16-base code word: complement every base (3-c == ~c
// on two bits) and reverse the base order.
LASV_HD uint32_t rc_code32(uint32_t code32) {
  uint32_t y = brev32(~code32);
  return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}

// ------------------------------------------------------------ tile geometry
// A tile is a window of the byte stream packed into shared memory (or a host
// array in the emulation).  Region base index j in [0, R): region position 0 is
// stream position t0 - HALO_L.  Three big-endian bit strings of 32-bit words:
//   P  forward 2-bit codes, 16 bases per word (base 16w in bits 31:30), PADW
//      pad words in front
//   PR reverse-complement string (RC position p <-> region R-1-p), same layout
//   BD bad bits, 32 positions per word (position 32w in bit 31)
// and, derived from BD once per tile,
//   VD valid-window bits: bit j set <=> the k bases from j on are all good.
// 32-bit words on purpose: a k-mer word is ONE funnel shift of two neighbouring
// string words (SHF), where 64-bit variable shifts cost four instructions each.
struct TileGeom {
  static constexpr int HALO_L = 16;   // bytes before the first window start (need >= 1)
  static constexpr int HALO_R = 112;  // bytes after the last window start   (need >= k+1, k<=95)
  static constexpr int PADW = 2;      // front pad words of P / PR
};

LASV_HD uint32_t funnel_l(uint32_t hi, uint32_t lo, uint32_t sh) {  // upper word of (hi:lo) << sh, 0<=sh<32
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, sh);
#else
  return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
#endif
}
LASV_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

LASV_HD bool base_bad(const uint32_t *BD, int j) { return (BD[j >> 5] >> (31 - (j & 31))) & 1u; }
LASV_HD uint32_t base_code(const uint32_t *P, int j) {
  return (P[TileGeom::PADW + (j >> 4)] >> (30 - 2 * (j & 15))) & 3u;
}

// 128-bit big-endian window (w[0] first) shifted left by s bits (0 <= s < 96),
// zeros shifted in behind.  No dynamically indexed arrays (they would live in
// local memory): the word shift is two conditional moves.
LASV_HD void shl128(const uint32_t (&w)[4], uint32_t s, uint32_t (&out)[4]) {
  uint32_t x0 = w[0], x1 = w[1], x2 = w[2], x3 = w[3];
  if (s & 64u) { x0 = x2; x1 = x3; x2 = 0u; x3 = 0u; }
  if (s & 32u) { x0 = x1; x1 = x2; x2 = x3; x3 = 0u; }
  const uint32_t bs = s & 31u;
  out[0] = funnel_l(x0, x1, bs);
  out[1] = funnel_l(x1, x2, bs);
  out[2] = funnel_l(x2, x3, bs);
  out[3] = funnel_l(x3, 0u, bs);
}

// Mask of the window-start positions of word `word` that belong to the tile
// proper, i.e. to region indices [HALO_L, HALO_L + T).
LASV_HD uint32_t tile_word_mask(int word, int T) {
  const int lo = TileGeom::HALO_L, hi = TileGeom::HALO_L + T;  // region index range
  const int b = 32 * word;
  if (b + 32 <= lo || b >= hi) return 0u;
  uint32_t m = ~0u;
  if (b < lo) m &= ~0u >> (lo - b);           // drop the first lo-b positions (top bits)
  if (b + 32 > hi) m &= ~0u << (b + 32 - hi); // drop the last positions (low bits)
  return m;
}

// Valid-window bits of the 32 positions [32*word, 32*word+32): position j is
// valid iff BD has no bad bit in [j, j+k), 1 <= k <= 95.  Word-parallel erosion:
// E_2n = E_n & (E_n << n), then E_k assembled from the binary digits of k.
// BD needs readable words up to index word+3 (pad words must be all-ones).
LASV_HD uint32_t valid_window_word(const uint32_t *BD, int word, int k) {
  uint32_t e[4] = {~BD[word], ~BD[word + 1], ~BD[word + 2], ~BD[word + 3]};  // E_1 = good bits
  uint32_t acc[4] = {~0u, ~0u, ~0u, ~0u};
  uint32_t off = 0;
  // digits of k from the lowest: consume E_(2^b) as it is produced
#pragma unroll
  for (int b = 0; b < 7; b++) {
    const uint32_t n = 1u << b;
    if ((uint32_t)k & n) {
      uint32_t t[4];
      shl128(e, off, t);
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] &= t[i];
      off += n;
    }
    if (b < 6) {
      uint32_t t[4];
      shl128(e, n, t);
#pragma unroll
      for (int i = 0; i < 4; i++) e[i] &= t[i];
    }
  }
  return acc[0];
}

// The k-mer whose first base is region index j, right-aligned in N 64-bit words
// with word 0 most significant (binary_kmer.h:41-47).  All 2N 32-bit halves come
// from 2N+1 consecutive string words through the same funnel shift.
template <int N>
LASV_HD void extract_kmer(const uint32_t *P, int j, int k, uint64_t (&out)[N]) {
  const int b0 = j + k - 32 * N;                 // first base of the top 32-bit half (may be < 0)
  const int w0 = (b0 >> 4) + TileGeom::PADW;     // arithmetic shift: floor
  const uint32_t sh = 2u * (uint32_t)(b0 & 15);
  uint32_t prev = P[w0];
#pragma unroll
  for (int t = 0; t < 2 * N; t++) {              // t = 0 is the most significant half
    const uint32_t next = P[w0 + t + 1];
    const uint32_t v = funnel_l(prev, next, sh);
    prev = next;
    if (t & 1) out[t >> 1] |= (uint64_t)v;
    else out[t >> 1] = (uint64_t)v << 32;
  }
  out[0] &= (~0ull) >> (64 - 2 * (k & 31));
}

template <int N>
LASV_HD bool kmer_less(const uint64_t (&a)[N], const uint64_t (&b)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (a[i] != b[i]) return a[i] < b[i];
  }
  return false;
}

template <int N>
LASV_HD bool kmer_equal(const uint64_t (&a)[N], const uint64_t (&b)[N]) {
  bool eq = true;
#pragma unroll
  for (int i = 0; i < N; i++) eq &= (a[i] == b[i]);
  return eq;
}

// One k-mer OCCURRENCE: the canonical key and the 8-bit edge contribution this
// occurrence makes to its node (element.c:519-560 evaluated on the read alone):
//   as target of (prev -> this):  bit (3 - f) of the nibble OPPOSITE to this
//       occurrence's orientation, f = base before the window;
//   as source of (this -> next):  bit b of the nibble OF this occurrence's
//       orientation, b = base after the window.
// Low nibble = forward edges, high nibble = reverse (element.c:501-513).
// prev/next exist iff the neighbouring window is good, i.e. (given this window
// is good) iff the base just before / just after it is good.
template <int N>
LASV_HD uint32_t kmer_occurrence(const uint32_t *P, const uint32_t *PR, const uint32_t *BD, int R,
                                 int j, int k, uint64_t (&key)[N]) {
  uint64_t fw[N], rc[N];
  extract_kmer<N>(P, j, k, fw);
  extract_kmer<N>(PR, R - j - k, k, rc);
  const bool rev = kmer_less<N>(rc, fw);  // k odd => fw != rc
#pragma unroll
  for (int i = 0; i < N; i++) key[i] = rev ? rc[i] : fw[i];
  const uint32_t o = rev ? 1u : 0u;
  uint32_t edge = 0;
  if (!base_bad(BD, j - 1)) edge |= (1u << (3u - base_code(P, j - 1))) << (4u * (1u - o));
  if (!base_bad(BD, j + k)) edge |= (1u << base_code(P, j + k)) << (4u * o);
  return edge;
}

// ------------------------------------------------------- device slot layout
// A slot has the reference Element layout (element.h:77-82): N key words, then
// one 8-byte word holding {uint32 coverage; int8 edges; int8 status; 2 pad}.
// While the table lives on the device the two pad bytes carry a 16-bit tag of
// the key (zeroed again when the table is finalised for the host).
constexpr uint32_t ST_EMPTY = 0;       // NodeStatus unassigned (element.h:57)
constexpr uint32_t ST_NONE = 1;        // NodeStatus none
constexpr uint32_t ST_PRUNED = 3;      // NodeStatus pruned (element.h:57-75): left by the cleaning, not part of the graph
constexpr uint32_t ST_LOCKED = 0x80;   // device-only: key words not yet published

LASV_HD uint64_t meta_pack(uint32_t covg, uint32_t edges, uint32_t status, uint32_t tag) {
  return (uint64_t)covg | ((uint64_t)(edges & 0xFFu) << 32) | ((uint64_t)(status & 0xFFu) << 40) |
         ((uint64_t)(tag & 0xFFFFu) << 48);
}
LASV_HD uint32_t meta_covg(uint64_t m) { return (uint32_t)m; }
LASV_HD uint32_t meta_edges(uint64_t m) { return (uint32_t)(m >> 32) & 0xFFu; }
LASV_HD uint32_t meta_status(uint64_t m) { return (uint32_t)(m >> 40) & 0xFFu; }
LASV_HD uint32_t meta_tag(uint64_t m) { return (uint32_t)(m >> 48); }

// ------------------------------------------------------------ line layout
// On the device a bucket is not W consecutive Elements but NL = ceil(W / G)
// LINES of 128 bytes (one L2 line, two 64-byte DRAM atoms):
//   bytes 0..7    header: one tag byte per slot of the line (0 = never claimed,
//                 else 8 hash bits of the key, 1..255); bytes >= G stay 0
//   bytes 8..     G slots of 8N+8 bytes in the reference Element layout
// so the tag probe and the slot it leads to share a line (mostly a 64-byte
// atom): one random DRAM location per lookup instead of two.  G = 7 / 5 / 3 for
// 1 / 2 / 3-word k-mers.  lasv_graph_finalize() rewrites every bucket, in place,
// as W hole-free consecutive Elements at the start of its NL*128 bytes.
constexpr uint32_t LINE_BYTES = 128;
template <int N>
struct LineLayout {
  static constexpr uint32_t SLOT = 8 * N + 8;
  static constexpr uint32_t G = (LINE_BYTES - 8) / SLOT;
};
LASV_HD uint32_t slots_per_line(int N) { return (LINE_BYTES - 8) / (8u * (uint32_t)N + 8u); }
LASV_HD uint32_t lines_per_bucket(int N, uint32_t W) {
  const uint32_t g = slots_per_line(N);
  return (W + g - 1) / g;
}
// header mask with 0x80 in the byte of every usable slot 0..n-1
LASV_HD uint64_t header_valid_mask(uint32_t n) {
  return n == 0 ? 0ull : 0x8080808080808080ull >> (8u * (8u - (n > 8u ? 8u : n)));
}

// Where a key probes inside its bucket: start line and tags.  Device layout only
// (invisible in any output), so a cheap multiplicative mix of the key words is
// enough; the bucket itself is the reference's lookup3 value (hash_value.c:767).
struct Probe {
  uint32_t h;      // 32 mixed bits: start line = mulhi(h, NL)
  uint32_t tag8;   // header tag, 1..255
  uint32_t tag16;  // kept in the slot's pad bytes while on the device
};
LASV_HD uint32_t fmix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
template <int N>
LASV_HD Probe probe_of_key(const uint64_t (&key)[N]) {
  constexpr uint32_t MA[6] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du};
  constexpr uint32_t MB[6] = {0xCC9E2D51u, 0x1B873593u, 0xE6546B64u, 0x38B34AE5u, 0xA1E38B93u, 0x7FEB352Du};
  uint32_t xa = 0x2545F491u, xb = 0x9E3779B9u;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t lo = (uint32_t)key[i], hi = (uint32_t)(key[i] >> 32);
    xa += lo * MA[2 * i] + hi * MA[2 * i + 1];
    xb += lo * MB[2 * i] + hi * MB[2 * i + 1];
  }
  const uint32_t ya = fmix32(xa ^ rot32(xb, 15));
  const uint32_t yb = fmix32(xb + ya);
  Probe p;
  p.h = ya;
  const uint32_t t = yb & 0xFFu;
  p.tag8 = t ? t : 0x55u;
  p.tag16 = yb >> 16;
  return p;
}

// bit 8i+7 of the result is set iff byte i of x is zero (exact, no carries across bytes)
LASV_HD uint64_t zero_bytes64(uint64_t x) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  const uint32_t zl = ~(((lo & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | lo | 0x7F7F7F7Fu);
  const uint32_t zh = ~(((hi & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | hi | 0x7F7F7F7Fu);
  return (uint64_t)zl | ((uint64_t)zh << 32);
}
LASV_HD uint32_t first_byte_index(uint64_t m) {  // m != 0, bits only at 8i+7
#if defined(__CUDA_ARCH__)
  return (uint32_t)(__ffsll((long long)m) - 1) >> 3;
#else
  return (uint32_t)__builtin_ctzll(m) >> 3;
#endif
}

// 64-bit mixer (splitmix64 finaliser) -- used by the order-independent table
// checksum and by the synthetic read generator.
LASV_HD uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Order-independent fingerprint of one node (key words, coverage, edges).
template <int N>
LASV_HD uint64_t node_fingerprint(const uint64_t (&key)[N], uint32_t covg, uint32_t edges) {
  uint64_t h = 0x243F6A8885A308D3ull;
#pragma unroll
  for (int i = 0; i < N; i++) h = mix64(h ^ key[i]);
  return mix64(h ^ ((uint64_t)covg | ((uint64_t)(edges & 0xFFu) << 32)));
}

}  // namespace lasv
