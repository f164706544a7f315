// table.cuh -- device side of the CORTEX-style open hash (2^L buckets x W slots):
// lock-free find-or-insert with coverage / edge accumulation.
//
// Replaces, on the device (reference file:line under /root/reference):
//   hash_table_find_or_insert / find_in_bucket  src/hash_table/open_hash/hash_table.c:69-122,270-327
//   element_initialise                          src/dB_graph_operations/element.c:339-362
//   db_node_update_coverage (saturating u32)    src/dB_graph_operations/element.c:395-417
//   add_edges (OR)                              src/dB_graph_operations/element.c:223-228
//
// Same bucket function and rehash chain as the reference (bucket r of a key is
// lookup3(key + r) & (2^L - 1), r = 0..max_rehash, a key moves to bucket r+1
// only when bucket r holds W keys), so capacity behaviour ("table full" is
// fatal) and the set of keys a bucket can hold are the reference's.  What
// differs is only the arrangement INSIDE a bucket while it lives on the device
// (kmer_core.cuh, "line layout"): a bucket is NL lines of 128 bytes, each an
// 8-byte header of tag bytes followed by G reference-layout slots.  A key starts
// at a hashed line and walks the lines of its bucket with wrap-around; inside a
// line the header tells which slots can hold it.  A steady-state lookup is one
// header read plus one slot read in the SAME line (the reference scans from
// slot 0: ~W*fill/2 slot reads).  lasv_graph_finalize() rewrites every bucket
// as W hole-free consecutive Elements, which is what hash_table_find expects.
//
// Concurrency protocol per slot (meta = the 8-byte word after the key words):
//   EMPTY  --CAS(meta: 0 -> {covg, edges, LOCKED, tag16})-->  LOCKED
//   LOCKED --key words stored, __threadfence, tag byte stored, status ^= LOCKED^NONE-->  NONE
// The tag byte is written AFTER the key words are visible: whoever reaches a
// slot through a non-zero tag byte may read its key words at once, without
// looking at the status.  (Why that holds: header and slot loads are strong GPU-scope
// loads served by L2, the point of coherence; the writer's fence makes the key stores
// reach L2 before the tag store; the reader cannot ISSUE the slot load before the
// header load has returned -- the slot address is computed from the header -- so the
// slot load reaches L2 after the tag, hence after the key words.  An acquire load on
// the header would say the same in memory-model terms but compiles to CCTL.IVALL,
// which throws away the L1 lines the record loads live in.)
// A slot reached through a zero tag byte is claimed by
// CAS; a failed CAS means somebody else got there first (tag byte not visible
// yet) and the slot is examined through its meta word: 16-bit tag differs ->
// other key; LOCKED -> wait for publication; then compare the key words.
// Slots are claimed in header order and never released, hence the invariant:
// a key sits behind claimed slots only, and every prober meets it before it
// meets a claimable slot.
// Updates of a slot are independent 32-bit L2 reductions: add on the coverage
// word, OR on the edges word (only when it adds a bit).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include "kmer_core.cuh"

namespace lasv {

struct TableView {
  uint8_t *lines;        // n_buckets * NL lines of 128 bytes (16-byte aligned at least)
  uint32_t bucket_mask;  // n_buckets - 1
  uint32_t W;            // slots per bucket (the reference's bucket_size)
  uint32_t NL;           // lines per bucket = ceil(W / G)
  uint32_t last_slots;   // usable slots of a bucket's last line = W - (NL-1)*G
  uint32_t max_rehash;   // 15 in laSV (graph_operations.c:640)
  uint32_t finalized;    // != 0: every bucket is W consecutive Elements at the start of its NL lines
};

LASV_HD TableView make_table_view(uint8_t *lines, int N, int L, int W, int max_rehash, bool finalized) {
  TableView t;
  t.lines = lines;
  t.bucket_mask = (uint32_t)((1ull << L) - 1);
  t.W = (uint32_t)W;
  t.NL = lines_per_bucket(N, (uint32_t)W);
  t.last_slots = (uint32_t)W - (t.NL - 1) * slots_per_line(N);
  t.max_rehash = (uint32_t)max_rehash;
  t.finalized = finalized ? 1u : 0u;
  return t;
}
LASV_HD uint64_t table_bytes(const TableView &t) {
  return ((uint64_t)t.bucket_mask + 1) * t.NL * LINE_BYTES;
}
LASV_HD uint8_t *bucket_base(const TableView &t, uint32_t bucket) {
  return t.lines + (uint64_t)bucket * t.NL * LINE_BYTES;
}
// slot s (0 <= s < W) of a bucket, in whichever arrangement the table is in
template <int N>
LASV_HD uint64_t *slot_ptr(const TableView &t, uint32_t bucket, uint32_t s) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  uint8_t *b = bucket_base(t, bucket);
  if (t.finalized) return reinterpret_cast<uint64_t *>(b + (uint64_t)s * SLOT);
  const uint32_t line = s / G, i = s - line * G;
  return reinterpret_cast<uint64_t *>(b + (uint64_t)line * LINE_BYTES + 8u + i * SLOT);
}

// Per-graph device counters (touched once per warp).
struct GraphCounters {
  unsigned long long unique_kmers;     // hash_table.c:305
  unsigned long long kmers_inserted;   // number of find_or_insert calls
  unsigned long long table_full;       // != 0: some key found 16 full buckets (reference die()s)
  unsigned long long rehash_hist[16];  // collisions[r], hash_table.c:325
  unsigned long long covg_saturated;   // element.c:402-415 warning condition
};

// ---------------------------------------------------------------------------
// Memory operations of the table protocol.  The kernels use DeviceOps (strong
// GPU-scope loads/stores that bypass L1, L2 atomics).  tests/host_emul supplies
// a sequential HostOps with the same interface so the probing / claiming /
// finalising LOGIC is exercised on the CPU; the product never instantiates it.
#if defined(__CUDACC__)
// L2::64B: a miss fetches one 64-byte DRAM atom instead of the default whole 128-byte
// line (tools/ubench "gran": 64 vs 128 bytes of DRAM read per random 8-byte load) --
// the table is touched one slot at a time, the other half of the line would be waste.
__device__ __forceinline__ uint64_t ld_strong_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ld_strong_v2u64(const uint64_t *p, uint64_t &a, uint64_t &b) {
  asm volatile("ld.relaxed.gpu.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_strong_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct DeviceOps {
  static __device__ __forceinline__ uint64_t load(const uint64_t *p) { return ld_strong_u64(p); }
  static __device__ __forceinline__ void load16(const uint64_t *p, uint64_t &a, uint64_t &b) {  // p 16-byte aligned
    ld_strong_v2u64(p, a, b);
  }
  static __device__ __forceinline__ void store(uint64_t *p, uint64_t v) { st_strong_u64(p, v); }
  static __device__ __forceinline__ void store_tag(uint8_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
  static __device__ __forceinline__ uint64_t cas(uint64_t *p, uint64_t expect, uint64_t want) {
    return atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)expect,
                     (unsigned long long)want);
  }
  static __device__ __forceinline__ void xor32(uint32_t *p, uint32_t v) { atomicXor(p, v); }
  static __device__ __forceinline__ void add32(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
  static __device__ __forceinline__ void or32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
  static __device__ __forceinline__ void and32(uint32_t *p, uint32_t v) { atomicAnd(p, v); }
  static __device__ __forceinline__ void fence() { __threadfence(); }
  static __device__ __forceinline__ void backoff() { __nanosleep(40); }
  static __device__ __forceinline__ void count(unsigned long long *p, unsigned long long v) {
    atomicAdd(p, v);
  }
  static __device__ __forceinline__ void flag(unsigned long long *p) { atomicExch(p, 1ull); }
};
#endif

// Saturating coverage add (slow path, only near UINT_MAX; element.c:395-417).
#pragma nv_exec_check_disable
template <class Ops>
LASV_HD void covg_add_saturating(uint64_t *meta_ptr, uint32_t add, GraphCounters *ctr) {
  uint64_t old = Ops::load(meta_ptr);
  for (;;) {
    const uint32_t c = (uint32_t)old;
    const uint32_t nc = (0xFFFFFFFFu - add >= c) ? c + add : 0xFFFFFFFFu;
    const uint64_t want = (old & 0xFFFFFFFF00000000ull) | nc;
    const uint64_t got = Ops::cas(meta_ptr, old, want);
    if (got == old) {
      if (nc == 0xFFFFFFFFu) Ops::count(&ctr->covg_saturated, 1ull);
      return;
    }
    old = got;
  }
}

// coverage += covg_add (saturating), edges |= edge, on a slot known to hold the
// key; m is a recent copy of its meta word.
#pragma nv_exec_check_disable
template <class Ops>
LASV_HD void slot_accumulate(uint64_t *meta_ptr, uint64_t m, uint32_t edge, uint32_t covg_add,
                             GraphCounters *ctr) {
  if (meta_covg(m) < 0xF0000000u && covg_add < 0x01000000u) {
    Ops::add32(reinterpret_cast<uint32_t *>(meta_ptr), covg_add);
  } else {
    covg_add_saturating<Ops>(meta_ptr, covg_add, ctr);
  }
  if ((meta_edges(m) | edge) != meta_edges(m)) {
    Ops::or32(reinterpret_cast<uint32_t *>(meta_ptr) + 1, edge & 0xFFu);
  }
}

// The N+1 words of a slot (key words, then meta) with as few memory requests as
// possible: 16-byte pieces from the enclosing 16-byte boundary on (scattered
// requests cost per request -- address translation -- not per byte).  Slots start
// 8 bytes past a 16-byte boundary or on one.  The pieces are independent loads:
// only for slots whose key words are known to be visible.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void slot_load_words(const uint64_t *slot, uint64_t (&w)[N + 1]) {
  constexpr int PIECES = (8 * N + 8 + 8 + 15) / 16;  // window of SLOT + 8 bytes
  const bool odd = (reinterpret_cast<uintptr_t>(slot) & 8u) != 0;
  const uint64_t *p16 = reinterpret_cast<const uint64_t *>(reinterpret_cast<uintptr_t>(slot) & ~(uintptr_t)15);
  uint64_t x[2 * PIECES];
#pragma unroll
  for (int i = 0; i < PIECES; i++) {
    // a slot on a 16-byte boundary does not reach into the last piece of the window
    if (i < PIECES - 1 || odd || (8 * N + 8) > 16 * (PIECES - 1)) Ops::load16(p16 + 2 * i, x[2 * i], x[2 * i + 1]);
    else { x[2 * i] = 0; x[2 * i + 1] = 0; }
  }
#pragma unroll
  for (int i = 0; i <= N; i++) w[i] = odd ? x[i + 1] : x[i];
}

// A slot reached through a visible tag byte: its key words are readable.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool slot_hit_tagged(uint64_t *slot, GraphCounters *ctr, const uint64_t (&key)[N],
                             uint32_t edge, uint32_t covg_add) {
  uint64_t sw[N + 1];
  slot_load_words<N, Ops>(slot, sw);
  bool same = true;
#pragma unroll
  for (int w = 0; w < N; w++) same &= (sw[w] == key[w]);
  if (!same) return false;
  slot_accumulate<Ops>(slot + N, sw[N], edge, covg_add, ctr);
  return true;
}

// A slot somebody else claimed before our CAS (m = its meta word, non-zero).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool slot_hit_claimed(uint64_t *slot, uint64_t m, const Probe &pr, GraphCounters *ctr,
                              const uint64_t (&key)[N], uint32_t edge, uint32_t covg_add) {
  uint64_t *meta_ptr = slot + N;
  if (meta_tag(m) != pr.tag16) return false;
  while (meta_status(m) == ST_LOCKED) {  // same tag, key words not published yet
    Ops::backoff();
    m = Ops::load(meta_ptr);
  }
  Ops::fence();
  bool same = true;
#pragma unroll
  for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
  if (!same) return false;
  slot_accumulate<Ops>(meta_ptr, m, edge, covg_add, ctr);
  return true;
}

// find-or-insert `key` in bucket `bucket` only (one step of the rehash chain).
// Returns 1 if the key was new, 0 if it existed, -1 if the bucket holds W other keys.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int bucket_upsert(const TableView &t, GraphCounters *ctr, uint32_t bucket, const Probe &pr,
                          const uint64_t (&key)[N], uint32_t edge, uint32_t covg_add) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  const uint64_t full_valid = header_valid_mask(G);
  const uint64_t last_valid = header_valid_mask(t.last_slots);
  const uint64_t rep = (uint64_t)pr.tag8 * 0x0101010101010101ull;
  uint8_t *base = bucket_base(t, bucket);
  uint32_t line = mulhi32(pr.h, t.NL);
#pragma unroll 1
  for (uint32_t step = 0; step < t.NL; ++step) {
    uint8_t *lp = base + (uint64_t)line * LINE_BYTES;
    const uint64_t hdr = Ops::load(reinterpret_cast<uint64_t *>(lp));
    const uint64_t valid = (line == t.NL - 1) ? last_valid : full_valid;
    uint64_t cand = zero_bytes64(hdr ^ rep) & valid;
    uint64_t empt = zero_bytes64(hdr) & valid;
    while (cand) {  // visible tags equal to ours, in slot order
      const uint32_t i = first_byte_index(cand);
      cand &= cand - 1;
      uint64_t *slot = reinterpret_cast<uint64_t *>(lp + 8u + i * SLOT);
      if (slot_hit_tagged<N, Ops>(slot, ctr, key, edge, covg_add)) return 0;
    }
    while (empt) {  // never-claimed slots (as far as the header told), in slot order
      const uint32_t i = first_byte_index(empt);
      empt &= empt - 1;
      uint64_t *slot = reinterpret_cast<uint64_t *>(lp + 8u + i * SLOT);
      const uint64_t want = meta_pack(covg_add, edge, ST_LOCKED, pr.tag16);
      const uint64_t m = Ops::cas(slot + N, 0ull, want);
      if (m == 0ull) {
#pragma unroll
        for (int w = 0; w < N; w++) Ops::store(slot + w, key[w]);
        Ops::fence();
        Ops::store_tag(lp + i, pr.tag8);
        Ops::xor32(reinterpret_cast<uint32_t *>(slot + N) + 1, (ST_LOCKED ^ ST_NONE) << 8);
        return 1;
      }
      if (slot_hit_claimed<N, Ops>(slot, m, pr, ctr, key, edge, covg_add)) return 0;
    }
    line = (line + 1 == t.NL) ? 0u : line + 1;  // line full without the key
  }
  return -1;
}

// find-or-insert `key`, then coverage += covg_add (saturating) and edges |= edge.
// c0 = lookup3(key, 0).c when the caller has it already (have_c0), so that the
// records path hashes once.  Returns 1 if the key was new, 0 if it existed, -1
// if the table is full.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int table_upsert_hashed(const TableView &t, GraphCounters *ctr, const uint64_t (&key)[N],
                                uint32_t c0, uint32_t edge, uint32_t covg_add) {
  const Probe pr = probe_of_key<N>(key);
  uint32_t c = c0;
  for (uint32_t r = 0;;) {
    const int res = bucket_upsert<N, Ops>(t, ctr, c & t.bucket_mask, pr, key, edge, covg_add);
    if (res >= 0) {
      if (r) Ops::count(&ctr->rehash_hist[r], 1ull);
      return res;
    }
    if (++r > t.max_rehash) break;
    c = lookup3_key<N>(key, r).c;  // bucket r-1 is full and does not hold the key
  }
  Ops::flag(&ctr->table_full);
  return -1;
}

#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int table_upsert(const TableView &t, GraphCounters *ctr, const uint64_t (&key)[N],
                         uint32_t edge, uint32_t covg_add) {
  return table_upsert_hashed<N, Ops>(t, ctr, key, lookup3_key<N>(key, 0).c, edge, covg_add);
}

// hash_table_find (hash_table.c:234-267) on the un-finalised device layout, with
// no insert in flight.  Returns true and the meta word if present.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool table_find(const TableView &t, const uint64_t (&key)[N], uint64_t &meta_out, uint64_t **slot_out = nullptr) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  const Probe pr = probe_of_key<N>(key);
  const uint64_t rep = (uint64_t)pr.tag8 * 0x0101010101010101ull;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    uint8_t *base = bucket_base(t, lookup3_key<N>(key, r).c & t.bucket_mask);
    uint32_t line = mulhi32(pr.h, t.NL);
    for (uint32_t step = 0; step < t.NL; ++step) {
      uint8_t *lp = base + (uint64_t)line * LINE_BYTES;
      const uint64_t hdr = Ops::load(reinterpret_cast<uint64_t *>(lp));
      const uint64_t valid = header_valid_mask((line == t.NL - 1) ? t.last_slots : G);
      uint64_t cand = zero_bytes64(hdr ^ rep) & valid;
      while (cand) {
        const uint32_t i = first_byte_index(cand);
        cand &= cand - 1;
        uint64_t *slot = reinterpret_cast<uint64_t *>(lp + 8u + i * SLOT);
        bool same = true;
#pragma unroll
        for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
        if (same) {
          meta_out = Ops::load(slot + N);
          if (slot_out) *slot_out = slot;
          return true;
        }
      }
      if (zero_bytes64(hdr) & valid) return false;  // a claimable slot ends the probe
      line = (line + 1 == t.NL) ? 0u : line + 1;
    }
    // bucket full without the key: next bucket of the chain
  }
  return false;
}

// hash_table_find on the FINALISED (reference) layout: scan each bucket of the
// rehash chain from slot 0 to the first empty slot (hash_table.c:84-117,247-264).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool table_find_finalized(const TableView &t, const uint64_t (&key)[N], uint64_t &meta_out) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    uint8_t *bucket = bucket_base(t, lookup3_key<N>(key, r).c & t.bucket_mask);
    for (uint32_t s = 0; s < t.W; ++s) {
      uint64_t *slot = reinterpret_cast<uint64_t *>(bucket + (uint64_t)s * SLOT);
      const uint64_t m = Ops::load(slot + N);
      if (meta_status(m) == ST_EMPTY) return false;  // hole-free buckets: not present
      bool same = true;
#pragma unroll
      for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
      if (same) { meta_out = m; return true; }
    }
    // bucket full and key absent: next bucket of the chain
  }
  return false;
}

}  // namespace lasv
