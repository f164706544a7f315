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
// only when bucket r is FULL), so capacity behaviour ("table full" is fatal) and
// the set of keys a bucket can hold are the reference's.  What differs is only
// the position INSIDE a bucket: instead of scanning from slot 0 (W/2 slot reads
// per lookup) a key starts at a hashed slot and probes linearly with wrap-around.
// lasv_graph_finalize() squeezes every bucket hole-free afterwards, which is
// exactly the layout hash_table_find expects.
//
// Probing does not read slots: a side array holds ONE TAG BYTE PER SLOT (0 =
// never claimed, else 8 hash bits of the key).  A lookup loads 16 tags with one
// 16-byte access (a 74 %-full bucket needs 2.4 probes on average, 16 tags cover
// the tail), compares them in-register, and touches only slots whose tag matches
// -- one or two 32-byte sectors per k-mer instead of the whole probe run, and no
// per-lane probe loop for the warp to diverge on.  The tag array is a HINT: the
// slot's meta word stays authoritative.
//
// Concurrency protocol per slot (meta = the 8-byte word after the key words):
//   EMPTY  --CAS(meta: 0 -> {covg=c, edges=e, LOCKED, tag16})-->  LOCKED
//   LOCKED --key words and tag byte stored, __threadfence, meta ^= LOCKED^NONE-->  NONE
// A reader that sees tag byte 0 claims the slot by CAS; if the CAS fails somebody
// else got there first (tag byte not visible yet) and the slot is examined like a
// tag match: 16-bit meta tag differs -> other key; LOCKED -> wait for publication;
// then compare key words.  A non-zero tag byte never changes (no deletions).
// Updates of a published slot are two independent L2 reductions: 32-bit add on
// the coverage word, 32-bit OR on the edges word (only when it adds a bit).
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif
#include "kmer_core.cuh"

namespace lasv {

struct TableView {
  uint8_t *slots;        // n_buckets * W slots of (8N+8) bytes
  uint8_t *tags;         // n_buckets * W tag bytes (+16 readable bytes of slack), 16-byte aligned
  uint32_t bucket_mask;  // n_buckets - 1
  uint32_t W;
  uint32_t max_rehash;   // 15 in laSV (graph_operations.c:640)
};

// Per-graph device counters (one cache line each would be nicer; they are only
// touched once per CTA).
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
__device__ __forceinline__ uint64_t ld_strong_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_strong_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct DeviceOps {
  static __device__ __forceinline__ uint64_t load(const uint64_t *p) { return ld_strong_u64(p); }
  static __device__ __forceinline__ void store(uint64_t *p, uint64_t v) { st_strong_u64(p, v); }
  static __device__ __forceinline__ void prefetch(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
  static __device__ __forceinline__ uint4 load_tags16(const uint8_t *p) {  // p 16-byte aligned
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
  static __device__ __forceinline__ void store_tag(uint8_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
  static __device__ __forceinline__ uint64_t cas(uint64_t *p, uint64_t expect, uint64_t want) {
    return atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)expect,
                     (unsigned long long)want);
  }
  static __device__ __forceinline__ void xor64(uint64_t *p, uint64_t v) {
    atomicXor(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
  }
  static __device__ __forceinline__ void add32(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
  static __device__ __forceinline__ void or32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
  static __device__ __forceinline__ void fence() { __threadfence(); }
  static __device__ __forceinline__ void backoff() { __nanosleep(40); }
  static __device__ __forceinline__ void count(unsigned long long *p, unsigned long long v) {
    atomicAdd(p, v);
  }
  static __device__ __forceinline__ void flag(unsigned long long *p) { atomicExch(p, 1ull); }
};
#endif

// Saturating coverage add on a published slot (slow path, only near UINT_MAX;
// element.c:395-417).
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

// A slot whose meta word `m` is known to be non-EMPTY: is it `key`?  If so
// accumulate.  Returns 0 (hit, updated) or 2 (some other key).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int slot_examine(uint64_t *slot, uint64_t m, const Probe &pr, GraphCounters *ctr,
                         const uint64_t (&key)[N], uint32_t edge, uint32_t covg_add) {
  uint64_t *meta_ptr = slot + N;
  if (meta_tag(m) != pr.tag) return 2;
  if (meta_status(m) == ST_LOCKED) {
    // same tag, key words not visible yet: wait for the owner to publish
    do {
      Ops::backoff();
      m = Ops::load(meta_ptr);
    } while (meta_status(m) == ST_LOCKED);
    Ops::fence();
  }
  // published: compare the key (loads issued after the meta load that saw the
  // slot published, and bypassing L1)
  bool same = true;
#pragma unroll
  for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
  if (!same) return 2;
  if (meta_covg(m) < 0xF0000000u && covg_add < 0x01000000u) {
    Ops::add32(reinterpret_cast<uint32_t *>(meta_ptr), covg_add);
  } else {
    covg_add_saturating<Ops>(meta_ptr, covg_add, ctr);
  }
  if ((meta_edges(m) | edge) != meta_edges(m)) {
    Ops::or32(reinterpret_cast<uint32_t *>(meta_ptr) + 1, edge & 0xFFu);
  }
  return 0;
}

// find-or-insert `key`, then coverage += covg_add (saturating) and edges |= edge.
// Returns 1 if the key was new, 0 if it existed, -1 if the table is full.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int table_upsert(const TableView &t, GraphCounters *ctr, const uint64_t (&key)[N],
                         uint32_t edge, uint32_t covg_add) {
  constexpr uint32_t STRIDE = 8 * N + 8;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    const Probe pr = probe_of(lookup3_key<N>(key, r), t.bucket_mask, t.W);
    const uint64_t b0 = (uint64_t)pr.bucket * t.W;  // global index of the bucket's slot 0
    // the bucket is walked as two runs: [start, W) then [0, start)
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const uint64_t lo = seg ? b0 : b0 + pr.start;
      const uint64_t hi = seg ? b0 + pr.start : b0 + t.W;
#pragma unroll 1
      for (uint64_t gb = lo & ~15ull; gb < hi; gb += 16) {
        const uint4 tg = Ops::load_tags16(t.tags + gb);
        // bytes of this group that belong to [lo, hi)
        uint32_t valid = 0xFFFFu;
        if (gb < lo) valid &= 0xFFFFu << (uint32_t)(lo - gb);
        if (gb + 16 > hi) valid &= 0xFFFFu >> (uint32_t)(gb + 16 - hi);
        const uint32_t empty = bytes_equal_mask16(tg, 0u) & valid;
        uint32_t ev = (bytes_equal_mask16(tg, pr.tag8) & valid) | empty;
        while (ev) {
          const uint32_t pbit = ev & (0u - ev);
          ev ^= pbit;
          uint32_t p = 0;
#if defined(__CUDA_ARCH__)
          p = (uint32_t)__ffs((int)pbit) - 1u;
#else
          p = (uint32_t)__builtin_ctz(pbit);
#endif
          const uint64_t sidx = gb + p;
          uint64_t *slot = reinterpret_cast<uint64_t *>(t.slots + sidx * STRIDE);
          uint64_t *meta_ptr = slot + N;
          uint64_t m;
          if (empty & pbit) {
            // first never-claimed slot of the probe order: the key is absent (unless a
            // concurrent claim is in flight, which the CAS detects)
            const uint64_t want = meta_pack(covg_add, edge, ST_LOCKED, pr.tag);
            m = Ops::cas(meta_ptr, 0ull, want);
            if (m == 0ull) {
#pragma unroll
              for (int w = 0; w < N; w++) Ops::store(slot + w, key[w]);
              Ops::store_tag(t.tags + sidx, pr.tag8);
              Ops::fence();
              Ops::xor64(meta_ptr, (uint64_t)(ST_LOCKED ^ ST_NONE) << 40);
              if (r) Ops::count(&ctr->rehash_hist[r], 1ull);
              return 1;
            }
          } else {
            m = Ops::load(meta_ptr);
          }
          if (slot_examine<N, Ops>(slot, m, pr, ctr, key, edge, covg_add) == 0) {
            if (r) Ops::count(&ctr->rehash_hist[r], 1ull);
            return 0;
          }
        }
      }
    }
    // bucket r is full and does not hold the key: follow the rehash chain
  }
  Ops::flag(&ctr->table_full);
  return -1;
}

// hash_table_find (hash_table.c:234-267) on the UN-finalised device layout, with
// no insert in flight.  Returns true and the meta word if present.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool table_find(const TableView &t, const uint64_t (&key)[N], uint64_t &meta_out) {
  constexpr uint32_t STRIDE = 8 * N + 8;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    const Probe pr = probe_of(lookup3_key<N>(key, r), t.bucket_mask, t.W);
    const uint64_t b0 = (uint64_t)pr.bucket * t.W;
    for (int seg = 0; seg < 2; ++seg) {
      const uint64_t lo = seg ? b0 : b0 + pr.start;
      const uint64_t hi = seg ? b0 + pr.start : b0 + t.W;
      for (uint64_t gb = lo & ~15ull; gb < hi; gb += 16) {
        const uint4 tg = Ops::load_tags16(t.tags + gb);
        uint32_t valid = 0xFFFFu;
        if (gb < lo) valid &= 0xFFFFu << (uint32_t)(lo - gb);
        if (gb + 16 > hi) valid &= 0xFFFFu >> (uint32_t)(gb + 16 - hi);
        const uint32_t empty = bytes_equal_mask16(tg, 0u) & valid;
        uint32_t ev = (bytes_equal_mask16(tg, pr.tag8) & valid) | empty;
        while (ev) {
          const uint32_t pbit = ev & (0u - ev);
          ev ^= pbit;
          if (empty & pbit) return false;  // an empty slot ends the probe run
          uint32_t p = 0;
          while (!((pbit >> p) & 1u)) p++;
          uint64_t *slot = reinterpret_cast<uint64_t *>(t.slots + (gb + p) * STRIDE);
          const uint64_t m = Ops::load(slot + N);
          if (meta_tag(m) != pr.tag) continue;
          bool same = true;
#pragma unroll
          for (int w = 0; w < N; w++) same &= (Ops::load(slot + w) == key[w]);
          if (same) { meta_out = m; return true; }
        }
      }
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
  constexpr uint32_t STRIDE = 8 * N + 8;
  for (uint32_t r = 0; r <= t.max_rehash; ++r) {
    const Probe pr = probe_of(lookup3_key<N>(key, r), t.bucket_mask, t.W);
    uint8_t *bucket = t.slots + (uint64_t)pr.bucket * t.W * STRIDE;
    uint32_t s = 0;
    for (; s < t.W; ++s) {
      uint64_t *slot = reinterpret_cast<uint64_t *>(bucket + (uint64_t)s * STRIDE);
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
