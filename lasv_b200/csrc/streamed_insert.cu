// streamed_insert.cu -- sm_100a kernels of the streamed insert (streamed_insert.cuh has the scheme).
//
//   sg_region_base_kernel                 records per region of a slab (from the stripe counters), exclusive prefix
//   sg_sort_kernel                        one CTA per REGION (<= 4096 buckets): pass 1 hashes the region's records once,
//                                         counts them per bucket in shared memory and keeps every record's bucket
//                                         there; prefix -> the slab's per-bucket offsets; pass 2 reads the records again
//                                         and stores each at its place in the slab's sorted buffer (inside the region's
//                                         own ~1 MB window, so the 16-byte stores merge in L2)
//   sg_count_kernel / sg_scatter_kernel   the same sort with global atomics, for regions too large for the above
//                                         (small tables with few stripes: tests)
//   sg_insert_kernel                      persistent; one WARP per unit (one or a few buckets, <= ~6.5 KB of lines): lines
//                                         global -> shared by cp.async.bulk (UBLKCP) behind an mbarrier, the unit's
//                                         records applied in shared memory, lines shared -> global by a bulk store; two
//                                         buffers per warp, so that the load of the next unit and the write-back of
//                                         the previous one overlap the updates of this one.  No CTA-wide barrier: a
//                                         warp never waits for another warp
//   sg_drain_kernel                       spilled records through the ordinary global-memory table_upsert
//
// Integer work only.  Bound: HBM streaming -- 2 x 128 bytes per table line per insert plus 4 x the record stream.
#include <cub/device/device_scan.cuh>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "../../include/lasv_b200.h"
#include "bin_records.cuh"
#include "groups.cuh"
#include "streamed_insert.cuh"

namespace lasv {

namespace {

constexpr int SG_THREADS = 256;        // direct sorting kernels, drain kernel
#ifndef LASV_SO_THREADS
#define LASV_SO_THREADS 512
#endif
#ifndef LASV_SO_SMEM_KB
#define LASV_SO_SMEM_KB 100
#endif
constexpr int SO_THREADS = LASV_SO_THREADS;  // fused sorting kernel
constexpr uint32_t SO_MAXB = 4096;     // buckets per region the fused sorting kernel handles
constexpr uint32_t SO_SMEM = (uint32_t)LASV_SO_SMEM_KB << 10;  // its shared memory (two CTAs per SM)
#ifndef LASV_SO_U
#define LASV_SO_U 4
#endif
constexpr int SO_U = LASV_SO_U;        // record loads in flight per thread
constexpr int SI_UNITS = 4;            // units a CTA of the insert kernel works on at a time
constexpr int SI_TEAM = 2;             // warps that share one unit (occupancy: the staging buffers bound the units per
                                       // SM at 16, and 16 warps leave every scheduler with 4 -- too few to cover the
                                       // shared-memory and record-load latencies of the state machine)
constexpr int SI_WARPS = SI_UNITS * SI_TEAM;  // warps per CTA of the insert kernel
constexpr uint32_t SI_MAX_UNIT = 8;    // buckets per unit at most
constexpr uint32_t SG_CHUNK = 4096;    // records per work item of the direct sorting passes
constexpr int SG_U = 4;                // records in flight per thread in the direct sorting passes

// The table protocol of table.cuh on shared memory: same loads / CAS / reductions as DeviceOps, CTA scope.
struct SmemOps {
  static __device__ __forceinline__ uint64_t load(const uint64_t *p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
  static __device__ __forceinline__ void load16(const uint64_t *p, uint64_t &a, uint64_t &b) {
    a = *reinterpret_cast<const volatile uint64_t *>(p);
    b = *reinterpret_cast<const volatile uint64_t *>(p + 1);
  }
  static __device__ __forceinline__ void store(uint64_t *p, uint64_t v) { *reinterpret_cast<volatile uint64_t *>(p) = v; }
  static __device__ __forceinline__ void store_tag(uint8_t *p, uint32_t v) { *reinterpret_cast<volatile uint8_t *>(p) = (uint8_t)v; }
  static __device__ __forceinline__ uint64_t cas(uint64_t *p, uint64_t expect, uint64_t want) {
    return atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)expect, (unsigned long long)want);
  }
  static __device__ __forceinline__ void xor32(uint32_t *p, uint32_t v) { atomicXor(p, v); }
  static __device__ __forceinline__ void add32(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
  static __device__ __forceinline__ void or32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
  static __device__ __forceinline__ void and32(uint32_t *p, uint32_t v) { atomicAnd(p, v); }
  static __device__ __forceinline__ void fence() { __threadfence_block(); }
  static __device__ __forceinline__ void backoff() { __nanosleep(20); }
  static __device__ __forceinline__ void count(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }  // global counters
  static __device__ __forceinline__ void flag(unsigned long long *p) { atomicExch(p, 1ull); }
};

struct SlabJob {
  BinView b;
  uint32_t bucket_mask;    // of the local table
  uint32_t rshift;         // log2(buckets per region)
  uint32_t region0;        // first (local) region of the slab
  uint32_t n_regions;      // regions in the slab
  uint32_t bucket0;        // first bucket of the slab = region0 << rshift
  uint32_t n_buckets;      // buckets in the slab
  uint32_t chunk;          // records per work item (direct passes)
  uint32_t chunks_per_stripe;
};

// records sender `src` holds for (local) region `region`
__device__ __forceinline__ uint32_t stripe_count(const SlabJob &j, uint32_t src, uint32_t region) {
  const uint32_t c = __ldg(&j.b.count[stripe_counter_at(j.b, src, region)]);
  return c > j.b.cap ? j.b.cap : c;  // what did not fit went to the spill area (or straight into the table)
}

// Work item `item` of a direct sorting pass: records [lo, hi) of one stripe of the slab.  Consecutive items lie in
// different stripes, so that the CTAs running at the same time spread over the whole slab.
struct SlabItem {
  uint32_t stripe;  // index for stripe_record_at
  uint32_t region;  // local region of the stripe
  uint32_t lo, hi;
};
__device__ __forceinline__ SlabItem slab_item(const SlabJob &j, uint32_t item) {
  const uint32_t n_stripes = j.n_regions << j.b.src_shift;
  const uint32_t s = item % n_stripes, ch = item / n_stripes;
  const uint32_t region = j.region0 + (s >> j.b.src_shift), src = s & ((1u << j.b.src_shift) - 1u);
  const uint32_t cnt = stripe_count(j, src, region);
  SlabItem it;
  it.stripe = src * j.b.bins_per_src + region;
  it.region = region;
  it.lo = ch * j.chunk;
  if (it.lo > cnt) it.lo = cnt;
  it.hi = it.lo + j.chunk < cnt ? it.lo + j.chunk : cnt;
  return it;
}

// The words of a bin record ...
__device__ __forceinline__ void load_record_words(const uint64_t *src, uint32_t rw, uint64_t (&w)[4]) {
  if (rw == 2) {
    asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(src));
    w[2] = w[3] = 0;
  } else {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(src));
  }
}
// ... and the first-choice hash of its key (stored in the record unless it is a 16-byte record of a two-word key)
template <int N>
__device__ __forceinline__ uint32_t record_hash(uint32_t rw, const uint64_t (&w)[4]) {
  if (N == 1) return (uint32_t)w[1];
  if (rw == 2) {
    const uint64_t key[2] = {w[0] & 0x0000FFFFFFFFFFFFull, w[1]};
    return lookup3_key<2>(key, 0).c;
  }
  return (uint32_t)w[N];  // aux word behind the key words
}
__device__ __forceinline__ void store_record_words(uint64_t *dst, uint32_t rw, const uint64_t (&w)[4]) {
  if (rw == 2) asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(w[0]), "l"(w[1]) : "memory");
  else asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(w[0]), "l"(w[1]), "l"(w[2]), "l"(w[3]) : "memory");
}
// key, edges and count of a record
template <int N>
__device__ __forceinline__ void unpack_record(uint32_t rw, const uint64_t (&w)[4], uint64_t (&key)[N], uint32_t &edge, uint32_t &count) {
  if (N == 1) {
    key[0] = w[0];
    edge = (uint32_t)(w[1] >> 32) & 0xFFu;
    count = (uint32_t)(w[1] >> 40);
  } else if (rw == 2) {
    key[0] = w[0] & 0x0000FFFFFFFFFFFFull;
    key[1 % N] = w[1];
    edge = (uint32_t)(w[0] >> 48) & 0xFFu;
    count = (uint32_t)(w[0] >> 56);
  } else {
#pragma unroll
    for (int x = 0; x < N; x++) key[x] = w[x];
    edge = (uint32_t)(w[N] >> 32) & 0xFFu;
    count = (uint32_t)(w[N] >> 40);
  }
}

// ---- direct sorting passes (regions of more than SO_MAXB buckets: small tables with few stripes, tests): one global
// atomic per record and pass, one scattered store per record
template <int N>
__global__ void __launch_bounds__(SG_THREADS) sg_count_kernel(const SlabJob j, uint32_t *__restrict__ counts,
                                                             unsigned long long *bad) {
  const uint32_t n_items = (j.n_regions << j.b.src_shift) * j.chunks_per_stripe;
  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const SlabItem it = slab_item(j, item);
    for (uint32_t i0 = it.lo + threadIdx.x; i0 < it.hi; i0 += SG_THREADS * SG_U) {
      uint32_t lb[SG_U];
#pragma unroll
      for (int u = 0; u < SG_U; u++) {
        const uint32_t i = i0 + (uint32_t)u * SG_THREADS;
        lb[u] = 0xFFFFFFFFu;
        uint64_t w[4];
        if (i < it.hi) {
          load_record_words(j.b.recs + stripe_record_at(j.b, it.stripe, i) * j.b.rec_words, j.b.rec_words, w);
          lb[u] = (record_hash<N>(j.b.rec_words, w) & j.bucket_mask) - j.bucket0;
        }
      }
#pragma unroll
      for (int u = 0; u < SG_U; u++) {
        if (lb[u] == 0xFFFFFFFFu) continue;
        if (lb[u] < j.n_buckets && ((lb[u] + j.bucket0) >> j.rshift) == it.region) atomicAdd(&counts[lb[u]], 1u);
        else atomicExch(bad, 1ull);  // a record outside its stripe's region: never, by construction
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(SG_THREADS) sg_scatter_kernel(const SlabJob j, const uint32_t *__restrict__ offsets,
                                                               uint32_t *cursor, uint64_t *__restrict__ sorted) {
  const uint32_t n_items = (j.n_regions << j.b.src_shift) * j.chunks_per_stripe;
  const uint32_t rw = j.b.rec_words;
  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const SlabItem it = slab_item(j, item);
    for (uint32_t i0 = it.lo + threadIdx.x; i0 < it.hi; i0 += SG_THREADS * SG_U) {
      uint64_t w[SG_U][4];
      uint32_t lb[SG_U], at[SG_U];
#pragma unroll
      for (int u = 0; u < SG_U; u++) {
        const uint32_t i = i0 + (uint32_t)u * SG_THREADS;
        lb[u] = 0xFFFFFFFFu;
        w[u][0] = w[u][1] = w[u][2] = w[u][3] = 0;
        if (i < it.hi) {
          load_record_words(j.b.recs + stripe_record_at(j.b, it.stripe, i) * rw, rw, w[u]);
          lb[u] = (record_hash<N>(rw, w[u]) & j.bucket_mask) - j.bucket0;
          if (lb[u] >= j.n_buckets || ((lb[u] + j.bucket0) >> j.rshift) != it.region) lb[u] = 0xFFFFFFFFu;  // counted as `bad`
        }
      }
#pragma unroll
      for (int u = 0; u < SG_U; u++) at[u] = lb[u] != 0xFFFFFFFFu ? offsets[lb[u]] + atomicAdd(&cursor[lb[u]], 1u) : 0u;
#pragma unroll
      for (int u = 0; u < SG_U; u++)
        if (lb[u] != 0xFFFFFFFFu) store_record_words(sorted + (uint64_t)at[u] * rw, rw, w[u]);
    }
  }
}

// ---- fused sort, one CTA per region
// region_base[r] = records of the slab's regions before region r (r = 0 .. n_regions), from the stripe counters the
// partition kernel left; offsets[n_buckets] = all records of the slab.  One CTA.
__global__ void __launch_bounds__(1024) sg_region_base_kernel(const SlabJob j, uint32_t *__restrict__ region_base,
                                                              uint32_t *__restrict__ offsets) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t tid = threadIdx.x, n_src = 1u << j.b.src_shift;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (uint32_t r0 = 0; r0 < j.n_regions; r0 += 1024) {
    const uint32_t r = r0 + tid;
    uint32_t c = 0;
    if (r < j.n_regions)
      for (uint32_t s = 0; s < n_src; s++) c += stripe_count(j, s, j.region0 + r);
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if ((int)(tid & 31) >= o) incl += y;
    }
    if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
    __syncthreads();
    uint32_t before = carry;
    for (uint32_t wi = 0; wi < (tid >> 5); wi++) before += warp_sums[wi];
    if (r < j.n_regions) region_base[r] = before + incl - c;
    __syncthreads();
    if (tid == 1023) carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) {
    region_base[j.n_regions] = carry;
    offsets[j.n_buckets] = carry;
  }
}

// Shared memory: cnt[Bk] | cur[Bk] | bucket-in-region of the region's first `lbcap` records (one byte each when the
// region has <= 256 buckets, else two); records beyond lbcap are hashed a second time in pass 2.
template <int N>
__global__ void __launch_bounds__(SO_THREADS) sg_sort_kernel(const SlabJob j, const uint32_t *__restrict__ region_base,
                                                            uint32_t *__restrict__ offsets, uint64_t *__restrict__ sorted,
                                                            unsigned long long *bad) {
  extern __shared__ __align__(16) unsigned char so_smem[];
  __shared__ uint32_t warp_sums[SO_THREADS / 32];
  const uint32_t Bk = 1u << j.rshift, rw = j.b.rec_words, n_src = 1u << j.b.src_shift;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(so_smem), *cur = cnt + Bk;
  uint8_t *lb8 = reinterpret_cast<uint8_t *>(cur + Bk);
  uint16_t *lb16 = reinterpret_cast<uint16_t *>(cur + Bk);
  const bool narrow = Bk <= 256;
  const uint32_t lbcap = (SO_SMEM - 8u * Bk) / (narrow ? 1u : 2u);
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (Bk + SO_THREADS - 1) / SO_THREADS;  // buckets per thread in the prefix
  for (uint32_t r = blockIdx.x; r < j.n_regions; r += gridDim.x) {
    const uint32_t region = j.region0 + r, rb0 = region << j.rshift;
    for (uint32_t b = tid; b < Bk; b += SO_THREADS) cnt[b] = 0;
    __syncthreads();
    // pass 1: hash, count, remember the bucket
    uint32_t q0 = 0;
    for (uint32_t s = 0; s < n_src; s++) {
      const uint32_t n = stripe_count(j, s, region), stripe = s * j.b.bins_per_src + region;
      for (uint32_t i0 = tid; i0 < n; i0 += SO_THREADS * SO_U) {
        uint64_t w[SO_U][4];
#pragma unroll
        for (int x = 0; x < SO_U; x++) {  // SO_U loads in flight per thread
          const uint32_t i = i0 + (uint32_t)x * SO_THREADS;
          w[x][0] = w[x][1] = w[x][2] = w[x][3] = 0;
          if (i < n) load_record_words(j.b.recs + stripe_record_at(j.b, stripe, i) * rw, rw, w[x]);
        }
#pragma unroll
        for (int x = 0; x < SO_U; x++) {
          const uint32_t i = i0 + (uint32_t)x * SO_THREADS;
          if (i >= n) continue;
          uint32_t b = (record_hash<N>(rw, w[x]) & j.bucket_mask) - rb0;
          if (b >= Bk) {  // a record outside its stripe's region: never, by construction; reported, and kept consistent
            atomicExch(bad, 1ull);
            b &= Bk - 1u;
          }
          atomicAdd(&cnt[b], 1u);
          const uint32_t q = q0 + i;
          if (q < lbcap) {
            if (narrow) lb8[q] = (uint8_t)b;
            else lb16[q] = (uint16_t)b;
          }
        }
      }
      q0 += n;
    }
    __syncthreads();
    // exclusive prefix over the region's buckets -> the slab's offsets, and the cursors of pass 2
    {
      uint32_t sum = 0;
      for (uint32_t x = 0; x < per; x++) {
        const uint32_t b = tid * per + x;
        if (b < Bk) sum += cnt[b];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if ((int)(tid & 31) >= o) incl += y;
      }
      if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
      __syncthreads();
      uint32_t run = region_base[r] + incl - sum;
      for (uint32_t wi = 0; wi < (tid >> 5); wi++) run += warp_sums[wi];
      for (uint32_t x = 0; x < per; x++) {
        const uint32_t b = tid * per + x;
        if (b < Bk) {
          cur[b] = run;
          offsets[rb0 - j.bucket0 + b] = run;
          run += cnt[b];
        }
      }
    }
    __syncthreads();
    // pass 2: every record to its place (the stores of a CTA fall into the region's own window of the sorted buffer)
    q0 = 0;
    for (uint32_t s = 0; s < n_src; s++) {
      const uint32_t n = stripe_count(j, s, region), stripe = s * j.b.bins_per_src + region;
      for (uint32_t i0 = tid; i0 < n; i0 += SO_THREADS * SO_U) {
        uint64_t w[SO_U][4];
#pragma unroll
        for (int x = 0; x < SO_U; x++) {
          const uint32_t i = i0 + (uint32_t)x * SO_THREADS;
          w[x][0] = w[x][1] = w[x][2] = w[x][3] = 0;
          if (i < n) load_record_words(j.b.recs + stripe_record_at(j.b, stripe, i) * rw, rw, w[x]);
        }
#pragma unroll
        for (int x = 0; x < SO_U; x++) {
          const uint32_t i = i0 + (uint32_t)x * SO_THREADS;
          if (i >= n) continue;
          const uint32_t q = q0 + i;
          uint32_t b;
          if (q < lbcap) b = narrow ? (uint32_t)lb8[q] : (uint32_t)lb16[q];
          else b = ((record_hash<N>(rw, w[x]) & j.bucket_mask) - rb0) & (Bk - 1u);
          const uint32_t at = atomicAdd(&cur[b], 1u);
          store_record_words(sorted + (uint64_t)at * rw, rw, w[x]);
        }
      }
      q0 += n;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Shared-memory accesses by 32-bit shared address (LDS / STS / ATOMS / REDS): through a generic pointer the compiler
// emits generic LD / ST / ATOM, which resolve the address space at run time and are several times slower on atomics.
__device__ __forceinline__ uint64_t lds64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint64_t atoms_cas64(uint32_t a, uint64_t expect, uint64_t want) {
  uint64_t old;
  asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(expect), "l"(want) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_or32(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

struct InsertJob {
  const uint64_t *sorted;   // the slab's records, sorted by bucket
  uint32_t rec_words;
  const uint32_t *offsets;  // [buckets of the slab + 1]: records of bucket b are sorted[offsets[b] .. offsets[b+1])
  uint32_t unit0;           // first unit of the slab (table-wide numbering)
  uint32_t n_units;
  uint32_t unit_shift;      // buckets per unit = 1 << unit_shift (<= SI_MAX_UNIT)
  uint32_t unit_bytes;
  uint32_t *spill;
  unsigned long long *n_spill;
};

// One WARP per unit at a time, two staging buffers per warp.  While the records of unit i are applied in one buffer,
// unit i+1 (the warp's next unit that has records) arrives in the other and unit i-1 is on its way back.  Lane 0
// issues the copies; an mbarrier per buffer counts the arriving bytes; the write-back of a buffer must have READ it
// before the buffer is refilled (wait_group.read), and the updates made through the generic proxy are fenced before
// the async proxy reads them.  Only this warp touches the unit.
//
// The records of a unit are applied by a STATE MACHINE, one line probe per lane and step: a lane walks its own records
// (lo + lane, lo + lane + 32, ...) and, in every step of the warp, examines ONE line for its current record --
//   a tag in the header equals mine   -> compare the slot's key: equal = accumulate, done; else never look at it again
//   no such tag, a free slot          -> claim it (CAS on the meta word; the loser looks again in the next step)
//   neither                           -> next line of the bucket (NL lines without the key or room: the record spills)
// and a lane that is done with a record starts its next one at once.  All lanes run the same few instructions per
// step whatever their records need; with one find-or-insert call per lane (bucket_upsert) a round of 32 records took
// as many passes as its slowest record and ran at 13 of 32 lanes.  Same protocol as table.cuh: slots are claimed
// in probe order and never released, a tag byte is written after the key words.  Within a step every lane reads its
// header BEFORE any lane claims (__syncwarp), so lanes holding the same new key see the same free slot, one wins,
// and the others find its tag in the next step.
template <int N>
__global__ void __launch_bounds__(SI_WARPS * 32) sg_insert_kernel(const InsertJob j, const TableView t, GraphCounters *ctr) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  extern __shared__ __align__(128) unsigned char staged[];
  __shared__ __align__(8) uint64_t mbar_all[SI_UNITS][2];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t team = warp / SI_TEAM, member = warp % SI_TEAM;  // the warps of a team share one unit; member 0 leads
  const bool leader = member == 0 && lane == 0;                   // issues the team's copies
  const uint32_t ubytes = j.unit_bytes, upu = 1u << j.unit_shift, bucket_bytes = t.NL * LINE_BYTES;
  unsigned char *bufs[2] = {staged + (size_t)(2 * team) * ubytes, staged + (size_t)(2 * team + 1) * ubytes};
  uint64_t *mbar = mbar_all[team];
  if (leader) {
    for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint64_t full_valid = header_valid_mask(G), last_valid = header_valid_mask(t.last_slots);
  const uint32_t n_warps = gridDim.x * SI_UNITS, gw = blockIdx.x * SI_UNITS + team;  // (teams, in fact)
  auto next_with_records = [&](uint32_t u) {  // the same value in every lane
    for (; u < j.n_units; u += n_warps)
      if (__ldg(&j.offsets[u << j.unit_shift]) != __ldg(&j.offsets[(u + 1) << j.unit_shift])) return u;
    return j.n_units;
  };
  auto load = [&](uint32_t u, int b) {  // one lane
    const uint8_t *src = t.lines + (uint64_t)(j.unit0 + u) * ubytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[b])), "r"(ubytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(bufs[b])),
                 "l"(src), "r"(ubytes), "r"(smem_addr(&mbar[b]))
                 : "memory");
  };
  unsigned long long my_new = 0;
  uint32_t u = next_with_records(gw);
  if (leader && u < j.n_units) load(u, 0);
  for (uint32_t it = 0; u < j.n_units; it++) {
    const int b = it & 1;
    const uint32_t un = next_with_records(u + n_warps);
    if (leader) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the other buffer's write-back has read it
      if (un < j.n_units) load(un, b ^ 1);
    }
    const uint32_t lo = __ldg(&j.offsets[u << j.unit_shift]), hi = __ldg(&j.offsets[(u + 1) << j.unit_shift]);
    // this lane's first record is on its way while the lines arrive
    uint32_t i = lo + member * 32u + lane;  // the team's lanes take the unit's records in turn
    bool have = i < hi;
    uint64_t w[4] = {0, 0, 0, 0};
    if (have) load_record_words(j.sorted + (uint64_t)i * j.rec_words, j.rec_words, w);
    const uint32_t parity = (it >> 1) & 1u;
    uint32_t done = 0;
    for (uint32_t tries = 0; !done; tries++) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_addr(&mbar[b])), "r"(parity) : "memory");
      if (tries > (1u << 24)) __trap();  // a copy that never lands must end the kernel with an error, not hang the GPU
    }
    // the record this lane is working on
    uint64_t key[N];
    uint32_t edge = 0, count = 0, cur = 0, tag16 = 0, line = 0, steps = 0;
    uint64_t rep = 0, rejected = 0;
    const uint32_t sbuf = smem_addr(bufs[b]);
    uint32_t sbase = sbuf;  // shared address of my record's bucket
    bool busy = false;
#pragma unroll
    for (int x = 0; x < N; x++) key[x] = 0;
    for (;;) {
      __syncwarp();
      if (!busy && have) {  // start my next record
        unpack_record<N>(j.rec_words, w, key, edge, count);
        cur = i;
        sbase = sbuf;
        if (upu > 1) {  // the bucket of a record is its place in the sorted buffer: no hashing here
          uint32_t bi = 0;
          for (uint32_t x = 1; x < upu; x++) bi += (i >= __ldg(&j.offsets[(u << j.unit_shift) + x]));
          sbase += bi * bucket_bytes;
        }
        const Probe pr = probe_of_key<N>(key);
        rep = (uint64_t)pr.tag8 * 0x0101010101010101ull;
        tag16 = pr.tag16;
        line = mulhi32(pr.h, t.NL);
        steps = 0;
        rejected = 0;
        busy = true;
        i += 32u * SI_TEAM;
        have = i < hi;
        if (have) load_record_words(j.sorted + (uint64_t)i * j.rec_words, j.rec_words, w);  // travels while this one is applied
      }
      if (!__any_sync(0xFFFFFFFFu, busy)) break;
      const uint32_t lp = sbase + line * LINE_BYTES;
      const uint64_t hdr = lds64(lp);
      __syncwarp();  // all headers read before any claim of this step
      if (busy) {
        const uint64_t valid = (line == t.NL - 1) ? last_valid : full_valid;
        const uint64_t cand = zero_bytes64(hdr ^ rep) & valid & ~rejected;
        if (cand) {
          const uint32_t slot = lp + 8u + first_byte_index(cand) * SLOT;
          bool same = true;
#pragma unroll
          for (int x = 0; x < N; x++) same &= (lds64(slot + 8u * x) == key[x]);
          if (same) {
            const uint32_t meta = slot + 8u * N;
            const uint64_t m = lds64(meta);
            if (meta_covg(m) < 0xF0000000u && count < 0x01000000u) reds_add32(meta, count);
            else covg_add_saturating<SmemOps>(reinterpret_cast<uint64_t *>(__cvta_shared_to_generic(meta)), count, ctr);
            if ((meta_edges(m) | edge) != meta_edges(m)) reds_or32(meta + 4u, edge & 0xFFu);
            busy = false;
          } else {
            rejected |= cand & (0 - cand);  // another key with my tag
          }
        } else {
          const uint64_t empt = zero_bytes64(hdr) & valid;
          if (empt) {
            const uint32_t si = first_byte_index(empt);
            const uint32_t slot = lp + 8u + si * SLOT;
            if (atoms_cas64(slot + 8u * N, 0ull, meta_pack(count, edge, ST_NONE, tag16)) == 0ull) {
#pragma unroll
              for (int x = 0; x < N; x++) sts64(slot + 8u * x, key[x]);
              __threadfence_block();
              sts8(lp + si, (uint32_t)(rep & 0xFFu));
              my_new++;
              busy = false;
            }
            // else a lane of this warp took the slot in this very step: its tag is in the header by the next one
          } else {
            rejected = 0;
            line = (line + 1 == t.NL) ? 0u : line + 1;
            if (++steps == t.NL) {  // W other keys in the bucket: the record leaves the unit
              const unsigned long long at = atomicAdd(j.n_spill, 1ull);
              j.spill[at] = cur;  // the list has room for every record of the slab
              busy = false;
            }
          }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my updates -> visible to the bulk copy
    asm volatile("bar.sync %0, %1;" ::"r"(1u + team), "r"(32u * SI_TEAM) : "memory");  // the whole team is done with the unit
    if (leader) {
      uint8_t *dst = t.lines + (uint64_t)(j.unit0 + u) * ubytes;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(bufs[b])), "r"(ubytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    u = un;
  }
  if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, o);
  if (lane == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

template <int N>
__global__ void __launch_bounds__(SG_THREADS) sg_drain_kernel(const uint64_t *__restrict__ sorted, const uint32_t rec_words,
                                                             const uint32_t *__restrict__ spill,
                                                             const unsigned long long *__restrict__ n_spill, const TableView t,
                                                             GraphCounters *ctr) {
  const uint64_t n = *n_spill;
  unsigned long long my_new = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * SG_THREADS + threadIdx.x; q < n; q += (uint64_t)gridDim.x * SG_THREADS) {
    uint64_t key[N];
    uint32_t c, edge, count;
    load_bin_record<N>(sorted + (uint64_t)spill[q] * rec_words, rec_words, key, c, edge, count);
    my_new += (table_upsert_hashed<N, DeviceOps>(t, ctr, key, c, edge, count) == 1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, o);
  if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

template <typename F>
void sg_dispatch(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

uint32_t log2_floor(uint64_t v) {
  uint32_t l = 0;
  while ((2ull << l) <= v) l++;
  return l;
}

}  // namespace

bool streamed_geometry(const TableView &t, const BinView &b, StreamedGeom *out) {
  const uint64_t bucket_bytes = (uint64_t)t.NL * LINE_BYTES;
  const uint64_t n_buckets = (uint64_t)t.bucket_mask + 1;
  // two staging buffers for each of the four warps of a CTA must fit in shared memory
  if (2ull * SI_UNITS * bucket_bytes > 200u * 1024u) return false;
  if (!b.bins_per_src || n_buckets % b.bins_per_src) return false;
  const uint64_t buckets_per_region = n_buckets / b.bins_per_src;
  // units of <= 6.5 KB (one bucket if a bucket is larger): 16 warps with two buffers each share the 227 KB of an SM
  uint32_t shift = 0;
  while ((2ull << shift) * bucket_bytes <= 6656u && (2u << shift) <= SI_MAX_UNIT && (2ull << shift) <= buckets_per_region) shift++;
  StreamedGeom g;
  g.unit_shift = shift;
  g.unit_bytes = (uint32_t)(bucket_bytes << shift);
  g.rshift = log2_floor(buckets_per_region);
  // a slab: a run of regions sorted and inserted together; its size only bounds the scratch memory (sorted
  // records + spill list: slab_cap records) and the 32-bit offsets
  uint32_t rps = 1;
  static const uint64_t slab_buckets = [] {  // experiments: LASV_B200_SLAB_BUCKETS
    const char *e = getenv("LASV_B200_SLAB_BUCKETS");
    const long v = e ? atol(e) : 0;
    return v > 0 ? (uint64_t)v : (uint64_t)(1u << 20);
  }();
  const uint64_t n_src = b.src_used ? b.src_used : (1ull << b.src_shift);  // senders that can hold records
  while (rps < b.bins_per_src && (uint64_t)(2 * rps) * buckets_per_region <= slab_buckets &&
         ((uint64_t)(2 * rps) * n_src * b.cap) < 0x7FFFFF00ull)
    rps <<= 1;
  g.regions_per_slab = rps;
  g.buckets_per_slab = (uint32_t)(rps * buckets_per_region);
  g.n_slabs = b.bins_per_src / rps;
  g.slab_cap = (uint64_t)rps * n_src * b.cap;
  g.fused = buckets_per_region <= SO_MAXB ? 1u : 0u;
  *out = g;
  return g.slab_cap < 0xFFFFFF00ull && (uint64_t)rps * buckets_per_region < 0xFFFFFF00ull;  // 32-bit offsets inside a slab
}

size_t streamed_scan_tmp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
  return bytes;
}

int launch_streamed_insert(int N, const BinView &b, const TableView &t, GraphCounters *ctr, unsigned long long *overflow,
                           const StreamedGeom &geom, const StreamedScratch &s, cudaStream_t st, uint64_t *launches,
                           const std::function<void(int, int)> &span) {
  const uint32_t units_per_slab = geom.buckets_per_slab >> geom.unit_shift;
  int rc = 0;
  auto mark = [&](int kind, int begin) {
    if (span) span(kind, begin);
  };
  sg_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    const size_t smem = 2 * (size_t)SI_UNITS * geom.unit_bytes;
    // kernel attributes belong to a device: cached per (device, N), set once per geometry, not per launch
    constexpr int MAX_DEV = 64;
    static int per_sm_all[MAX_DEV][4] = {};
    static size_t smem_all[MAX_DEV][4] = {};
    static int sort_all[MAX_DEV][4] = {};
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= MAX_DEV) {
      rc = -1;
      return;
    }
    int *per_sm_cached = per_sm_all[dev_id], *sort_per_sm = sort_all[dev_id];
    size_t *smem_cached = smem_all[dev_id];
    if (smem_cached[NN] != smem) {
      int per_sm = 0;
      if (cudaFuncSetAttribute(sg_insert_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sg_insert_kernel<NN>, SI_WARPS * 32, smem) != cudaSuccess) {
        rc = -1;
        return;
      }
      per_sm_cached[NN] = per_sm < 1 ? 1 : per_sm;
      smem_cached[NN] = smem;
    }
    if (!sort_per_sm[NN]) {
      int per_sm = 0;
      if (cudaFuncSetAttribute(sg_sort_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SO_SMEM) != cudaSuccess ||
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sg_sort_kernel<NN>, SO_THREADS, SO_SMEM) != cudaSuccess) {
        rc = -1;
        return;
      }
      sort_per_sm[NN] = per_sm < 1 ? 1 : per_sm;
    }
    for (uint32_t slab = 0; slab < geom.n_slabs; slab++) {
      SlabJob j;
      j.b = b;
      j.bucket_mask = t.bucket_mask;
      j.rshift = geom.rshift;
      j.region0 = slab * geom.regions_per_slab;
      j.n_regions = geom.regions_per_slab;
      j.bucket0 = j.region0 << geom.rshift;
      j.n_buckets = geom.buckets_per_slab;
      j.chunk = SG_CHUNK;
      j.chunks_per_stripe = (b.cap + SG_CHUNK - 1) / SG_CHUNK;
      mark(LASV_T_SORT, 1);
      if (geom.fused) {
        sg_region_base_kernel<<<1, 1024, 0, st>>>(j, s.counts, s.offsets);
        static const int sort_ctas = [] { const char *e = getenv("LASV_B200_SORT_CTAS"); return e ? atoi(e) : 0; }();  // experiments
        const uint32_t grid = std::min<uint32_t>(j.n_regions, sort_ctas > 0 ? (uint32_t)sort_ctas : (uint32_t)(sm_count() * sort_per_sm[NN]));
        sg_sort_kernel<NN><<<grid, SO_THREADS, SO_SMEM, st>>>(j, s.counts, s.offsets, s.sorted, overflow);
        if (launches) *launches += 2;
      } else {
        const uint32_t n_items = (j.n_regions << b.src_shift) * j.chunks_per_stripe;
        const uint32_t grid = std::min<uint32_t>(n_items, (uint32_t)sm_count() * 8u);
        cudaMemsetAsync(s.counts, 0, ((size_t)geom.buckets_per_slab + 1) * sizeof(uint32_t), st);
        sg_count_kernel<NN><<<grid, SG_THREADS, 0, st>>>(j, s.counts, overflow);
        size_t tmp = s.scan_tmp_bytes;
        cub::DeviceScan::ExclusiveSum(s.scan_tmp, tmp, s.counts, s.offsets, (int)(geom.buckets_per_slab + 1), st);
        cudaMemsetAsync(s.cursor, 0, (size_t)geom.buckets_per_slab * sizeof(uint32_t), st);
        sg_scatter_kernel<NN><<<grid, SG_THREADS, 0, st>>>(j, s.offsets, s.cursor, s.sorted);
        if (launches) *launches += 3;
      }
      cudaMemsetAsync(s.n_spill, 0, sizeof(unsigned long long), st);
      mark(LASV_T_SORT, 0);
      InsertJob ij;
      ij.sorted = s.sorted;
      ij.rec_words = b.rec_words;
      ij.offsets = s.offsets;
      ij.unit0 = j.bucket0 >> geom.unit_shift;
      ij.n_units = units_per_slab;
      ij.unit_shift = geom.unit_shift;
      ij.unit_bytes = geom.unit_bytes;
      ij.spill = s.spill;
      ij.n_spill = s.n_spill;
      uint32_t igrid = (uint32_t)per_sm_cached[NN] * (uint32_t)sm_count();  // persistent: resident CTAs x SMs
      igrid = std::min<uint32_t>(igrid, (units_per_slab + SI_UNITS - 1) / SI_UNITS);
      mark(LASV_T_STREAM, 1);
      sg_insert_kernel<NN><<<igrid, SI_WARPS * 32, smem, st>>>(ij, t, ctr);
      mark(LASV_T_STREAM, 0);
      mark(LASV_T_DRAIN, 1);
      sg_drain_kernel<NN><<<(uint32_t)sm_count() * 2u, SG_THREADS, 0, st>>>(s.sorted, b.rec_words, s.spill, s.n_spill, t, ctr);
      mark(LASV_T_DRAIN, 0);
      if (launches) *launches += 2;
    }
  });
  if (rc) return rc;
  if (b.spill_cap) {  // the senders' spill stripes (records of any region): the random-access insert, spill stripes only
    BinView sp = b;
    sp.n_bins = 0;
    mark(LASV_T_DRAIN, 1);
    launch_insert_bins(N, sp, t, ctr, 0, st);
    mark(LASV_T_DRAIN, 0);
    if (launches) *launches += 1;
  }
  return 0;
}

}  // namespace lasv
