// graph_kernels.cu -- hand-written sm_100a kernels of the de Bruijn graph build.
//
//   build_kernel<N,MODE>   K1+K2: stream tile -> classify/pack (128-bit coalesced
//                          loads, shared-memory bit strings) -> window validity ->
//                          O(1) forward / reverse-complement k-mer extraction ->
//                          canonical key + edge byte -> then, by MODE:
//                            BIN     one record per k-mer into the stripe of the
//                                    (owner GPU, 64 MB table region) it will touch
//                                    -- the production path (partition kernel)
//                            FUSED   find-or-insert straight into the table (K3
//                                    in the same pass; small tables)
//                            RECORDS / ROUTE  (2N+1)-word records, flat or per
//                                    owner (older variable-length exchange)
//   insert_bins_kernel     K3: sweeps the stripes region by region with the whole
//                          grid, find-or-insert + coverage / edge accumulation
//                          (upsert_batch: lock-step fast path, deferred 32-wide
//                          slow path).
//   insert_records_kernel  K3 on (2N+1)-word records.
//   read_stats_kernel      loader bookkeeping: bad reads, bases loaded, contig
//                          length histogram (file_reader.c:460-470,579).
//   finalize / dump / ingest / find / summary kernels: K4 and the .ctx wire
//                          format (element.c:894-910, file_reader.c:1686-1703).
//
// Integer hashing only: no tensor cores.  What bounds the kernels, and why they
// look the way they do, is measured in tools/ubench.  Grids
// are persistent: (#SMs x resident CTAs) CTAs looping over tiles / chunks.
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <type_traits>
#include "graph_kernels.cuh"
#include "bin_records.cuh"
#include "synth.cuh"

namespace lasv {

namespace {

constexpr int THREADS = 256;
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Warp-level pre-aggregation of identical keys (reads with homopolymer or
// tandem runs put the same k-mer in neighbouring lanes; 30x coverage makes hot
// k-mers).  On return `lead` tells whether this lane performs the table update,
// carrying the group's summed count and OR-ed edges.
template <int N>
__device__ __forceinline__ bool warp_aggregate(bool active, const uint64_t (&key)[N],
                                               uint32_t &edge, uint32_t &count) {
  // Common case: no two NEIGHBOURING lanes hold the same key (consecutive windows of a read are
  // neighbours; a homopolymer or tandem run is what this function is for) -- one shuffle per key
  // word and a vote.  Only then the general grouping below (match.any costs ~20 % of the
  // partition kernel if run every round; a repeated key in non-neighbouring lanes just stays
  // two updates).
  {
    bool same_prev = active && lane_id() > 0;
#pragma unroll
    for (int i = 0; i < N; i++) same_prev &= (__shfl_up_sync(FULL, key[i], 1) == key[i]);
    same_prev &= (__shfl_up_sync(FULL, active ? 1 : 0, 1) != 0);
    bool same_prev2 = active && lane_id() > 1;   // period-2 repeats: (AC)n alternates two keys
#pragma unroll
    for (int i = 0; i < N; i++) same_prev2 &= (__shfl_up_sync(FULL, key[i], 2) == key[i]);
    same_prev2 &= (__shfl_up_sync(FULL, active ? 1 : 0, 2) != 0);
    if (!__any_sync(FULL, same_prev | same_prev2)) return active;
  }
  uint64_t sig = key[N - 1];
#pragma unroll
  for (int i = 0; i < N - 1; i++) sig = sig * 0x9E3779B97F4A7C15ull + key[i];
  if (!active) sig = ~0ull - lane_id();
  const unsigned grp = __match_any_sync(FULL, sig);
  // (The reductions below run once per distinct lane mask, i.e. 32 times for 32 singleton
  // groups -- never enter them for nothing.)
  if (!__any_sync(FULL, (grp & (grp - 1u)) != 0u)) return active;
  const int leader = __ffs(grp) - 1;
  bool same = active;
#pragma unroll
  for (int i = 0; i < N; i++) same &= (__shfl_sync(FULL, key[i], leader) == key[i]);
  const unsigned same_mask = __ballot_sync(FULL, same);
  if (!same) return active;  // signature clash with a different key: go alone
  const unsigned g = grp & same_mask;  // lanes holding exactly this key (leader included)
  const uint32_t e = __reduce_or_sync(g, edge);
  const uint32_t c = __reduce_add_sync(g, count);
  if ((int)lane_id() != __ffs(g) - 1) return false;
  edge = e;
  count = c;
  return true;
}

template <int N>
__device__ __forceinline__ void store_record(uint32_t *dst, const uint64_t (&key)[N], uint32_t edge,
                                             uint32_t count) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    dst[2 * i] = (uint32_t)key[i];
    dst[2 * i + 1] = (uint32_t)(key[i] >> 32);
  }
  dst[2 * N] = (edge & 0xFFu) | (count << 8);
}

template <int N>
__device__ __forceinline__ void load_record(const uint32_t *src, uint64_t (&key)[N], uint32_t &edge,
                                            uint32_t &count) {
#pragma unroll
  for (int i = 0; i < N; i++) key[i] = (uint64_t)src[2 * i] | ((uint64_t)src[2 * i + 1] << 32);
  const uint32_t m = src[2 * N];
  edge = m & 0xFFu;
  count = m >> 8;
}

// --------------------------------------------------------------------------
// upsert_batch: U independent find-or-inserts per lane with the memory accesses
// of the common case issued together; the uncommon cases are set aside and run
// 32 at a time
// --------------------------------------------------------------------------
// Steady state of a 30x build: ~99 % of the operations find their key, ~90 % of
// them behind a visible tag in the first line probed.  That case is two dependent
// round trips (header, then meta + key words of the tagged slot) and one
// reduction; the U items of a lane take it in lock step so that U headers, then U
// slots, are in flight at once.  Whatever does not finish there (no tag match in
// the first line, tag collision, new key) needs the complete protocol of
// table.cuh, whose loops would run with two or three lanes of the warp alive if
// entered on the spot.  Those items are pushed to a per-warp queue in shared
// memory instead and drained 32 at a time (operations on the table commute, so
// the order does not matter).
template <int N>
struct DeferItem {
  uint64_t key[N];
  uint32_t c;
  uint32_t em;  // edges | count << 8
};
constexpr int DEFER_CAP = 64;  // < 32 queued between calls, + at most 32 per push

template <int N>
struct DeferQueue {
  DeferItem<N> *q;  // DEFER_CAP items of this warp, shared memory
  uint32_t n;       // warp-uniform

  __device__ __forceinline__ void drain32(const TableView &t, GraphCounters *ctr, unsigned long long &my_new,
                                          uint32_t take) {
    // `take` <= 32 items off the top; lanes >= take idle
    const unsigned lane = lane_id();
    n -= take;
    __syncwarp();
    if (lane < take) {
      const DeferItem<N> it = q[n + lane];
      uint64_t key[N];
#pragma unroll
      for (int w = 0; w < N; w++) key[w] = it.key[w];
      const int r = table_upsert_hashed<N, DeviceOps>(t, ctr, key, it.c, it.em & 0xFFu, it.em >> 8);
      my_new += (r == 1);
    }
    __syncwarp();
  }
  __device__ __forceinline__ void push(const TableView &t, GraphCounters *ctr, unsigned long long &my_new,
                                       bool want, const uint64_t (&key)[N], uint32_t c, uint32_t edge,
                                       uint32_t count) {
    const unsigned m = __ballot_sync(FULL, want);
    if (m == 0u) return;
    if (want) {
      DeferItem<N> &d = q[n + __popc(m & lanemask_lt())];
#pragma unroll
      for (int w = 0; w < N; w++) d.key[w] = key[w];
      d.c = c;
      d.em = (edge & 0xFFu) | (count << 8);
    }
    n += __popc(m);
    if (n >= 32u) drain32(t, ctr, my_new, 32u);
  }
  __device__ __forceinline__ void flush(const TableView &t, GraphCounters *ctr, unsigned long long &my_new) {
    while (n) drain32(t, ctr, my_new, n < 32u ? n : 32u);
  }
};

template <int N, int U>
__device__ __forceinline__ void upsert_batch(const TableView &t, GraphCounters *ctr, DeferQueue<N> &dq,
                                             const uint64_t (&key)[U][N], const uint32_t (&c)[U],
                                             const uint32_t (&edge)[U], const uint32_t (&count)[U],
                                             const bool (&act)[U], unsigned long long &my_new) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT, G = LineLayout<N>::G;
  uint8_t *lp[U];
  uint64_t hdr[U];
  uint32_t tag8[U];
  bool last[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const Probe pr = probe_of_key<N>(key[u]);
    const uint32_t line = mulhi32(pr.h, t.NL);
    tag8[u] = pr.tag8;
    last[u] = line == t.NL - 1;
    lp[u] = bucket_base(t, c[u] & t.bucket_mask) + (uint64_t)line * LINE_BYTES;
  }
#pragma unroll
  for (int u = 0; u < U; u++) hdr[u] = act[u] ? DeviceOps::load(reinterpret_cast<uint64_t *>(lp[u])) : 0ull;
  uint64_t *slot[U];
  bool has[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const uint64_t valid = header_valid_mask(last[u] ? t.last_slots : G);
    const uint64_t cand = zero_bytes64(hdr[u] ^ ((uint64_t)tag8[u] * 0x0101010101010101ull)) & valid;
    has[u] = act[u] && cand != 0ull;
    const uint32_t i = cand ? first_byte_index(cand) : 0u;
    slot[u] = reinterpret_cast<uint64_t *>(lp[u] + 8u + i * SLOT);
  }
  uint64_t sw[U][N + 1];
#pragma unroll
  for (int u = 0; u < U; u++) {
#pragma unroll
    for (int w = 0; w < N; w++) sw[u][w] = ~key[u][w];
    sw[u][N] = 0;
    if (has[u]) slot_load_words<N, DeviceOps>(slot[u], sw[u]);
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    bool same = has[u];
#pragma unroll
    for (int w = 0; w < N; w++) same &= (sw[u][w] == key[u][w]);
    if (same) slot_accumulate<DeviceOps>(slot[u] + N, sw[u][N], edge[u], count[u], ctr);
    dq.push(t, ctr, my_new, act[u] && !same, key[u], c[u], edge[u], count[u]);
  }
}

// --------------------------------------------------------------------------
// build_kernel
// --------------------------------------------------------------------------
// One WARP owns one tile: it stages, classifies, erodes, compacts and inserts on
// its own, synchronising with __syncwarp only -- the eight warps of a CTA never
// wait for each other, so staging in one warp overlaps table probes in another.
#ifndef LASV_BIN_U
#define LASV_BIN_U 2     // rounds of 32 k-mers in flight per warp in the partition kernel
#endif
#ifndef LASV_BIN_CTAS
#define LASV_BIN_CTAS 4  // resident CTAs per SM the partition kernel is compiled for
#endif
template <int MODE>
struct TileCfg {
  // window starts per tile; region = T + 128 bases, a whole number of 32-lane rounds
  static constexpr int T = (MODE == MODE_ROUTE) ? 384 : 896;
  static constexpr int R = T + TileGeom::HALO_L + TileGeom::HALO_R;
  static constexpr int CHUNKS = R / 16;  // 16-base chunks: 32 or 64 -> one or two per lane
  static constexpr int WORDS = R / 32;   // validity words: 16 or 32 -> at most one per lane
  // k-mers per lane in flight in the heavy phase (independent table probes / bin reservations)
  static constexpr int U = (MODE == MODE_BIN) ? LASV_BIN_U : (MODE == MODE_FUSED) ? 2 : 1;
  static constexpr int MIN_CTAS = (MODE == MODE_ROUTE) ? 3 : (MODE == MODE_BIN) ? LASV_BIN_CTAS : 4;
};

template <int N, int MODE>
struct WarpTile {
  using Cfg = TileCfg<MODE>;
  uint32_t P[Cfg::CHUNKS + TileGeom::PADW + 2];
  uint32_t PR[Cfg::CHUNKS + TileGeom::PADW + 2];
  uint32_t BD[Cfg::WORDS + 4];
  uint16_t pos[Cfg::T];
  // ROUTE only: the tile's records wait here until per-destination space is reserved
  uint32_t stash[(MODE == MODE_ROUTE) ? Cfg::T * (2 * N + 1) : 1];
  uint8_t dest[(MODE == MODE_ROUTE) ? Cfg::T : 4];
  uint32_t bin_count[16];
  uint32_t bin_rank[16];
  unsigned long long tile_base[16];
  DeferItem<N> defer[(MODE == MODE_FUSED) ? DEFER_CAP : 1];
};

template <int N, int MODE>
__global__ void __launch_bounds__(THREADS, TileCfg<MODE>::MIN_CTAS)
build_kernel(const BuildParams p, const uint64_t n_tiles) {
  using Cfg = TileCfg<MODE>;
  constexpr int T = Cfg::T, R = Cfg::R, CHUNKS = Cfg::CHUNKS, WORDS = Cfg::WORDS, U = Cfg::U;
  static_assert(CHUNKS % 32 == 0 && WORDS <= 32 && R % 64 == 0, "tile must map onto whole warp rounds");
  constexpr int RECW = 2 * N + 1;
  constexpr int WARPS = THREADS / 32;

  extern __shared__ __align__(16) unsigned char sTilesRaw[];  // WARPS tiles (dynamic: ROUTE exceeds 48 KB)
  WarpTile<N, MODE> &wt = reinterpret_cast<WarpTile<N, MODE> *>(sTilesRaw)[threadIdx.x >> 5];
  const unsigned lane = lane_id();
  const bool has_qual = p.qual != nullptr;
  const int k = p.k;

  if (lane < TileGeom::PADW + 2) {  // pad words: zeros around the code strings, all-bad behind BD
    if (lane < TileGeom::PADW) { wt.P[lane] = 0; wt.PR[lane] = 0; }
    else { wt.P[CHUNKS + lane] = 0; wt.PR[CHUNKS + lane] = 0; }
    wt.BD[WORDS + lane] = ~0u;
  }

  unsigned long long my_new = 0;    // keys this thread created
  unsigned long long my_kmers = 0;  // lane 0: occurrences in this warp's tiles
  DeferQueue<N> dq{wt.defer, 0u};

  const uint64_t warp_global = (uint64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const uint64_t warp_stride = (uint64_t)gridDim.x * WARPS;

  for (uint64_t tile = warp_global; tile < n_tiles; tile += warp_stride) {
    __syncwarp();  // previous tile fully consumed before its strings are overwritten
    // ---- stage: 16 bytes of both streams per lane and round, classify, pack
#pragma unroll
    for (int c = (int)lane; c < CHUNKS; c += 32) {
      const long long g = (long long)(tile * (uint64_t)T) - TileGeom::HALO_L + 16ll * c;
      const uint32_t inr = chunk_inrange16(g, p.lead, p.n_bytes);
      uint4 sv = make_uint4(0, 0, 0, 0), qv = make_uint4(0, 0, 0, 0);
      if (inr) {
        sv = __ldg(reinterpret_cast<const uint4 *>(p.seq + g));
        if (has_qual) qv = __ldg(reinterpret_cast<const uint4 *>(p.qual + g));
      }
      uint32_t code32, bad16;
      classify16(sv, qv, has_qual, p.cutoff, inr, code32, bad16);
      wt.P[TileGeom::PADW + c] = code32;
      wt.PR[TileGeom::PADW + (CHUNKS - 1 - c)] = rc_code32(code32);
      reinterpret_cast<uint16_t *>(wt.BD)[c ^ 1] = (uint16_t)bad16;  // big-endian halves of a 32-bit word
    }
    if (MODE == MODE_ROUTE && lane < 16) { wt.bin_count[lane] = 0; wt.bin_rank[lane] = 0; }
    __syncwarp();

    // ---- window validity by word-parallel erosion (one word per lane), then an
    //      ordered dense list of the good window starts (warp scan of popcounts)
    uint32_t vbits = 0;
    if ((int)lane < WORDS) vbits = valid_window_word(wt.BD, (int)lane, k) & tile_word_mask((int)lane, T);
    const uint32_t vcnt = (uint32_t)__popc(vbits);
    uint32_t incl = vcnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, incl, o);
      if ((int)lane >= o) incl += y;
    }
    const int M = (int)__shfl_sync(FULL, incl, 31);
    {
      uint32_t idx = incl - vcnt;
      while (vbits) {
        const int b = __clz((int)vbits);
        wt.pos[idx++] = (uint16_t)(32 * (int)lane + b - TileGeom::HALO_L);
        vbits &= ~(0x80000000u >> b);
      }
    }
    __syncwarp();
    if (lane == 0 && (MODE != MODE_FUSED || p.pass == 0)) my_kmers += (unsigned long long)M;

    unsigned long long rec_base = 0;
    if (MODE == MODE_RECORDS) {
      if (lane == 0 && M) rec_base = atomicAdd(p.rec_count, (unsigned long long)M);
      rec_base = __shfl_sync(FULL, rec_base, 0);
    }

    // ---- heavy phase over the dense list of good windows, U rounds of 32 at a time
    for (int base = 0; base < M; base += 32 * U) {
      __syncwarp();
      uint64_t key[U][N];
      uint32_t edge[U], count[U], hc[U];
      bool lead[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int i = base + 32 * u + (int)lane;
        const bool active = i < M;
        edge[u] = 0;
        count[u] = 1;
        hc[u] = 0;
        if (active) {
          const int j = TileGeom::HALO_L + (int)wt.pos[i];
          edge[u] = kmer_occurrence<N>(wt.P, wt.PR, wt.BD, R, j, k, key[u]);
        } else {
#pragma unroll
          for (int w = 0; w < N; w++) key[u][w] = 0;
        }
        lead[u] = active;
        if (MODE != MODE_RECORDS) hc[u] = lookup3_key<N>(key[u], 0).c;
        if (MODE == MODE_FUSED || MODE == MODE_BIN) {
          bool mine = active;
          if (MODE == MODE_FUSED && p.n_pass > 1)
            mine = active && (int)((hc[u] & p.table.bucket_mask) >> p.pass_shift) == p.pass;
          // identical keys of a round (homopolymer and tandem runs) become one update
          lead[u] = warp_aggregate<N>(mine, key[u], edge[u], count[u]);
        }
      }
      if (MODE == MODE_FUSED) {
        upsert_batch<N, U>(p.table, p.ctr, dq, key, hc, edge, count, lead, my_new);
      } else if (MODE == MODE_BIN) {
        uint32_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const uint32_t bin = (hc[u] & p.bins.bin_mask) >> p.bins.bin_shift;
          const uint32_t ci = stripe_counter_at(p.bins, bin >> p.bins.block_shift, bin & ((1u << p.bins.block_shift) - 1u));
          idx[u] = lead[u] ? atomicAdd(&p.bins.count[ci], 1u) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (!lead[u]) continue;
          const uint32_t bin = (hc[u] & p.bins.bin_mask) >> p.bins.bin_shift;
          if (idx[u] < p.bins.cap) {
            store_bin_record<N>(stripe_record_ptr(p.bins, bin, idx[u]), p.bins.rec_words, key[u], hc[u], edge[u], count[u]);
          } else if (p.bins.spill_ok) {  // stripe full (skewed batch): straight into the table
            const int r = table_upsert_hashed<N, DeviceOps>(p.table, p.ctr, key[u], hc[u], edge[u], count[u]);
            my_new += (r == 1);
          } else {  // the owner's spill area
            const uint32_t owner = bin >> p.bins.block_shift;
            const uint32_t si = atomicAdd(&p.bins.count[stripe_counter_at(p.bins, owner, 1u << p.bins.block_shift)], 1u);
            if (si < p.bins.spill_cap)
              store_bin_record<N>(spill_record_ptr(p.bins, owner, si), p.bins.rec_words, key[u], hc[u], edge[u], count[u]);
            else
              atomicExch(p.rec_overflow, 1ull);
          }
        }
        __syncwarp();
      } else if (MODE == MODE_RECORDS) {
        const int i = base + (int)lane;
        if (lead[0]) {
          const unsigned long long idx = rec_base + (unsigned long long)i;
          if (idx < p.rec_capacity) store_record<N>(p.rec_out + idx * RECW, key[0], edge[0], 1u);
          else atomicExch(p.rec_overflow, 1ull);
        }
      } else {  // MODE_ROUTE
        const int i = base + (int)lane;
        if (lead[0]) {
          const uint32_t d = (hc[0] >> p.owner_shift) & (uint32_t)(p.n_dest - 1);
          store_record<N>(wt.stash + i * RECW, key[0], edge[0], 1u);
          wt.dest[i] = (uint8_t)d;
          atomicAdd(&wt.bin_count[d], 1u);
        }
      }
    }

    if (MODE == MODE_ROUTE) {
      __syncwarp();
      if ((int)lane < p.n_dest) {
        const uint32_t c = wt.bin_count[lane];
        wt.tile_base[lane] = c ? atomicAdd(p.rec_count + lane, (unsigned long long)c) : 0ull;
      }
      __syncwarp();
      for (int i = (int)lane; i < M; i += 32) {
        const uint32_t d = wt.dest[i];
        const unsigned long long idx = wt.tile_base[d] + atomicAdd(&wt.bin_rank[d], 1u);
        if (idx < p.rec_capacity) {
          uint32_t *dst = p.rec_out + ((uint64_t)d * p.rec_capacity + idx) * RECW;
#pragma unroll
          for (int w = 0; w < RECW; w++) dst[w] = wt.stash[i * RECW + w];
        } else {
          atomicExch(p.rec_overflow, 1ull);
        }
      }
    }
  }

  if (MODE == MODE_FUSED) dq.flush(p.table, p.ctr, my_new);

  // ---- warp-level counters
  if (MODE == MODE_FUSED || MODE == MODE_BIN) {
#pragma unroll
    for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(FULL, my_new, o);
    if (lane == 0 && my_new) atomicAdd(&p.ctr->unique_kmers, my_new);
  }
  if (lane == 0 && my_kmers) atomicAdd(&p.ctr->kmers_inserted, my_kmers);
}

// --------------------------------------------------------------------------
// partitioned insert: prefix over the bins, then K3 region by region
// --------------------------------------------------------------------------
// All warps of the grid sweep the bins in the same order, bin by bin, so that at
// any moment the whole GPU works inside one or two table regions (<= 64 MB each:
// TLB-, L2- and DRAM-row-friendly).  A bin is cut into chunks of 32 records and
// chunk j of bin b belongs to warp (j + first_b) mod n_warps, first_b = number of
// chunks of all earlier bins (a bin = the stripes of all senders for one region): a round robin over the concatenation of the bins
// that every warp derives by itself from the bin counters -- no prefix pass, no
// search.  A warp collects U of its chunks into a batch, loads the records together
// and runs them through upsert_batch.  (A three-stage software pipeline with L2
// prefetches of records and table lines was tried and was slower: the kernel is
// bound by scattered DRAM write-backs -- tools/ubench "sweep": ~21 G dirty
// sectors/s whatever the instruction that dirties them -- not by load latency,
// and FEWER resident warps do better, hence the small grid.)
template <int U>
struct BinBatch {
  uint64_t at[U];  // record index of this lane in each chunk
  bool live[U];
};

template <int N, int U>
__global__ void __launch_bounds__(THREADS, 4) insert_bins_kernel(const BinView b, const TableView t,
                                                             GraphCounters *ctr) {
  __shared__ DeferItem<N> sDefer[THREADS / 32][DEFER_CAP];
  DeferQueue<N> dq{sDefer[threadIdx.x >> 5], 0u};
  unsigned long long my_new = 0;
  const unsigned lane = lane_id();
  const uint32_t n_warps = (gridDim.x * THREADS) >> 5;
  const uint32_t warp = (blockIdx.x * THREADS + threadIdx.x) >> 5;

  BinBatch<U> s0;  // being collected
  int n_have = 0;  // chunks in s0, warp-uniform
#pragma unroll
  for (int u = 0; u < U; u++) { s0.at[u] = 0; s0.live[u] = false; }

  auto advance = [&]() {
    uint64_t key[U][N];
    uint32_t c[U], edge[U], count[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      c[u] = 0; edge[u] = 0; count[u] = 0;
#pragma unroll
      for (int w = 0; w < N; w++) key[u][w] = 0;
      if (s0.live[u]) load_bin_record<N>(b.recs + s0.at[u] * b.rec_words, b.rec_words, key[u], c[u], edge[u], count[u]);
    }
    upsert_batch<N, U>(t, ctr, dq, key, c, edge, count, s0.live, my_new);
#pragma unroll
    for (int u = 0; u < U; u++) s0.live[u] = false;
    n_have = 0;
  };

  // sweep position g < n_bins: region g / n_src of sender g % n_src; then one spill stripe per sender
  const uint32_t src_mask = (1u << b.src_shift) - 1u;
  const uint32_t n_sweep = b.n_bins + (b.spill_cap ? (1u << b.src_shift) : 0u);
  auto count_of = [&](uint32_t g) -> uint32_t {
    if (g < b.n_bins) {
      const uint32_t c = __ldg(&b.count[stripe_counter_at(b, g & src_mask, g >> b.src_shift)]);
      return c > b.cap ? b.cap : c;
    }
    const uint32_t c = __ldg(&b.count[stripe_counter_at(b, g - b.n_bins, b.bins_per_src)]);
    return c > b.spill_cap ? b.spill_cap : c;
  };
  auto record_of = [&](uint32_t g, uint32_t i) -> uint64_t {
    if (g < b.n_bins) return stripe_record_at(b, (g & src_mask) * b.bins_per_src + (g >> b.src_shift), i);
    return spill_record_at(b, g - b.n_bins, i);
  };
  uint32_t first = 0;  // (chunks of all earlier stripes) mod n_warps, warp-uniform
  for (uint32_t b0 = 0; b0 < n_sweep; b0 += 32) {
    __syncwarp();
    // 32 stripes at a time, one per lane: chunk counts, their running sum, and whether THIS warp owns a
    // chunk of the lane's stripe -- the warp then visits only the stripes where it has work (a sweep
    // that looks at every stripe costs 0.2 ms per launch, which small batches cannot amortise)
    const uint32_t cnt_l = b0 + lane < n_sweep ? count_of(b0 + lane) : 0u;
    const uint32_t chunks_l = (cnt_l + 31u) >> 5;
    uint32_t incl = chunks_l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, incl, o);
      if ((int)lane >= o) incl += y;
    }
    const uint32_t first_l = (first + (incl - chunks_l)) % n_warps;         // first chunk index base of my stripe
    const uint32_t j0_l = warp >= first_l ? warp - first_l : warp + n_warps - first_l;
    unsigned todo = __ballot_sync(FULL, j0_l < chunks_l);
    while (todo) {
      const int tb = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t cnt = __shfl_sync(FULL, cnt_l, tb);
      const uint32_t n_chunks = (cnt + 31u) >> 5;
      for (uint32_t j = __shfl_sync(FULL, j0_l, tb); j < n_chunks; j += n_warps) {
        const uint32_t i = j * 32u + lane;
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (u == n_have) {
            s0.live[u] = i < cnt;
            s0.at[u] = record_of(b0 + (uint32_t)tb, i);
          }
        }
        if (++n_have == U) advance();
      }
    }
    first = (first + __shfl_sync(FULL, incl, 31)) % n_warps;
  }
  if (n_have) advance();
  dq.flush(t, ctr, my_new);
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(FULL, my_new, o);
  if (lane == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// --------------------------------------------------------------------------
// insert_records_kernel: K3 over (2N+1)-word records
// --------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(THREADS, 4) insert_records_kernel(const uint32_t *__restrict__ recs,
                                                                const uint64_t n_recs,
                                                                const TableView t,
                                                                GraphCounters *ctr) {
  constexpr int RECW = 2 * N + 1;
  constexpr int U = 2;
  __shared__ DeferItem<N> sDefer[THREADS / 32][DEFER_CAP];
  DeferQueue<N> dq{sDefer[threadIdx.x >> 5], 0u};
  unsigned long long my_new = 0;
  const uint64_t warp_global = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * THREADS) >> 5;
  for (uint64_t base = warp_global * 32 * U; base < n_recs; base += n_warps * 32 * U) {
    __syncwarp();
    const uint64_t i0 = base + lane_id();
    uint64_t key[U][N];
    uint32_t c[U], edge[U], count[U];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = i0 + 32ull * u;
      act[u] = i < n_recs;
      edge[u] = 0; count[u] = 0;
#pragma unroll
      for (int w = 0; w < N; w++) key[u][w] = 0;
      if (act[u]) load_record<N>(recs + i * RECW, key[u], edge[u], count[u]);
      c[u] = lookup3_key<N>(key[u], 0).c;
      act[u] = warp_aggregate<N>(act[u], key[u], edge[u], count[u]);
    }
    upsert_batch<N, U>(t, ctr, dq, key, c, edge, count, act, my_new);
  }
  dq.flush(t, ctr, my_new);
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(FULL, my_new, o);
  if (lane_id() == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// --------------------------------------------------------------------------
// read_stats_kernel: one LANE per read
// --------------------------------------------------------------------------
// Loader bookkeeping (file_reader.c:460-470,579): reads without a run of k good
// bases, bases in such runs, number of runs, run-length histogram.  A lane walks
// its read in aligned 8-byte words of both streams; per base: class lookup in a
// 32-bit constant, signed quality compare, run counter.  (One warp per read with
// ballots over 32 bases cost 49 thread instructions per base, this costs ~12.)
constexpr int HIST_SMEM = 1024;

__global__ void __launch_bounds__(THREADS) read_stats_kernel(
    const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual,
    const uint64_t *__restrict__ offsets, const uint64_t n_reads, const int sep, const int k,
    const int cutoff, ReadStats *stats, unsigned long long *hist, const uint32_t hist_size) {
  __shared__ uint32_t sHist[HIST_SMEM];
  for (int i = threadIdx.x; i < HIST_SMEM; i += THREADS) sHist[i] = 0;
  __syncthreads();
  const bool has_qual = qual != nullptr;
  // A=0x41 C=0x43 G=0x47 T=0x54 after folding case: bits 1, 3, 7, 20 of the 0x40..0x5F block
  constexpr uint32_t ACGT = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 20);
  unsigned long long bad_reads = 0, bases_read = 0, bases_loaded = 0, contigs = 0, reads = 0;

  auto end_run = [&](uint32_t run, bool &has) {
    if (run >= (uint32_t)k) {
      has = true;
      bases_loaded += run;
      contigs++;
      const uint32_t b = run < hist_size - 1 ? run : hist_size - 1;
      if (b < HIST_SMEM) atomicAdd(&sHist[b], 1u);
      else atomicAdd(&hist[b], 1ull);
    }
  };

  for (uint64_t r = (uint64_t)blockIdx.x * THREADS + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * THREADS) {
    const uint64_t beg = offsets[r];
    const uint64_t len = offsets[r + 1] - beg - (uint64_t)sep;
    reads++;
    bases_read += len;
    uint32_t run = 0;
    bool has = false;
    const uintptr_t s0 = reinterpret_cast<uintptr_t>(seq) + beg, s1 = s0 + len;
    const uintptr_t q0 = reinterpret_cast<uintptr_t>(qual) + beg;
    // seq and qual share their alignment modulo 16 (checked by the host), so one word walk serves both
    for (uintptr_t wa = s0 & ~(uintptr_t)7; wa < s1; wa += 8) {
      const uint64_t sw = __ldg(reinterpret_cast<const uint64_t *>(wa));
      const uint64_t qw = has_qual ? __ldg(reinterpret_cast<const uint64_t *>(wa - s0 + q0)) : 0ull;
      const int lo = wa < s0 ? (int)(s0 - wa) : 0;
      const int hi = (s1 - wa) < 8 ? (int)(s1 - wa) : 8;
      for (int i = lo; i < hi; i++) {
        const uint32_t u = (uint32_t)(sw >> (8 * i)) & 0xDFu;
        bool good = ((u & 0xE0u) == 0x40u) && ((ACGT >> (u & 31u)) & 1u);
        if (has_qual) good = good && ((int)(int8_t)(qw >> (8 * i)) > cutoff);
        if (good) {
          run++;
        } else {
          end_run(run, has);
          run = 0;
        }
      }
    }
    end_run(run, has);
    if (!has) bad_reads++;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HIST_SMEM; i += THREADS) {
    const uint32_t v = sHist[i];
    if (v && (uint32_t)i < hist_size) atomicAdd(&hist[i], (unsigned long long)v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    reads += __shfl_xor_sync(FULL, reads, o);
    bad_reads += __shfl_xor_sync(FULL, bad_reads, o);
    bases_read += __shfl_xor_sync(FULL, bases_read, o);
    bases_loaded += __shfl_xor_sync(FULL, bases_loaded, o);
    contigs += __shfl_xor_sync(FULL, contigs, o);
  }
  if (lane_id() == 0) {
    if (reads) atomicAdd(&stats->n_reads, reads);
    if (bad_reads) atomicAdd(&stats->bad_reads, bad_reads);
    if (bases_read) atomicAdd(&stats->bases_read, bases_read);
    if (bases_loaded) atomicAdd(&stats->bases_loaded, bases_loaded);
    if (contigs) atomicAdd(&stats->n_contigs, contigs);
  }
}

// --------------------------------------------------------------------------
// table-wide kernels (streaming over slots / buckets)
// --------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(THREADS) table_summary_kernel(const TableView t, TableSummary *out) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  unsigned long long n = 0, sc = 0, se = 0, fp = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_slots;
       i += (uint64_t)gridDim.x * THREADS) {
    const uint32_t bucket = (uint32_t)(i / t.W);
    const uint64_t *slot = slot_ptr<N>(t, bucket, (uint32_t)(i - (uint64_t)bucket * t.W));
    const uint64_t m = slot[N];
    if (meta_status(m) == ST_EMPTY || meta_status(m) == ST_PRUNED) continue;  // db_node_check_status_not_pruned
    uint64_t key[N];
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = slot[w];
    n++;
    sc += meta_covg(m);
    se += __popc(meta_edges(m));
    fp += node_fingerprint<N>(key, meta_covg(m), meta_edges(m));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    n += __shfl_xor_sync(FULL, n, o);
    sc += __shfl_xor_sync(FULL, sc, o);
    se += __shfl_xor_sync(FULL, se, o);
    fp += __shfl_xor_sync(FULL, fp, o);
  }
  if (lane_id() == 0 && n) {
    atomicAdd(&out->n_nodes, n);
    atomicAdd(&out->sum_covg, sc);
    atomicAdd(&out->sum_edge_bits, se);
    atomicAdd(&out->fingerprint, fp);
  }
}

// slots per bucket that hold a node of the graph (occupied and not pruned): one warp per bucket
template <int N>
__global__ void __launch_bounds__(THREADS) bucket_counts_kernel(const TableView t, const uint64_t first,
                                                               const uint64_t n_buckets, uint32_t *counts) {
  // buckets [first, first + n_buckets); counts[] is relative to `first`
  const uint64_t warp0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
  for (uint64_t b = warp0; b < n_buckets; b += nwarps) {
    uint32_t c = 0;
    for (uint32_t s = lane_id(); s < t.W; s += 32) {
      const uint64_t m = slot_ptr<N>(t, (uint32_t)(first + b), s)[N];
      c += (meta_status(m) != ST_EMPTY && meta_status(m) != ST_PRUNED);  // what dump_ctx_kernel writes
    }
    c = __reduce_add_sync(FULL, c);
    if (lane_id() == 0) counts[b] = c;
  }
}

// Rewrite every bucket, in place, from lines of tagged slots to W hole-free
// consecutive Elements in slot order (status none -- pruned stays pruned --, pad bytes zero) at the start of
// its NL*128 bytes: afterwards a bucket IS a reference bucket (hash_table.c:69-122
// finds every key) and the table is copied out bucket by bucket.  One warp per
// bucket; 32 slots are read into registers before any of them is overwritten, and
// a destination never runs ahead of a source: slot s of the line layout starts at
// byte 128*(s/G) + 8 + SLOT*(s%G) >= SLOT*s + 8, its destination at SLOT*d, d <= s.
template <int N>
__global__ void __launch_bounds__(THREADS) finalize_kernel(const TableView t, int16_t *next_element) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT;
  const uint64_t n_buckets = (uint64_t)t.bucket_mask + 1;
  const uint64_t warp0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
  const unsigned lane = lane_id();
  for (uint64_t b = warp0; b < n_buckets; b += nwarps) {
    uint8_t *bucket = bucket_base(t, (uint32_t)b);
    uint32_t filled = 0;
    for (uint32_t s0 = 0; s0 < t.W; s0 += 32) {
      const uint32_t s = s0 + lane;
      uint64_t key[N];
      uint64_t m = 0;
      if (s < t.W) {
        const uint64_t *slot = slot_ptr<N>(t, (uint32_t)b, s);
        m = slot[N];
#pragma unroll
        for (int w = 0; w < N; w++) key[w] = slot[w];
      }
      const bool occ = meta_status(m) != ST_EMPTY;
      const unsigned om = __ballot_sync(FULL, occ);
      __syncwarp();
      if (occ) {
        const uint32_t d = filled + __popc(om & lanemask_lt());
        uint64_t *dst = reinterpret_cast<uint64_t *>(bucket + (uint64_t)d * SLOT);
#pragma unroll
        for (int w = 0; w < N; w++) dst[w] = key[w];
        dst[N] = meta_pack(meta_covg(m), meta_edges(m), meta_status(m) == ST_PRUNED ? ST_PRUNED : ST_NONE, 0);
      }
      filled += __popc(om);
      __syncwarp();
    }
    // everything behind the last Element, line slack included, becomes zero
    uint64_t *words = reinterpret_cast<uint64_t *>(bucket);
    const uint32_t w_end = t.NL * (LINE_BYTES / 8);
    for (uint32_t w = filled * (SLOT / 8) + lane; w < w_end; w += 32) words[w] = 0;
    if (lane == 0 && next_element) next_element[b] = (int16_t)filled;
  }
}

// A shard of a hash-sharded graph after finalize_kernel: keys that do not sit in their first-choice bucket were
// rehashed inside the shard; their place in the merged reference table is elsewhere (capi.cu:
// lasv_multi_graph_export_table).  One warp per bucket: such Elements are copied to `out` (whole Elements, in no
// particular order), the others are squeezed together, next_element is corrected.
template <int N>
__global__ void __launch_bounds__(THREADS) extract_misplaced_kernel(const TableView t, int16_t *next_element, uint8_t *out,
                                                                   unsigned long long *n_out, const uint64_t cap) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT;
  const uint64_t n_buckets = (uint64_t)t.bucket_mask + 1;
  const uint64_t warp0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
  const unsigned lane = lane_id();
  for (uint64_t b = warp0; b < n_buckets; b += nwarps) {
    uint8_t *bucket = bucket_base(t, (uint32_t)b);
    const uint32_t filled = (uint32_t)next_element[b];
    uint32_t kept = 0;
    for (uint32_t s0 = 0; s0 < filled; s0 += 32) {
      const uint32_t s = s0 + lane;
      uint64_t key[N];
      uint64_t m = 0;
      bool have = s < filled;
      if (have) {
        const uint64_t *slot = reinterpret_cast<const uint64_t *>(bucket + (uint64_t)s * SLOT);
        m = slot[N];
#pragma unroll
        for (int w = 0; w < N; w++) key[w] = slot[w];
      } else {
#pragma unroll
        for (int w = 0; w < N; w++) key[w] = 0;
      }
      const bool home = have && (lookup3_key<N>(key, 0).c & t.bucket_mask) == (uint32_t)b;
      const unsigned km = __ballot_sync(FULL, home), mm = __ballot_sync(FULL, have && !home);
      unsigned long long at = 0;
      if (lane == 0 && mm) at = atomicAdd(n_out, (unsigned long long)__popc(mm));
      at = __shfl_sync(FULL, at, 0);
      __syncwarp();
      if (home) {
        uint64_t *dst = reinterpret_cast<uint64_t *>(bucket + (uint64_t)(kept + __popc(km & lanemask_lt())) * SLOT);
#pragma unroll
        for (int w = 0; w < N; w++) dst[w] = key[w];
        dst[N] = m;
      } else if (have) {
        const unsigned long long i = at + __popc(mm & lanemask_lt());
        if (i < cap) {
          uint64_t *dst = reinterpret_cast<uint64_t *>(out + i * SLOT);
#pragma unroll
          for (int w = 0; w < N; w++) dst[w] = key[w];
          dst[N] = m;
        }
      }
      kept += __popc(km);
      __syncwarp();
    }
    if (kept != filled) {
      uint64_t *words = reinterpret_cast<uint64_t *>(bucket);
      for (uint32_t w = kept * (SLOT / 8) + lane; w < filled * (SLOT / 8); w += 32) words[w] = 0;
      if (lane == 0) next_element[b] = (int16_t)kept;
    }
  }
}

// .ctx records (element.c:894-910) in table order: kmer[N] | uint32 covg | uint8 edges
template <int N>
__global__ void __launch_bounds__(THREADS) dump_ctx_kernel(const TableView t, const uint64_t first,
                                                          const uint64_t n_buckets,
                                                          const uint64_t *__restrict__ bucket_offsets,
                                                          uint8_t *__restrict__ out) {
  // buckets [first, first + n_buckets); bucket_offsets[] and out are relative to `first`
  constexpr uint32_t REC = 8 * N + 5;
  const uint64_t warp0 = ((uint64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * THREADS) >> 5;
  const unsigned lane = lane_id();
  for (uint64_t bb = warp0; bb < n_buckets; bb += nwarps) {
    const uint64_t b = first + bb;
    uint64_t o = bucket_offsets[bb];
    for (uint32_t s0 = 0; s0 < t.W; s0 += 32) {
      const uint32_t s = s0 + lane;
      uint64_t key[N];
      uint64_t m = 0;
      if (s < t.W) {
        const uint64_t *slot = slot_ptr<N>(t, (uint32_t)b, s);
        m = slot[N];
#pragma unroll
        for (int w = 0; w < N; w++) key[w] = slot[w];
      }
      const bool occ = meta_status(m) != ST_EMPTY && meta_status(m) != ST_PRUNED;  // db_node_check_status_not_pruned (element.c:795-802)
      const unsigned om = __ballot_sync(FULL, occ);
      if (occ) {
        uint8_t *dst = out + (o + __popc(om & lanemask_lt())) * REC;
#pragma unroll
        for (int w = 0; w < N; w++) {
#pragma unroll
          for (int by = 0; by < 8; by++) dst[8 * w + by] = (uint8_t)(key[w] >> (8 * by));
        }
        const uint32_t c = meta_covg(m);
        dst[8 * N + 0] = (uint8_t)c;
        dst[8 * N + 1] = (uint8_t)(c >> 8);
        dst[8 * N + 2] = (uint8_t)(c >> 16);
        dst[8 * N + 3] = (uint8_t)(c >> 24);
        dst[8 * N + 4] = (uint8_t)meta_edges(m);
      }
      o += __popc(om);
    }
  }
}

// load packed .ctx records into the table (file_reader.c:1686-1703: insert key,
// add coverage, OR edges)
template <int N>
__global__ void __launch_bounds__(THREADS, 5) ingest_ctx_kernel(const uint8_t *__restrict__ recs,
                                                            const uint64_t n_recs, const TableView t,
                                                            GraphCounters *ctr) {
  constexpr uint32_t REC = 8 * N + 5;
  unsigned long long my_new = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_recs;
       i += (uint64_t)gridDim.x * THREADS) {
    const uint8_t *src = recs + i * REC;
    uint64_t key[N];
#pragma unroll
    for (int w = 0; w < N; w++) {
      uint64_t v = 0;
#pragma unroll
      for (int by = 0; by < 8; by++) v |= (uint64_t)src[8 * w + by] << (8 * by);
      key[w] = v;
    }
    const uint32_t c = (uint32_t)src[8 * N] | ((uint32_t)src[8 * N + 1] << 8) |
                       ((uint32_t)src[8 * N + 2] << 16) | ((uint32_t)src[8 * N + 3] << 24);
    const int r = table_upsert<N, DeviceOps>(t, ctr, key, src[8 * N + 4], c);
    my_new += (r == 1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(FULL, my_new, o);
  if (lane_id() == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// merge an uploaded reference-layout Element[] (any hole pattern) into the table
template <int N>
__global__ void __launch_bounds__(THREADS, 5) ingest_elements_kernel(const uint8_t *__restrict__ elems,
                                                                 const uint64_t n_slots,
                                                                 const TableView t, GraphCounters *ctr) {
  constexpr uint32_t STRIDE = 8 * N + 8;
  unsigned long long my_new = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_slots;
       i += (uint64_t)gridDim.x * THREADS) {
    const uint64_t *slot = reinterpret_cast<const uint64_t *>(elems + i * STRIDE);
    const uint64_t m = slot[N];
    if (meta_status(m) == ST_EMPTY) continue;
    uint64_t key[N];
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = slot[w];
    const int r = table_upsert<N, DeviceOps>(t, ctr, key, meta_edges(m), meta_covg(m));
    my_new += (r == 1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(FULL, my_new, o);
  if (lane_id() == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// Second pass of the import: an Element the caller's cleaning left `pruned` (element.h:57-75; edges reset, not part
// of the graph any more) stays pruned -- the reference keeps such nodes pruned when more reads are loaded onto a
// cleaned graph and leaves them out of dumps and supernodes.  Runs after ingest_elements_kernel (no insert in flight).
template <int N>
__global__ void __launch_bounds__(THREADS) mark_pruned_elements_kernel(const uint8_t *__restrict__ elems, const uint64_t n_slots,
                                                                      const TableView t) {
  constexpr uint32_t STRIDE = 8 * N + 8;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * THREADS) {
    const uint64_t *src = reinterpret_cast<const uint64_t *>(elems + i * STRIDE);
    if (meta_status(src[N]) != ST_PRUNED) continue;
    uint64_t key[N];
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = src[w];
    uint64_t m = 0, *slot = nullptr;
    if (table_find<N, DeviceOps>(t, key, m, &slot) && meta_status(m) == ST_NONE)
      DeviceOps::xor32(reinterpret_cast<uint32_t *>(slot + N) + 1, (ST_NONE ^ ST_PRUNED) << 8);
  }
}

// hash_table_find for a batch of canonical keys (keys: n x N words).
// finalized != 0: reference layout (scan each bucket from slot 0 to the first
// empty slot, hash_table.c:84-111); else the hashed-start device layout.
template <int N>
__global__ void __launch_bounds__(THREADS) find_kernel(const TableView t,
                                                      const uint64_t *__restrict__ keys,
                                                      const uint64_t n,
                                                      uint32_t *covg, uint8_t *edges, uint8_t *found) {
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * THREADS) {
    uint64_t key[N];
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = keys[i * N + w];
    uint64_t m = 0;
    bool hit = false;
    if (!t.finalized) {
      hit = table_find<N, DeviceOps>(t, key, m);
    } else {
      hit = table_find_finalized<N, DeviceOps>(t, key, m);
    }
    found[i] = hit ? 1 : 0;
    covg[i] = hit ? meta_covg(m) : 0;
    edges[i] = hit ? (uint8_t)meta_edges(m) : 0;
  }
}

__global__ void widen_u32_kernel(const uint32_t *in, unsigned long long *out, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i];
}

// --------------------------------------------------------------------------
// synthetic reads + random-access microbenchmark (bench / test infrastructure)
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) synth_kernel(const SynthParams p, const uint64_t first_read,
                                                       const uint64_t n_reads, uint8_t *seq,
                                                       uint8_t *qual) {
  const uint64_t stride = (uint64_t)p.read_len + 1;
  const uint64_t total = n_reads * stride;
  for (uint64_t t = (uint64_t)blockIdx.x * THREADS + threadIdx.x; t < total;
       t += (uint64_t)gridDim.x * THREADS) {
    const uint64_t r = t / stride;
    const uint32_t i = (uint32_t)(t - r * stride);
    uint8_t s, q;
    synth_byte(p, first_read + r, i, s, q);
    seq[t] = s;
    qual[t] = q;
  }
}

// Uniform random 32-byte read-modify-write over a buffer of n_sectors sectors:
// the random-access ceiling the insert kernel is compared with.
__global__ void __launch_bounds__(THREADS) random_rmw_kernel(uint64_t *buf, const uint64_t n_sectors,
                                                            const uint64_t n_ops, const uint64_t seed,
                                                            const int mode) {
  // mode 0: 16-byte read + 32-bit atomic (one slot probe + coverage update)
  // mode 1: read only     mode 2: atomic only     mode 3: 64-byte read (two sectors) + atomic
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_ops;
       i += (uint64_t)gridDim.x * THREADS) {
    const uint64_t h = mix64(seed + i);
    const uint64_t s = __umul64hi(h, n_sectors);
    uint64_t *p = buf + s * 4;
    if (mode == 2) {
      atomicAdd(reinterpret_cast<unsigned int *>(p + 2), 1u);
      continue;
    }
    uint64_t a = ld_strong_u64(p), b = ld_strong_u64(p + 1);
    if (mode == 3 && s + 1 < n_sectors) { a ^= ld_strong_u64(p + 4); b ^= ld_strong_u64(p + 5); }
    if (mode == 1) { acc += a ^ b; continue; }
    if (a + b != 0x123456789ull)  // data-dependent so the loads cannot be dropped
      atomicAdd(reinterpret_cast<unsigned int *>(p + 2), 1u);
  }
  if (mode == 1 && acc == 0x123456789ull) buf[0] = acc;
}

// bucket (rehash 0) of every record: scaffolding for locality experiments
template <int N>
__global__ void __launch_bounds__(THREADS) record_buckets_kernel(const uint32_t *__restrict__ recs,
                                                                const uint64_t n_recs,
                                                                const uint32_t bucket_mask,
                                                                uint32_t *__restrict__ buckets) {
  constexpr int RECW = 2 * N + 1;
  for (uint64_t i = (uint64_t)blockIdx.x * THREADS + threadIdx.x; i < n_recs;
       i += (uint64_t)gridDim.x * THREADS) {
    uint64_t key[N];
    uint32_t e, c;
    load_record<N>(recs + i * RECW, key, e, c);
    buckets[i] = lookup3_key<N>(key, 0).c & bucket_mask;
  }
}

int g_sm_count = 0;

template <typename F>
void dispatch_n(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

template <typename K>
int resident_grid(K kernel, int cap_per_sm = 8, size_t dyn_smem = 0) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, dyn_smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > cap_per_sm) per_sm = cap_per_sm;
  return per_sm * sm_count();
}

}  // namespace

int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

void launch_build(int N, BuildMode mode, const BuildParams &p, cudaStream_t st) {
  if (p.n_bytes == 0) return;
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    auto go = [&](auto kernel, int T, size_t smem) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      const uint64_t n_tiles = (p.n_bytes + T - 1) / T;  // one tile per warp and round
      uint64_t grid = (uint64_t)resident_grid(kernel, p.ctas_per_sm > 0 ? p.ctas_per_sm : 8, smem);
      const uint64_t need = (n_tiles + THREADS / 32 - 1) / (THREADS / 32);
      if (grid > need) grid = need;
      kernel<<<(unsigned)grid, THREADS, smem, st>>>(p, n_tiles);
    };
    constexpr int W8 = THREADS / 32;
    if (mode == MODE_FUSED)
      go(build_kernel<NN, MODE_FUSED>, TileCfg<MODE_FUSED>::T, W8 * sizeof(WarpTile<NN, MODE_FUSED>));
    else if (mode == MODE_BIN)
      go(build_kernel<NN, MODE_BIN>, TileCfg<MODE_BIN>::T, W8 * sizeof(WarpTile<NN, MODE_BIN>));
    else if (mode == MODE_RECORDS)
      go(build_kernel<NN, MODE_RECORDS>, TileCfg<MODE_RECORDS>::T, W8 * sizeof(WarpTile<NN, MODE_RECORDS>));
    else
      go(build_kernel<NN, MODE_ROUTE>, TileCfg<MODE_ROUTE>::T, W8 * sizeof(WarpTile<NN, MODE_ROUTE>));
  });
}

void launch_insert_records(int N, const uint32_t *recs, uint64_t n_recs, const TableView &t,
                           GraphCounters *ctr, cudaStream_t st) {
  if (!n_recs) return;
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    uint64_t grid = (uint64_t)resident_grid(insert_records_kernel<NN>);
    const uint64_t need = (n_recs + THREADS - 1) / THREADS;
    if (grid > need) grid = need;
    insert_records_kernel<NN><<<(unsigned)grid, THREADS, 0, st>>>(recs, n_recs, t, ctr);
  });
}

void launch_insert_bins(int N, const BinView &b, const TableView &t, GraphCounters *ctr, int ctas_per_sm,
                        cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    // few resident warps on purpose (see the kernel's comment); LASV_B200_INSERT_CTAS overrides
    int per_sm = ctas_per_sm > 0 ? ctas_per_sm : 3;
    if (const char *e = getenv("LASV_B200_INSERT_CTAS")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
    int u = 2;
    if (const char *e = getenv("LASV_B200_INSERT_U")) u = atoi(e);
    if (u == 1) insert_bins_kernel<NN, 1><<<resident_grid(insert_bins_kernel<NN, 1>, per_sm), THREADS, 0, st>>>(b, t, ctr);
    else insert_bins_kernel<NN, 2><<<resident_grid(insert_bins_kernel<NN, 2>, per_sm), THREADS, 0, st>>>(b, t, ctr);
  });
}

void launch_read_stats(const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                       uint64_t n_reads, int sep, int k, int cutoff, ReadStats *stats,
                       unsigned long long *hist, uint32_t hist_size, cudaStream_t st) {
  if (!n_reads) return;
  uint64_t grid = (uint64_t)resident_grid(read_stats_kernel);
  const uint64_t need = (n_reads + THREADS - 1) / THREADS;  // one lane per read
  if (grid > need) grid = need;
  read_stats_kernel<<<(unsigned)grid, THREADS, 0, st>>>(seq, qual, offsets, n_reads, sep, k, cutoff,
                                                       stats, hist, hist_size);
}

void launch_table_summary(int N, const TableView &t, TableSummary *out, cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    table_summary_kernel<NN><<<resident_grid(table_summary_kernel<NN>), THREADS, 0, st>>>(t, out);
  });
}

void launch_bucket_counts(int N, const TableView &t, uint64_t first, uint64_t n_buckets, uint32_t *counts,
                          cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    bucket_counts_kernel<NN><<<resident_grid(bucket_counts_kernel<NN>), THREADS, 0, st>>>(t, first, n_buckets, counts);
  });
}

void launch_finalize(int N, const TableView &t, int16_t *next_element, cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    finalize_kernel<NN><<<resident_grid(finalize_kernel<NN>), THREADS, 0, st>>>(t, next_element);
  });
}

void launch_extract_misplaced(int N, const TableView &t, int16_t *next_element, uint8_t *out, unsigned long long *n_out,
                              uint64_t cap, cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    extract_misplaced_kernel<NN><<<resident_grid(extract_misplaced_kernel<NN>), THREADS, 0, st>>>(t, next_element, out, n_out, cap);
  });
}

void launch_dump_ctx(int N, const TableView &t, uint64_t first, uint64_t n_buckets, const uint64_t *bucket_offsets,
                     uint8_t *out, cudaStream_t st) {
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    dump_ctx_kernel<NN><<<resident_grid(dump_ctx_kernel<NN>), THREADS, 0, st>>>(t, first, n_buckets, bucket_offsets, out);
  });
}

void launch_ingest_ctx(int N, const uint8_t *recs, uint64_t n_recs, const TableView &t,
                       GraphCounters *ctr, cudaStream_t st) {
  if (!n_recs) return;
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    ingest_ctx_kernel<NN><<<resident_grid(ingest_ctx_kernel<NN>), THREADS, 0, st>>>(recs, n_recs, t, ctr);
  });
}

void launch_ingest_elements(int N, const uint8_t *elements, uint64_t n_slots, const TableView &t,
                            GraphCounters *ctr, cudaStream_t st) {
  if (!n_slots) return;
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    ingest_elements_kernel<NN><<<resident_grid(ingest_elements_kernel<NN>), THREADS, 0, st>>>(
        elements, n_slots, t, ctr);
    mark_pruned_elements_kernel<NN><<<resident_grid(mark_pruned_elements_kernel<NN>), THREADS, 0, st>>>(elements, n_slots, t);
  });
}

void launch_find(int N, const TableView &t, const uint64_t *keys, uint64_t n, uint32_t *covg,
                 uint8_t *edges, uint8_t *found, cudaStream_t st) {
  if (!n) return;
  dispatch_n(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    uint64_t grid = (n + THREADS - 1) / THREADS;
    const uint64_t cap = (uint64_t)resident_grid(find_kernel<NN>);
    if (grid > cap) grid = cap;
    find_kernel<NN><<<(unsigned)grid, THREADS, 0, st>>>(t, keys, n, covg, edges, found);
  });
}

size_t exclusive_scan_tmp_bytes(uint64_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const unsigned long long *)nullptr,
                                (unsigned long long *)nullptr, (long long)n);
  return bytes + n * sizeof(unsigned long long);
}

// out[i] = sum_{j<i} in[j]; tmp must hold exclusive_scan_tmp_bytes(n)
void launch_exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *tmp,
                                      size_t tmp_bytes, cudaStream_t st) {
  if (!n) return;
  unsigned long long *wide = reinterpret_cast<unsigned long long *>(tmp);
  widen_u32_kernel<<<sm_count() * 4, 256, 0, st>>>(in, wide, n);
  void *cub_tmp = reinterpret_cast<uint8_t *>(tmp) + n * sizeof(unsigned long long);
  size_t cub_bytes = tmp_bytes - n * sizeof(unsigned long long);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, wide, reinterpret_cast<unsigned long long *>(out),
                                (long long)n, st);
}

void launch_synth(const SynthParams &p, uint64_t first_read, uint64_t n_reads, uint8_t *seq,
                  uint8_t *qual, cudaStream_t st) {
  if (!n_reads) return;
  synth_kernel<<<sm_count() * 8, THREADS, 0, st>>>(p, first_read, n_reads, seq, qual);
}

void synth_host(const SynthParams &p, uint64_t first_read, uint64_t n_reads, uint8_t *seq,
                uint8_t *qual) {
  const uint64_t stride = (uint64_t)p.read_len + 1;
  for (uint64_t r = 0; r < n_reads; r++)
    for (uint32_t i = 0; i <= p.read_len; i++)
      synth_byte(p, first_read + r, i, seq[r * stride + i], qual[r * stride + i]);
}

void launch_random_rmw(uint64_t *buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed, int mode,
                       cudaStream_t st) {
  random_rmw_kernel<<<sm_count() * 8, THREADS, 0, st>>>(buf, n_sectors, n_ops, seed, mode);
}

void launch_record_buckets(int N, const uint32_t *recs, uint64_t n_recs, uint32_t bucket_mask,
                           uint32_t *buckets, cudaStream_t st) {
  if (!n_recs) return;
  dispatch_n(N, [&](auto n) {
    constexpr int NN = decltype(n)::value;
    record_buckets_kernel<NN><<<sm_count() * 8, THREADS, 0, st>>>(recs, n_recs, bucket_mask, buckets);
  });
}

}  // namespace lasv
