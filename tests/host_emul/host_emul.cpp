// tests/host_emul/host_emul.cpp -- TEST HARNESS ONLY.
//
// Runs the kernels' own __host__ __device__ arithmetic (lasv_b200/csrc/
// kmer_core.cuh: classify16, tile bit-string layout, window_good, extract_kmer,
// kmer_occurrence, lookup3_key; table.cuh: table_upsert / table_find* with a
// sequential HostOps) on the CPU, tile by tile in the order build_kernel walks
// them, so that index arithmetic, packing, canonicalisation, edge formula,
// probing and finalisation are checked against the oracle WITHOUT a GPU.
// What it cannot check -- atomics under contention, warp intrinsics -- is
// covered by the -m gpu parity tests.  The product never links this file.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../lasv_b200/csrc/kmer_core.cuh"
#include "../../lasv_b200/csrc/table.cuh"
#include "../../lasv_b200/csrc/supernodes.cuh"
#include "../../lasv_b200/csrc/supernodes_host.h"
#include "../../lasv_b200/csrc/cleaning.cuh"
#include "../../lasv_b200/csrc/cleaning_host.h"
#include "../../lasv_b200/csrc/branches.cuh"
#include "../../lasv_b200/csrc/branches_host.h"
#include "../../lasv_b200/csrc/groups.cuh"
#include "../../lasv_b200/csrc/ranking.cuh"

using namespace lasv;

namespace {

struct HostOps {
  static uint64_t load(const uint64_t *p) { return *p; }
  static void load16(const uint64_t *p, uint64_t &a, uint64_t &b) { a = p[0]; b = p[1]; }
  static void store(uint64_t *p, uint64_t v) { *p = v; }
  static void store_tag(uint8_t *p, uint32_t v) { *p = (uint8_t)v; }
  static uint64_t cas(uint64_t *p, uint64_t expect, uint64_t want) {
    uint64_t old = *p;
    if (old == expect) *p = want;
    return old;
  }
  static void xor32(uint32_t *p, uint32_t v) { *p ^= v; }
  static void add32(uint32_t *p, uint32_t v) { *p += v; }
  static void or32(uint32_t *p, uint32_t v) { *p |= v; }
  static void and32(uint32_t *p, uint32_t v) { *p &= v; }
  static void fence() {}
  static void backoff() {}
  static void count(unsigned long long *p, unsigned long long v) { *p += v; }
  static void flag(unsigned long long *p) { *p = 1; }
};

struct Tile {
  int T, R, CHUNKS, WORDS;
  std::vector<uint32_t> P, PR, BD;
  explicit Tile(int t)
      : T(t), R(t + TileGeom::HALO_L + TileGeom::HALO_R), CHUNKS(R / 16), WORDS(R / 32) {
    P.assign(R / 16 + TileGeom::PADW + 2, 0);
    PR.assign(R / 16 + TileGeom::PADW + 2, 0);
    BD.assign(WORDS + 4, ~0u);
  }
  // mirrors build_kernel's issue_stage + stage phase
  void stage(const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes, uint32_t lead, uint64_t tile,
             int cutoff) {
    for (int c = 0; c < CHUNKS; c++) {
      uint4 sv = {0, 0, 0, 0}, qv = {0, 0, 0, 0};
      const long long g = (long long)(tile * (uint64_t)T) - TileGeom::HALO_L + 16ll * c;
      const uint32_t inr = chunk_inrange16(g, lead, n_bytes);
      if (inr) {
        const uint64_t left = n_bytes - (uint64_t)g;
        const size_t take = left >= 16 ? 16 : (size_t)left;
        uint8_t sb[16], qb[16];
        memset(sb, 0xAB, 16);  // bytes past the stream are garbage on the device too
        memset(qb, 0xCD, 16);
        memcpy(sb, seq + g, take);
        if (qual) memcpy(qb, qual + g, take);
        memcpy(&sv, sb, 16);
        memcpy(&qv, qb, 16);
      }
      uint32_t code32, bad16;
      classify16(sv, qv, qual != nullptr, cutoff, inr, code32, bad16);
      P[TileGeom::PADW + c] = code32;
      PR[TileGeom::PADW + (CHUNKS - 1 - c)] = rc_code32(code32);
      reinterpret_cast<uint16_t *>(BD.data())[c ^ 1] = (uint16_t)bad16;
    }
  }
};

template <int N, class Sink>
void walk(const uint8_t *seq_in, const uint8_t *qual_in, uint64_t n_in, int k, int cutoff, int T,
          Sink &&sink, uint32_t lead = 0) {
  // a caller pointer that is `lead` bytes past a 16-byte boundary: the kernels see
  // the aligned-down buffer, whose first `lead` bytes are foreign (here: valid-looking
  // bases, so that a wrong in-range mask would show up as extra k-mers)
  std::vector<uint8_t> sbuf(lead + n_in + 32, (uint8_t)'A'), qbuf(lead + n_in + 32, (uint8_t)'I');
  if (n_in) {
    memcpy(sbuf.data() + lead, seq_in, n_in);
    if (qual_in) memcpy(qbuf.data() + lead, qual_in, n_in);
  }
  const uint8_t *seq = sbuf.data();
  const uint8_t *qual = qual_in ? qbuf.data() : nullptr;
  const uint64_t n_bytes = n_in + lead;
  Tile tile(T);
  const uint64_t n_tiles = (n_bytes + T - 1) / T;
  for (uint64_t t = 0; t < n_tiles; t++) {
    tile.stage(seq, qual, n_bytes, lead, t, cutoff);
    // mirrors the validity + ordered compaction phase: one word per "thread"
    for (int w = 0; w < tile.WORDS; w++) {
      uint32_t vbits = valid_window_word(tile.BD.data(), w, k) & tile_word_mask(w, T);
      while (vbits) {
        const int b = clz32(vbits);
        vbits &= ~(0x80000000u >> b);
        const int j = 32 * w + b;  // == HALO_L + pos
        uint64_t key[N];
        const uint32_t edge = kmer_occurrence<N>(tile.P.data(), tile.PR.data(), tile.BD.data(), tile.R, j, k, key);
        sink(key, edge);
      }
    }
  }
}

template <int N>
int build_fused(const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes, int k, int cutoff, int L,
                int W, int max_rehash, int T, uint8_t *table, GraphCounters *ctr) {
  const TableView tv = make_table_view(table, N, L, W, max_rehash, false);
  walk<N>(seq, qual, n_bytes, k, cutoff, T & 0xFFFF, [&](const uint64_t(&key)[N], uint32_t edge) {
    ctr->kmers_inserted++;
    const int r = table_upsert<N, HostOps>(tv, ctr, key, edge, 1u);
    if (r == 1) ctr->unique_kmers++;
  }, (uint32_t)(T >> 16) & 15u);  // bits 16..19 of T: emulated pointer misalignment
  return ctr->table_full ? -3 : 0;
}

template <int N>
uint64_t extract(const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes, int k, int cutoff, int T,
                 uint32_t *out, uint64_t cap) {
  uint64_t n = 0;
  walk<N>(seq, qual, n_bytes, k, cutoff, T, [&](const uint64_t(&key)[N], uint32_t edge) {
    if (n < cap) {
      uint32_t *dst = out + n * (2 * N + 1);
      for (int i = 0; i < N; i++) {
        dst[2 * i] = (uint32_t)key[i];
        dst[2 * i + 1] = (uint32_t)(key[i] >> 32);
      }
      dst[2 * N] = (edge & 0xFFu) | (1u << 8);
    }
    n++;
  });
  return n;
}

template <int N>
int insert_records(const uint32_t *recs, uint64_t n, int L, int W, int max_rehash, uint8_t *table,
                   GraphCounters *ctr) {
  const TableView tv = make_table_view(table, N, L, W, max_rehash, false);
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t *src = recs + i * (2 * N + 1);
    uint64_t key[N];
    for (int w = 0; w < N; w++) key[w] = (uint64_t)src[2 * w] | ((uint64_t)src[2 * w + 1] << 32);
    const uint32_t m = src[2 * N];
    // the records path hands the first-bucket hash over instead of hashing twice
    const int r = table_upsert_hashed<N, HostOps>(tv, ctr, key, lookup3_key<N>(key, 0).c, m & 0xFFu, m >> 8);
    if (r == 1) ctr->unique_kmers++;
  }
  return ctr->table_full ? -3 : 0;
}

// finalize_kernel restated with its 32-slot chunking (read a chunk, then write)
template <int N>
void finalize(uint8_t *table, int L, int W, int16_t *next) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT;
  const TableView tv = make_table_view(table, N, L, W, 15, false);
  const uint64_t nb = 1ull << L;
  for (uint64_t b = 0; b < nb; b++) {
    uint8_t *bucket = bucket_base(tv, (uint32_t)b);
    uint32_t filled = 0;
    for (uint32_t s0 = 0; s0 < (uint32_t)W; s0 += 32) {
      uint64_t keys[32][N], metas[32];
      bool occ[32];
      for (uint32_t l = 0; l < 32; l++) {
        const uint32_t s = s0 + l;
        metas[l] = 0;
        if (s < (uint32_t)W) {
          const uint64_t *slot = slot_ptr<N>(tv, (uint32_t)b, s);
          metas[l] = slot[N];
          for (int w = 0; w < N; w++) keys[l][w] = slot[w];
        }
        occ[l] = meta_status(metas[l]) != ST_EMPTY;
      }
      uint32_t rank = 0;
      for (uint32_t l = 0; l < 32; l++) {
        if (!occ[l]) continue;
        uint64_t *dst = reinterpret_cast<uint64_t *>(bucket + (uint64_t)(filled + rank) * SLOT);
        for (int w = 0; w < N; w++) dst[w] = keys[l][w];
        dst[N] = meta_pack(meta_covg(metas[l]), meta_edges(metas[l]),
                           meta_status(metas[l]) == ST_PRUNED ? ST_PRUNED : ST_NONE, 0);
        rank++;
      }
      filled += rank;
    }
    memset(bucket + (uint64_t)filled * SLOT, 0, (size_t)tv.NL * LINE_BYTES - (size_t)filled * SLOT);
    if (next) next[b] = (int16_t)filled;
  }
}

// slot s of bucket b -> out[b*W + s] (reference Element bytes), either arrangement
template <int N>
void gather(uint8_t *table, int L, int W, int finalized, uint8_t *out) {
  constexpr uint32_t SLOT = LineLayout<N>::SLOT;
  const TableView tv = make_table_view(table, N, L, W, 15, finalized != 0);
  for (uint64_t b = 0; b < (1ull << L); b++)
    for (uint32_t s = 0; s < (uint32_t)W; s++)
      memcpy(out + (b * (uint64_t)W + s) * SLOT, slot_ptr<N>(tv, (uint32_t)b, s), SLOT);
}

template <int N>
void find(uint8_t *table, int L, int W, int max_rehash, int finalized, const uint64_t *keys,
          uint64_t n, uint32_t *covg, uint8_t *edges, uint8_t *found) {
  const TableView tv = make_table_view(table, N, L, W, max_rehash, finalized != 0);
  for (uint64_t i = 0; i < n; i++) {
    uint64_t key[N], m = 0;
    for (int w = 0; w < N; w++) key[w] = keys[i * N + w];
    const bool hit = finalized ? table_find_finalized<N, HostOps>(tv, key, m)
                               : table_find<N, HostOps>(tv, key, m);
    found[i] = hit;
    covg[i] = hit ? meta_covg(m) : 0;
    edges[i] = hit ? (uint8_t)meta_edges(m) : 0;
  }
}

// the supernode kernels of supernode_kernels.cu, slot by slot / start by start, + the replay
// sn_paths_ranked (ranking_kernels.cu): the path records by list ranking instead of one walk per start
bool g_ranked = false;
struct RankedState {  // what sn_paths_ranked leaves on the device for the per-node writer
  std::vector<uint64_t> bits, rank;
  std::vector<RkElem> el;
};
template <int N>
void ranked_enumerate(const TableView &tv, int k, const std::vector<uint64_t> &starts, std::vector<SnPath> &recs,
                      std::vector<uint8_t> &covered, uint64_t &dangling, RankedState *keep = nullptr) {
  const uint64_t n_slots = ((uint64_t)tv.bucket_mask + 1) * tv.W, n_words = (n_slots + 63) / 64;
  std::vector<uint64_t> bits(n_words, 0), rank(n_words, 0);
  for (uint64_t q = 0; q < n_slots; q++)  // rk_flags_kernel
    if ((sn_classify<N, HostOps>(tv, q) & 3u) == 3u) bits[q >> 6] |= 1ull << (q & 63);
  uint64_t n_int = 0;
  for (uint64_t w = 0; w < n_words; w++) { rank[w] = n_int; n_int += rk_popc64(bits[w]); }  // rk_popc_kernel + scan
  const RkMap map{bits.data(), rank.data()};
  std::vector<RkElem> el[2] = {std::vector<RkElem>(2 * n_int + 1), std::vector<RkElem>(2 * n_int + 1)};
  for (uint64_t q = 0; q < n_slots; q++) {  // rk_init_kernel
    if (!((bits[q >> 6] >> (q & 63)) & 1ull)) continue;
    for (uint32_t o = 0; o < 2; o++)
      if (!rk_init<N, HostOps>(tv, k, map, q, o, el[0][2u * rk_cid(map, q) + o])) dangling++;
  }
  int cur = 0;
  uint64_t prev = ~0ull;
  for (int round = 0; round < 64 && n_int; round++) {  // rk_jump_kernel
    uint64_t active = 0;
    for (uint64_t i = 0; i < 2 * n_int; i++) active += rk_jump(el[cur].data(), (uint32_t)i, el[cur ^ 1][i]);
    cur ^= 1;
    if (active == 0 || active == prev) break;
    prev = active;
  }
  for (uint64_t s : starts) {  // rk_paths_kernel
    const uint64_t q = s >> 3;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    SnPath rec;
    bool d = false, u = false;
    if (rk_path_of_start<N, HostOps>(tv, k, map, el[cur].data(), q, key, meta_edges(m), (uint32_t)(s >> 2) & 1u,
                                     (uint32_t)s & 3u, rec, d, u))
      recs.push_back(rec);
    dangling += d || u;
  }
  for (uint64_t q = 0; q < n_slots; q++)  // rk_covered_kernel
    if ((bits[q >> 6] >> (q & 63)) & 1ull) covered[q] = el[cur][2u * rk_cid(map, q)].succ == RK_DONE ? 1 : 0;
  if (keep) {
    keep->el.swap(el[cur]);
    keep->bits.swap(bits);
    keep->rank.swap(rank);
  }
}

template <int N>
int supernodes(uint8_t *table, int k, int L, int W, int finalized, const char *path, uint64_t *counts8) {
  const TableView tv = make_table_view(table, N, L, W, 15, finalized != 0);
  const uint64_t n_slots = (1ull << L) * (uint64_t)W;
  const uint64_t max_steps = 2 * n_slots + 2;
  std::vector<uint8_t> covered(n_slots, 0);
  std::vector<SnPath> recs;
  uint64_t nodes = 0, interior = 0, starts = 0, dangling = 0;
  std::vector<uint64_t> rk_starts;
  for (uint64_t q = 0; q < n_slots; q++) {  // sn_count_kernel + sn_starts_kernel + sn_paths_kernel
    const uint32_t v = sn_classify<N, HostOps>(tv, q);
    nodes += v & 1u;
    interior += (v >> 1) & 1u;
    starts += v >> 2;
    if (!(v & 1u) || (v & 2u)) continue;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    const uint32_t e = meta_edges(m);
    for (uint32_t bit = 0; bit < 8; bit++) {
      if (!((e >> bit) & 1u)) continue;
      if (g_ranked) { rk_starts.push_back((q << 3) | bit); continue; }
      SnPath rec;
      bool d = false;
      if (sn_walk_path<N, HostOps>(tv, k, q, key, e, bit >> 2, bit & 3u, max_steps, covered.data(), rec, d))
        recs.push_back(rec);
      dangling += d;
    }
  }
  RankedState ranked;
  if (g_ranked) ranked_enumerate<N>(tv, k, rk_starts, recs, covered, dangling, &ranked);
  const uint64_t n_paths = recs.size();
  uint64_t uncovered = 0;
  for (uint64_t q = 0; q < n_slots; q++) {  // sn_uncovered_kernel + sn_cycles_kernel
    if (covered[q] || (sn_classify<N, HostOps>(tv, q) & 3u) != 3u) continue;
    uncovered++;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    SnPath rec;
    bool d = false;
    if (sn_walk_cycle<N, HostOps>(tv, k, q, key, meta_edges(m), max_steps, rec, d)) recs.push_back(rec);
    dangling += d;
  }
  std::vector<SnPrint> prints;
  std::vector<uint64_t> end_keys;
  const SnReplayStats st = sn_replay(recs, k, n_slots, prints, g_ranked ? &end_keys : nullptr);
  std::vector<char> image(sn_layout_fasta(prints, k) + 1);
  if (g_ranked) {  // rk_write_frames_kernel + rk_write_nodes_kernel: one thread per record frame, one per interior node
    std::vector<RkPrintKey> keys;
    for (size_t i = 0; i < prints.size(); i++)
      if (end_keys[i] != SN_NONE) keys.push_back({end_keys[i], (uint64_t)i});
    std::sort(keys.begin(), keys.end(), [](const RkPrintKey &a, const RkPrintKey &b) { return a.key < b.key; });
    const RkMap map{ranked.bits.data(), ranked.rank.data()};
    for (size_t i = 0; i < prints.size(); i++)
      if (!rk_write_frame<N, HostOps>(tv, k, prints[i], end_keys[i], i, image.data())) dangling++;
    for (uint64_t q = 0; q < n_slots; q++)
      if ((ranked.bits[q >> 6] >> (q & 63)) & 1ull)
        rk_write_node<N, HostOps>(tv, k, map, ranked.el.data(), keys.data(), keys.size(), prints.data(), q, image.data());
  } else {
    for (size_t i = 0; i < prints.size(); i++)  // sn_write_kernel
      if (!sn_write_record<N, HostOps>(tv, k, prints[i], i, image.data())) dangling++;
  }
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  const bool ok = fwrite(image.data(), 1, image.size() - 1, fp) == image.size() - 1;
  if (fclose(fp) != 0 || !ok) return -1;
  counts8[0] = st.printed; counts8[1] = nodes; counts8[2] = interior; counts8[3] = starts;
  counts8[4] = n_paths; counts8[5] = recs.size() - n_paths; counts8[6] = st.never_printed; counts8[7] = dangling;
  (void)uncovered;
  return dangling ? -2 : 0;
}

// The cleaning as the library will run it: enumerate the paths (the supernode kernels' walks),
// take the reference's sequential decisions on the compressed graph (cleaning_host.h), apply the
// actions to the table (cleaning.cuh); once for the tips, once more for the low-coverage supernodes.
template <int N>
bool build_cl_graph(const TableView &tv, int k, uint32_t T, ClGraph &g) {
  const uint64_t n_slots = ((uint64_t)tv.bucket_mask + 1) * tv.W;
  const uint64_t max_steps = 2 * n_slots + 2;
  std::vector<uint8_t> covered(n_slots, 0);
  for (uint64_t q = 0; q < n_slots; q++) {
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    if (!m || meta_status(m) == ST_PRUNED) continue;
    const uint32_t e = meta_edges(m);
    if (e != 0 && !sn_interior(e)) g.add_node(q, meta_covg(m), e);
  }
  g.seal_nodes(n_slots);
  bool ok = true;
  std::vector<uint64_t> rk_starts;
  std::vector<SnPath> p_recs;
  std::vector<uint32_t> p_max;
  std::vector<uint64_t> p_low;
  const size_t n_nodes = g.nodes.size();
  for (size_t i = 0; i < n_nodes; i++) {
    const uint64_t q = g.nodes[i].q;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    const uint32_t e = meta_edges(m);
    for (uint32_t bit = 0; bit < 8; bit++) {
      if (!((e >> bit) & 1u)) continue;
      if (g_ranked) { rk_starts.push_back((q << 3) | bit); continue; }
      SnPath rec;
      bool d = false;
      if (sn_walk_path<N, HostOps>(tv, k, q, key, e, bit >> 2, bit & 3u, max_steps, covered.data(), rec, d)) {
        uint32_t mx = 0;
        uint64_t lo = SN_NONE;
        ok = ok && cl_path_coverage<N, HostOps>(tv, k, rec, T, mx, lo);
        p_recs.push_back(rec);
        p_max.push_back(mx);
        p_low.push_back(lo);
      }
      ok = ok && !d;
    }
  }
  if (g_ranked) {
    uint64_t dangling = 0;
    ranked_enumerate<N>(tv, k, rk_starts, p_recs, covered, dangling);
    ok = ok && !dangling;
    for (const SnPath &rec : p_recs) {
      uint32_t mx = 0;
      uint64_t lo = SN_NONE;
      ok = ok && cl_path_coverage<N, HostOps>(tv, k, rec, T, mx, lo);
      p_max.push_back(mx);
      p_low.push_back(lo);
    }
  }
  ok = g.add_paths(p_recs.data(), p_max.data(), p_low.data(), p_recs.size()) && ok;  // as the library does
  for (uint64_t q = 0; q < n_slots; q++) {
    if (covered[q] || (sn_classify<N, HostOps>(tv, q) & 3u) != 3u) continue;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, HostOps>(tv, q, key);
    if (meta_status(m) == ST_PRUNED) continue;
    SnPath rec;
    bool d = false;
    if (sn_walk_cycle<N, HostOps>(tv, k, q, key, meta_edges(m), max_steps, rec, d)) {
      ClCycle c;
      c.s_q = q;
      c.s_on = ((rec.bits >> SNB_S_N) & 3u) << 1;
      c.len = rec.len;
      c.s_covg = meta_covg(m);
      uint64_t lo = SN_NONE;
      ok = ok && cl_path_coverage<N, HostOps>(tv, k, rec, T, c.max_covg, lo);
      if (c.s_covg > 0 && c.s_covg <= T && q < lo) lo = q;
      c.min_low_q = lo;
      c.both_ways = (rec.bits & SNB_BOTH_WAYS) ? 1u : 0u;
      g.cycles.push_back(c);
    }
    ok = ok && !d;
  }
  return ok;
}

template <int N>
bool apply_cl_actions(const TableView &tv, int k, const ClActions &act) {
  bool ok = true;
  for (const auto &p : act.paths) ok = cl_prune_path<N, HostOps>(tv, k, p.q, p.on & 1u, p.on >> 1, p.len, p.keep_q) && ok;
  if (!act.paths.empty()) {  // cl_settle_kernel
    const uint64_t n_slots = ((uint64_t)tv.bucket_mask + 1) * tv.W;
    for (uint64_t q = 0; q < n_slots; q++) cl_settle_node<N, HostOps>(tv, q);
  }
  for (uint64_t q : act.nodes) cl_prune_node<N, HostOps>(tv, q);
  for (const auto &r : act.resets) cl_reset_edges<N, HostOps>(tv, r.q, r.mask);
  return ok;
}

template <int N>
int clean(uint8_t *table, int k, int L, int W, int finalized, int threshold, uint64_t *counts4) {
  const TableView tv = make_table_view(table, N, L, W, 15, finalized != 0);
  ClActions tips, sups;
  {
    ClGraph g;
    if (!build_cl_graph<N>(tv, k, (uint32_t)threshold, g)) return -2;
    g.clip_tips(k, tips);
    if (!apply_cl_actions<N>(tv, k, tips)) return -3;
  }
  {
    ClGraph g;  // the paths of the clipped graph
    if (!build_cl_graph<N>(tv, k, (uint32_t)threshold, g)) return -4;
    g.remove_low_coverage_supernodes(k, (uint32_t)threshold, sups);
    if (!apply_cl_actions<N>(tv, k, sups)) return -5;
  }
  counts4[0] = tips.tips;
  counts4[1] = tips.tip_edges;
  counts4[2] = sups.supernodes_pruned;
  counts4[3] = tips.nodes_pruned() + sups.nodes_pruned();
  return 0;
}

// Branch printing as the library runs it: the paths (supernode kernels' walks) -> compressed
// graph -> replay of the reference's traversal (branches_host.h) -> one record writer per branch
// (branches.cuh), optionally leaving the reference's `visited` marks in the table.
template <int N>
int branches(uint8_t *table, int k, int L, int W, int finalized, const char *path, int mark, uint64_t *counts6) {
  const TableView tv = make_table_view(table, N, L, W, 15, finalized != 0);
  if (mark) {  // br_reset_marks_kernel
    const uint64_t n_slots = ((uint64_t)tv.bucket_mask + 1) * tv.W;
    for (uint64_t q = 0; q < n_slots; q++) br_reset_mark<N, HostOps>(tv, q);
  }
  ClGraph g;
  if (!build_cl_graph<N>(tv, k, 0u, g)) return -2;
  std::vector<BrPrint> prints;
  BrReplay replay(g, k);
  const BrStats st = replay.run(prints);
  std::vector<char> image(st.image_bytes + 8);
  bool ok = true;
  for (const BrPrint &p : prints) ok = br_write_record<N, HostOps>(tv, k, p, image.data(), mark) && ok;  // br_write_kernel
  FILE *fp = fopen(path, "w");
  if (!fp) return -1;
  const bool wrote = fwrite(image.data(), 1, st.image_bytes, fp) == st.image_bytes;
  if (fclose(fp) != 0 || !wrote) return -1;
  counts6[0] = st.branches; counts6[1] = st.branching_nodes; counts6[2] = st.nodes; counts6[3] = st.refused;
  counts6[4] = st.cut; counts6[5] = st.inconsistent;
  return (ok && !st.inconsistent) ? 0 : -3;
}

// hc_check_kernel: `--health_check`
template <int N>
void health(uint8_t *table, int k, int L, int W, int finalized, uint64_t *counts5) {
  const TableView tv = make_table_view(table, N, L, W, 15, finalized != 0);
  const uint64_t n_slots = (1ull << L) * (uint64_t)W;
  for (int i = 0; i < 5; i++) counts5[i] = 0;
  for (uint64_t q = 0; q < n_slots; q++) {
    if (meta_status(HostOps::load(cl_meta_ptr<N, HostOps>(tv, q))) == ST_EMPTY) continue;
    uint32_t a = 0, b = 0;
    bool z = false;
    counts5[1] += hc_check_node<N, HostOps>(tv, k, q, a, b, z);
    counts5[0]++;
    counts5[2] += a;
    counts5[3] += b;
    counts5[4] += z;
  }
}

// insert_groups_kernel + insert_spilled_kernel (groups.cuh): records grouped by their first-choice bucket,
// every group updated in a STAGED COPY of its lines (the kernel's shared memory), copied back, and the
// records that found their bucket full inserted through the ordinary path afterwards.
template <int N>
int insert_grouped(const uint32_t *recs, uint64_t n, int L, int W, int max_rehash, uint32_t bpg, uint8_t *table,
                   GraphCounters *ctr, uint64_t *n_spilled) {
  constexpr int RECW = 2 * N + 1;
  const TableView tv = make_table_view(table, N, L, W, max_rehash, false);
  const uint64_t n_groups = ((uint64_t)tv.bucket_mask + 1) / bpg, gbytes = group_bytes(tv, bpg);
  std::vector<std::vector<uint64_t>> by_group(n_groups);  // the caller's sort by group
  for (uint64_t i = 0; i < n; i++) by_group[group_of_record<N>(tv, recs + i * RECW, bpg)].push_back(i);
  std::vector<uint8_t> staged(gbytes);
  std::vector<uint64_t> spill;
  for (uint64_t g = 0; g < n_groups; g++) {
    if (by_group[g].empty()) continue;
    memcpy(staged.data(), tv.lines + g * gbytes, gbytes);
    const TableView local = group_local_view(tv, staged.data(), g, bpg);
    for (uint64_t i : by_group[g]) {
      const int r = group_upsert<N, HostOps>(local, ctr, recs + i * RECW);
      if (r == 1) ctr->unique_kmers++;
      if (r < 0) spill.push_back(i);
    }
    memcpy(tv.lines + g * gbytes, staged.data(), gbytes);
  }
  for (uint64_t i : spill) {
    uint64_t key[N];
    uint32_t edge, count;
    group_load_record<N>(recs + i * RECW, key, edge, count);
    if (table_upsert<N, HostOps>(tv, ctr, key, edge, count) == 1) ctr->unique_kmers++;
  }
  *n_spilled = spill.size();
  return ctr->table_full ? -3 : 0;
}

}  // namespace

#define DISPATCH(N, call) ((N) == 1 ? call<1> : (N) == 2 ? call<2> : call<3>)

extern "C" {

// bytes of the device table for (k, L, W): 2^L buckets x ceil(W/G) lines x 128
uint64_t emul_table_bytes(int k, int L, int W) {
  const int N = 2 * k / 64 + 1;
  return table_bytes(make_table_view(nullptr, N, L, W, 15, false));
}

int emul_build_fused(const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes, int k, int cutoff,
                     int L, int W, int max_rehash, int T, uint8_t *table, uint64_t *counters20) {
  GraphCounters ctr;
  memset(&ctr, 0, sizeof ctr);
  const int N = 2 * k / 64 + 1;
  const int rc = DISPATCH(N, build_fused)(seq, qual, n_bytes, k, cutoff, L, W, max_rehash, T, table, &ctr);
  counters20[0] = ctr.unique_kmers;
  counters20[1] = ctr.kmers_inserted;
  counters20[2] = ctr.table_full;
  for (int r = 0; r < 16; r++) counters20[3 + r] = ctr.rehash_hist[r];
  return rc;
}

uint64_t emul_extract_records(const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes, int k,
                              int cutoff, int T, uint32_t *out, uint64_t cap) {
  const int N = 2 * k / 64 + 1;
  return DISPATCH(N, extract)(seq, qual, n_bytes, k, cutoff, T, out, cap);
}

int emul_insert_records(const uint32_t *recs, uint64_t n, int k, int L, int W, int max_rehash,
                        uint8_t *table, uint64_t *counters20) {
  GraphCounters ctr;
  memset(&ctr, 0, sizeof ctr);
  const int N = 2 * k / 64 + 1;
  const int rc = DISPATCH(N, insert_records)(recs, n, L, W, max_rehash, table, &ctr);
  counters20[0] = ctr.unique_kmers;
  counters20[2] = ctr.table_full;
  for (int r = 0; r < 16; r++) counters20[3 + r] = ctr.rehash_hist[r];
  return rc;
}

void emul_finalize(uint8_t *table, int k, int L, int W, int16_t *next) {
  const int N = 2 * k / 64 + 1;
  DISPATCH(N, finalize)(table, L, W, next);
}

void emul_gather(uint8_t *table, int k, int L, int W, int finalized, uint8_t *out) {
  const int N = 2 * k / 64 + 1;
  DISPATCH(N, gather)(table, L, W, finalized, out);
}

void emul_find(uint8_t *table, int k, int L, int W, int max_rehash, int finalized,
               const uint64_t *keys, uint64_t n, uint32_t *covg, uint8_t *edges, uint8_t *found) {
  const int N = 2 * k / 64 + 1;
  DISPATCH(N, find)(table, L, W, max_rehash, finalized, keys, n, covg, edges, found);
}

int emul_supernodes(uint8_t *table, int k, int L, int W, int finalized, const char *path, uint64_t *counts8) {
  const int N = 2 * k / 64 + 1;
  return DISPATCH(N, supernodes)(table, k, L, W, finalized, path, counts8);
}

int emul_clean(uint8_t *table, int k, int L, int W, int finalized, int threshold, uint64_t *counts4) {
  const int N = 2 * k / 64 + 1;
  return DISPATCH(N, clean)(table, k, L, W, finalized, threshold, counts4);
}

int emul_branches(uint8_t *table, int k, int L, int W, int finalized, const char *path, int mark, uint64_t *counts6) {
  const int N = 2 * k / 64 + 1;
  return DISPATCH(N, branches)(table, k, L, W, finalized, path, mark, counts6);
}

void emul_health(uint8_t *table, int k, int L, int W, int finalized, uint64_t *counts5) {
  const int N = 2 * k / 64 + 1;
  DISPATCH(N, health)(table, k, L, W, finalized, counts5);
}

int emul_insert_grouped(const uint32_t *recs, uint64_t n, int k, int L, int W, int max_rehash, uint32_t buckets_per_group,
                        uint8_t *table, uint64_t *counters20, uint64_t *n_spilled) {
  GraphCounters ctr;
  memset(&ctr, 0, sizeof ctr);
  const int N = 2 * k / 64 + 1;
  const int rc = DISPATCH(N, insert_grouped)(recs, n, L, W, max_rehash, buckets_per_group, table, &ctr, n_spilled);
  counters20[0] = ctr.unique_kmers;
  counters20[2] = ctr.table_full;
  for (int r = 0; r < 16; r++) counters20[3 + r] = ctr.rehash_hist[r];
  return rc;
}

// path enumeration by list ranking (ranking.cuh) instead of one walk per start, for everything that follows
void emul_set_ranking(int on) { g_ranked = on != 0; }

// threads for ClGraph::add_paths from this many path records on (tests force the threaded branch)
void emul_set_parallel_min(uint64_t n) { ClGraph::parallel_min() = n; }

// kmer_revcomp of supernodes.cuh, for known-answer checks
void emul_revcomp(const uint64_t *in, int k, uint64_t *out) {
  const int N = 2 * k / 64 + 1;
  if (N == 1) { uint64_t a[1] = {in[0]}, b[1]; kmer_revcomp<1>(a, k, b); out[0] = b[0]; }
  else if (N == 2) { uint64_t a[2] = {in[0], in[1]}, b[2]; kmer_revcomp<2>(a, k, b); out[0] = b[0]; out[1] = b[1]; }
  else { uint64_t a[3] = {in[0], in[1], in[2]}, b[3]; kmer_revcomp<3>(a, k, b); out[0] = b[0]; out[1] = b[1]; out[2] = b[2]; }
}

uint32_t emul_hash_c(const uint64_t *key, int N, uint32_t rehash) {
  if (N == 1) { uint64_t k1[1] = {key[0]}; return lookup3_key<1>(k1, rehash).c; }
  if (N == 2) { uint64_t k2[2] = {key[0], key[1]}; return lookup3_key<2>(k2, rehash).c; }
  uint64_t k3[3] = {key[0], key[1], key[2]};
  return lookup3_key<3>(k3, rehash).c;
}

}  // extern "C"
