// ranking.cuh -- path enumeration by LIST RANKING (pointer jumping), the alternative to one walk per path
// start (sn_walk_path, supernodes.cuh) for graphs whose paths are long: a cleaned graph is a handful of
// million-edge paths, and one thread per path then walks for seconds.  Opt-in (LASV_B200_SN_RANKING=1) until it has
// run on a device; the records it produces are the same SnPath records, so everything downstream -- the replay of
// supernodes_host.h, the cleaning and branch decisions, the writers -- is unchanged.
//
//   element   (interior node q, orientation o in which a walk crosses it); index 2 * cid(q) + o, cid = the rank
//             of q among the interior nodes in slot order (one bit per slot + a running count per 64 slots)
//   init      one table lookup per element: the node the walk steps to.  Interior -> successor element;
//             else the element is DONE and remembers that terminal (slot, arrival orientation, nucleotide back)
//   jump      ~log2(longest path) rounds, double buffered, no atomics: e takes over its successor's successor,
//             adds its length, keeps the lower "lowest slot" (ties: the earlier crossing, as the walk's `<`)
//   paths     one record per start of a non-interior node from the element its first step lands on
//   cycles    elements that never become DONE; their nodes are the "uncovered" ones the cycle walker takes
//
// Shared by ranking_kernels.cu and tests/host_emul (which checks the records against the walks').
#pragma once
#include "supernodes.cuh"

namespace lasv {

constexpr uint32_t RK_DONE = 0xFFFFFFFFu;

struct RkMap {           // the interior nodes
  const uint64_t *bits;  // one bit per slot
  const uint64_t *rank;  // interior nodes in the slots before each 64-slot word
};
LASV_HD uint32_t rk_popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popcll(x);
#else
  return (uint32_t)__builtin_popcountll(x);
#endif
}
LASV_HD uint32_t rk_cid(const RkMap &m, uint64_t q) {
  const uint64_t w = q >> 6;
  return (uint32_t)(m.rank[w] + rk_popc64(m.bits[w] & ((1ull << (q & 63)) - 1ull)));
}

struct RkElem {
  uint32_t succ;  // element the chain continues with, RK_DONE once the terminal is known
  uint32_t len;   // edges from this element's node to where succ points (DONE: to the terminal node)
  uint64_t minv;  // (lowest interior slot on that stretch, this node included) << 1 | orientation it is crossed in
  uint64_t term;  // DONE: terminal slot << 3 | arrival orientation << 2 | nucleotide that leads back
};

// element (q, o) of interior node q; false if its edge leads nowhere (the reference die()s)
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool rk_init(const TableView &t, int k, const RkMap &map, uint64_t q, uint32_t o, RkElem &e) {
  uint64_t key[N], key2[N];
  const uint64_t m = sn_load_node<N, Ops>(t, q, key);
  uint32_t o2 = 0, back = 0;
  uint64_t m2 = 0;
  const uint64_t q2 = sn_step<N, Ops>(t, k, key, o, low_bit4(sn_side(meta_edges(m), o)), key2, m2, o2, back);
  e.len = 1;
  e.minv = (q << 1) | o;
  e.term = 0;
  e.succ = RK_DONE;
  if (q2 == SN_NONE) return false;
  if (sn_interior(meta_edges(m2))) e.succ = 2u * rk_cid(map, q2) + o2;
  else e.term = (q2 << 3) | ((uint64_t)o2 << 2) | back;
  return true;
}

// one round for element i; true if it is still on its way afterwards
LASV_HD bool rk_jump(const RkElem *cur, uint32_t i, RkElem &out) {
  RkElem e = cur[i];
  if (e.succ != RK_DONE) {
    const RkElem s = cur[e.succ];
    const uint64_t len = (uint64_t)e.len + s.len;
    e.len = len > 0x7FFFFFFFull ? 0x7FFFFFFFu : (uint32_t)len;  // only cycles get there
    if ((s.minv >> 1) < (e.minv >> 1)) e.minv = s.minv;
    if (s.succ == RK_DONE) e.term = s.term;
    e.succ = s.succ;
  }
  out = e;
  return e.succ != RK_DONE;
}

// The record of the path that leaves non-interior node (s_q, key, edges) in orientation o along nuc -- what
// sn_walk_path returns, read off the ranked elements.  unfinished: the chain never reached a terminal
// (cannot happen for a chain entered from a non-interior node of a consistent graph).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool rk_path_of_start(const TableView &t, int k, const RkMap &map, const RkElem *el, uint64_t s_q,
                              const uint64_t (&s_key)[N], uint32_t s_edges, uint32_t o, uint32_t nuc, SnPath &rec,
                              bool &dangling, bool &unfinished) {
  uint64_t key2[N];
  uint32_t o2 = 0, back = 0, min_o = 0;
  uint64_t m2 = 0, min_q = SN_NONE, len = 1;
  dangling = unfinished = false;
  uint64_t q2 = sn_step<N, Ops>(t, k, s_key, o, nuc, key2, m2, o2, back);
  if (q2 == SN_NONE) { dangling = true; return false; }
  uint32_t e2 = meta_edges(m2);
  if (sn_interior(e2)) {
    const RkElem e = el[2u * rk_cid(map, q2) + o2];
    if (e.succ != RK_DONE) { unfinished = true; return false; }
    len += e.len;
    min_q = e.minv >> 1;
    min_o = (uint32_t)(e.minv & 1u);
    q2 = e.term >> 3;
    o2 = (uint32_t)(e.term >> 2) & 1u;
    back = (uint32_t)e.term & 3u;
    const uint32_t bucket = (uint32_t)(q2 / t.W);
    e2 = meta_edges(Ops::load(slot_ptr<N>(t, bucket, (uint32_t)(q2 - (uint64_t)bucket * t.W)) + N));
  }
  const uint64_t mine = ((2 * s_q + o) << 2) | nuc, twin = ((2 * q2 + (o2 ^ 1u)) << 2) | back;
  if (mine > twin) return false;
  const uint32_t s_side = sn_trigger_side(s_edges), t_side = sn_trigger_side(e2);
  rec.s_q = s_q;
  rec.t_q = q2;
  rec.min_q = min_q;
  rec.len = (uint32_t)len;
  rec.bits = (o ? SNB_S_O : 0u) | (nuc << SNB_S_N) | ((o2 ^ 1u) ? SNB_T_O : 0u) | (back << SNB_T_N) |
             (min_o ? SNB_MIN_O : 0u) | (s_side == o ? SNB_TRIG_S : 0u) | (t_side == (o2 ^ 1u) ? SNB_TRIG_T : 0u) |
             (s_side << SNB_S_SIDE) | (t_side << SNB_T_SIDE);
  return true;
}

// ------------------------------------------------------------------ writing long supernodes, one thread per NODE
// A supernode is one path; printed from one of its ends it is a walk that ends in a known state: the node at the
// other end, the orientation it is reached in, the nucleotide that leads back (SN_NONE for cycles, which keep the
// per-record writer).  Every ranked element (q, o) that walks in that direction carries exactly this state in
// `term` and its distance to it in `len`, so it can place the base node q contributes -- the last base of its
// k-mer read in orientation o -- without anyone walking the path: sequence index k - 1 + (edges - len).
struct RkPrintKey {
  uint64_t key;    // end state of a printed supernode
  uint64_t print;  // its index in the print list
};
LASV_HD uint64_t rk_end_key(uint64_t end_q, uint32_t arrival_o, uint32_t back) {
  return (end_q << 3) | ((uint64_t)(arrival_o & 1u) << 2) | (back & 3u);
}
LASV_HD bool rk_find_print(const RkPrintKey *keys, uint64_t n, uint64_t key, uint64_t &print) {
  uint64_t lo = 0, hi = n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (keys[mid].key < key) lo = mid + 1;
    else hi = mid;
  }
  if (lo < n && keys[lo].key == key) { print = keys[lo].print; return true; }
  return false;
}
template <int N>
LASV_HD uint32_t rk_last_base(const uint64_t (&key)[N], int k, uint32_t o) {  // last base of the k-mer read in orientation o
  return o ? 3u - kmer_first_base<N>(key, k) : (uint32_t)(key[N - 1] & 3u);
}

// the bases interior node q contributes (it may be crossed in both orientations by a path that turns round)
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void rk_write_node(const TableView &t, int k, const RkMap &map, const RkElem *el, const RkPrintKey *keys,
                           uint64_t n_keys, const SnPrint *prints, uint64_t q, char *image) {
  uint64_t key[N];
  if (!sn_load_node<N, Ops>(t, q, key)) return;
  const uint32_t cid = rk_cid(map, q);
  for (uint32_t o = 0; o < 2; o++) {
    const RkElem e = el[2u * cid + o];
    uint64_t i = 0;
    if (e.succ != RK_DONE || !rk_find_print(keys, n_keys, e.term, i)) continue;
    const SnPrint p = prints[i];
    if (e.len > p.len) continue;  // cannot happen: the element lies on that path
    const uint64_t bases = (uint64_t)k + p.len;
    image[p.offset + sn_header_bytes(i, bases, k) + (uint64_t)(k - 1) + (p.len - e.len)] = base_char(rk_last_base<N>(key, k, o));
  }
}

// everything else of record i: header line, first k-mer, the base of the end node, newline; a cycle (end_key SN_NONE)
// is written whole by the per-record walker
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool rk_write_frame(const TableView &t, int k, const SnPrint &p, uint64_t end_key, uint64_t i, char *image) {
  if (end_key == SN_NONE) return sn_write_record<N, Ops>(t, k, p, i, image);
  uint64_t key[N], km[N];
  if (!sn_load_node<N, Ops>(t, p.q, key)) return false;
  if (p.on & 1u) kmer_revcomp<N>(key, k, km);
  else {
#pragma unroll
    for (int w = 0; w < N; w++) km[w] = key[w];
  }
  char hdr[72];
  const uint64_t bases = (uint64_t)k + p.len;
  const uint32_t hn = sn_write_header(hdr, i, bases, k);
  char *out = image + p.offset;
  for (uint32_t b = 0; b < hn; b++) out[b] = hdr[b];
  for (int b = 0; b < k; b++) out[hn + b] = base_char(kmer_base_at<N>(km, k, b));
  uint64_t ekey[N];
  if (!sn_load_node<N, Ops>(t, end_key >> 3, ekey)) return false;
  out[hn + bases - 1] = base_char(rk_last_base<N>(ekey, k, (uint32_t)(end_key >> 2) & 1u));
  out[hn + bases] = '\n';
  return true;
}

#if defined(__CUDACC__)
// What the ranking leaves on the device for the per-node writer (rk_release frees it).
struct RkState {
  unsigned long long *bits = nullptr;
  uint64_t *rank = nullptr;
  RkElem *el = nullptr;
};
void rk_release(RkState &s);
// Drop-in for launch_sn_paths (same arguments plus the number of interior nodes from sn_count): allocates its
// working arrays (48 bytes per element while it runs), fills recs / covered / the counters in c; with `keep` the
// ranked elements stay on the device for rk_write_supernodes.  Returns cudaSuccess or the error of a failed
// allocation / launch.
cudaError_t sn_paths_ranked(int N, const TableView &t, int k, const uint64_t *starts, uint64_t n_starts,
                            uint64_t n_interior, uint8_t *covered, SnPath *recs, uint64_t cap, SnCounters *c,
                            cudaStream_t st, RkState *keep = nullptr);
// The FASTA image of the printed supernodes without anyone walking a path: one thread per record writes its frame
// (cycles: the whole record), one thread per interior node its base(s).  keys: the end states of the non-cycle
// prints, sorted; end_keys[i]: the end state of print i (SN_NONE: cycle).  Failures are counted in c->n_dangling.
void rk_write_supernodes(int N, const TableView &t, int k, const RkState &s, const RkPrintKey *keys, uint64_t n_keys,
                         const SnPrint *prints, const uint64_t *end_keys, uint64_t n_prints, char *image, SnCounters *c,
                         cudaStream_t st);
#endif

}  // namespace lasv
