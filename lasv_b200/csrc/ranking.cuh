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

#if defined(__CUDACC__)
// Drop-in for launch_sn_paths (same arguments plus the number of interior nodes from sn_count): allocates its
// working arrays (48 bytes per element while it runs), fills recs / covered / the counters in c.  Returns
// cudaSuccess or the error of a failed allocation / launch.
cudaError_t sn_paths_ranked(int N, const TableView &t, int k, const uint64_t *starts, uint64_t n_starts,
                            uint64_t n_interior, uint8_t *covered, SnPath *recs, uint64_t cap, SnCounters *c,
                            cudaStream_t st);
#endif

}  // namespace lasv
