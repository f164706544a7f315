// cleaning.cuh -- per-path arithmetic of the graph cleaning that laSV always runs
// (`--remove_low_coverage_supernodes T`, graph_operations.c:1069-1090): tip
// clipping (dB_graph_population.c:365-470, 2563-2608) and removal of low-coverage
// supernodes (:3371-3462, 3595-3649, node pruning :485-530).  Shared by the
// (future) kernels and tests/host_emul; the sequential decisions are taken on the
// path records by cleaning_host.h.
//
// STATUS: host-emulated only (tests/test_host_emul.py against the oracle's
// restatement of the reference); not yet part of the C ABI -- see DESIGN 6d.
#pragma once
#include "supernodes.cuh"

namespace lasv {

constexpr uint32_t ST_PRUNED = 3;  // NodeStatus pruned (element.h:57-75)

LASV_HD uint64_t meta_with(uint64_t m, uint32_t edges, uint32_t status) {
  return (m & 0xFFFF0000FFFFFFFFull) | ((uint64_t)(edges & 0xFFu) << 32) | ((uint64_t)(status & 0xFFu) << 40);
}

#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint64_t *cl_meta_ptr(const TableView &t, uint64_t q) {
  const uint32_t bucket = (uint32_t)(q / t.W);
  return slot_ptr<N>(t, bucket, (uint32_t)(q - (uint64_t)bucket * t.W)) + N;
}

// db_node_action_set_status_pruned + db_node_reset_edges on one node
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void cl_prune_node(const TableView &t, uint64_t q) {
  uint64_t *mp = cl_meta_ptr<N, Ops>(t, q);
  Ops::store(mp, meta_with(Ops::load(mp), 0u, ST_PRUNED));
}

// reset_one_edge (element.c:249-267) for every bit of `mask` on node q.  Several
// paths may end in the same node: the kernels apply this with an atomic AND.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void cl_reset_edges(const TableView &t, uint64_t q, uint32_t mask) {
  uint64_t *mp = cl_meta_ptr<N, Ops>(t, q);
  const uint64_t m = Ops::load(mp);
  Ops::store(mp, meta_with(m, meta_edges(m) & ~mask, meta_status(m)));
}

// Coverage summary of the interior nodes of a path record (walked from its s end):
// the largest coverage, and the lowest slot among the interior nodes with
// 0 < coverage <= threshold (SN_NONE if none) -- the node at which the reference's
// traversal first looks at this path (dB_graph_population.c:3404).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool cl_path_coverage(const TableView &t, int k, const SnPath &p, uint32_t threshold, uint32_t &max_covg,
                              uint64_t &min_low_q) {
  uint64_t key[N], key2[N];
  max_covg = 0;
  min_low_q = SN_NONE;
  if (!sn_load_node<N, Ops>(t, p.s_q, key)) return false;
  uint32_t co = (p.bits & SNB_S_O) ? 1u : 0u, cn = (p.bits >> SNB_S_N) & 3u, o2 = 0, back = 0;
  for (uint32_t e = 0; e + 1 < p.len; e++) {
    uint64_t m2 = 0;
    const uint64_t q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) return false;
    const uint32_t c = meta_covg(m2);
    if (c > max_covg) max_covg = c;
    if (c > 0 && c <= threshold && q2 < min_low_q) min_low_q = q2;
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(meta_edges(m2), o2));
  }
  return true;
}

// Prune the interior nodes of a path, walked from (q, orientation, nucleotide) over
// len edges; the node `keep_q` (SN_NONE: none) is spared -- the start node of a
// cycle the reference printed from one of its own interior nodes -- and only loses
// its edges.  The next node is located BEFORE the current one loses its edges.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool cl_prune_path(const TableView &t, int k, uint64_t q, uint32_t o, uint32_t n, uint32_t len,
                           uint64_t keep_q) {
  uint64_t key[N], key2[N];
  if (!sn_load_node<N, Ops>(t, q, key)) return false;
  uint32_t co = o, cn = n, o2 = 0, back = 0;
  uint64_t prev_q = SN_NONE;  // interior node whose successor is known by now
  for (uint32_t e = 0; e + 1 < len; e++) {
    uint64_t m2 = 0;
    const uint64_t q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) return false;
    if (prev_q != SN_NONE) {
      if (prev_q == keep_q) cl_reset_edges<N, Ops>(t, prev_q, 0xFFu);
      else cl_prune_node<N, Ops>(t, prev_q);
    }
    prev_q = q2;
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(meta_edges(m2), o2));
  }
  if (prev_q != SN_NONE) {
    if (prev_q == keep_q) cl_reset_edges<N, Ops>(t, prev_q, 0xFFu);
    else cl_prune_node<N, Ops>(t, prev_q);
  }
  return true;
}

}  // namespace lasv
