// cleaning.cuh -- per-path arithmetic of the graph cleaning that laSV always runs
// (`--remove_low_coverage_supernodes T`, graph_operations.c:1069-1090): tip
// clipping (dB_graph_population.c:365-470, 2563-2608) and removal of low-coverage
// supernodes (:3371-3462, 3595-3649, node pruning :485-530).  Shared by the
// kernels (cleaning_kernels.cu) and tests/host_emul; the sequential decisions are taken on the
// path records by cleaning_host.h.
//
// Entry points (capi.cu): lasv_ref_hash_table_clean on the caller's reference-layout table,
// lasv_graph_clean on a graph handle, lasv_graph_health_check (`--health_check`, below).
#pragma once
#include "supernodes.cuh"

namespace lasv {

// ST_PRUNED (NodeStatus pruned) lives in kmer_core.cuh: the dump and summary kernels skip such nodes too

// What the sequential decisions (cleaning_host.h) ask the device to do.
struct ClPathPrune {
  uint64_t q;       // walk from this node ...
  uint32_t on;      // ... in this orientation | nucleotide << 1 ...
  uint32_t len;     // ... over len edges, pruning the len - 1 nodes in between,
  uint64_t keep_q;  // except this one, which only loses its edges (SN_NONE: none)
};
struct ClEdgeReset {
  uint64_t q;
  uint32_t mask;
  uint32_t pad;
};
struct ClNodeRec {  // a non-interior node as the decisions see it
  uint64_t q;
  uint32_t covg;
  uint32_t edges;
};
struct ClCounters {
  unsigned long long n_nodes;
  unsigned long long n_failed;  // walks that left the graph
};

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
  // the edges byte is the low byte of the meta word's upper half (kmer_core.cuh, meta_pack)
  Ops::and32(reinterpret_cast<uint32_t *>(mp) + 1, ~(mask & 0xFFu));
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

// Mark the interior nodes of a path as pruned, walked from (q, orientation, nucleotide) over
// len edges; the node `keep_q` (SN_NONE: none) is spared -- the start node of a cycle the
// reference printed from one of its own interior nodes -- and only loses its edges.  Only the
// STATUS changes here: a path that is its own reverse complement turns round in the middle and
// runs back over its own nodes, so the edges must stay readable until every walk is done;
// cl_settle_node clears them afterwards, in one sweep over the table.
constexpr uint32_t ST_CLEARED = 0x7E;  // transient, device only: keeps status none, loses its edges

#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool cl_prune_path(const TableView &t, int k, uint64_t q, uint32_t o, uint32_t n, uint32_t len,
                           uint64_t keep_q) {
  uint64_t key[N], key2[N];
  if (!sn_load_node<N, Ops>(t, q, key)) return false;
  uint32_t co = o, cn = n, o2 = 0, back = 0;
  for (uint32_t e = 0; e + 1 < len; e++) {
    uint64_t m2 = 0;
    const uint64_t q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) return false;
    const uint32_t want = q2 == keep_q ? ST_CLEARED : ST_PRUNED;
    if (meta_status(m2) != want) Ops::store(cl_meta_ptr<N, Ops>(t, q2), meta_with(m2, meta_edges(m2), want));
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    co = o2;
    cn = low_bit4(sn_side(meta_edges(m2), o2));
  }
  return true;
}

// After all walks: pruned nodes lose their edges (db_node_reset_edges), spared ones too but stay `none`.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void cl_settle_node(const TableView &t, uint64_t q) {
  uint64_t *mp = cl_meta_ptr<N, Ops>(t, q);
  const uint64_t m = Ops::load(mp);
  const uint32_t st = meta_status(m);
  if (st == ST_CLEARED) Ops::store(mp, meta_with(m, 0u, ST_NONE));
  else if (st == ST_PRUNED && meta_edges(m) != 0) Ops::store(mp, meta_with(m, 0u, ST_PRUNED));
}

// db_graph_health_check (dB_graph_population.c:6274-6350; `--health_check`) for one node: every edge
// must lead to a node that is in the graph (present, not pruned, with at least one edge:
// db_graph_get_next_node_for_specific_person_or_pop, :100-152) and that node must hold the return
// arrow.  Returns the number of edges checked; adds to the two problem counts.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD uint32_t hc_check_node(const TableView &t, int k, uint64_t q, uint32_t &missing_target, uint32_t &missing_return,
                               bool &zero_kmer) {
  uint64_t key[N], key2[N];
  const uint64_t m = sn_load_node<N, Ops>(t, q, key);
  if (!m) return 0u;
  zero_kmer = true;
#pragma unroll
  for (int w = 0; w < N; w++) zero_kmer = zero_kmer && key[w] == 0ull;
  const uint32_t e = meta_edges(m);
  uint32_t n = 0;
  for (uint32_t bit = 0; bit < 8; bit++) {
    if (!((e >> bit) & 1u)) continue;
    n++;
    uint32_t o2 = 0, back = 0;
    uint64_t m2 = 0;
    const uint64_t q2 = sn_step<N, Ops>(t, k, key, bit >> 2, bit & 3u, key2, m2, o2, back);
    if (q2 == SN_NONE || meta_status(m2) == ST_PRUNED || meta_edges(m2) == 0u) { missing_target++; continue; }
    if (!((meta_edges(m2) >> (4u * (o2 ^ 1u) + back)) & 1u)) missing_return++;
  }
  return n;
}

struct HcCounters {
  unsigned long long nodes, edges, missing_target, missing_return, zero_kmers;
};

#if defined(__CUDACC__)
void launch_hc_check(int N, const TableView &t, int k, HcCounters *c, cudaStream_t st);
// launch interface (cleaning_kernels.cu)
void launch_cl_nodes(int N, const TableView &t, ClNodeRec *out, uint64_t cap, ClCounters *c, cudaStream_t st);
void launch_cl_coverage(int N, const TableView &t, int k, const SnPath *recs, uint64_t n, uint32_t threshold,
                        uint32_t *max_covg, uint64_t *min_low_q, uint32_t *s_covg, ClCounters *c, cudaStream_t st);
void launch_cl_apply(int N, const TableView &t, int k, const ClPathPrune *paths, uint64_t n_paths, const uint64_t *nodes,
                     uint64_t n_nodes, const ClEdgeReset *resets, uint64_t n_resets, ClCounters *c, cudaStream_t st);
#endif

}  // namespace lasv
