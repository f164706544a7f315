// branches.cuh -- the branch sequences `--mode build` prints after the cleaning
// (`./branches.fa`, graph_operations.c:1145-1152): the per-record arithmetic shared by
// the kernel (cleaning_kernels.cu) and tests/host_emul.
//
// Replaces (reference file:line under /root/reference):
//   db_graph_print_all_branches_for_specific_person_or_pop          src/dB_graph_operations/print_branches.c:131-199
//   db_graph_get_sequence_for_a_single_branch_for_specific_person_or_pop   print_branches.c:16-127
//   db_graph_get_next_node_for_specific_person_or_pop               dB_graph_population.c:100-152
//   db_node_get_degree                                              element.c:1429-1447
//
// The reference walks the table slot by slot on one thread.  Every node with more than one
// edge on a side hands each neighbour X on that side to the branch walk, unless X is already
// `visited`: the walk marks X, and if X has exactly one way on it follows nodes with one edge
// each way (at most 300, stopping at a marked one), marking them; the node that ends the run
// is included (and marked) iff the walk is its only way in and it does not have exactly one
// way on.  Printed: ">b_<k-mer of X>\n<sequence of the L nodes>\n".  The marks make the result
// depend on the slot order -- but a walk only ever runs along one path between non-interior
// nodes (supernodes.cuh), marking a prefix of it, so the whole traversal is replayed exactly on
// the compressed graph in O(1) per branch (branches_host.h), and all the device has to do is
// enumerate the paths (the supernode kernels) and write the sequences (below).
#pragma once
#include "cleaning.cuh"

namespace lasv {

constexpr uint32_t ST_VISITED = 2;   // NodeStatus visited (element.h:57-75)
constexpr uint32_t BR_LIMIT = 300;   // print_branches.c:82: `length <= 300`

// One printed branch: X is the node one step from `q` in orientation (on & 1) along
// nucleotide (on >> 1); the record covers `nodes` nodes starting with X.
struct BrPrint {
  uint64_t q;
  uint64_t offset;  // byte offset of the record in the FASTA image
  uint32_t nodes;
  uint32_t on;
};

// ">b_" + k + "\n" + (k + nodes - 1) + "\n"
LASV_HD uint64_t br_record_bytes(int k, uint32_t nodes) { return 2ull * (uint64_t)k + nodes + 4ull; }

// Writes record p into the image; mark != 0 additionally leaves the `visited` marks of the
// reference on the nodes of the branch (print_branches.c:41,88,114,120).  False if the walk
// leaves the graph (cannot happen in a consistent graph).
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD bool br_write_record(const TableView &t, int k, const BrPrint &p, char *image, int mark) {
  uint64_t key[N], key2[N], km[N];
  if (!sn_load_node<N, Ops>(t, p.q, key)) return false;
  uint32_t o2 = 0, back = 0;
  uint64_t m2 = 0;
  uint64_t q2 = sn_step<N, Ops>(t, k, key, p.on & 1u, p.on >> 1, key2, m2, o2, back);
  if (q2 == SN_NONE) return false;
  if (o2) kmer_revcomp<N>(key2, k, km);
  else {
#pragma unroll
    for (int w = 0; w < N; w++) km[w] = key2[w];
  }
  ByteWriter out(image + p.offset);
  out.put('>');
  out.put('b');
  out.put('_');
  for (int b = 0; b < k; b++) out.put(base_char(kmer_base_at<N>(km, k, b)));
  out.put('\n');
  for (int b = 0; b < k; b++) out.put(base_char(kmer_base_at<N>(km, k, b)));
  bool ok = true;
  for (uint32_t e = 1;; e++) {
    if (mark && meta_status(m2) != ST_VISITED)
      Ops::store(cl_meta_ptr<N, Ops>(t, q2), meta_with(m2, meta_edges(m2), ST_VISITED));
    if (e >= p.nodes) break;
#pragma unroll
    for (int w = 0; w < N; w++) key[w] = key2[w];
    const uint32_t co = o2, cn = low_bit4(sn_side(meta_edges(m2), o2));
    out.put(base_char(cn));
    q2 = sn_step<N, Ops>(t, k, key, co, cn, key2, m2, o2, back);
    if (q2 == SN_NONE) { ok = false; break; }
  }
  out.put('\n');
  out.flush();
  return ok;
}

// print_branches.c:133 (db_node_action_unset_status_visited_...): visited -> none
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD void br_reset_mark(const TableView &t, uint64_t q) {
  uint64_t *mp = cl_meta_ptr<N, Ops>(t, q);
  const uint64_t m = Ops::load(mp);
  if (meta_status(m) == ST_VISITED) Ops::store(mp, meta_with(m, meta_edges(m), ST_NONE));
}

#if defined(__CUDACC__)
// launch interface (cleaning_kernels.cu); failures are counted in c->n_failed
void launch_br_write(int N, const TableView &t, int k, const BrPrint *prints, uint64_t n, char *image, int mark,
                     ClCounters *c, cudaStream_t st);
void launch_br_reset_marks(int N, const TableView &t, cudaStream_t st);
#endif

}  // namespace lasv
