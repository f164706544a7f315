/* integration/clean_shim.c -- reference-side binding of the GPU graph cleaning
 * (SURVEY 8f rank 2).  Compiled WITH the reference's headers and
 * linked into the reference's own main() in place of the two functions that
 * `--remove_low_coverage_supernodes T` calls (graph_operations.c:1069-1090), both moved
 * out of the way by -D renames on dB_graph_population.c (see Makefile):
 *   db_graph_clip_tips_in_union_of_all_colours            (dB_graph_population.c:2602)
 *   db_graph_remove_errors_considering_covg_and_topology  (dB_graph_population.c:3595)
 * Single colour only, so the union-of-colours accessors main() passes are the plain
 * edges / coverage and the two reset callbacks are reset_one_edge / db_node_reset_edges
 * on colour 0 -- what the GPU cleaning does to the table it hands back. */
#include <stdio.h>

#include "dB_graph.h"
#include "dB_graph_population.h"

/* include/lasv_b200.h (not included: its mirror of the HashTable struct clashes with the reference's) */
typedef struct {
  unsigned long long tips_clipped, tip_nodes_pruned, supernodes_pruned, nodes_pruned;
} lasv_clean_stats;
int lasv_ref_hash_table_clean(void *reference_hash_table, int what, int threshold, lasv_clean_stats *stats);
int lasv_ref_hash_table_clean_limited(void *reference_hash_table, int what, int threshold, int max_length,
                                      lasv_clean_stats *stats);
const char *lasv_last_error(void);
#define LASV_ERR_LIMIT (-7)

/* the reference's own function, moved out of the way by the -D rename (Makefile) */
long long cpu_db_graph_remove_errors_considering_covg_and_topology(
    Covg coverage, dBGraph *db_graph, Covg (*sum_of_covgs_in_desired_colours)(const Element *),
    Edges (*get_edge_of_interest)(const Element *),
    void (*apply_reset_to_specified_edges)(dBNode *, Orientation, Nucleotide),
    void (*apply_reset_to_specified_edges_2)(dBNode *), int max_expected_sup);

void db_graph_clip_tips_in_union_of_all_colours(dBGraph *db_graph)
{
  lasv_clean_stats st;
  if (lasv_ref_hash_table_clean(db_graph, 1 /* LASV_CLEAN_TIPS */, 0, &st) != 0)
    die("%s\n", lasv_last_error());
  printf("%llu tips clipped (%llu nodes)\n", st.tips_clipped, st.tip_nodes_pruned);
}

long long db_graph_remove_errors_considering_covg_and_topology(
    Covg coverage, dBGraph *db_graph, Covg (*sum_of_covgs_in_desired_colours)(const Element *),
    Edges (*get_edge_of_interest)(const Element *),
    void (*apply_reset_to_specified_edges)(dBNode *, Orientation, Nucleotide),
    void (*apply_reset_to_specified_edges_2)(dBNode *), int max_expected_sup)
{
  lasv_clean_stats st;
  /* max_expected_sup = --max_var_len: where the reference would judge a PIECE of a longer supernode the GPU call
   * changes nothing and says so; the reference's own function then does the pass (same table, same result) */
  const int rc = lasv_ref_hash_table_clean_limited(db_graph, 2 /* LASV_CLEAN_SUPERNODES */, (int)coverage, max_expected_sup, &st);
  if (rc == LASV_ERR_LIMIT) {
    printf("lasv_b200: a low-coverage supernode exceeds max length [%i]; the reference's own pass takes over\n", max_expected_sup);
    return cpu_db_graph_remove_errors_considering_covg_and_topology(coverage, db_graph, sum_of_covgs_in_desired_colours,
                                                                    get_edge_of_interest, apply_reset_to_specified_edges,
                                                                    apply_reset_to_specified_edges_2, max_expected_sup);
  }
  if (rc != 0) die("%s\n", lasv_last_error());
  return (long long)st.supernodes_pruned;
}
