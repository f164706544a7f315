/* integration/branch_shim.c -- reference-side binding of the GPU branch printing
 * (SURVEY 8f rank 3).  Compiled WITH the reference's headers and linked
 * into the reference's own main() in place of
 *   db_graph_print_all_branches_for_specific_person_or_pop   (print_branches.c:131-199)
 * which is moved out of the way by a -D rename on print_branches.c (see Makefile).
 * `--mode build` (graph_operations.c:1145-1152, run_laSV.sh:178) then writes ./branches.fa
 * from the GPU: the file the reference's own traversal of db_graph->table would write,
 * byte for byte, and the same `visited` marks left in the table.  Single colour only. */
#include <stdio.h>

#include "dB_graph.h"
#include "print_branches.h"

/* include/lasv_b200.h (not included: its mirror of the HashTable struct clashes with the reference's) */
typedef struct {
  unsigned long long n_branches, n_branching_nodes, n_nodes, n_refused, n_cut, bytes;
} lasv_branch_stats;
int lasv_ref_hash_table_write_branches(void *reference_hash_table, const char *fasta_path, lasv_branch_stats *stats);
const char *lasv_last_error(void);

void db_graph_print_all_branches_for_specific_person_or_pop(char *out_file_name, dBGraph *db_graph, int index)
{
  if (index != 0) die("lasv_b200: the GPU branch printer handles colour 0 only (got %d)\n", index);
  lasv_branch_stats st;
  if (lasv_ref_hash_table_write_branches(db_graph, out_file_name, &st) != 0)
    die("%s\n", lasv_last_error());
  printf("%llu branches printed from %llu branching nodes\n", st.n_branches, st.n_branching_nodes);
}
