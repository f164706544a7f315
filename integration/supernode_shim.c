/* integration/supernode_shim.c -- reference-side binding of the GPU supernode
 * extraction.  Compiled WITH the reference's
 * headers and linked into the reference's own main() in place of its
 * db_graph_print_supernodes_defined_by_func_of_colours (dB_graph_population.c:
 * 2696-2803, moved out of the way by a -D rename, see Makefile): `--output_supernodes f`
 * (graph_operations.c:1276-1293) then enumerates the supernodes on the GPU and writes
 * the FASTA the reference's own traversal of db_graph->table would write.
 * Single colour only, so the union-of-colours accessors main() passes are the plain
 * edges / coverage; the extra-info callback prints nothing unless
 * --print_colour_coverages is given, which this binding does not support. */
#include <stdio.h>
#include <string.h>

#include "dB_graph.h"
#include "dB_graph_population.h"

/* include/lasv_b200.h (not included: its mirror of the HashTable struct clashes with the reference's) */
typedef struct {
  unsigned long long n_supernodes, n_nodes, n_interior, n_paths, n_cycles, n_not_printed, longest_edges,
      total_bases, over_max_length;
} lasv_supernode_stats;
int lasv_ref_hash_table_write_supernodes(void *reference_hash_table, const char *fasta_path, int max_length,
                                         lasv_supernode_stats *stats);
const char *lasv_last_error(void);

void db_graph_print_supernodes_defined_by_func_of_colours(char *filename_sups, char *filename_sings, int max_length,
                                                          dBGraph *db_graph, Edges (*get_colour)(const dBNode *),
                                                          Covg (*get_covg)(const dBNode *),
                                                          void (*print_extra_info)(dBNode **, Orientation *, int, FILE *))
{
  (void)get_colour;
  (void)get_covg;
  (void)print_extra_info;
  if (strcmp(filename_sings, "") != 0)
    die("lasv_b200: the GPU supernode printer does not write a singletons file (%s)\n", filename_sings);
  printf("Only printing supernodes consisting of >1 node "
         "(ie contigs longer than %d bases)", db_graph->kmer_size);
  lasv_supernode_stats st;
  if (lasv_ref_hash_table_write_supernodes(db_graph, filename_sups, max_length, &st) != 0)
    die("%s\n", lasv_last_error());
  if (st.over_max_length)
    printf("%llu contigs are longer than max length [%i] (printed whole)\n", st.over_max_length, max_length);
  printf("%llu nodes visted [%llu supernodes]\n", st.n_nodes, st.n_supernodes);
}
