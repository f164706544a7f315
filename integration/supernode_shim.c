/* integration/supernode_shim.c -- reference-side binding of the GPU supernode
 * extraction.  Compiled WITH the reference's
 * headers and linked into the reference's own main() in place of its
 * db_graph_print_supernodes_defined_by_func_of_colours (dB_graph_population.c:
 * 2696-2803, moved out of the way by a -D rename, see Makefile): `--output_supernodes f`
 * (graph_operations.c:1276-1293) then enumerates the supernodes on the GPU and writes
 * the FASTA the reference's own traversal of db_graph->table would write.
 * Single colour only, so the union-of-colours accessors main() passes are the plain
 * edges / coverage.  Three cases stay with the reference's own CPU printer (renamed cpu_..., same file,
 * same arguments), so that the output is the reference's in EVERY case:
 *   a singletons file is asked for (never by laSV: graph_operations.c:1283 passes "");
 *   --print_colour_coverages (the extra-info callback then prints per-node coverages; LASV_B200_SN_EXTRA=1
 *     tells the shim, as it cannot see cmd_line);
 *   a supernode has more than max_length (--max_var_len, default 10 000) edges: the reference cuts its walks
 *     there and prints overlapping pieces (dB_graph_population.c:2739-2767) -- the GPU call reports
 *     LASV_ERR_LIMIT and writes nothing. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dB_graph.h"
#include "dB_graph_population.h"

/* include/lasv_b200.h (not included: its mirror of the HashTable struct clashes with the reference's) */
typedef struct {
  unsigned long long n_supernodes, n_nodes, n_interior, n_paths, n_cycles, n_not_printed, longest_edges,
      total_bases, over_max_length;
} lasv_supernode_stats;
int lasv_ref_hash_table_write_supernodes_strict(void *reference_hash_table, const char *fasta_path, int max_length,
                                                lasv_supernode_stats *stats);
const char *lasv_last_error(void);
#define LASV_ERR_LIMIT (-7)

/* the reference's own printer, moved out of the way by the -D rename (Makefile) */
void cpu_db_graph_print_supernodes_defined_by_func_of_colours(char *filename_sups, char *filename_sings, int max_length,
                                                              dBGraph *db_graph, Edges (*get_colour)(const dBNode *),
                                                              Covg (*get_covg)(const dBNode *),
                                                              void (*print_extra_info)(dBNode **, Orientation *, int, FILE *));

void db_graph_print_supernodes_defined_by_func_of_colours(char *filename_sups, char *filename_sings, int max_length,
                                                          dBGraph *db_graph, Edges (*get_colour)(const dBNode *),
                                                          Covg (*get_covg)(const dBNode *),
                                                          void (*print_extra_info)(dBNode **, Orientation *, int, FILE *))
{
  const char *extra = getenv("LASV_B200_SN_EXTRA");
  if (strcmp(filename_sings, "") != 0 || (extra && extra[0] == '1')) {
    cpu_db_graph_print_supernodes_defined_by_func_of_colours(filename_sups, filename_sings, max_length, db_graph, get_colour,
                                                             get_covg, print_extra_info);
    return;
  }
  lasv_supernode_stats st;
  const int rc = lasv_ref_hash_table_write_supernodes_strict(db_graph, filename_sups, max_length, &st);
  if (rc == LASV_ERR_LIMIT) {
    printf("lasv_b200: %llu supernodes exceed max length [%i]; the reference's own printer takes over\n",
           st.over_max_length, max_length);
    cpu_db_graph_print_supernodes_defined_by_func_of_colours(filename_sups, filename_sings, max_length, db_graph, get_colour,
                                                             get_covg, print_extra_info);
    return;
  }
  if (rc != 0) die("%s\n", lasv_last_error());
  printf("Only printing supernodes consisting of >1 node "
         "(ie contigs longer than %d bases)", db_graph->kmer_size);
  printf("%llu nodes visted [%llu supernodes]\n", st.n_nodes, st.n_supernodes);
}
