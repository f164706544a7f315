/* integration/ctx_loader_shim.c -- reference-side binding of the GPU .ctx reload
 * (SURVEY 8f rank 1).  Compiled WITH the reference's headers and
 * linked into the reference's own main() in place of its
 * load_multicolour_binary_from_filename_into_graph (file_reader.c:1652-1710, moved
 * out of the way by a -D rename, see Makefile).  The header is parsed by the
 * reference's own check_binary_signature_NEW -- so GraphInfo is updated exactly as
 * before -- and only the per-record hash_table_insert loop is replaced by
 * lasv_ref_hash_table_load_ctx_records (liblasv_b200.so). */
#include <stdio.h>

#include "dB_graph.h"
#include "file_reader.h"
#include "graph_info.h"

/* include/lasv_b200.h (not included: it re-declares the two reference-named loaders with its own
 * mirror of the HashTable struct, which clashes with file_reader.h in one translation unit) */
long long lasv_ref_hash_table_load_ctx_records(void *reference_hash_table, const char *path, long record_offset);
const char *lasv_last_error(void);

long long load_multicolour_binary_from_filename_into_graph(char *filename, dBGraph *db_graph, GraphInfo *ginfo,
                                                           int *num_cols_in_loaded_binary)
{
  FILE *fp = fopen(filename, "r");
  if (fp == NULL)
    die("load_multicolour_binary_from_filename_into_graph cannot open file:%s\n", filename);
  BinaryHeaderErrorCode ecode = EValid;
  BinaryHeaderInfo binfo;
  initialise_binary_header_info(&binfo, ginfo);
  if (!check_binary_signature_NEW(fp, db_graph->kmer_size, &binfo, &ecode, 0))
    die("Cannot load this binary(%s) - signature check fails. Wrong max kmer, number of colours, or "
        "binary version. Exiting, error code %d\n", filename, ecode);
  *num_cols_in_loaded_binary = binfo.number_of_colours;
  if (binfo.number_of_colours != 1)
    die("lasv_b200: the GPU binary loader handles single-colour binaries only (%s has %d colours)\n", filename,
        binfo.number_of_colours);
  const long first_record = ftell(fp);
  fclose(fp);
  const long long n = lasv_ref_hash_table_load_ctx_records(db_graph, filename, first_record);
  if (n < 0)
    die("%s\n", lasv_last_error());
  return n * db_graph->kmer_size;
}
