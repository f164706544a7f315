/*
 * lasv_b200.h -- C ABI of the B200-native de Bruijn graph build for laSV.
 *
 * One shared library (lasv_b200/liblasv_b200.so) serves MAXK 31/63/95: the
 * number of 64-bit words per k-mer, N = 2k/64 + 1 (graph_operations.c:620), is
 * picked at run time from kmer_size where the reference picks it at compile
 * time (Makefile:8-43).  Plain pointers and sizes only; no torch types.
 *
 * Three groups of entry points, each naming the reference interface it stands
 * in for (paths relative to /root/reference):
 *
 *  (A) L-loader drop-ins   : same names and signatures as
 *        include/dB_graph_operations/file_reader.h:86-118, operating on the
 *        caller's reference-layout HashTable (include/hash_table/open_hash/
 *        hash_table.h:40-50).  INTEGRATION.md shows the two-line change that
 *        makes the reference's own main() (graph_operations.c:746-765) call
 *        them.
 *  (B) graph handle        : the hash_table / element / binary_kmer surface
 *        (hash_table.h:53-86, element.h:89-292) over a device-resident table.
 *  (C) .ctx wire format    : file_reader.c:2930-2994 (writer), :1652-1715
 *        (loader), element.c:894-910 (record).
 *
 * Error convention: group (A) behaves like the reference -- fatal conditions
 * print "Error: ..." to stderr and exit(EXIT_FAILURE) (src/basic/global.c:44-63).
 * Groups (B),(C) return 0 on success or a negative LASV_ERR_* code, with the
 * message available from lasv_last_error().  There is NO CPU fallback: every
 * compute entry point fails with LASV_ERR_CUDA if no CUDA device is usable.
 */
#ifndef LASV_B200_H_
#define LASV_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LASV_OK 0
#define LASV_ERR_ARG (-1)        /* bad argument (k even/out of range, NULL, sizes) */
#define LASV_ERR_CUDA (-2)       /* CUDA runtime failure, or no device */
#define LASV_ERR_TABLE_FULL (-3) /* a key met 16 full buckets: hash_table.c:309-317 die() */
#define LASV_ERR_IO (-4)         /* file could not be opened / parsed */
#define LASV_ERR_STATE (-5)      /* e.g. inserting into a finalised table */
#define LASV_ERR_CAPACITY (-6)   /* caller-provided output buffer too small */
#define LASV_ERR_LIMIT (-7)      /* *_strict / *_limited calls: a walk exceeds the reference's --max_var_len; nothing was done */

typedef struct lasv_graph lasv_graph;

/* Loader bookkeeping, the quantities file_reader.c prints / returns
 * (file_reader.c:1062-1075) plus the insert counter behind the k-mers/s metric
 * (= sum of hash_table.c:325 collisions[]). */
typedef struct {
  uint64_t n_reads;
  uint64_t bad_reads;      /* reads with no k consecutive good bases */
  uint64_t bases_read;     /* all bases seen */
  uint64_t bases_loaded;   /* sum of contig lengths that entered the graph */
  uint64_t n_contigs;
  uint64_t kmers_inserted; /* find_or_insert calls */
  uint64_t unique_kmers;   /* hash_table_get_unique_kmers */
} lasv_load_stats;

/* Order-independent digest of the node multiset {(key, coverage, edges)}. */
typedef struct {
  uint64_t n_nodes;
  uint64_t sum_covg;
  uint64_t sum_edge_bits;
  uint64_t fingerprint;
} lasv_table_summary;

/* Reference HashTable layout (hash_table.h:40-50), for the (A) drop-ins. */
typedef struct {
  short kmer_size;
  long long number_buckets;
  int bucket_size;
  void *table;             /* Element[number_buckets*bucket_size], element.h:77-82 */
  short *next_element;
  long long *collisions;
  long long unique_kmers;
  int max_rehash_tries;
} lasv_ref_hash_table;

const char *lasv_last_error(void);
int lasv_device_count(void);
const char *lasv_version(void);

/* ------------------------------------------------------------------ (B) -- */
/* hash_table_new (hash_table.c:13-53): 2^number_bits buckets x bucket_size
 * slots of sizeof(Element) = 8N+8 bytes, zeroed, on CUDA device `device`.
 * Returns NULL on failure (as the reference does on OOM). */
lasv_graph *lasv_graph_new(int kmer_size, int number_bits, int bucket_size,
                           int max_rehash_tries, int device);
/* hash_table_free (hash_table.c:55-62): frees and NULLs the caller's pointer. */
void lasv_graph_free(lasv_graph **g);
/* Back to the state right after lasv_graph_new (zero table and counters). */
int lasv_graph_clear(lasv_graph *g, void *cuda_stream);
/* Pipelined inserts (large tables): with `on`, batch i is inserted on an internal
 * stream while the caller's stream already partitions batch i+1.  Library calls that
 * read the table join the internal stream themselves; a caller that reads the table
 * with its own kernels, or times with events on its stream, calls lasv_graph_flush
 * (the stream then waits for all queued inserts).  Default off; the environment
 * variable LASV_B200_PIPELINE=1 turns it on at lasv_graph_new. */
int lasv_graph_set_pipeline(lasv_graph *g, int on);
int lasv_graph_flush(lasv_graph *g, void *cuda_stream);
/* Measurement aid: CUDA events around every launch of the partitioned path's two
 * kernels (partition = build_kernel<MODE_BIN>, insert = insert_bins_kernel), each on
 * the stream the kernel is launched on.  enable(1) starts and resets the collection;
 * kernel_times synchronises the device and sums the durations collected so far. */
int lasv_graph_enable_kernel_timing(lasv_graph *g, int on);
int lasv_graph_kernel_times(lasv_graph *g, double *partition_ms, double *insert_ms, uint64_t *n_batches);
/* The same collection broken down by kind of launch: total milliseconds and number of timed spans per kind.
 * LASV_T_INSERT spans a whole insert (random-access kernel, or every launch of a streamed insert); the
 * streamed insert's parts are timed inside it: LASV_T_SORT (count + scan + scatter of a slab's records by
 * group), LASV_T_STREAM (sg_insert_kernel: group lines through shared memory) and LASV_T_DRAIN (records that
 * left their group + the senders' spill stripes). */
#define LASV_T_PARTITION 0
#define LASV_T_INSERT 1
#define LASV_T_SORT 2
#define LASV_T_STREAM 3
#define LASV_T_DRAIN 4
#define LASV_T_KINDS 5
int lasv_graph_kernel_times_by_kind(lasv_graph *g, double ms[LASV_T_KINDS], uint64_t n[LASV_T_KINDS]);
/* Which insert follows the partition kernel on large tables (hash_table_find_or_insert,
 * hash_table.c:270-327, either way -- same buckets, same rehash chain, same multiset):
 *   LASV_INSERT_RANDOM    one random table access per record, region by region (insert_bins_kernel)
 *   LASV_INSERT_STREAMED  the table streamed through shared memory group by group, the records of a group
 *                         applied there (streamed_insert.cuh): two sequential passes over the table's lines
 *                         per insert, whatever the number of records
 *   LASV_INSERT_AUTO      streamed once the records to insert number `min_records_per_line` (default 0.75;
 *                         <= 0 keeps the current value) per 128-byte table line, random access below that;
 *                         batches too small to partition (< 256 bytes per region) or tables below 256 MB
 *                         take the fused find-or-insert without a partition
 * RANDOM and STREAMED, asked for by name, partition every batch whatever its size.
 * Default AUTO; LASV_B200_INSERT=random|streamed at lasv_graph_new overrides. */
#define LASV_INSERT_AUTO 0
#define LASV_INSERT_RANDOM 1
#define LASV_INSERT_STREAMED 2
int lasv_graph_set_insert_mode(lasv_graph *g, int mode, double min_records_per_line);
/* Inserts issued so far, by kind. */
int lasv_graph_insert_counts(lasv_graph *g, uint64_t *n_streamed, uint64_t *n_random);
/* Accumulation: with max_stream_bytes != 0, lasv_graph_load_reads_device only PARTITIONS each batch into the
 * graph's region stripes; the stripes are inserted once they hold batches of max_stream_bytes stream bytes
 * in total (clamped to what the free device memory can hold; UINT64_MAX = as much as fits), on
 * lasv_graph_flush, and by every library call that reads the table or its counters (stats, find, dumps,
 * finalize, ...; NOT by lasv_graph_stats_async, whose k-mer counters then lag behind the reads).  The cost of
 * an insert per record falls with the number of records per table line, so few large inserts beat many small
 * ones.  Returns the clamped limit through *granted (may be NULL).  0 switches accumulation off (after a
 * flush).  The host-buffer and file loaders accumulate the chunks of ONE call by themselves. */
int lasv_graph_set_accumulation(lasv_graph *g, uint64_t max_stream_bytes, uint64_t *granted);
/* hash_table_get_unique_kmers / hash_table_get_capacity (hash_table.c:397-408). */
long long lasv_graph_unique_kmers(lasv_graph *g);
long long lasv_graph_capacity(lasv_graph *g);
int lasv_graph_words_per_kmer(lasv_graph *g);
size_t lasv_graph_element_bytes(lasv_graph *g);
/* Device address and size of the table memory.  Until lasv_graph_finalize a
 * bucket is ceil(W/G) lines of 128 bytes (8 tag bytes + G Elements, G = 7/5/3
 * for N = 1/2/3); afterwards it is W consecutive reference Elements at the start
 * of the same bytes.  lasv_graph_export_table packs the buckets for the host. */
void *lasv_graph_device_table(lasv_graph *g);
uint64_t lasv_graph_device_table_bytes(lasv_graph *g);

/* THE HOT PATH.  One batch of reads as two byte streams (sequence, quality):
 * reads back to back, every read followed by exactly `1` separator byte '\n'
 * in BOTH streams; offsets[i] = start of read i, offsets[n_reads] = n_bytes.
 * Does, for every read: quality/N filter -> k-mers -> canonical key ->
 * find-or-insert, coverage += 1, edges |= ... (file_reader.c:322-473,589-773).
 * quality_cutoff is the RAW ASCII cutoff = threshold + offset
 * (file_reader.c:798,919); qual == NULL means "no quality scores" (FASTA).
 *
 * _device: buffers already in HBM, 16-byte aligned, readable up to
 *   round_up(n_bytes,16); asynchronous on `cuda_stream`; d_offsets may be NULL
 *   (then the read-level bookkeeping -- bad reads, bases loaded, contig
 *   histogram -- is skipped; k-mer counters are always kept).
 * _host:   host buffers (pinned for full speed); copies are chunked and
 *   overlapped with the kernels; synchronous; returns the batch statistics. */
int lasv_graph_load_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                 uint64_t n_bytes, const uint64_t *d_offsets, uint64_t n_reads,
                                 int quality_cutoff, void *cuda_stream);
int lasv_graph_load_reads_host(lasv_graph *g, const uint8_t *seq, const uint8_t *qual,
                               uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                               int quality_cutoff, lasv_load_stats *batch_stats);
/* Asynchronous host-buffer call: the same chunked copies and partition launches, but nothing waits for a
 * kernel.  It returns once the last copy has LEFT the caller's buffers (they may be reused at once); the
 * kernels of the call -- and, with lasv_graph_set_accumulation in force, the insert of stripes that filled
 * up -- are still queued or running.  The copies run ahead of the kernels through the staging ring
 * (lasv_graph_set_staging), so that the next call's host->device traffic overlaps this call's insert.
 * pinned_stats9 (may be NULL): PINNED host memory that receives the statistics as they stand behind this
 * call's kernels (layout of lasv_graph_stats_async; valid after lasv_graph_flush + a stream synchronisation,
 * or any synchronous statistics call).  Table-full and overflow conditions are reported by the next
 * synchronous call (lasv_graph_stats, lasv_graph_load_reads_host, ...).  Earlier work queued on the caller's
 * own streams (lasv_graph_clear, device-buffer loads) must be complete before the first asynchronous call.
 * Replaces the reference's one-read-at-a-time loop of load_se/pe_seq_data_into_graph_colour
 * (file_reader.c:475-773) for callers that parse ahead of the graph. */
int lasv_graph_load_reads_host_async(lasv_graph *g, const uint8_t *seq, const uint8_t *qual,
                                     uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                                     int quality_cutoff, uint64_t *pinned_stats9);
/* Size of the device staging ring of the host-buffer calls in bytes (legs of 2 x 64 MB; at least three legs,
 * at most 256, never more than half of the free device memory).  Three legs (the default) overlap a copy with
 * the partition kernel of the previous chunk; a ring of a few GB lets lasv_graph_load_reads_host_async keep
 * copying while an insert of accumulated stripes (~0.1 s on a human-scale table) occupies the GPU.  Call it
 * BEFORE lasv_graph_set_accumulation (which sizes the stripes from the memory that is left). */
int lasv_graph_set_staging(lasv_graph *g, uint64_t bytes, uint64_t *granted);
/* Cumulative statistics since new/clear (synchronises `cuda_stream`).  Returns
 * LASV_ERR_TABLE_FULL if any insert overflowed its whole rehash chain.
 * contig_hist (may be NULL) receives the contig-length histogram,
 * readlen_count_array of file_reader.c:467-471, hist_size bins. */
int lasv_graph_stats(lasv_graph *g, lasv_load_stats *out, uint64_t *contig_hist,
                     uint64_t hist_size, void *cuda_stream);
/* Non-blocking variant for pipelined callers: the statistics as they stand when
 * `cuda_stream` reaches this point, copied to PINNED host memory: n_reads, bad_reads,
 * bases_read, bases_loaded, n_contigs, unique_kmers, kmers_inserted, table_full,
 * exchange_overflow (9 x uint64).  Valid once the stream has passed the copy. */
int lasv_graph_stats_async(lasv_graph *g, uint64_t *pinned_out9, void *cuda_stream);
/* collisions[r] of hash_table.c:325 (r = 0..15). */
int lasv_graph_rehash_histogram(lasv_graph *g, long long out16[16]);

/* Hash-sharded build.  A rank's local table has 2^number_bits
 * buckets; owner(key) = (lookup3(key) >> number_bits) & (n_dest-1).
 * route: reads -> records binned by owner: d_bins holds n_dest bins of
 *   bin_capacity records of (2N+1) uint32; d_bin_counts[n_dest] is zeroed by the
 *   caller and receives the counts (a count above bin_capacity means overflow:
 *   LASV_ERR_CAPACITY is reported by lasv_graph_stats).
 * insert_records: the receive side: find-or-insert every record. */
int lasv_graph_route_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                  uint64_t n_bytes, int quality_cutoff, int n_dest,
                                  uint32_t *d_bins, uint64_t bin_capacity,
                                  unsigned long long *d_bin_counts, void *cuda_stream);
int lasv_graph_extract_records_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                      uint64_t n_bytes, int quality_cutoff, uint32_t *d_records,
                                      uint64_t capacity, unsigned long long *d_count,
                                      void *cuda_stream);
int lasv_graph_insert_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                     void *cuda_stream);

/* The same exchange for large tables, region by region (the production path of a
 * hash-sharded build).  bin_reads: K1+K2 over this rank's reads, each k-mer
 * occurrence written as one fixed-size record (key words, lookup3 value, edges,
 * count) into the stripe of the 64 MB TABLE REGION it will touch on its owner:
 * stripe index = owner * bins_per_owner + region, `stripe_capacity_records` records
 * each, counters one per 128-byte line.  The stripes of one owner are one contiguous
 * block, so the all-to-all moves fixed-size blocks (no host round trip for sizes).
 * insert_bins: the receive side: sweeps the stripes of all senders region by region
 * (stripe index = sender * bins_per_src + region) and find-or-inserts every record.
 * bin_geometry sizes the buffers for a batch (reads known: k-mers <= n_bytes -
 * n_reads*k).  Buffer layout, per owner / sender, contiguous:
 *   records: bins_per_owner * stripe_capacity records (the stripes, interleaved in
 *            runs of 128 records) followed by spill_capacity records
 *   counts : bins_per_owner + 1 counters, count_stride_bytes apart (the last one
 *            counts the spill area)
 * A record that finds its stripe full (a k-mer occurring very often in one batch)
 * goes to its owner's spill area, which the owner inserts after its regions; only if
 * that is full too is the overflow flag raised (reported by lasv_graph_stats).  With a
 * single owner and spill_capacity 0 the excess goes straight into the table. */
int lasv_graph_bin_geometry(lasv_graph *g, int n_owners, uint64_t n_bytes, uint64_t n_reads,
                            uint32_t *bins_per_owner, uint32_t *stripe_capacity_records,
                            uint32_t *spill_capacity_records, uint32_t *record_bytes,
                            uint32_t *count_stride_bytes);
int lasv_graph_bin_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                uint64_t n_bytes, int quality_cutoff, int n_owners,
                                uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                uint32_t spill_capacity_records, void *d_records, void *d_counts,
                                void *cuda_stream);
/* The same, APPENDING to stripes that already hold earlier batches (the counters are not reset): a rank of a
 * hash-sharded build accumulates several batches in stripes sized for all of them (lasv_graph_bin_geometry
 * with the summed bytes and reads), exchanges them once and the owners insert them once -- the insert's
 * cost per record falls with the records per table line (see lasv_graph_set_accumulation). */
int lasv_graph_bin_reads_append_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                       uint64_t n_bytes, int quality_cutoff, int n_owners,
                                       uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                       uint32_t spill_capacity_records, void *d_records, void *d_counts,
                                       void *cuda_stream);
/* FUSED producer -> exchange: the partition kernel stores every record straight into its OWNER's
 * receive stripes -- owner_records[o] (host array of n_owners device pointers) is the block that sender `this rank`
 * has in owner o's receive buffer, mapped into this process (CUDA peer access / symmetric memory), so the stores to
 * remote owners travel over NVLink while the kernel runs and no send buffer or block copy exists.  The slot
 * reservations stay local (the counters are this sender's; they are copied to the owners afterwards, a few KB).
 * The owners must not read the blocks before every sender's kernel has completed (a barrier between the ranks),
 * and no sender may start before the owner has finished inserting the previous contents.  append as above. */
int lasv_graph_bin_reads_peers_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                      uint64_t n_bytes, int quality_cutoff, int n_owners,
                                      uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                      uint32_t spill_capacity_records, void *const *owner_records, void *d_counts,
                                      void *cuda_stream, int append);
int lasv_graph_insert_bins_device(lasv_graph *g, const void *d_records, const void *d_counts, int n_src,
                                  uint32_t bins_per_src, uint32_t stripe_capacity_records,
                                  uint32_t spill_capacity_records, void *cuda_stream);
int lasv_graph_read_stats_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                 const uint64_t *d_offsets, uint64_t n_reads, int quality_cutoff,
                                 void *cuda_stream);

/* hash_table_find (hash_table.c:234-267) for n canonical keys (n x N words,
 * word 0 most significant).  Host arrays. */
int lasv_graph_find(lasv_graph *g, const uint64_t *keys, uint64_t n, uint32_t *covg,
                    uint8_t *edges, uint8_t *found);
int lasv_graph_summary(lasv_graph *g, lasv_table_summary *out);

/* Squeeze every bucket hole-free and restore pure reference Element bytes: the
 * device table becomes a valid HashTable->table.  No inserts afterwards. */
int lasv_graph_finalize(lasv_graph *g);
/* finalize + copy to host: `elements` receives capacity*element_bytes bytes,
 * next_element (may be NULL) the per-bucket fill (hash_table.h:46). */
int lasv_graph_export_table(lasv_graph *g, void *elements, short *next_element);
/* Merge an existing reference-layout table (e.g. from an earlier CPU load). */
int lasv_graph_import_table(lasv_graph *g, const void *elements);

/* Supernodes = maximal unambiguous contigs of the graph on the device
 * (`--output_supernodes FILE`, graph_operations.c:1276-1293 ->
 * db_graph_print_supernodes_defined_by_func_of_colours, dB_graph_population.c:
 * 2696-2803, with db_graph_supernode_... :1454-1527 and ..._get_perfect_path_
 * with_first_edge_... :783-896).  Writes the reference's FASTA (">node_<i>
 * length:<bases> INFO:KMER=<k>", print_ultra_minimal_fasta_from_path :6125-6178)
 * for a traversal in table order -- the order of the .ctx records this library
 * dumps -- byte for byte what the reference prints after reloading that .ctx;
 * singletons are not printed (filename_sings == "").  The reference's walk limit
 * (--max_var_len, default 10000 edges, after which it prints overlapping pieces)
 * is not applied: a supernode is always whole, `over_max_length` counts those
 * longer than `max_length`.  The graph is left unchanged. */
typedef struct {
  uint64_t n_supernodes;     /* records written */
  uint64_t n_nodes;          /* nodes in the graph */
  uint64_t n_interior;       /* nodes with exactly one edge on each side */
  uint64_t n_paths;          /* distinct paths between non-interior nodes */
  uint64_t n_cycles;         /* cycles of interior nodes */
  uint64_t n_not_printed;    /* one-edge paths whose end nodes were both visited first (the reference skips them) */
  uint64_t longest_edges;    /* edges of the longest supernode */
  uint64_t total_bases;      /* sum of sequence lengths */
  uint64_t over_max_length;  /* supernodes with more than max_length edges */
} lasv_supernode_stats;
int lasv_graph_write_supernodes(lasv_graph *g, const char *fasta_path, int max_length,
                                lasv_supernode_stats *stats /* may be NULL */);
/* The same, but if any supernode has more than max_length edges nothing is written and LASV_ERR_LIMIT is
 * returned (stats->over_max_length is set): that is where the reference's output turns into overlapping pieces
 * of max_length edges (dB_graph_population.c:2739-2767, "contig length equals max length").  The reference-side
 * binding (integration/supernode_shim.c) then runs the reference's own CPU printer, so the file is the
 * reference's in every case. */
int lasv_graph_write_supernodes_strict(lasv_graph *g, const char *fasta_path, int max_length,
                                       lasv_supernode_stats *stats /* may be NULL */);

/* `--health_check` (db_graph_health_check, dB_graph_population.c:6274-6350) on the device graph, in
 * either arrangement: every edge of every node (pruned ones have none) must lead to a node that is in
 * the graph -- present, not pruned, with at least one edge (db_graph_get_next_node_for_specific_person_
 * or_pop, :100-152) -- and that node must hold the return arrow.  Problems are counted (the reference
 * prints one line each); nothing is fixed.  A graph built by the loaders, and one left by the cleaning,
 * has none: a size-independent invariant for graphs too large to compare record by record. */
typedef struct {
  uint64_t nodes_checked;         /* the reference's return value */
  uint64_t edges_checked;
  uint64_t missing_target;        /* "didnt find node in hash table" */
  uint64_t missing_return_arrow;  /* "inconsitency return arrow missing" */
  uint64_t zero_kmers;            /* nodes whose k-mer is all A (more than one is an error in the reference) */
} lasv_health_stats;
int lasv_graph_health_check(lasv_graph *g, lasv_health_stats *out);

/* Branch sequences of the device graph: what `--mode build` writes to ./branches.fa after
 * the cleaning (graph_operations.c:1145-1152 -> db_graph_print_all_branches_for_specific_
 * person_or_pop, print_branches.c:131-199, with db_graph_get_sequence_for_a_single_branch_...
 * :16-127): every neighbour X of a node with more than one edge on a side starts a record
 * ">b_<k-mer of X>\n<X and the nodes with one edge each way that follow, at most 300>\n",
 * unless an earlier record left X `visited`.  Byte for byte the reference's file for a
 * traversal in table order.  mark_visited != 0 also leaves the reference's `visited` marks
 * (status 2) on the nodes of the printed branches; pruned nodes are skipped either way. */
typedef struct {
  uint64_t n_branches;         /* records written */
  uint64_t n_branching_nodes;  /* nodes with more than one edge on a side */
  uint64_t n_nodes;            /* sum of the record lengths in nodes */
  uint64_t n_refused;          /* neighbours an earlier branch had marked */
  uint64_t n_cut;              /* walks stopped by the 300-node limit */
  uint64_t bytes;              /* size of the file */
} lasv_branch_stats;
int lasv_graph_write_branches(lasv_graph *g, const char *fasta_path, int mark_visited,
                              lasv_branch_stats *stats /* may be NULL */);

/* ------------------------------------------------------------------ (C) -- */
/* Records in table order, each kmer[N] | uint32 covg | uint8 edges = 8N+5
 * bytes (element.c:894-910).  Returns the record count, or a negative error. */
long long lasv_graph_dump_ctx_records(lasv_graph *g, uint8_t *out, uint64_t capacity_records);
/* db_graph_dump_single_colour_binary_of_colour0 (dB_graph_population.c:3730):
 * BINVERSION 6 header (file_reader.c:2930-2994) + records. */
int lasv_graph_write_ctx(lasv_graph *g, const char *path, uint32_t mean_read_len,
                         uint64_t total_seq);
int lasv_ctx_header(int kmer_size, uint32_t mean_read_len, uint64_t total_seq, uint8_t *out128);
/* The records alone, appended to an existing file: the ranks of a hash-sharded build write one `.ctx` by
 * appending their shards in rank order behind rank 0's header (the shards hold disjoint key sets and, laid end to
 * end, are the reference's table order -- SURVEY 8e).  Returns the number of records written or a negative error. */
long long lasv_graph_append_ctx_records(lasv_graph *g, const char *path);
/* load_multicolour_binary_from_filename_into_graph (file_reader.c:1652): */
int lasv_graph_load_ctx_records(lasv_graph *g, const uint8_t *records, uint64_t n_records);
int lasv_graph_load_ctx_file(lasv_graph *g, const char *path, uint32_t *mean_read_len,
                             uint64_t *total_seq);

/* Sequence files -> graph.  load_se/pe_seq_data_into_graph_colour
 * (file_reader.c:475,589): FASTQ / FASTA / plain, optionally gzipped; the
 * quality cutoff is RAW ASCII here, as in the reference's inner loaders. */
int lasv_graph_load_se_file(lasv_graph *g, const char *path, int quality_cutoff,
                            lasv_load_stats *stats);
int lasv_graph_load_pe_files(lasv_graph *g, const char *path1, const char *path2,
                             int quality_cutoff, lasv_load_stats *stats);
/* load_se_filelist / load_pe_filelists (file_reader.c:789,910): qual_thresh +
 * ascii_fq_offset is the cutoff; paths inside a list are relative to it. */
int lasv_graph_load_se_filelist(lasv_graph *g, const char *list, int qual_thresh,
                                int ascii_fq_offset, unsigned int *files_loaded,
                                lasv_load_stats *stats);
int lasv_graph_load_pe_filelists(lasv_graph *g, const char *list1, const char *list2,
                                 int qual_thresh, int ascii_fq_offset, unsigned int *pairs_loaded,
                                 lasv_load_stats *stats);

/* ------------------------------------------------------------------ (A) -- */
/* The record loop of load_multicolour_binary_from_filename_into_graph
 * (file_reader.c:1652-1710: hash_table_insert + add_edges + db_node_update_coverage
 * per record, single-threaded -- SURVEY 8f rank 1) on the GPU.  The caller parses the
 * header with the reference's own check_binary_signature_NEW (which also updates its
 * GraphInfo) and passes the file offset of the first record; the caller's
 * reference-layout table is merged in first if it already holds k-mers and is then
 * replaced by the result (hole-free buckets, next_element[], unique_kmers,
 * collisions[]).  Returns the number of records read or a negative LASV_ERR_*.
 * integration/ctx_loader_shim.c is the reference-side binding. */
long long lasv_ref_hash_table_load_ctx_records(lasv_ref_hash_table *t, const char *path, long record_offset);

/* db_graph_print_supernodes_defined_by_func_of_colours (dB_graph_population.c:2696-2803)
 * on the caller's reference-layout table: the table is copied to the device as it is
 * (same slot order, so the traversal order is the caller's), the supernodes are
 * enumerated there (lasv_graph_write_supernodes above) and the reference's FASTA is
 * written to fasta_path.  Nodes whose edges were reset by the reference's cleaning
 * (pruned) are isolated and print nothing, as in the reference.  The table is not
 * modified.  integration/supernode_shim.c is the reference-side binding. */
int lasv_ref_hash_table_write_supernodes(lasv_ref_hash_table *t, const char *fasta_path, int max_length,
                                         lasv_supernode_stats *stats /* may be NULL */);
/* ... failing with LASV_ERR_LIMIT instead of printing a supernode of more than max_length edges whole
 * (lasv_graph_write_supernodes_strict). */
int lasv_ref_hash_table_write_supernodes_strict(lasv_ref_hash_table *t, const char *fasta_path, int max_length,
                                                lasv_supernode_stats *stats /* may be NULL */);

/* The cleaning laSV always runs, `--remove_low_coverage_supernodes T` (graph_operations.c:
 * 1069-1090), on the caller's reference-layout table:
 *   LASV_CLEAN_TIPS        db_graph_clip_tips_in_union_of_all_colours (dB_graph_population.c:
 *                          2563-2608 with :365-470; tips of up to kmer_size + 1 edges)
 *   LASV_CLEAN_SUPERNODES  db_graph_remove_errors_considering_covg_and_topology (:3595-3649
 *                          with :3371-3462 and :485-530), coverage threshold `threshold`
 * Both passes of the reference mutate the graph while they traverse it; the result here is the
 * reference's for the caller's slot order: the table is copied to the device as it is, the paths
 * between non-interior nodes are enumerated there, the reference's traversal is replayed on those
 * path records (cleaning_host.h) and the decisions are applied on the device; the table comes
 * back with pruned nodes marked `pruned` (status 3) and their edges reset, everything else
 * `none`, as the reference leaves it.  Not reproduced: the --max_var_len walk limit inside the
 * supernode computation.  integration/clean_shim.c is the reference-side binding. */
#define LASV_CLEAN_TIPS 1
#define LASV_CLEAN_SUPERNODES 2
typedef struct {
  uint64_t tips_clipped;
  uint64_t tip_nodes_pruned;
  uint64_t supernodes_pruned;
  uint64_t nodes_pruned;     /* all nodes that went from none to pruned */
} lasv_clean_stats;
int lasv_ref_hash_table_clean(lasv_ref_hash_table *t, int what, int threshold, lasv_clean_stats *stats /* may be NULL */);
/* The same with the reference's walk limit (`max_expected_sup` = --max_var_len, dB_graph_population.c:3595-3649):
 * the reference judges a supernode from a walk of at most max_length edges each way, starting at a node with
 * 0 < coverage <= threshold.  If such a walk would be cut short -- a path of more than max_length edges with a
 * low-coverage node inside or at an end -- LASV_ERR_LIMIT is returned and the table is left as it was
 * (integration/clean_shim.c then runs the reference's own CPU function).  max_length <= 0: no limit. */
int lasv_ref_hash_table_clean_limited(lasv_ref_hash_table *t, int what, int threshold, int max_length,
                                      lasv_clean_stats *stats /* may be NULL */);
/* The same on a graph handle, in either arrangement of the device table -- no host copy of the table.
 * Pruned nodes stay in the table with status `pruned` and no edges, as in the reference: they are
 * skipped by lasv_graph_dump_ctx_records / lasv_graph_write_ctx (db_node_check_status_not_pruned,
 * element.c:795-802), lasv_graph_summary, the supernode and branch printers, and exported as `pruned`
 * by lasv_graph_export_table; lasv_graph_unique_kmers keeps counting them (hash_table->unique_kmers
 * does too).  lasv_graph_write_ctx afterwards writes the cleaning flags and the threshold into the
 * header as the reference's --dump_binary does after --remove_low_coverage_supernodes
 * (file_reader.c:2930-2994, graph_operations.c:1084-1088).  With lasv_graph_write_branches this is the
 * whole of `--mode build` after the load (run_laSV.sh:178) on the device. */
int lasv_graph_clean(lasv_graph *g, int what, int threshold, lasv_clean_stats *stats /* may be NULL */);
int lasv_graph_clean_limited(lasv_graph *g, int what, int threshold, int max_length, lasv_clean_stats *stats /* may be NULL */);

/* db_graph_print_all_branches_for_specific_person_or_pop (print_branches.c:131-199) on the
 * caller's reference-layout table (lasv_graph_write_branches above, in the caller's slot
 * order).  As in the reference, `visited` marks are cleared first (:133) and the nodes of the
 * printed branches come back `visited`; nothing else changes.  integration/branch_shim.c is
 * the reference-side binding. */
int lasv_ref_hash_table_write_branches(lasv_ref_hash_table *t, const char *fasta_path,
                                       lasv_branch_stats *stats /* may be NULL */);

/* ------------------------------------------- one process, several GPUs -- */
/* The hash-sharded build of SURVEY 8e for callers that are ONE process and one thread, like the reference's
 * main() (graph_operations.c:746-765): a graph of 2^number_bits buckets x bucket_size slots split over
 * n_devices GPUs (a power of two <= 16; the same device may be named more than once), shard o = the buckets
 * whose index has high bits o, on devices[o].  Host batches go to the devices in turn: copy + partition kernel
 * there (records by owner and table region into that device's send stripes), one exchange of fixed-size blocks
 * per `round_bytes` of reads and device (cudaMemcpyPeerAsync: copy engines over NVLink between peers), one insert
 * per exchange on every owner (streamed when dense).  round_bytes = 0: 2 GB, less if memory is short.
 * Replaces hash_table_find_or_insert's single table (hash_table.c:270-327) by n of them with the same
 * bucket-of-key rule; the result is handed back as ONE reference-layout table by
 * lasv_multi_graph_export_table (keys that left their first-choice bucket are re-inserted along the
 * reference's own rehash chain over all 2^number_bits buckets).  Everything is synchronous for the caller
 * except the queued GPU work; statistics / export / flush wait for it. */
typedef struct lasv_multi_graph lasv_multi_graph;
lasv_multi_graph *lasv_multi_graph_new(int kmer_size, int number_bits, int bucket_size, int max_rehash_tries,
                                       const int *devices, int n_devices, uint64_t round_bytes);
void lasv_multi_graph_free(lasv_multi_graph **m);
int lasv_multi_graph_n_shards(lasv_multi_graph *m);
/* One batch in the layout of lasv_graph_load_reads_host (pinned buffers for full speed); the buffers may be
 * reused when the call returns. */
int lasv_multi_graph_load_reads_host(lasv_multi_graph *m, const uint8_t *seq, const uint8_t *qual,
                                     uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                                     int quality_cutoff);
/* load_se_filelist_into_graph_colour / load_pe_filelists_into_graph_colour (file_reader.c:789,910) over all shards. */
int lasv_multi_graph_load_se_filelist(lasv_multi_graph *m, const char *list, int qual_thresh, int ascii_fq_offset,
                                      unsigned int *files_loaded, lasv_load_stats *stats);
int lasv_multi_graph_load_pe_filelists(lasv_multi_graph *m, const char *list1, const char *list2, int qual_thresh,
                                       int ascii_fq_offset, unsigned int *pairs_loaded, lasv_load_stats *stats);
/* Exchange and insert what the send stripes still hold; wait for all devices. */
int lasv_multi_graph_flush(lasv_multi_graph *m);
/* Sums over the shards (lasv_graph_stats, lasv_graph_rehash_histogram, lasv_graph_summary). */
int lasv_multi_graph_stats(lasv_multi_graph *m, lasv_load_stats *out, uint64_t *contig_hist, uint64_t hist_size);
int lasv_multi_graph_rehash_histogram(lasv_multi_graph *m, long long out16[16]);
int lasv_multi_graph_summary(lasv_multi_graph *m, lasv_table_summary *out);
/* The whole graph as one reference-layout HashTable: `elements` receives 2^number_bits * bucket_size Elements,
 * next_element (may be NULL) the fill of every bucket, *unique_kmers (may be NULL) the node count.  The shards
 * are finalised: no inserts afterwards. */
int lasv_multi_graph_export_table(lasv_multi_graph *m, void *elements, short *next_element, long long *unique_kmers);

/* Exact reference signatures (file_reader.h:86-118); `boolean` is signed char
 * (global.h:37).  With LASV_B200_DEVICES=0,1,... (or "all") in the environment and an empty table they
 * build on several GPUs (lasv_multi_graph above), else on LASV_B200_DEVICE (default 0).
 * They build the graph on the GPU and leave db_graph->table,
 * next_element, unique_kmers and collisions[] as the reference loader would
 * (hole-free buckets, reference rehash chain), so every consumer in
 * src/dB_graph_operations runs unchanged.  Unsupported arguments
 * (homopol_limit != 0, remove_dups == true, colour lists) die() with a clear
 * message: laSV never passes them (graph_operations.c:741-765). */
void load_se_filelist_into_graph_colour(
    char *se_filelist_path, int qual_thresh, int homopol_limit, signed char remove_dups_se,
    char ascii_fq_offset, int colour, lasv_ref_hash_table *db_graph, char is_colour_list,
    unsigned int *total_files_loaded, unsigned long long *total_bad_reads,
    unsigned long long *total_dup_reads, unsigned long long *total_bases_read,
    unsigned long long *total_bases_loaded, unsigned long *readlen_count_array,
    unsigned long readlen_count_array_size);
void load_pe_filelists_into_graph_colour(
    char *pe_filelist_path1, char *pe_filelist_path2, int qual_thresh, int homopol_limit,
    signed char remove_dups_pe, char ascii_fq_offset, int colour, lasv_ref_hash_table *db_graph,
    char is_colour_lists, unsigned int *total_file_pairs_loaded,
    unsigned long long *total_bad_reads, unsigned long long *total_dup_reads,
    unsigned long long *total_bases_read, unsigned long long *total_bases_loaded,
    unsigned long *readlen_count_array, unsigned long readlen_count_array_size);

/* ------------------------------------------------- bench / test utilities -- */
typedef struct {
  uint64_t seed;
  uint64_t genome_len;
  uint32_t read_len;
  uint32_t insert;
  uint32_t err_q20;      /* probabilities are scaled by 2^20 */
  uint32_t n_q20;
  uint32_t lowq_q20;
  uint32_t err_lowq_pct;
  uint32_t tail_min;
  uint32_t tail_max;
  uint32_t tiling;
  uint32_t tiling_step;
} lasv_synth_params;
/* Deterministic synthetic reads, stride read_len+1 ('\n' after every read).
 * The host and device versions produce identical bytes. */
int lasv_synth_reads_device(const lasv_synth_params *p, uint64_t first_read, uint64_t n_reads,
                            uint8_t *d_seq, uint8_t *d_qual, void *cuda_stream);
int lasv_synth_reads_host(const lasv_synth_params *p, uint64_t first_read, uint64_t n_reads,
                          uint8_t *seq, uint8_t *qual);
/* n_ops uniform random 32-byte read-modify-writes over n_sectors sectors: the
 * random-access HBM ceiling quoted next to the roofline. */
int lasv_random_access_probe(void *d_buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed,
                             void *cuda_stream);
/* mode 0: 16 B read + 32-bit atomic; 1: read only; 2: atomic only; 3: 64 B read + atomic. */
int lasv_random_access_probe_mode(void *d_buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed,
                                  int mode, void *cuda_stream);
/* First-choice bucket (hash_value.c:767-773, rehash 0) of each record. */
int lasv_graph_record_buckets_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                     uint32_t *d_buckets, void *cuda_stream);
/* PROTOTYPE, not used by the loaders: the streamed insert planned for the next round.  The
 * records of lasv_graph_extract_records_device, SORTED by group = first-choice bucket / buckets_per_group
 * (buckets from lasv_graph_record_buckets_device), with d_group_offsets[g] .. [g+1] delimiting the records of
 * group g (n_buckets / buckets_per_group + 1 entries): one CTA per group brings the group's lines into shared
 * memory, applies its records there (same find-or-insert protocol, shared-memory atomics) and streams the lines
 * back; a record whose bucket holds W other keys leaves the group and is inserted through the ordinary path
 * afterwards (*n_spilled counts them).  Same graph as lasv_graph_insert_records_device. */
int lasv_graph_insert_grouped_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                             const uint64_t *d_group_offsets, uint32_t buckets_per_group,
                                             uint64_t *n_spilled /* may be NULL */, void *cuda_stream);
/* The sort it needs, in the library (counting sort by group: histogram, scan, scatter): d_sorted gets the n_records
 * records grouped, d_group_offsets the n_buckets / buckets_per_group + 1 group boundaries. */
int lasv_graph_group_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                    uint32_t buckets_per_group, uint32_t *d_sorted, uint64_t *d_group_offsets,
                                    void *cuda_stream);
/* The parallel inflate of ordinary gzip files on its own (host only, no GPU): what the file loaders use for .gz
 * input of 8 MB and more that is not BGZF (the reference: zlib's gzgetc, one byte at a time,
 * libs/seq_file/seq_common.h:102-103).  The compressed stream is cut into chunks of chunk_kb KB; every thread but
 * the first finds the start of a deflate block in its chunk and inflates from there without knowing the 32 KB of
 * text before it (16-bit symbols; resolved when the chunk before it has met that start on the same bit, which also
 * proves the start was one).  out4 = {bytes of text, CRC-32 of the text, chunks confirmed, chunks dropped}.
 * LASV_ERR_IO: not a gzip file, or corrupt (a member's CRC-32 / length is checked like zlib does). */
int lasv_seqfile_inflate_check(const char *path, int threads, uint64_t chunk_kb, uint64_t *out4);
/* The file loaders' parser on its own (host only, no GPU): reads per file, bases and
 * an order-independent checksum of every (sequence, quality) pair, parsed with
 * `threads` parser threads per file exactly as the loaders do (large uncompressed
 * FASTQ files are cut at record boundaries and parsed slice by slice; anything else
 * gets one sequential parser).  out4 = {reads of file 1, reads of file 2, bases,
 * checksum}.  Used by the tests to check the parallel parse against the sequential. */
int lasv_seqfile_parse_check(const char *path1, const char *path2, int threads, uint64_t batch_bytes,
                             uint64_t *out4, int *producers_used);
/* The same with the producers set up as the file loaders set them up for the device parse (fastq_parse.cu): batches
 * of FASTQ TEXT -- slices of mapped plain files, windows of inflated BGZF / gzip text cut at record boundaries with
 * the tail carried over.  The text batches are taken apart by the host parser here, so the producers' side of that
 * path is checked without a GPU; *raw_batches (may be NULL) counts them. */
int lasv_seqfile_parse_check_text(const char *path1, const char *path2, int threads, uint64_t batch_bytes,
                                  uint64_t *out4, int *producers_used, uint64_t *raw_batches);

/* Host-only: parse a whole sequence file with the loader's reader (FASTQ /
 * FASTA / plain, optionally gzipped; grammar of libs/seq_file) into the stream
 * layout of lasv_graph_load_reads_host.  Returns the number of reads, or a
 * negative error; *has_qual tells whether the file carries quality scores. */
long long lasv_seqfile_to_streams(const char *path, uint8_t *seq, uint8_t *qual, uint64_t cap_bytes,
                                  uint64_t *offsets, uint64_t cap_reads, int *has_qual);
/* Number of kernels this library has launched since load (bench gpu_launches). */
uint64_t lasv_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* LASV_B200_H_ */
