// graph_kernels.cuh -- launch interface between the host runtime (capi.cu) and
// the sm_100a kernels (graph_kernels.cu).  Plain structs, no torch types.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "table.cuh"

namespace lasv {

// Records travelling between the extract kernel, the all-to-all and the insert
// kernel: (2N+1) uint32 each = key words as (lo,hi) pairs in word order, then
// one word {edges : 8, count : 24}.
__host__ __device__ inline int record_u32(int N) { return 2 * N + 1; }

enum BuildMode { MODE_FUSED = 0, MODE_RECORDS = 1, MODE_ROUTE = 2, MODE_BIN = 3 };

// Region bins of the partitioned insert: k-mer occurrences of a batch grouped by
// the slice of the table they will touch (bin = bucket >> bin_shift), so that the
// insert kernel walks the table region by region (TLB-, L2- and DRAM-row-friendly)
// instead of at random.  One fixed-capacity stripe of records per bin.  A record is
// written with ONE vector store: scattered stores cost per request, not per byte
// (tools/ubench: 3 x 8 B = 15 ms, 1 x 32 B = 5 ms per 180 M records) -- but every
// record byte is DRAM traffic three times over (partition write, exchange copy,
// insert read) in kernels that are bound by exactly that, so records are as small as
// the key allows:
//   16 bytes  N = 1: {key, aux};  N = 2 and k <= 56: {key0 | edges << 48 | count << 56, key1}
//             (the top key word has 16 spare bits; the insert side hashes again)
//   32 bytes  N = 2 and k > 56: {key0, key1, aux, 0};  N = 3: {key0, key1, key2, aux}
// aux = {lookup3 c : 32, edges : 8, count : 24}.
__host__ __device__ inline uint32_t bin_record_words(int N, int k) { return (N == 1 || (N == 2 && k <= 56)) ? 2u : 4u; }
constexpr uint32_t COUNT_STRIDE = 32;  // uint32 units = 128 bytes
struct BinView {
  uint64_t *recs;       // [n_bins*cap*rec_words], see stripe_record_at
  uint32_t rec_words;   // 2 or 4 64-bit words per record (bin_record_words)
  unsigned int *count;  // [(block+1)*n_owners*COUNT_STRIDE] records wanted per stripe (stripe_counter_at), ONE COUNTER PER 128-BYTE LINE: the L2
                        // atomic unit serialises per line, 8 counters in a line would share one queue
                        // (a count may exceed cap: see spill_ok)
  uint32_t cap;
  uint32_t n_bins;
  // producer side (build_kernel<MODE_BIN>): bin = (lookup3 c & bin_mask) >> bin_shift.  In a hash-sharded
  // job bin_mask covers the GLOBAL bucket index, so the owner GPU is the top bits of the bin and the
  // stripes of one owner are contiguous (they travel as one block of the all-to-all).
  uint32_t bin_mask;
  uint32_t bin_shift;
  uint32_t spill_ok;    // 1: a record that finds its stripe full goes straight into the (local) table;
                        // 0: it goes to its owner's spill area (overflow flag if that is full too)
  // consumer side (insert_bins_kernel): the stripes are swept region by region; with n_src senders the
  // stripe of (sender s, local region r) is at index s*bins_per_src + r and the sweep takes all senders
  // of a region before the next region (n_src = 1 << src_shift, n_bins = n_src*bins_per_src)
  uint32_t src_shift;
  uint32_t bins_per_src;
  // Record placement, both sides.  The stripes of one owner (producer) / one sender (consumer) form a
  // BLOCK of `1 << block_shift` stripes that is stored interleaved in runs of STRIPE_RUN records:
  //   record idx of stripe (o, r)  ->  o*owner_stride + ((idx / RUN)*block + r)*RUN + idx % RUN
  // Bins fill at the same pace, so at any moment the write frontiers of ALL stripes of a block lie in
  // the same few pages (block * RUN * 32 B = 1-8 MB) instead of one page per stripe: the scatter of the
  // partition kernel stops missing the TLB on every store.  The consumer reads a stripe in 4 KB runs.
  // Behind the block of every owner sit `spill_cap` records for whatever found its stripe full (a k-mer
  // that occurs a million times in one batch lands in ONE stripe): the owner inserts them after its
  // regions.  owner_stride = (cap << block_shift) + spill_cap records; every owner has block+1 counters.
  uint32_t block_shift;
  uint32_t spill_cap;
  // Fused producer -> exchange (hash-sharded build over peer memory): owner_recs[o] != nullptr for all owners means
  // that owner o's block (its stripes + spill area) does not live at recs + o * owner_stride but at owner_recs[o] --
  // the slot of THIS sender in owner o's receive stripes, mapped into this process (symmetric memory / peer access),
  // so that the partition kernel's stores ARE the exchange (NVLink stores for the remote owners).
  uint64_t *owner_recs[16];
  uint32_t direct;      // 1: use owner_recs
  uint32_t src_used;    // consumer side: senders that can hold records (0: all 1 << src_shift of them) -- sizes the scratch
};
constexpr uint32_t STRIPE_RUN = 128;  // records; cap and spill_cap are multiples of it
__host__ __device__ inline uint64_t owner_stride_records(const BinView &b) {
  return ((uint64_t)b.cap << b.block_shift) + b.spill_cap;
}
__host__ __device__ inline uint64_t stripe_record_at(const BinView &b, uint32_t stripe, uint32_t idx) {
  const uint32_t o = stripe >> b.block_shift, r = stripe & ((1u << b.block_shift) - 1u);
  return (uint64_t)o * owner_stride_records(b) + (((uint64_t)(idx / STRIPE_RUN) * STRIPE_RUN) << b.block_shift) +
         (uint64_t)r * STRIPE_RUN + (idx % STRIPE_RUN);
}
// address of record idx of stripe `stripe` (producer side): through the owner's own base pointer in the fused mode
__device__ __forceinline__ uint64_t *stripe_record_ptr(const BinView &b, uint32_t stripe, uint32_t idx) {
  if (!b.direct) return b.recs + stripe_record_at(b, stripe, idx) * b.rec_words;
  const uint32_t o = stripe >> b.block_shift, r = stripe & ((1u << b.block_shift) - 1u);
  const uint64_t in_block = (((uint64_t)(idx / STRIPE_RUN) * STRIPE_RUN) << b.block_shift) + (uint64_t)r * STRIPE_RUN + (idx % STRIPE_RUN);
  return b.owner_recs[o] + in_block * b.rec_words;
}
__device__ __forceinline__ uint64_t *spill_record_ptr(const BinView &b, uint32_t owner, uint32_t idx) {
  const uint64_t in_block = ((uint64_t)b.cap << b.block_shift) + idx;
  if (!b.direct) return b.recs + ((uint64_t)owner * owner_stride_records(b) + in_block) * b.rec_words;
  return b.owner_recs[owner] + in_block * b.rec_words;
}
__host__ __device__ inline uint64_t spill_record_at(const BinView &b, uint32_t owner, uint32_t idx) {
  return (uint64_t)owner * owner_stride_records(b) + ((uint64_t)b.cap << b.block_shift) + idx;
}
// counter of stripe (o, r), r == block: the owner's spill area
__host__ __device__ inline uint32_t stripe_counter_at(const BinView &b, uint32_t owner, uint32_t r) {
  return (owner * ((1u << b.block_shift) + 1u) + r) * COUNT_STRIDE;
}

__host__ __device__ inline uint64_t bin_aux(uint32_t c, uint32_t edge, uint32_t count) {
  return (uint64_t)c | ((uint64_t)(edge & 0xFFu) << 32) | ((uint64_t)count << 40);
}

struct BuildParams {
  const uint8_t *seq;   // sequence stream: reads back to back, each followed by >= 1 separator byte
  const uint8_t *qual;  // quality stream, byte-aligned with seq; nullptr = no quality filter (FASTA)
  uint64_t n_bytes;     // lead + stream length; buffers must be readable up to round_up(n_bytes, 16)
  uint32_t lead;        // seq/qual are aligned DOWN to 16 bytes; the stream starts `lead` bytes in
  int k;
  int cutoff;           // RAW ascii cutoff (threshold + offset), signed-char compare, '<=' is bad
  TableView table;
  GraphCounters *ctr;
  // MODE_RECORDS: one output array; MODE_ROUTE: n_dest bins of bin_capacity records each
  uint32_t *rec_out;
  unsigned long long *rec_count;  // [1] (RECORDS) or [n_dest] (ROUTE)
  uint64_t rec_capacity;          // records (RECORDS) / records per bin (ROUTE)
  unsigned long long *rec_overflow;
  int n_dest;       // power of two
  int owner_shift;  // owner = (hash.c >> owner_shift) & (n_dest-1)
  // MODE_FUSED over a table larger than the TLB reach: the batch is walked n_pass
  // times and pass `pass` only inserts the k-mers whose bucket lies in its
  // 1/n_pass slice of the table (n_pass power of two, pass_shift = L - log2 n_pass)
  int n_pass, pass, pass_shift;
  BinView bins;  // MODE_BIN
  int ctas_per_sm;  // > 0: cap on resident CTAs per SM (kernels that share the GPU with another kernel)
};

struct ReadStats {  // loader bookkeeping (file_reader.c:460-470,579,584)
  unsigned long long n_reads;
  unsigned long long bad_reads;     // reads without k consecutive good bases
  unsigned long long bases_read;
  unsigned long long bases_loaded;  // sum of contig lengths
  unsigned long long n_contigs;
};

struct TableSummary {  // order-independent digest of the node multiset
  unsigned long long n_nodes;
  unsigned long long sum_covg;
  unsigned long long sum_edge_bits;
  unsigned long long fingerprint;  // sum over nodes of node_fingerprint() mod 2^64
};

struct SynthParams {  // synthetic paired-end reads (bench / tests input, not reference behaviour)
  uint64_t seed;
  uint64_t genome_len;
  uint32_t read_len;
  uint32_t insert;        // outer fragment length
  uint32_t err_q20;       // P(substitution) * 2^20 per base
  uint32_t n_q20;         // P(N) * 2^20 per base
  uint32_t lowq_q20;      // P(isolated low-quality base, Phred 2..5) * 2^20
  uint32_t err_lowq_pct;  // % of substituted bases that also get a low quality
  uint32_t tail_min;      // low-quality 3' tail length range (Phred 2); 0,0 = none
  uint32_t tail_max;
  uint32_t tiling;        // 1: error-free forward reads tiling the genome with k-1 overlap
  uint32_t tiling_step;   // read_len - k + 1 in tiling mode
};

int sm_count();

void launch_build(int N, BuildMode mode, const BuildParams &p, cudaStream_t st);
void launch_insert_records(int N, const uint32_t *recs, uint64_t n_recs, const TableView &t,
                           GraphCounters *ctr, cudaStream_t st);
void launch_insert_bins(int N, const BinView &b, const TableView &t, GraphCounters *ctr, int ctas_per_sm,
                        cudaStream_t st);
void launch_read_stats(const uint8_t *seq, const uint8_t *qual, const uint64_t *offsets,
                       uint64_t n_reads, int sep, int k, int cutoff, ReadStats *stats,
                       unsigned long long *hist, uint32_t hist_size, cudaStream_t st);
void launch_table_summary(int N, const TableView &t, TableSummary *out, cudaStream_t st);
void launch_bucket_counts(int N, const TableView &t, uint64_t first, uint64_t n_buckets, uint32_t *counts,
                          cudaStream_t st);
void launch_finalize(int N, const TableView &t, int16_t *next_element, cudaStream_t st);
void launch_extract_misplaced(int N, const TableView &t, int16_t *next_element, uint8_t *out, unsigned long long *n_out,
                              uint64_t cap, cudaStream_t st);
void launch_dump_ctx(int N, const TableView &t, uint64_t first, uint64_t n_buckets, const uint64_t *bucket_offsets,
                     uint8_t *out, cudaStream_t st);
void launch_ingest_ctx(int N, const uint8_t *recs, uint64_t n_recs, const TableView &t,
                       GraphCounters *ctr, cudaStream_t st);
void launch_ingest_elements(int N, const uint8_t *elements, uint64_t n_slots, const TableView &t,
                            GraphCounters *ctr, cudaStream_t st);
void launch_find(int N, const TableView &t, const uint64_t *keys, uint64_t n, uint32_t *covg,
                 uint8_t *edges, uint8_t *found, cudaStream_t st);
void launch_exclusive_scan_u32_to_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *tmp,
                                      size_t tmp_bytes, cudaStream_t st);
size_t exclusive_scan_tmp_bytes(uint64_t n);
void launch_synth(const SynthParams &p, uint64_t first_read, uint64_t n_reads, uint8_t *seq,
                  uint8_t *qual, cudaStream_t st);
void synth_host(const SynthParams &p, uint64_t first_read, uint64_t n_reads, uint8_t *seq,
                uint8_t *qual);
void launch_random_rmw(uint64_t *buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed, int mode,
                       cudaStream_t st);
void launch_record_buckets(int N, const uint32_t *recs, uint64_t n_recs, uint32_t bucket_mask,
                           uint32_t *buckets, cudaStream_t st);

}  // namespace lasv
