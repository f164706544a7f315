// capi.cu -- host runtime behind include/lasv_b200.h: graph handle, chunked
// host->device pipeline, sequence-file loaders, .ctx reader/writer and the
// reference-named drop-in loaders.  All compute is in graph_kernels.cu; there
// is no CPU implementation of any of it here.
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <libgen.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/lasv_b200.h"
#include "graph_kernels.cuh"
#include "supernodes_host.h"
#include "cleaning_host.h"
#include "branches_host.h"
#include "groups.cuh"
#include "streamed_insert.cuh"
#include "fastq_parse.cuh"
#include "ranking.cuh"
#include "seq_reader.h"

using namespace lasv;

namespace {

thread_local std::string g_err;
thread_local int g_err_code = 0;  // of the last fail(), for callers that got a NULL handle back
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  g_err_code = code;
  return code;
}

int last_error_code() { return g_err_code < 0 ? g_err_code : LASV_ERR_CUDA; }

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(LASV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),   \
                  __FILE__, __LINE__);                                                      \
  } while (0)

const char *TABLE_FULL_MSG =
    "Dear user - you have not allocated enough memory to contain your sequence data. Either "
    "allocate more memory (have you done your calculations right? have you allowed for sequencing "
    "errors?), or threshold more harshly on quality score, and try again. Aborting mission.";

constexpr uint32_t HIST_BINS = 1u << 16;       // device contig-length histogram (last bin = longer)
constexpr uint64_t CHUNK_BYTES = 64ull << 20;  // host->device pipeline granularity (staging buffers are this big)
constexpr uint64_t CHUNK_READS = 4ull << 20;

// Environment knobs of the insert path, read ONCE per graph (lasv_graph_new), never on a hot entry point.
//   LASV_B200_INSERT = direct | binned | random | streamed   (tests, experiments)
//       direct    fused find-or-insert from the partition kernel's tiles, no partition
//       binned    force the partition even on small tables; the insert kind is chosen as usual
//       random    partition + random-access insert (insert_bins_kernel), never streamed
//       streamed  partition + streamed insert (streamed_insert.cuh), whatever the batch size
//   LASV_B200_BINS = regions to partition into (power of two), LASV_B200_PASSES, LASV_B200_CHUNK_MB / _KB
struct Knobs {
  int insert = 0;  // 0 none, 1 direct, 2 binned
  long bins = 0;
  int passes = 0;
  long chunk_mb = 0;
  long chunk_kb = 0;  // tests: chunks of a few KB make a small input wrap the staging ring many times
  int pipe_ctas = 2;  // LASV_B200_PIPE_CTAS: CTAs per SM of the partition kernel when an insert may run beside it
  int acc_slots = 32; // LASV_B200_ACC_SLOTS: sets of stripes the accumulation is divided into at most (1: one set, as in
                      // the first version of the accumulation)
};
struct Staging {  // one leg of the double-buffered host->device pipeline
  uint8_t *d_seq = nullptr, *d_qual = nullptr;
  uint64_t *d_offsets = nullptr;
  cudaEvent_t copied = nullptr, consumed = nullptr;
  uint64_t *h_offsets = nullptr;  // pinned: a pageable source would make the "async" copy block the host
  // device-side FASTQ parse (fastq_parse.cuh), allocated on the first raw chunk: the chunk's text and the parse scratch
  uint8_t *d_raw = nullptr;
  FqScratch fq;
  FqResult *h_fq = nullptr;  // pinned
};

}  // namespace

// Pinned host buffers for one batch of parsed reads (file loaders).  Kept by the graph between loader
// calls: pinning memory costs more than parsing a few hundred MB of FASTQ.
struct PinnedBatch : lasv::HostBatch {
  bool ok = false;
  PinnedBatch(uint64_t bytes, uint64_t reads) {
    cap_bytes = bytes;
    cap_reads = reads;
    ok = cudaMallocHost(&seq, bytes) == cudaSuccess && cudaMallocHost(&qual, bytes) == cudaSuccess &&
         cudaMallocHost(&offsets, (reads + 1) * sizeof(uint64_t)) == cudaSuccess;
    if (ok) reset();
  }
  ~PinnedBatch() {
    if (seq) cudaFreeHost(seq);
    if (qual) cudaFreeHost(qual);
    if (offsets) cudaFreeHost(offsets);
  }
};

struct lasv_graph {
  std::vector<std::unique_ptr<PinnedBatch>> host_pool;  // loader batches, all of one size
  int k = 0, N = 0, L = 0, W = 0, max_rehash = 15, device = 0;
  uint64_t n_buckets = 0, n_slots = 0;
  size_t slot_bytes = 0;
  uint8_t *d_lines = nullptr;  // 2^L buckets x NL lines of 128 bytes (kmer_core.cuh, "line layout")
  uint64_t table_bytes = 0;
  GraphCounters *d_ctr = nullptr;
  ReadStats *d_stats = nullptr;
  unsigned long long *d_hist = nullptr;
  unsigned long long *d_overflow = nullptr;
  bool finalized = false;
  // set by lasv_graph_clean, written into the .ctx header (graph_info.c:147-173, graph_operations.c:1084-1088)
  bool cleaned_tips = false, cleaned_sups = false;
  int clean_threshold = -1;
  cudaStream_t compute = nullptr, copy = nullptr;
  // >= 3 legs: the copy of chunk i+1 may start while chunk i-1 is still being consumed.  More legs (a ring,
  // lasv_graph_set_staging) let the copies of the asynchronous host call run ahead of the kernels for as long as an
  // insert of the accumulated stripes keeps the compute stream busy.
  std::vector<Staging> stage = std::vector<Staging>(3);
  uint64_t stage_reads = CHUNK_READS;  // reads a leg's offset arrays hold
  bool staging_ready = false;
  int stage_turn = 0;  // the leg the next chunk goes to (the ring carries over from call to call)
  // scratch of the partitioned insert (region bins), grown on demand.  Pipelined mode uses both legs:
  // batch i+1 is binned on the caller's stream while batch i is inserted on `ins`.
  struct BinLeg {
    BinView v{};
    uint64_t alloc_records = 0;
    uint32_t alloc_n = 0;
    cudaEvent_t binned = nullptr, inserted = nullptr;
    bool in_flight = false;
  } leg[2];
  int turn = 0;
  bool pipeline = false;
  cudaStream_t ins = nullptr;
  // which insert follows the partition kernel (read once from LASV_B200_INSERT at creation; lasv_graph_set_insert_mode):
  // AUTO = streamed (streamed_insert.cuh) when the batch puts ~a record on every table line, else random access
  int insert_mode = LASV_INSERT_AUTO;
  double streamed_min_per_line = 0.75;  // records (stripe capacity) per 128-byte table line from which streaming pays
  StreamedScratch ss;
  // stripes that accumulate several partition launches before ONE insert (host-buffer and file loaders): stream bytes
  // binned into leg 0 since its last insert, and the batch size its stripes are sized for (0 = not accumulating)
  // acc_limit: stream bytes the caller (lasv_graph_set_accumulation: acc_user) or a loader call wants per insert;
  // acc_cap_records: records the stripes were sized for; acc_records: sum of the k-mer bounds of the batches binned
  // since the last insert (an insert is due before a batch whose bound does not fit any more); acc_stream: the stream
  // those partition launches were queued on (the insert follows on it)
  uint64_t acc_bytes = 0, acc_limit = 0, acc_cap_records = 0, acc_records = 0;
  // How full the stripes really get: a batch is booked with its k-mer BOUND (bytes - reads * k) times acc_fill, the
  // ratio of k-mers to bound that the previous accumulation showed (+3 %), read back without a synchronisation (a
  // probe of the k-mer counter queued in front of every insert).  Quality filtering, Ns and errors keep real reads
  // ~12 % under the bound, which is 12 % more batches per pass over the table.  An underestimate is harmless: a
  // record that finds its stripe full goes straight into the table (spill_ok).
  // The accumulating stripes are divided into up to 32 SLOTS, each a complete set of stripes of its own (the consumer
  // sees them as 32 "senders" of one region each, the layout of the multi-GPU receive side): a batch goes into the
  // current slot, the next slot is taken when that one is booked out.  The write frontiers of a slot's stripes start
  // together and drift apart only as far as one slot's worth of records lets them, so the partition kernel's scattered
  // stores stay inside a few hundred 2 MB pages however much the accumulation holds (with ONE set of stripes the
  // kernel went from 3.2 ms for the first batch of an accumulation to 4.65 ms for the 17th).
  uint32_t acc_slots = 0, acc_slot = 0;
  uint64_t acc_slot_records = 0, acc_slot_booked = 0;
  double acc_fill_sized = 0.0;  // acc_fill the slots were sized with
  double acc_fill = 1.0;
  uint64_t acc_bound_sum = 0, probe_bound = 0, probe_prev_kmers = 0;
  unsigned long long *h_probe = nullptr;  // pinned
  cudaEvent_t probe_ev = nullptr;
  bool probe_pending = false;
  bool acc_user = false, acc_stream_used = false;
  cudaStream_t acc_stream = nullptr;
  uint64_t n_streamed = 0, n_random = 0;  // inserts by kind (lasv_graph_insert_counts)
  // One shard of a lasv_multi_graph: the chunks of the host-buffer calls are binned by (owner, region) into the
  // multi graph's send stripes of this device instead of this graph's own stripes
  struct ShardSink {
    int n_owners = 0;
    uint32_t bins_per_owner = 0, cap = 0, spill = 0;
    void *records = nullptr, *counts = nullptr;
    bool fresh = true;          // the next launch resets the counters (first batch after an exchange)
    uint64_t bound = 0;         // k-mer bounds of the batches binned since the last exchange
  } *sink = nullptr;
  struct lasv_multi_graph *multi = nullptr;  // set on shard 0 while a file loader drives the multi graph through it
  // the *_strict / *_limited entry points: the reference's walk limit (--max_var_len).  Where a walk of the reference
  // would be cut short by it (and the output would become the reference's overlapping pieces), they return
  // LASV_ERR_LIMIT and change / write nothing -- the reference-side shims then run the reference's own CPU function
  bool sn_strict = false;
  int walk_limit = 0;
  Knobs knobs;
  // optional per-launch timing of the two kernels of the partitioned path (bench: roofline of the
  // dominant kernel from CUDA events on the stream the kernel is launched on)
  bool timing = false;
  std::vector<cudaEvent_t> t_events;  // pairs (begin, end), kind of pair i in t_kind[i] (LASV_T_*)
  std::vector<int> t_kind;
  size_t t_used = 0;

  TableView view() const {
    return make_table_view(d_lines, N, L, W, max_rehash, finalized);
  }
};

namespace {

// Entry points of the load path itself: no side effects beyond selecting the device.
int use_hot(lasv_graph *g) {
  if (!g) return fail(LASV_ERR_ARG, "NULL graph");
  CUDA_TRY(cudaSetDevice(g->device));
  return LASV_OK;
}
int flush_accumulated(lasv_graph *g, cudaStream_t st);
int queue_fill_probe(lasv_graph *g, cudaStream_t st);
void poll_fill_probe(lasv_graph *g);
int load_raw_fastq_host(lasv_graph *g, const uint8_t *text, uint64_t n_bytes, int quality_cutoff, uint64_t *n_reads_out);
int bin_reads_impl(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual, uint64_t n_bytes, int quality_cutoff,
                   int n_owners, uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                   uint32_t spill_capacity_records, void *d_records, void *d_counts, void *cuda_stream, bool append,
                   void *const *owner_records = nullptr);
// Every other entry point: records that were partitioned but not inserted yet (lasv_graph_set_accumulation) go into
// the table first, so that whatever reads the table or its counters sees every read loaded so far.
int use(lasv_graph *g) {
  if (int e = use_hot(g)) return e;
  if (g->acc_records) {
    cudaStream_t st = g->acc_stream;
    if (int e = flush_accumulated(g, st)) return e;
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return LASV_OK;
}

// A NULL stream argument means the CUDA default stream, as everywhere in CUDA (torch's
// default stream is that stream too, so CUDA events recorded there bracket our launches).
cudaStream_t stream_of(lasv_graph *, void *s) { return (cudaStream_t)s; }

void free_staging(lasv_graph *g) {
  for (auto &s : g->stage) {
    cudaFree(s.d_seq);
    cudaFree(s.d_qual);
    cudaFree(s.d_offsets);
    if (s.h_offsets) cudaFreeHost(s.h_offsets);
    cudaFree(s.d_raw); cudaFree(s.fq.nl); cudaFree(s.fq.n_nl); cudaFree(s.fq.lens); cudaFree(s.fq.result); cudaFree(s.fq.temp);
    if (s.h_fq) cudaFreeHost(s.h_fq);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.consumed) cudaEventDestroy(s.consumed);
    s = Staging{};
  }
  g->staging_ready = false;
}

int ensure_staging(lasv_graph *g) {
  if (g->staging_ready) return LASV_OK;
  for (auto &s : g->stage) {
    CUDA_TRY(cudaMalloc(&s.d_seq, CHUNK_BYTES + 256));
    CUDA_TRY(cudaMalloc(&s.d_qual, CHUNK_BYTES + 256));
    CUDA_TRY(cudaMalloc(&s.d_offsets, (g->stage_reads + 1) * sizeof(uint64_t)));
    CUDA_TRY(cudaMallocHost(&s.h_offsets, (g->stage_reads + 1) * sizeof(uint64_t)));
    CUDA_TRY(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
  }
  g->staging_ready = true;
  return LASV_OK;
}

void fill_stats(lasv_load_stats *o, const ReadStats &rs, const GraphCounters &c) {
  o->n_reads = rs.n_reads;
  o->bad_reads = rs.bad_reads;
  o->bases_read = rs.bases_read;
  o->bases_loaded = rs.bases_loaded;
  o->n_contigs = rs.n_contigs;
  o->kmers_inserted = c.kmers_inserted;
  o->unique_kmers = c.unique_kmers;
}

int fetch_counters(lasv_graph *g, cudaStream_t st, ReadStats *rs, GraphCounters *c,
                   unsigned long long *overflow) {
  CUDA_TRY(cudaMemcpyAsync(rs, g->d_stats, sizeof *rs, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(c, g->d_ctr, sizeof *c, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(overflow, g->d_overflow, sizeof *overflow, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return LASV_OK;
}

BuildParams build_params(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                         uint64_t n_bytes, int cutoff) {
  BuildParams p;
  memset(&p, 0, sizeof p);
  // the kernels read 16-byte chunks: align the pointers down, skip the lead bytes
  const uint32_t lead = (uint32_t)((uintptr_t)d_seq & 15u);
  p.seq = d_seq - lead;
  p.qual = d_qual ? d_qual - lead : nullptr;
  p.lead = lead;
  p.n_bytes = n_bytes + lead;
  p.k = g->k;
  p.cutoff = cutoff;
  p.table = g->view();
  p.ctr = g->d_ctr;
  p.rec_overflow = g->d_overflow;
  p.n_dest = 1;
  p.owner_shift = g->L;
  p.n_pass = 1;
  p.pass = 0;
  p.pass_shift = g->L;
  return p;
}

void free_streamed(lasv_graph *g) {
  StreamedScratch &s = g->ss;
  cudaFree(s.counts);
  cudaFree(s.offsets);
  cudaFree(s.cursor);
  cudaFree(s.scan_tmp);
  cudaFree(s.sorted);
  cudaFree(s.spill);
  cudaFree(s.n_spill);
  s = StreamedScratch{};
}

void free_bins(lasv_graph *g) {
  free_streamed(g);
  for (auto &l : g->leg) {
    cudaFree(l.v.recs);
    cudaFree(l.v.count);
    l.v = BinView{};
    l.alloc_records = 0;
    l.alloc_n = 0;
  }
}

// next timing event (created on demand, reused after a reset); nullptr when timing is off
cudaEvent_t timing_event(lasv_graph *g) {
  if (g->t_used == g->t_events.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    g->t_events.push_back(e);
  }
  return g->t_events[g->t_used++];
}
// A timed span of kind `kind` on stream st: records its begin event on construction, its end event on end().
struct Span {
  cudaEvent_t e1 = nullptr;
  cudaStream_t st;
  Span(lasv_graph *g, int kind, cudaStream_t st_) : st(st_) {
    if (!g->timing) return;
    cudaEvent_t e0 = timing_event(g);
    e1 = timing_event(g);
    if (!e0 || !e1) { e1 = nullptr; return; }
    g->t_kind.resize(g->t_used / 2);
    g->t_kind[g->t_used / 2 - 1] = kind;
    cudaEventRecord(e0, st);
  }
  void end() {
    if (e1) cudaEventRecord(e1, st);
    e1 = nullptr;
  }
};

// The caller's stream waits for every insert queued on the internal stream.
int join_inserts(lasv_graph *g, cudaStream_t st) {
  for (auto &l : g->leg)
    if (l.in_flight) CUDA_TRY(cudaStreamWaitEvent(st, l.inserted, 0));
  return LASV_OK;
}

Knobs read_knobs(int *insert_mode) {
  Knobs k;
  if (const char *e = getenv("LASV_B200_INSERT")) {
    if (!strcmp(e, "direct")) k.insert = 1;
    else if (!strcmp(e, "binned")) k.insert = 2;
    else if (!strcmp(e, "random")) { k.insert = 2; *insert_mode = LASV_INSERT_RANDOM; }
    else if (!strcmp(e, "streamed")) { k.insert = 2; *insert_mode = LASV_INSERT_STREAMED; }
  }
  if (const char *e = getenv("LASV_B200_BINS")) k.bins = atol(e);
  if (const char *e = getenv("LASV_B200_PASSES")) k.passes = atoi(e);
  if (const char *e = getenv("LASV_B200_CHUNK_MB")) k.chunk_mb = atol(e);
  if (const char *e = getenv("LASV_B200_CHUNK_KB")) k.chunk_kb = atol(e);
  if (const char *e = getenv("LASV_B200_PIPE_CTAS")) k.pipe_ctas = std::max(0, std::min(8, atoi(e)));
  if (const char *e = getenv("LASV_B200_ACC_SLOTS")) k.acc_slots = std::max(1, std::min(32, atoi(e)));
  return k;
}

// dense: the stripes are meant for the streamed insert (about a record or more per table line); a sparse batch goes
// through the random-access insert, which sweeps the stripes one by one and pays for every stripe it looks at.
uint32_t region_count(const lasv_graph *g, bool dense = true) {
  uint32_t n_bins = 1;
  // regions of <= 64 MB: the random-access insert doubles its rate once the region all SMs work on at a time is
  // ~100 MB or less and is flat below that (profiles/: bins sweep), while fewer bins keep the scatter's open
  // write streams few
  while ((uint64_t)n_bins * (64ull << 20) < g->table_bytes && n_bins < 65536u && n_bins < g->n_buckets) n_bins <<= 1;
  // ... and of <= 256 buckets on large tables: the sorting passes of the streamed insert rank a chunk of one
  // region's records by bucket in shared memory, and their output runs are chunk / buckets-per-region records long
  // (the partition kernel costs the same from 2 K to 16 K stripes, +11 % at 64 K: profiles/r2)
  if (dense && g->table_bytes >= (256ull << 20))
    while ((uint64_t)n_bins * 256u < g->n_buckets && n_bins < 65536u) n_bins <<= 1;
  const long v = g->knobs.bins;  // tests: force a bin count on small tables
  if (v >= 1 && v <= 65536 && (v & (v - 1)) == 0 && (uint64_t)v <= g->n_buckets) n_bins = (uint32_t)v;
  return n_bins;
}

// Partitioned insert applies when the table is much larger than L2 and the batch
// puts a useful number of records into every region.
bool want_binned(const lasv_graph *g, uint64_t n_bytes, uint32_t *n_bins_out, bool dense = true) {
  const uint32_t n_bins = region_count(g, dense);
  *n_bins_out = n_bins;
  bool binned = g->table_bytes >= (256ull << 20) && n_bytes >= (uint64_t)n_bins * 256u;
  if (g->knobs.insert == 1) binned = false;
  // an insert kind asked for by name (LASV_B200_INSERT, lasv_graph_set_insert_mode) implies the partition
  else if (g->knobs.insert == 2 || g->insert_mode != LASV_INSERT_AUTO) binned = n_bins > 1 || g->n_buckets > 1;
  return binned;
}

int log2_u32(uint32_t v) {
  int lg = 0;
  while ((1u << lg) < v) lg++;
  return lg;
}

// Upper bound on the k-mers of a batch of n_bytes stream bytes in n_reads reads: a read of length l yields at
// most l-k+1 k-mers and occupies l+1 stream bytes; without the read count, n_bytes itself (ANY batch of that
// many bytes: few long reads, FASTA contigs, small k).
uint64_t kmer_bound(const lasv_graph *g, uint64_t n_bytes, uint64_t n_reads) {
  if (n_reads && n_reads * (uint64_t)g->k < n_bytes) return n_bytes - n_reads * (uint64_t)g->k;
  return n_bytes;
}

// Records per stripe for `bound` records spread over n_bins stripes.  Stripes fill like balls into bins by a
// well-mixed hash: mean + 8 sigma (+ a little) is never reached in practice.
uint64_t stripe_capacity(uint64_t bound, uint32_t n_bins) {
  const double mean = (double)bound / n_bins;
  const uint64_t cap = (uint64_t)(mean * 1.01 + 8.0 * sqrt(mean) + 64.0);
  return (cap + STRIPE_RUN - 1) / STRIPE_RUN * STRIPE_RUN;  // whole runs (stripe_record_at)
}

int ensure_bins(lasv_graph *g, lasv_graph::BinLeg &l, uint64_t bound, uint32_t n_bins) {
  // a fuller stripe spills its excess straight into the table, so this is a sizing choice, not a limit
  const uint64_t cap = stripe_capacity(bound, n_bins);
  if (cap > 0xFFFFFFF0ull) return fail(LASV_ERR_ARG, "batch too large for the bin stripes");
  const uint64_t want = cap * n_bins;
  if (l.alloc_records < want || l.alloc_n != n_bins) {
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(l.v.recs);
    cudaFree(l.v.count);
    l.v = BinView{};
    l.alloc_records = 0;
    CUDA_TRY(cudaMalloc(&l.v.recs, want * (size_t)bin_record_words(g->N, g->k) * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(&l.v.count, (size_t)(n_bins + 1) * COUNT_STRIDE * sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(l.v.count, 0, (size_t)(n_bins + 1) * COUNT_STRIDE * sizeof(unsigned int)));
    l.alloc_records = want;
    l.alloc_n = n_bins;
  }
  l.v.cap = (uint32_t)(l.alloc_records / n_bins / STRIPE_RUN * STRIPE_RUN);
  l.v.block_shift = (uint32_t)log2_u32(n_bins);
  l.v.rec_words = bin_record_words(g->N, g->k);
  l.v.spill_cap = 0;  // one owner: an overfull stripe spills straight into the table
  l.v.n_bins = n_bins;
  l.v.bin_mask = (uint32_t)(g->n_buckets - 1);
  l.v.bin_shift = (uint32_t)(g->L - log2_u32(n_bins));
  l.v.spill_ok = 1;
  l.v.src_shift = 0;
  l.v.src_used = 0;
  l.v.direct = 0;
  l.v.bins_per_src = n_bins;
  return LASV_OK;
}

// bytes of the counters a view's consumer reads: (stripes + spill counter) per sender
size_t view_counter_bytes(const BinView &v) {
  const uint32_t per = 1u << v.block_shift;
  return (size_t)(v.n_bins / per) * (per + 1u) * COUNT_STRIDE * sizeof(unsigned int);
}

// Records per stripe of one SLOT of the accumulation (see lasv_graph::acc_slots): mean + 4 sigma -- a few dozen of the
// 65 536 x slots stripes overflow by a few records per accumulation, and a record that finds its stripe full goes
// straight into the table (spill_ok), so this is a sizing choice, not a limit; every per cent of slack here is a per
// cent fewer batches per pass over the table.
uint64_t slot_stripe_capacity(uint64_t slot_records, uint32_t n_bins) {
  const double mean = (double)slot_records / n_bins;
  const uint64_t cap = (uint64_t)(mean * 1.005 + 4.0 * sqrt(mean) + 8.0);
  return (cap + STRIPE_RUN - 1) / STRIPE_RUN * STRIPE_RUN;
}

// The accumulating stripes of leg 0: n_slots complete sets of n_bins stripes for slot_records records each.  l.v is
// the CONSUMER's view (n_slots rounded up to a power of two "senders"; the extra ones never receive anything).
int ensure_acc_bins(lasv_graph *g, lasv_graph::BinLeg &l, uint64_t slot_records, uint32_t n_bins, uint32_t n_slots) {
  const uint64_t cap = slot_stripe_capacity(slot_records, n_bins);
  if (cap > 0xFFFFFFF0ull) return fail(LASV_ERR_ARG, "batch too large for the bin stripes");
  uint32_t s2 = 1;
  while (s2 < n_slots) s2 <<= 1;
  const uint64_t want = cap * n_bins * n_slots;
  const uint32_t want_n = n_bins * s2;
  if (l.alloc_records < want || l.alloc_n != want_n) {
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(l.v.recs);
    cudaFree(l.v.count);
    l.v = BinView{};
    l.alloc_records = 0;
    CUDA_TRY(cudaMalloc(&l.v.recs, want * (size_t)bin_record_words(g->N, g->k) * sizeof(uint64_t)));
    const size_t cbytes = (size_t)(n_bins + 1) * s2 * COUNT_STRIDE * sizeof(unsigned int);
    CUDA_TRY(cudaMalloc(&l.v.count, cbytes));
    CUDA_TRY(cudaMemset(l.v.count, 0, cbytes));
    l.alloc_records = want;
    l.alloc_n = want_n;
  }
  l.v.cap = (uint32_t)cap;
  l.v.block_shift = (uint32_t)log2_u32(n_bins);
  l.v.rec_words = bin_record_words(g->N, g->k);
  l.v.spill_cap = 0;  // one owner: an overfull stripe spills straight into the table
  l.v.n_bins = want_n;
  l.v.bin_mask = (uint32_t)(g->n_buckets - 1);
  l.v.bin_shift = (uint32_t)(g->L - log2_u32(n_bins));
  l.v.spill_ok = 1;
  l.v.src_shift = (uint32_t)log2_u32(s2);
  l.v.src_used = n_slots;
  l.v.bins_per_src = n_bins;
  return LASV_OK;
}

// the PRODUCER's view of slot `slot` of those stripes: one set of n_bins stripes, its own counters
BinView acc_slot_view(const BinView &c, uint32_t slot) {
  BinView p = c;
  const uint32_t n_bins = c.bins_per_src;
  p.recs = c.recs + (uint64_t)slot * ((uint64_t)c.cap << c.block_shift) * c.rec_words;
  p.count = c.count + (size_t)slot * (n_bins + 1u) * COUNT_STRIDE;
  p.n_bins = n_bins;
  p.src_shift = 0;
  p.src_used = 0;
  return p;
}

// Streamed or random-access insert for these stripes?  Streaming costs two passes over the table's lines
// whatever the batch; it pays from about 0.75 records per line on.
bool want_streamed(const lasv_graph *g, const BinView &v, uint64_t n_records, StreamedGeom *geom) {
  if (g->insert_mode == LASV_INSERT_RANDOM) return false;
  if (!streamed_geometry(g->view(), v, geom)) return false;
  if (g->insert_mode == LASV_INSERT_STREAMED) return true;
  const double lines = (double)g->table_bytes / LINE_BYTES;
  return (double)n_records >= g->streamed_min_per_line * lines;
}

int ensure_streamed(lasv_graph *g, const StreamedGeom &geom, uint32_t rec_words) {
  StreamedScratch &s = g->ss;
  const uint32_t buckets = geom.buckets_per_slab;
  if (s.alloc_records >= geom.slab_cap && s.alloc_buckets >= buckets && s.n_spill) return LASV_OK;
  CUDA_TRY(cudaDeviceSynchronize());
  free_streamed(g);
  const uint64_t recs = std::max<uint64_t>(geom.slab_cap, 1);
  CUDA_TRY(cudaMalloc(&s.counts, ((size_t)buckets + 1) * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&s.offsets, ((size_t)buckets + 1) * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&s.cursor, (size_t)buckets * sizeof(uint32_t)));
  s.scan_tmp_bytes = std::max<size_t>(streamed_scan_tmp_bytes(buckets + 1), 16);
  CUDA_TRY(cudaMalloc(&s.scan_tmp, s.scan_tmp_bytes));
  CUDA_TRY(cudaMalloc(&s.sorted, recs * rec_words * sizeof(uint64_t)));
  CUDA_TRY(cudaMalloc(&s.spill, recs * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&s.n_spill, sizeof(unsigned long long)));
  s.alloc_records = recs;
  s.alloc_buckets = buckets;
  return LASV_OK;
}

int check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CUDA_TRY(cudaGetLastError());
  return LASV_OK;
}

int check_streams(const void *seq, const void *qual) {
  if (!seq) return fail(LASV_ERR_ARG, "sequence stream is NULL");
  if (qual && (((uintptr_t)seq ^ (uintptr_t)qual) & 15u) != 0)
    return fail(LASV_ERR_ARG, "sequence and quality streams must share their alignment modulo 16");
  return LASV_OK;
}

}  // namespace

extern "C" {

const char *lasv_last_error(void) { return g_err.c_str(); }
const char *lasv_version(void) { return "lasv_b200 0.1 (sm_100a)"; }
uint64_t lasv_kernel_launches(void) { return g_launches.load(); }

int lasv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

lasv_graph *lasv_graph_new(int kmer_size, int number_bits, int bucket_size, int max_rehash_tries,
                           int device) {
  if (kmer_size < 1 || kmer_size > 95 || (kmer_size & 1) == 0) {
    fail(LASV_ERR_ARG, "kmer_size must be odd and in [1,95] (cmd_line.c:805; MAXK 31/63/95), got %d",
         kmer_size);
    return nullptr;
  }
  if (number_bits < 1 || number_bits > 31 || bucket_size < 1 || bucket_size > 32767 ||
      max_rehash_tries < 0 || max_rehash_tries > 15) {
    fail(LASV_ERR_ARG, "bad table geometry: number_bits=%d bucket_size=%d max_rehash_tries=%d",
         number_bits, bucket_size, max_rehash_tries);
    return nullptr;
  }
  if (lasv_device_count() <= device || device < 0) {
    fail(LASV_ERR_CUDA, "no usable CUDA device %d (this library has no CPU fallback)", device);
    return nullptr;
  }
  if (const char *e = getenv("LASV_B200_L2FETCH")) {  // experiment: L2 fetch granularity (32/64/128)
    cudaSetDevice(device);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
  }
  lasv_graph *g = new lasv_graph();
  g->k = kmer_size;
  g->N = (2 * kmer_size) / 64 + 1;  // graph_operations.c:620
  g->L = number_bits;
  g->W = bucket_size;
  g->max_rehash = max_rehash_tries;
  g->device = device;
  g->n_buckets = 1ull << number_bits;
  g->n_slots = g->n_buckets * (uint64_t)bucket_size;
  g->slot_bytes = 8 * (size_t)g->N + 8;
  g->table_bytes = table_bytes(g->view());
  bool ok = cudaSetDevice(device) == cudaSuccess &&
            cudaMalloc(&g->d_lines, g->table_bytes) == cudaSuccess &&
            cudaMalloc(&g->d_ctr, sizeof(GraphCounters)) == cudaSuccess &&
            cudaMalloc(&g->d_stats, sizeof(ReadStats)) == cudaSuccess &&
            cudaMalloc(&g->d_hist, HIST_BINS * sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc(&g->d_overflow, sizeof(unsigned long long)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&g->compute, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&g->copy, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&g->ins, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMallocHost(&g->h_probe, sizeof(unsigned long long)) == cudaSuccess &&
       cudaEventCreateWithFlags(&g->probe_ev, cudaEventDisableTiming) == cudaSuccess;
  for (auto &l : g->leg)
    ok = ok && cudaEventCreateWithFlags(&l.binned, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&l.inserted, cudaEventDisableTiming) == cudaSuccess;
  if (const char *e = getenv("LASV_B200_PIPELINE")) g->pipeline = atoi(e) != 0;
  g->knobs = read_knobs(&g->insert_mode);
  if (ok) ok = lasv_graph_clear(g, nullptr) == LASV_OK && cudaDeviceSynchronize() == cudaSuccess;
  if (!ok) {
    if (g_err.empty() || g_err.find("failed") == std::string::npos)
      fail(LASV_ERR_CUDA, "could not allocate hash table of size %llu: %s",
           (unsigned long long)g->n_slots, cudaGetErrorString(cudaGetLastError()));
    lasv_graph_free(&g);
    return nullptr;
  }
  return g;
}

void lasv_graph_free(lasv_graph **gp) {
  if (!gp || !*gp) return;
  lasv_graph *g = *gp;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  cudaFree(g->d_lines);
  cudaFree(g->d_ctr);
  cudaFree(g->d_stats);
  cudaFree(g->d_hist);
  cudaFree(g->d_overflow);
  free_bins(g);
  free_staging(g);
  for (auto &l : g->leg) {
    if (l.binned) cudaEventDestroy(l.binned);
    if (l.inserted) cudaEventDestroy(l.inserted);
  }
  for (cudaEvent_t e : g->t_events) cudaEventDestroy(e);
  if (g->h_probe) cudaFreeHost(g->h_probe);
  if (g->probe_ev) cudaEventDestroy(g->probe_ev);
  if (g->ins) cudaStreamDestroy(g->ins);
  if (g->compute) cudaStreamDestroy(g->compute);
  if (g->copy) cudaStreamDestroy(g->copy);
  delete g;
  *gp = nullptr;
}

int lasv_graph_clear(lasv_graph *g, void *cuda_stream) {
  if (int e = use_hot(g)) return e;
  cudaStream_t st = stream_of(g, cuda_stream);
  if (int e = join_inserts(g, st)) return e;
  CUDA_TRY(cudaMemsetAsync(g->d_lines, 0, g->table_bytes, st));
  CUDA_TRY(cudaMemsetAsync(g->d_ctr, 0, sizeof(GraphCounters), st));
  CUDA_TRY(cudaMemsetAsync(g->d_stats, 0, sizeof(ReadStats), st));
  CUDA_TRY(cudaMemsetAsync(g->d_hist, 0, HIST_BINS * sizeof(unsigned long long), st));
  CUDA_TRY(cudaMemsetAsync(g->d_overflow, 0, sizeof(unsigned long long), st));
  if (g->acc_records) {  // records binned but not yet inserted belong to the graph that is being cleared
    CUDA_TRY(cudaStreamSynchronize(g->acc_stream));
    CUDA_TRY(cudaMemsetAsync(g->leg[0].v.count, 0, view_counter_bytes(g->leg[0].v), st));
    g->acc_records = 0;
    g->acc_bytes = 0;
    g->acc_slot = 0;
    g->acc_slot_booked = 0;
  }
  if (g->probe_pending) {  // a probe of the counters that are being reset: let it land and use it (the read set may go on)
    CUDA_TRY(cudaEventSynchronize(g->probe_ev));
    poll_fill_probe(g);
  }
  g->probe_prev_kmers = 0;
  g->acc_bound_sum = 0;
  g->finalized = false;
  return LASV_OK;
}

long long lasv_graph_unique_kmers(lasv_graph *g) {
  if (use(g)) return -1;
  unsigned long long v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;  // work may be queued on any stream
  if (cudaMemcpy(&v, &g->d_ctr->unique_kmers, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (long long)v;
}

/* Pipelined inserts: with `on`, the partitioned path inserts batch i on an internal stream while
 * the caller's stream already bins batch i+1 (both kernels are then sized to share the SMs).  Every
 * entry point of this library that reads the table joins the internal stream itself; a caller that
 * reads the table with its OWN kernels, or times with events on its stream, calls lasv_graph_flush. */
int lasv_graph_set_pipeline(lasv_graph *g, int on) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  for (auto &l : g->leg) l.in_flight = false;
  g->pipeline = on != 0;
  return LASV_OK;
}

/* Per-launch CUDA-event timing of the partitioned path's two kernels.  enable: start (and reset) or
 * stop collecting; kernel_times: device-synchronises, then sums the durations collected so far. */
int lasv_graph_enable_kernel_timing(lasv_graph *g, int on) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  g->timing = on != 0;
  g->t_used = 0;
  g->t_kind.clear();
  return LASV_OK;
}

int lasv_graph_kernel_times_by_kind(lasv_graph *g, double ms[LASV_T_KINDS], uint64_t n[LASV_T_KINDS]) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  for (int k = 0; k < LASV_T_KINDS; k++) { ms[k] = 0; n[k] = 0; }
  for (size_t i = 0; 2 * i + 1 < g->t_used && i < g->t_kind.size(); i++) {
    float x = 0;
    CUDA_TRY(cudaEventElapsedTime(&x, g->t_events[2 * i], g->t_events[2 * i + 1]));
    const int k = g->t_kind[i];
    if (k < 0 || k >= LASV_T_KINDS) continue;
    ms[k] += x;
    n[k]++;
  }
  return LASV_OK;
}

int lasv_graph_kernel_times(lasv_graph *g, double *partition_ms, double *insert_ms, uint64_t *n_batches) {
  double ms[LASV_T_KINDS];
  uint64_t n[LASV_T_KINDS];
  if (int e = lasv_graph_kernel_times_by_kind(g, ms, n)) return e;
  if (partition_ms) *partition_ms = ms[LASV_T_PARTITION];
  if (insert_ms) *insert_ms = ms[LASV_T_INSERT];
  if (n_batches) *n_batches = n[LASV_T_INSERT];
  return LASV_OK;
}

int lasv_graph_set_insert_mode(lasv_graph *g, int mode, double min_records_per_line) {
  if (int e = use(g)) return e;
  if (mode != LASV_INSERT_AUTO && mode != LASV_INSERT_RANDOM && mode != LASV_INSERT_STREAMED)
    return fail(LASV_ERR_ARG, "insert mode must be LASV_INSERT_AUTO, _RANDOM or _STREAMED");
  g->insert_mode = mode;
  if (min_records_per_line > 0) g->streamed_min_per_line = min_records_per_line;
  return LASV_OK;
}

int lasv_graph_insert_counts(lasv_graph *g, uint64_t *n_streamed, uint64_t *n_random) {
  if (!g) return fail(LASV_ERR_ARG, "NULL graph");
  if (n_streamed) *n_streamed = g->n_streamed;
  if (n_random) *n_random = g->n_random;
  return LASV_OK;
}

int lasv_graph_flush(lasv_graph *g, void *cuda_stream) {
  if (int e = use_hot(g)) return e;
  cudaStream_t st = stream_of(g, cuda_stream);
  if (g->acc_records) {  // accumulated stripes: the insert is queued on the stream of their partition launches
    cudaStream_t as = g->acc_stream;
    if (int e = flush_accumulated(g, as)) return e;
    if (as != st) {
      CUDA_TRY(cudaEventRecord(g->leg[0].inserted, as));
      CUDA_TRY(cudaStreamWaitEvent(st, g->leg[0].inserted, 0));
    }
  }
  return join_inserts(g, st);
}

long long lasv_graph_capacity(lasv_graph *g) { return g ? (long long)g->n_slots : -1; }
int lasv_graph_words_per_kmer(lasv_graph *g) { return g ? g->N : -1; }
size_t lasv_graph_element_bytes(lasv_graph *g) { return g ? g->slot_bytes : 0; }
void *lasv_graph_device_table(lasv_graph *g) { return g ? g->d_lines : nullptr; }
uint64_t lasv_graph_device_table_bytes(lasv_graph *g) { return g ? g->table_bytes : 0; }

// ------------------------------------------------------------- the hot path
namespace {

// K3 for everything the stripes `v` hold (about n_records of them), queued on st: the streamed insert when the
// batch is dense enough on the table's lines, else the random-access insert.  own_counters: the stripes are the
// graph's (counters zeroed afterwards, ready for the next partition launch).
int insert_stripes(lasv_graph *g, const BinView &v, uint64_t n_records, int ctas_per_sm, bool allow_streamed,
                   bool own_counters, cudaStream_t st) {
  StreamedGeom geom;
  Span total(g, LASV_T_INSERT, st);
  if (allow_streamed && want_streamed(g, v, n_records, &geom)) {
    if (int e = ensure_streamed(g, geom, v.rec_words)) return e;
    std::function<void(int, int)> span;
    std::vector<Span> open;
    if (g->timing)
      span = [&](int kind, int begin) {
        if (begin) open.emplace_back(g, kind, st);
        else { open.back().end(); open.pop_back(); }
      };
    uint64_t launches = 0;
    if (launch_streamed_insert(g->N, v, g->view(), g->d_ctr, g->d_overflow, geom, g->ss, st, &launches, span))
      return fail(LASV_ERR_CUDA, "streamed insert: kernel attributes: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    g->n_streamed++;
  } else {
    launch_insert_bins(g->N, v, g->view(), g->d_ctr, ctas_per_sm, st);
    if (int e = check_launch()) return e;
    g->n_random++;
  }
  total.end();
  if (own_counters)
    CUDA_TRY(cudaMemsetAsync(v.count, 0, view_counter_bytes(v), st));
  return LASV_OK;
}

// Records of bin records + streamed-insert scratch that fit in the free device memory, at most `want`.
uint64_t accumulation_records(lasv_graph *g, uint64_t want) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
  const uint64_t rec_bytes = 8ull * bin_record_words(g->N, g->k);
  // what the graph holds already for this purpose is reallocated, so it counts as free
  const uint64_t have = g->leg[0].alloc_records * rec_bytes + g->ss.alloc_records * (rec_bytes + 4);
  const uint64_t reserve = 3ull << 30;  // staging legs, statistics, the caller's next buffers
  if (free_b + have <= reserve) return 0;
  const double budget = (double)(free_b + have - reserve);
  // stripes (1.09: the slots' stripes hold mean + 4 sigma in whole runs of 128 records) and one slab (<= 1/8 of them)
  // of sorted records + spill indices
  const double per_record = 1.09 * ((double)rec_bytes + ((double)rec_bytes + 4.0) / 8.0);
  const uint64_t fit = (uint64_t)(budget / per_record);
  return std::min(want, fit);
}

}  // namespace

extern "C++" {
namespace {
// Insert what leg 0 has accumulated since its last insert.
int flush_accumulated(lasv_graph *g, cudaStream_t st) {
  if (!g->acc_records) return LASV_OK;
  const uint64_t n = g->acc_records;
  g->acc_records = 0;
  g->acc_bytes = 0;
  g->acc_slot = 0;
  g->acc_slot_booked = 0;
  if (int e = queue_fill_probe(g, st)) return e;
  return insert_stripes(g, g->leg[0].v, n, 0, true, true, st);
}

// Queue a read-back of the k-mer counter as it stands behind the partition launches since the last probe (the
// counter is kept by the partition kernel, not by the insert): k-mers / bound of those launches is the fill estimate.
int queue_fill_probe(lasv_graph *g, cudaStream_t st) {
  if (g->h_probe && g->probe_ev && !g->probe_pending && g->acc_bound_sum) {
    CUDA_TRY(cudaMemcpyAsync(g->h_probe, &g->d_ctr->kmers_inserted, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(g->probe_ev, st));
    g->probe_bound = g->acc_bound_sum;
    g->probe_pending = true;
    g->acc_bound_sum = 0;
  }
  return LASV_OK;
}

// a finished probe updates the fill estimate (never waits)
void poll_fill_probe(lasv_graph *g) {
  if (!g->probe_pending || cudaEventQuery(g->probe_ev) != cudaSuccess) {
    cudaGetLastError();  // (cudaErrorNotReady is not an error)
    return;
  }
  g->probe_pending = false;
  static const bool adapt = [] { const char *e = getenv("LASV_B200_ACC_FILL"); return !e || atoi(e) != 0; }();  // experiments
  if (!adapt) return;
  const uint64_t now = *g->h_probe;
  if (now >= g->probe_prev_kmers && g->probe_bound) {
    const double ratio = (double)(now - g->probe_prev_kmers) / (double)g->probe_bound;
    g->acc_fill = std::min(1.0, std::max(0.05, ratio * 1.02));
  }
  g->probe_prev_kmers = now;
}
}  // namespace
}

int lasv_graph_set_accumulation(lasv_graph *g, uint64_t max_stream_bytes, uint64_t *granted) {
  if (int e = use(g)) return e;  // inserts what the previous setting left in the stripes
  g->acc_limit = max_stream_bytes;
  g->acc_user = max_stream_bytes != 0;
  g->acc_cap_records = 0;  // sized by the first batch
  if (granted) {
    // a stream byte yields at most one k-mer
    const uint64_t fit = accumulation_records(g, max_stream_bytes);
    *granted = std::min(max_stream_bytes, fit);
  }
  return LASV_OK;
}

int lasv_graph_load_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                 uint64_t n_bytes, const uint64_t *d_offsets, uint64_t n_reads,
                                 int quality_cutoff, void *cuda_stream) {
  if (int e = use_hot(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (int e = check_streams(d_seq, d_qual)) return e;
  cudaStream_t st = stream_of(g, cuda_stream);
  if (n_bytes) {
    BuildParams p = build_params(g, d_seq, d_qual, n_bytes, quality_cutoff);
    uint32_t n_bins = 1;
    const bool accumulate = g->acc_limit != 0 && !g->pipeline;
    // what the batch is expected to put into the stripes: its bound times the fill earlier batches showed (a
    // filter-heavy read set stays far below bytes - reads * k); a batch that will be inserted by random access (too
    // few records per table line for the streamed insert) is partitioned into the few 64 MB regions of round 1
    poll_fill_probe(g);
    const uint64_t bound = kmer_bound(g, n_bytes, d_offsets ? n_reads : 0);
    const uint64_t booked = std::min<uint64_t>(bound, (uint64_t)((double)bound * g->acc_fill) + 1);
    const bool dense = accumulate || g->pipeline || g->insert_mode == LASV_INSERT_STREAMED ||
                       (g->insert_mode == LASV_INSERT_AUTO &&
                        (double)booked >= g->streamed_min_per_line * ((double)g->table_bytes / LINE_BYTES));
    if (want_binned(g, accumulate ? std::max(g->acc_limit, n_bytes) : n_bytes, &n_bins, dense)) {
      // Partitioned insert: group the batch's k-mers by 64 MB table region, then insert region by
      // region: streamed through shared memory when the batch is dense on the table's lines
      // (streamed_insert.cuh), at random inside one region at a time otherwise.
      lasv_graph::BinLeg &l = g->leg[(g->pipeline && !accumulate) ? (g->turn++ & 1) : 0];
      if (l.in_flight) CUDA_TRY(cudaStreamWaitEvent(st, l.inserted, 0));  // the insert that read this leg
      if (accumulate) {
        // several partition launches fill the stripes, slot by slot; ONE insert when the next batch does not fit
        // (the slots are re-cut, while they are empty, when the fill estimate has moved since they were sized)
        const bool stale = g->acc_slots && fabs(g->acc_fill - g->acc_fill_sized) > 0.03 * g->acc_fill_sized;
        bool fits = g->acc_slots && booked <= g->acc_slot_records &&
                    g->acc_bytes + n_bytes <= std::max(g->acc_limit, n_bytes) &&  // the caller's limit, in stream bytes
                    (g->acc_slot_booked + booked <= g->acc_slot_records || g->acc_slot + 1 < g->acc_slots);
        if (g->acc_records && (g->acc_stream != st || !fits)) {
          static const bool trace = getenv("LASV_B200_ACC_TRACE") != nullptr;
          if (trace)
            fprintf(stderr, "[acc] flush: booked %llu slot %u/%u booked-in-slot %llu slot-records %llu bytes %llu+%llu limit %llu fill %.3f\n",
                    (unsigned long long)booked, g->acc_slot, g->acc_slots, (unsigned long long)g->acc_slot_booked,
                    (unsigned long long)g->acc_slot_records, (unsigned long long)g->acc_bytes, (unsigned long long)n_bytes,
                    (unsigned long long)g->acc_limit, g->acc_fill);
          if (int e = flush_accumulated(g, g->acc_stream)) return e;
        }
        const bool resize = !g->acc_records && stale;
        if (!g->acc_records) {
          if (g->acc_stream != st && g->acc_stream_used) {  // the insert just queued there reads the stripes
            CUDA_TRY(cudaEventRecord(l.inserted, g->acc_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, l.inserted, 0));
          }
          if (!g->acc_cap_records || !g->acc_slots || resize || booked > g->acc_slot_records) {
            // as many records as `acc_limit` stream bytes of batches like this one hold, if they fit in memory
            const double ratio = (double)booked / (double)n_bytes;
            const double want_d = (double)std::max(g->acc_limit, n_bytes) * ratio;
            const uint64_t want = want_d >= 1.8e19 ? ~0ull : (uint64_t)want_d + 1;
            const uint64_t fit = accumulation_records(g, ~0ull);  // what the memory holds
            g->acc_cap_records = std::max(booked, std::min(want, fit));
            // slots: one per batch like this one (a little more: batches cut at read boundaries differ by a read or
            // two), at most 32 (smaller batches then share a slot); when memory is not the limit, the last slot may
            // reach beyond what was asked for
            // (a slot shared by smaller batches holds a whole number of them)
            const uint64_t one = booked + booked / 100 + 1024, max_slots = (uint64_t)g->knobs.acc_slots;
            g->acc_slot_records = one * std::max<uint64_t>(1, ((g->acc_cap_records + max_slots - 1) / max_slots + one - 1) / one);
            // (a slot holds whole batches: count batches, not records, when memory is not the limit)
            const uint64_t per_slot = std::max<uint64_t>(1, g->acc_slot_records / std::max<uint64_t>(booked, 1));
            const uint64_t n_batches = (g->acc_cap_records + booked - 1) / std::max<uint64_t>(booked, 1);
            const uint64_t n_sl = want <= fit ? (n_batches + per_slot - 1) / per_slot : g->acc_cap_records / g->acc_slot_records;
            g->acc_slots = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(max_slots, n_sl));
            g->acc_fill_sized = g->acc_fill;
          }
          if (int e = ensure_acc_bins(g, l, g->acc_slot_records, n_bins, g->acc_slots)) return e;
        }
        if (g->acc_slot_booked + booked > g->acc_slot_records) {  // (there is a next slot: `fits`)
          g->acc_slot++;
          g->acc_slot_booked = 0;
        }
        g->acc_stream = st;
        g->acc_stream_used = true;
      } else {
        if (int e = ensure_bins(g, l, bound, n_bins)) return e;
      }
      p.bins = accumulate ? acc_slot_view(l.v, g->acc_slot) : l.v;
      p.ctas_per_sm = g->pipeline ? 2 : 0;  // pipelined: this kernel shares the SMs with the previous batch's insert
      {
        Span sp(g, LASV_T_PARTITION, st);
        launch_build(g->N, MODE_BIN, p, st);
        if (int e = check_launch()) return e;
        sp.end();
      }
      if (accumulate) {
        g->acc_bytes += n_bytes;
        g->acc_records += booked;
        g->acc_slot_booked += booked;
        g->acc_bound_sum += bound;
      } else if (!g->pipeline) {
        // (the stripes are sized for the bound; the CHOICE of the insert goes by what they are expected to hold)
        g->acc_bound_sum += bound;
        if (int e = queue_fill_probe(g, st)) return e;
        if (int e = insert_stripes(g, l.v, booked, 0, true, true, st)) return e;
      } else {
        // (random access only: the next batch's partition kernel may spill straight into the table while
        // this insert runs, which the streamed insert's staged copies would overwrite)
        CUDA_TRY(cudaEventRecord(l.binned, st));
        CUDA_TRY(cudaStreamWaitEvent(g->ins, l.binned, 0));
        if (int e = insert_stripes(g, l.v, bound, 2, false, true, g->ins)) return e;
        CUDA_TRY(cudaEventRecord(l.inserted, g->ins));
        l.in_flight = true;
      }
    } else {
      // Direct fused insert.  Random access over more than ~64 GB falls off the
      // TLB reach of a B200 (profiles/random_access_sweep_r1.jsonl): walk the batch
      // once per <=50 GB slice of the table.  LASV_B200_PASSES overrides.
      if (g->acc_records)  // (a batch too small to partition arrives while larger ones wait in the stripes)
        if (int e = flush_accumulated(g, g->acc_stream)) return e;
      int n_pass = 1;
      while (g->table_bytes / (uint64_t)n_pass > (50ull << 30) && (1 << g->L) > 2 * n_pass) n_pass *= 2;
      {
        const int v = g->knobs.passes;
        if (v >= 1 && v <= 64 && (v & (v - 1)) == 0 && v <= (1 << (g->L - 1))) n_pass = v;
      }
      int lg = 0;
      while ((1 << lg) < n_pass) lg++;
      p.n_pass = n_pass;
      p.pass_shift = g->L - lg;
      for (int pass = 0; pass < n_pass; pass++) {
        p.pass = pass;
        launch_build(g->N, MODE_FUSED, p, st);
        if (int e = check_launch()) return e;
      }
    }
  }
  if (d_offsets && n_reads) {
    launch_read_stats(d_seq, d_qual, d_offsets, n_reads, 1, g->k, quality_cutoff, g->d_stats,
                      g->d_hist, HIST_BINS, st);
    if (int e = check_launch()) return e;
  }
  return LASV_OK;
}

int lasv_graph_read_stats_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                 const uint64_t *d_offsets, uint64_t n_reads, int quality_cutoff,
                                 void *cuda_stream) {
  if (int e = use_hot(g)) return e;
  if (!n_reads) return LASV_OK;
  if (int e = check_streams(d_seq, d_qual)) return e;
  launch_read_stats(d_seq, d_qual, d_offsets, n_reads, 1, g->k, quality_cutoff, g->d_stats,
                    g->d_hist, HIST_BINS, stream_of(g, cuda_stream));
  return check_launch();
}

int lasv_graph_set_staging(lasv_graph *g, uint64_t bytes, uint64_t *granted) {
  if (int e = use_hot(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  const uint64_t leg_bytes = 2 * CHUNK_BYTES;
  size_t legs = (size_t)std::min<uint64_t>(256, std::max<uint64_t>(3, bytes / leg_bytes));
  // never more than half of what is free (the ring must not starve the stripes of the insert path)
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    const uint64_t have = g->staging_ready ? g->stage.size() * leg_bytes : 0;
    legs = (size_t)std::min<uint64_t>(legs, std::max<uint64_t>(3, (free_b + have) / 2 / leg_bytes));
  }
  if (legs != g->stage.size() || !g->staging_ready) {
    free_staging(g);
    g->stage.assign(legs, Staging{});
    // a ring's legs hold the offsets of 2^20 reads each (64 MB of reads of >= 63 bytes); three legs keep the 2^22
    g->stage_reads = legs > 3 ? (1ull << 20) : CHUNK_READS;
    if (int e = ensure_staging(g)) return e;
  }
  if (granted) *granted = (uint64_t)legs * leg_bytes;
  return LASV_OK;
}

extern "C++" {
namespace {
// The chunk that sits in staging leg `s` (streams + offsets, its copy or device parse ordered before s.copied) goes
// through the hot path on the compute stream: into this graph, or -- a shard of a multi-GPU graph -- into the send
// stripes of its device.
int consume_staged(lasv_graph *g, Staging &s, bool has_qual, uint64_t bytes, uint64_t nr, int quality_cutoff) {
  CUDA_TRY(cudaStreamWaitEvent(g->compute, s.copied, 0));
  if (g->sink) {
    lasv_graph::ShardSink &k = *g->sink;
    if (int e = bin_reads_impl(g, s.d_seq, has_qual ? s.d_qual : nullptr, bytes, quality_cutoff, k.n_owners, k.bins_per_owner,
                               k.cap, k.spill, k.records, k.counts, g->compute, !k.fresh))
      return e;
    k.fresh = false;
    k.bound += kmer_bound(g, bytes, nr);
    launch_read_stats(s.d_seq, has_qual ? s.d_qual : nullptr, s.d_offsets, nr, 1, g->k, quality_cutoff, g->d_stats, g->d_hist,
                      HIST_BINS, g->compute);
    if (int e = check_launch()) return e;
  } else if (int e = lasv_graph_load_reads_device(g, s.d_seq, has_qual ? s.d_qual : nullptr, bytes, s.d_offsets, nr,
                                                  quality_cutoff, g->compute)) {
    return e;
  }
  CUDA_TRY(cudaEventRecord(s.consumed, g->compute));
  return LASV_OK;
}

// Both host-buffer calls.  async: nothing here waits for a kernel -- the call returns once its last copy has left
// the host buffers (copy stream drained); statistics, if wanted, are copied to pinned memory behind the kernels.
int load_reads_host_impl(lasv_graph *g, const uint8_t *seq, const uint8_t *qual,
                         uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                         int quality_cutoff, lasv_load_stats *batch_stats, bool async, uint64_t *pinned_stats9) {
  if (int e = use_hot(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (!seq || !offsets) return fail(LASV_ERR_ARG, "seq and offsets are required");
  if (offsets[n_reads] != n_bytes) return fail(LASV_ERR_ARG, "offsets[n_reads] != n_bytes");
  if (async && g->pipeline) return fail(LASV_ERR_STATE, "the asynchronous host call does not combine with lasv_graph_set_pipeline");
  if (int e = ensure_staging(g)) return e;
  if (!async) CUDA_TRY(cudaDeviceSynchronize());  // earlier device-API work may sit on a caller's stream

  ReadStats rs0, rs1;
  GraphCounters c0, c1;
  unsigned long long ov = 0;
  if (batch_stats) if (int e = fetch_counters(g, g->compute, &rs0, &c0, &ov)) return e;

  // The chunks of this call are partitioned one by one (each as soon as its copy has landed) into the SAME
  // stripes, which are inserted once they hold as much as the device memory allows, and at the end of the call:
  // the insert's cost per record falls with the number of records per table line (streamed_insert.cuh).
  // With lasv_graph_set_accumulation in force the chunks join the caller's accumulation instead (and stay there).
  struct Accumulate {
    lasv_graph *g;
    bool mine;
    Accumulate(lasv_graph *g_, uint64_t limit) : g(g_), mine(g_->acc_limit == 0 && limit != 0) {
      if (mine) {
        g->acc_limit = limit;
        g->acc_cap_records = 0;
      }
    }
    ~Accumulate() {
      if (!mine) return;
      if (g->acc_records) {  // error exit: drop what was binned but not inserted
        cudaStreamSynchronize(g->compute);
        cudaMemset(g->leg[0].v.count, 0, view_counter_bytes(g->leg[0].v));
        g->acc_records = 0;
        g->acc_bytes = 0;
        g->acc_slot = 0;
        g->acc_slot_booked = 0;
      }
      g->acc_limit = 0;
      g->acc_cap_records = 0;
    }
  };
  uint32_t n_bins_call = 1;
  const bool binned_call = !g->sink && !g->pipeline && want_binned(g, n_bytes, &n_bins_call);
  Accumulate acc(g, binned_call ? n_bytes : 0);

  // chunks are cut at read boundaries, so no k-mer window straddles two chunks
  uint64_t r0 = 0;
  int leg = g->stage_turn % (int)g->stage.size(), n_chunk = async ? 3 : 0;
  while (r0 < n_reads) {
    uint64_t r1 = r0;
    const uint64_t base = offsets[r0];
    // largest r1 with offsets[r1]-base <= limit and r1-r0 <= CHUNK_READS.  The first chunks are short so
    // that the kernels start early (nothing overlaps the first copy): 1/8, 1/4, 1/2, then whole chunks.
    {
      uint64_t chunk = CHUNK_BYTES;
      if (g->knobs.chunk_mb >= 1 && (uint64_t)g->knobs.chunk_mb <= (CHUNK_BYTES >> 20))  // experiments: 1..64
        chunk = (uint64_t)g->knobs.chunk_mb << 20;
      if (g->knobs.chunk_kb >= 1 && (uint64_t)g->knobs.chunk_kb <= (CHUNK_BYTES >> 10))
        chunk = (uint64_t)g->knobs.chunk_kb << 10;
      const uint64_t limit = n_chunk < 3 ? (chunk >> (3 - n_chunk)) : chunk;
      uint64_t lo = r0 + 1, hi = std::min(n_reads, r0 + g->stage_reads);
      if (offsets[lo] - base > CHUNK_BYTES)
        return fail(LASV_ERR_ARG, "read %llu is longer than the %llu-byte pipeline chunk",
                    (unsigned long long)r0, (unsigned long long)CHUNK_BYTES);
      while (lo < hi) {
        const uint64_t mid = (lo + hi + 1) / 2;
        if (offsets[mid] - base <= limit) lo = mid;
        else hi = mid - 1;
      }
      n_chunk++;
      r1 = lo;
    }
    const uint64_t bytes = offsets[r1] - base, nr = r1 - r0;
    Staging &s = g->stage[leg];
    CUDA_TRY(cudaEventSynchronize(s.consumed));  // the kernels that read this leg are done
    for (uint64_t i = 0; i <= nr; i++) s.h_offsets[i] = offsets[r0 + i] - base;
    CUDA_TRY(cudaMemcpyAsync(s.d_seq, seq + base, bytes, cudaMemcpyHostToDevice, g->copy));
    if (qual) CUDA_TRY(cudaMemcpyAsync(s.d_qual, qual + base, bytes, cudaMemcpyHostToDevice, g->copy));
    CUDA_TRY(cudaMemcpyAsync(s.d_offsets, s.h_offsets, (nr + 1) * sizeof(uint64_t),
                             cudaMemcpyHostToDevice, g->copy));
    CUDA_TRY(cudaEventRecord(s.copied, g->copy));
    if (int e = consume_staged(g, s, qual != nullptr, bytes, nr, quality_cutoff)) return e;
    // h_offsets of this leg must stay alive until its copy completed: the next
    // use of the leg waits on `consumed`, which is ordered after `copied`.
    g->stage_turn = (leg + 1) % (int)g->stage.size();
    leg = g->stage_turn;
    r0 = r1;
  }
  if (acc.mine)
    if (int e = flush_accumulated(g, g->compute)) return e;
  if (async) {
    if (pinned_stats9)
      if (int e = lasv_graph_stats_async(g, pinned_stats9, g->compute)) return e;
    CUDA_TRY(cudaStreamSynchronize(g->copy));  // the caller may reuse its buffers
    return LASV_OK;
  }
  if (int e = join_inserts(g, g->compute)) return e;  // pipelined inserts of the last chunks
  if (int e = fetch_counters(g, g->compute, &rs1, &c1, &ov)) return e;
  if (c1.table_full) return fail(LASV_ERR_TABLE_FULL, "%s", TABLE_FULL_MSG);
  if (batch_stats) {
    ReadStats d;
    d.n_reads = rs1.n_reads - rs0.n_reads;
    d.bad_reads = rs1.bad_reads - rs0.bad_reads;
    d.bases_read = rs1.bases_read - rs0.bases_read;
    d.bases_loaded = rs1.bases_loaded - rs0.bases_loaded;
    d.n_contigs = rs1.n_contigs - rs0.n_contigs;
    GraphCounters dc = c1;
    dc.kmers_inserted = c1.kmers_inserted - c0.kmers_inserted;
    fill_stats(batch_stats, d, dc);
  }
  return LASV_OK;
}
}  // namespace
}

int lasv_graph_load_reads_host(lasv_graph *g, const uint8_t *seq, const uint8_t *qual,
                               uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                               int quality_cutoff, lasv_load_stats *batch_stats) {
  return load_reads_host_impl(g, seq, qual, n_bytes, offsets, n_reads, quality_cutoff, batch_stats, false, nullptr);
}

int lasv_graph_load_reads_host_async(lasv_graph *g, const uint8_t *seq, const uint8_t *qual,
                                     uint64_t n_bytes, const uint64_t *offsets, uint64_t n_reads,
                                     int quality_cutoff, uint64_t *pinned_stats9) {
  return load_reads_host_impl(g, seq, qual, n_bytes, offsets, n_reads, quality_cutoff, nullptr, true, pinned_stats9);
}

extern "C++" {
namespace {
constexpr uint64_t RAW_CHUNK_BYTES = 32ull << 20;  // largest chunk of FASTQ text the device parse takes (= a loader batch)

int ensure_raw(lasv_graph *g, Staging &s) {
  if (s.d_raw) return LASV_OK;
  const uint32_t rec_cap = (uint32_t)g->stage_reads;
  CUDA_TRY(cudaMalloc(&s.d_raw, RAW_CHUNK_BYTES + 256));
  CUDA_TRY(cudaMalloc(&s.fq.nl, RAW_CHUNK_BYTES * sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&s.fq.n_nl, sizeof(uint32_t)));
  CUDA_TRY(cudaMalloc(&s.fq.lens, ((size_t)rec_cap + 1) * sizeof(uint64_t)));
  CUDA_TRY(cudaMalloc(&s.fq.result, sizeof(FqResult)));
  s.fq.temp_bytes = fq_temp_bytes((uint32_t)RAW_CHUNK_BYTES, rec_cap);
  CUDA_TRY(cudaMalloc(&s.fq.temp, s.fq.temp_bytes));
  CUDA_TRY(cudaMallocHost(&s.h_fq, sizeof(FqResult)));
  s.fq.nl_cap = (uint32_t)RAW_CHUNK_BYTES;
  s.fq.rec_cap = rec_cap;
  return LASV_OK;
}

// One chunk of plain four-line FASTQ text (whole records, ending with '\n', pinned for full speed): copied to the
// device as it is, taken apart there (fastq_parse.cu) and sent through the hot path.  Returns LASV_OK and the number
// of reads, or 1 if the text is not plain four-line FASTQ throughout -- nothing was loaded then and the caller's
// host parser (the reference's full grammar) takes the chunk.  The buffer may be reused when the call returns.
int load_raw_fastq_host(lasv_graph *g, const uint8_t *text, uint64_t n_bytes, int quality_cutoff, uint64_t *n_reads_out) {
  if (int e = use_hot(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (!n_bytes) { *n_reads_out = 0; return LASV_OK; }
  if (n_bytes > RAW_CHUNK_BYTES || text[n_bytes - 1] != '\n') return 1;
  if (int e = ensure_staging(g)) return e;
  Staging &s = g->stage[(size_t)g->stage_turn % g->stage.size()];
  if (int e = ensure_raw(g, s)) return e;
  CUDA_TRY(cudaEventSynchronize(s.consumed));  // the kernels that read this leg's streams are done
  CUDA_TRY(cudaMemcpyAsync(s.d_raw, text, n_bytes, cudaMemcpyHostToDevice, g->copy));
  const cudaError_t pe = fq_parse_chunk(s.d_raw, (uint32_t)n_bytes, s.fq, s.d_seq, s.d_qual, s.d_offsets, CHUNK_BYTES, g->copy);
  if (pe != cudaSuccess) return fail(LASV_ERR_CUDA, "device FASTQ parse: %s", cudaGetErrorString(pe));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  CUDA_TRY(cudaMemcpyAsync(s.h_fq, s.fq.result, sizeof(FqResult), cudaMemcpyDeviceToHost, g->copy));
  CUDA_TRY(cudaEventRecord(s.copied, g->copy));
  CUDA_TRY(cudaStreamSynchronize(g->copy));  // (a copy and four small kernels: the streams are still to be consumed)
  const FqResult r = *s.h_fq;
  if (r.bad) return 1;
  *n_reads_out = r.n_records;
  if (!r.n_records) return LASV_OK;
  // the chunks of a file accumulate like the chunks of a host-buffer call (the loaders set the accumulation up)
  if (int e = consume_staged(g, s, true, r.n_bytes, r.n_records, quality_cutoff)) return e;
  g->stage_turn = (g->stage_turn + 1) % (int)g->stage.size();
  return LASV_OK;
}
}  // namespace
}

int lasv_graph_stats(lasv_graph *g, lasv_load_stats *out, uint64_t *contig_hist, uint64_t hist_size,
                     void *cuda_stream) {
  if (int e = use(g)) return e;
  (void)cuda_stream;
  ReadStats rs;
  GraphCounters c;
  unsigned long long ov = 0;
  CUDA_TRY(cudaDeviceSynchronize());  // whatever stream the loads were queued on
  if (int e = fetch_counters(g, g->compute, &rs, &c, &ov)) return e;
  if (out) fill_stats(out, rs, c);
  if (contig_hist && hist_size) {
    std::vector<unsigned long long> h(HIST_BINS);
    CUDA_TRY(cudaMemcpy(h.data(), g->d_hist, HIST_BINS * sizeof(unsigned long long),
                        cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < hist_size; i++) contig_hist[i] = 0;
    for (uint32_t i = 0; i < HIST_BINS; i++) {
      if (!h[i]) continue;
      const uint64_t b = std::min<uint64_t>(i, hist_size - 1);  // file_reader.c:469
      contig_hist[b] += h[i];
    }
  }
  if (c.table_full) return fail(LASV_ERR_TABLE_FULL, "%s", TABLE_FULL_MSG);
  if (ov) return fail(LASV_ERR_CAPACITY, "record buffer / exchange bin overflow");
  return LASV_OK;
}

/* Asynchronous read-back of the load statistics as they stand when `cuda_stream` reaches this
 * point: 7 uint64 (lasv_load_stats order) + table_full + overflow into pinned host memory. */
int lasv_graph_stats_async(lasv_graph *g, uint64_t *pinned_out9, void *cuda_stream) {
  if (int e = use_hot(g)) return e;
  if (!pinned_out9) return fail(LASV_ERR_ARG, "NULL output");
  cudaStream_t st = stream_of(g, cuda_stream);
  // ReadStats = {n_reads, bad_reads, bases_read, bases_loaded, n_contigs}; GraphCounters starts with
  // {unique_kmers, kmers_inserted, table_full}
  CUDA_TRY(cudaMemcpyAsync(pinned_out9, g->d_stats, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(pinned_out9 + 5, g->d_ctr, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(pinned_out9 + 8, g->d_overflow, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  return LASV_OK;
}

int lasv_graph_rehash_histogram(lasv_graph *g, long long out16[16]) {
  if (int e = use(g)) return e;
  GraphCounters c;
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(&c, g->d_ctr, sizeof c, cudaMemcpyDeviceToHost));
  unsigned long long later = 0;
  for (int r = 1; r < 16; r++) {
    out16[r] = (long long)c.rehash_hist[r];
    later += c.rehash_hist[r];
  }
  out16[0] = (long long)(c.kmers_inserted - later);
  return LASV_OK;
}

// ---------------------------------------------------------- sharded pieces
int lasv_graph_route_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                  uint64_t n_bytes, int quality_cutoff, int n_dest, uint32_t *d_bins,
                                  uint64_t bin_capacity, unsigned long long *d_bin_counts,
                                  void *cuda_stream) {
  if (int e = use(g)) return e;
  if (n_dest < 1 || n_dest > 16 || (n_dest & (n_dest - 1)))
    return fail(LASV_ERR_ARG, "n_dest must be a power of two <= 16, got %d", n_dest);
  if (int e = check_streams(d_seq, d_qual)) return e;
  if (!n_bytes) return LASV_OK;
  BuildParams p = build_params(g, d_seq, d_qual, n_bytes, quality_cutoff);
  p.rec_out = d_bins;
  p.rec_count = d_bin_counts;
  p.rec_capacity = bin_capacity;
  p.n_dest = n_dest;
  p.owner_shift = g->L;
  launch_build(g->N, MODE_ROUTE, p, stream_of(g, cuda_stream));
  return check_launch();
}

int lasv_graph_extract_records_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                      uint64_t n_bytes, int quality_cutoff, uint32_t *d_records,
                                      uint64_t capacity, unsigned long long *d_count,
                                      void *cuda_stream) {
  if (int e = use(g)) return e;
  if (int e = check_streams(d_seq, d_qual)) return e;
  if (!n_bytes) return LASV_OK;
  BuildParams p = build_params(g, d_seq, d_qual, n_bytes, quality_cutoff);
  p.rec_out = d_records;
  p.rec_count = d_count;
  p.rec_capacity = capacity;
  launch_build(g->N, MODE_RECORDS, p, stream_of(g, cuda_stream));
  return check_launch();
}

int lasv_graph_insert_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                     void *cuda_stream) {
  if (int e = use(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (!n_records) return LASV_OK;
  launch_insert_records(g->N, d_records, n_records, g->view(), g->d_ctr, stream_of(g, cuda_stream));
  return check_launch();
}

// PROTOTYPE of the streamed insert (groups.cuh): records sorted by group, one CTA per group
// updates the group's lines in shared memory; records whose bucket is full follow through the ordinary path.
int lasv_graph_insert_grouped_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                             const uint64_t *d_group_offsets, uint32_t buckets_per_group,
                                             uint64_t *n_spilled, void *cuda_stream) {
  if (int e = use(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (!buckets_per_group || (buckets_per_group & (buckets_per_group - 1)) || buckets_per_group > g->n_buckets)
    return fail(LASV_ERR_ARG, "buckets_per_group must be a power of two <= the number of buckets");
  const TableView tv = g->view();
  if (group_bytes(tv, buckets_per_group) > insert_groups_smem_limit())
    return fail(LASV_ERR_ARG, "a group of %u buckets (%llu bytes) does not fit in shared memory", buckets_per_group,
                (unsigned long long)group_bytes(tv, buckets_per_group));
  if (n_records >= (1ull << 32)) return fail(LASV_ERR_ARG, "at most 2^32 - 1 records per call");
  if (n_spilled) *n_spilled = 0;
  if (!n_records) return LASV_OK;
  if (!d_records || !d_group_offsets) return fail(LASV_ERR_ARG, "NULL argument");
  cudaStream_t st = stream_of(g, cuda_stream);
  uint32_t *d_spill = nullptr;
  unsigned long long *d_n = nullptr, h_n = 0;
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_spill, n_records * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d_n, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), st));
    launch_insert_groups(g->N, d_records, d_group_offsets, g->n_buckets / buckets_per_group, buckets_per_group, tv, g->d_ctr,
                         d_spill, d_n, n_records, st);
    if (int e = check_launch()) return e;
    CUDA_TRY(cudaMemcpyAsync(&h_n, d_n, sizeof h_n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    launch_insert_spilled(g->N, d_records, d_spill, h_n, tv, g->d_ctr, st);  // after ALL groups are back in the table
    if (h_n)
      if (int e = check_launch()) return e;
    CUDA_TRY(cudaStreamSynchronize(st));
    return LASV_OK;
  }();
  cudaFree(d_spill);
  cudaFree(d_n);
  if (n_spilled) *n_spilled = h_n;
  return rc;
}

// The grouping the prototype needs, in the library: counting sort of the records by group.
int lasv_graph_group_records_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records, uint32_t buckets_per_group,
                                    uint32_t *d_sorted, uint64_t *d_group_offsets, void *cuda_stream) {
  if (int e = use(g)) return e;
  if (!buckets_per_group || (buckets_per_group & (buckets_per_group - 1)) || buckets_per_group > g->n_buckets)
    return fail(LASV_ERR_ARG, "buckets_per_group must be a power of two <= the number of buckets");
  if (!d_group_offsets || (n_records && (!d_records || !d_sorted))) return fail(LASV_ERR_ARG, "NULL argument");
  cudaStream_t st = stream_of(g, cuda_stream);
  const uint64_t n_groups = g->n_buckets / buckets_per_group;
  uint32_t *d_counts = nullptr;
  void *d_tmp = nullptr;
  const size_t tmp_bytes = exclusive_scan_tmp_bytes(n_groups);
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_counts, n_groups * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    launch_group_records(g->N, d_records, n_records, buckets_per_group, g->view(), d_counts, d_group_offsets, d_tmp, tmp_bytes,
                         d_sorted, st);
    if (int e = check_launch()) return e;
    CUDA_TRY(cudaStreamSynchronize(st));
    return LASV_OK;
  }();
  cudaFree(d_counts);
  cudaFree(d_tmp);
  return rc;
}

int lasv_graph_bin_geometry(lasv_graph *g, int n_owners, uint64_t n_bytes, uint64_t n_reads,
                            uint32_t *bins_per_owner, uint32_t *stripe_capacity_out,
                            uint32_t *spill_capacity_out, uint32_t *record_bytes,
                            uint32_t *count_stride_bytes) {
  if (!g) return fail(LASV_ERR_ARG, "NULL graph");
  if (n_owners < 1 || n_owners > 16 || (n_owners & (n_owners - 1)))
    return fail(LASV_ERR_ARG, "n_owners must be a power of two <= 16, got %d", n_owners);
  const uint32_t per = region_count(g);
  const uint64_t cap = stripe_capacity(kmer_bound(g, n_bytes, n_reads), per * (uint32_t)n_owners);
  if (cap > 0xFFFFFFF0ull) return fail(LASV_ERR_ARG, "batch too large for the bin stripes");
  if (bins_per_owner) *bins_per_owner = per;
  if (stripe_capacity_out) *stripe_capacity_out = (uint32_t)cap;
  // spill area per owner: 1/16 of its stripes (a single k-mer may be that share of a batch), >= 4096 records
  if (spill_capacity_out) {
    const uint64_t sp = std::max<uint64_t>(4096, cap * per / 16);
    *spill_capacity_out = n_owners == 1 ? 0u : (uint32_t)std::min<uint64_t>(0xFFFFFF80ull, (sp + STRIPE_RUN - 1) / STRIPE_RUN * STRIPE_RUN);
  }
  if (record_bytes) *record_bytes = 8u * bin_record_words(g->N, g->k);
  if (count_stride_bytes) *count_stride_bytes = COUNT_STRIDE * (uint32_t)sizeof(unsigned int);
  return LASV_OK;
}

extern "C++" {
namespace {
int bin_reads_impl(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                   uint64_t n_bytes, int quality_cutoff, int n_owners,
                   uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                   uint32_t spill_capacity_records, void *d_records, void *d_counts,
                   void *cuda_stream, bool append, void *const *owner_records) {
  if (int e = use(g)) return e;
  if (n_owners < 1 || n_owners > 16 || (n_owners & (n_owners - 1)) || !bins_per_owner ||
      (bins_per_owner & (bins_per_owner - 1)) || bins_per_owner > g->n_buckets || !stripe_capacity_records)
    return fail(LASV_ERR_ARG, "bad bin geometry: %d owners x %u bins x %u records", n_owners, bins_per_owner,
                stripe_capacity_records);
  if ((!d_records && !owner_records) || !d_counts) return fail(LASV_ERR_ARG, "NULL bin buffers");
  if (owner_records)
    for (int o = 0; o < n_owners; o++)
      if (!owner_records[o]) return fail(LASV_ERR_ARG, "NULL receive block of owner %d", o);
  if (stripe_capacity_records % STRIPE_RUN || spill_capacity_records % STRIPE_RUN)
    return fail(LASV_ERR_ARG, "stripe and spill capacities must be multiples of %u records", STRIPE_RUN);
  if (int e = check_streams(d_seq, d_qual)) return e;
  cudaStream_t st = stream_of(g, cuda_stream);
  const uint32_t n_bins = bins_per_owner * (uint32_t)n_owners;
  if (!append)
    CUDA_TRY(cudaMemsetAsync(d_counts, 0, (size_t)(n_bins + n_owners) * COUNT_STRIDE * sizeof(unsigned int), st));
  if (!n_bytes) return LASV_OK;
  BuildParams p = build_params(g, d_seq, d_qual, n_bytes, quality_cutoff);
  p.bins.recs = (uint64_t *)d_records;
  if (owner_records) {  // fused producer -> exchange: every owner's block at its own (peer) address
    p.bins.direct = 1;
    for (int o = 0; o < n_owners; o++) p.bins.owner_recs[o] = (uint64_t *)owner_records[o];
  }
  p.bins.count = (unsigned int *)d_counts;
  p.bins.cap = stripe_capacity_records;
  p.bins.n_bins = n_bins;
  // this graph is one of n_owners shards: the global bucket index has log2(n_owners) more bits
  const int Lg = g->L + log2_u32((uint32_t)n_owners);
  p.bins.bin_mask = Lg >= 32 ? 0xFFFFFFFFu : ((1u << Lg) - 1u);
  p.bins.bin_shift = (uint32_t)(Lg - log2_u32(n_bins));
  p.bins.spill_ok = (n_owners == 1 && !spill_capacity_records) ? 1u : 0u;
  p.bins.spill_cap = spill_capacity_records;
  p.bins.src_shift = 0;
  p.bins.bins_per_src = n_bins;
  p.bins.block_shift = (uint32_t)log2_u32(bins_per_owner);
  p.bins.rec_words = bin_record_words(g->N, g->k);
  p.ctas_per_sm = g->pipeline ? g->knobs.pipe_ctas : 0;
  Span sp(g, LASV_T_PARTITION, st);
  launch_build(g->N, MODE_BIN, p, st);
  if (int e = check_launch()) return e;
  sp.end();
  return LASV_OK;
}
}  // namespace
}

int lasv_graph_bin_reads_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                uint64_t n_bytes, int quality_cutoff, int n_owners,
                                uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                uint32_t spill_capacity_records, void *d_records, void *d_counts,
                                void *cuda_stream) {
  return bin_reads_impl(g, d_seq, d_qual, n_bytes, quality_cutoff, n_owners, bins_per_owner, stripe_capacity_records,
                        spill_capacity_records, d_records, d_counts, cuda_stream, false);
}

int lasv_graph_bin_reads_append_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                       uint64_t n_bytes, int quality_cutoff, int n_owners,
                                       uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                       uint32_t spill_capacity_records, void *d_records, void *d_counts,
                                       void *cuda_stream) {
  return bin_reads_impl(g, d_seq, d_qual, n_bytes, quality_cutoff, n_owners, bins_per_owner, stripe_capacity_records,
                        spill_capacity_records, d_records, d_counts, cuda_stream, true);
}

int lasv_graph_bin_reads_peers_device(lasv_graph *g, const uint8_t *d_seq, const uint8_t *d_qual,
                                      uint64_t n_bytes, int quality_cutoff, int n_owners,
                                      uint32_t bins_per_owner, uint32_t stripe_capacity_records,
                                      uint32_t spill_capacity_records, void *const *owner_records, void *d_counts,
                                      void *cuda_stream, int append) {
  if (!owner_records) return fail(LASV_ERR_ARG, "NULL owner_records");
  return bin_reads_impl(g, d_seq, d_qual, n_bytes, quality_cutoff, n_owners, bins_per_owner, stripe_capacity_records,
                        spill_capacity_records, nullptr, d_counts, cuda_stream, append != 0, owner_records);
}

int lasv_graph_insert_bins_device(lasv_graph *g, const void *d_records, const void *d_counts, int n_src,
                                  uint32_t bins_per_src, uint32_t stripe_capacity_records,
                                  uint32_t spill_capacity_records, void *cuda_stream) {
  if (int e = use(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  if (n_src < 1 || n_src > 16 || (n_src & (n_src - 1)) || !bins_per_src || (bins_per_src & (bins_per_src - 1)))
    return fail(LASV_ERR_ARG, "bad bin geometry: %d senders x %u bins", n_src, bins_per_src);
  if (!d_records || !d_counts) return fail(LASV_ERR_ARG, "NULL bin buffers");
  if (stripe_capacity_records % STRIPE_RUN || spill_capacity_records % STRIPE_RUN)
    return fail(LASV_ERR_ARG, "stripe and spill capacities must be multiples of %u records", STRIPE_RUN);
  BinView b{};
  b.spill_cap = spill_capacity_records;
  b.recs = (uint64_t *)d_records;
  b.count = (unsigned int *)d_counts;
  b.cap = stripe_capacity_records;
  b.n_bins = bins_per_src * (uint32_t)n_src;
  b.src_shift = (uint32_t)log2_u32((uint32_t)n_src);
  b.bins_per_src = bins_per_src;
  b.block_shift = (uint32_t)log2_u32(bins_per_src);
  b.rec_words = bin_record_words(g->N, g->k);
  // (the stripes hold about their capacity less the 8-sigma margin; the counts themselves are on the device)
  const uint64_t about = (uint64_t)((double)b.n_bins * (double)b.cap / 1.03);
  // Streamed only when no partition kernel of this graph can write into the table meanwhile: with several
  // owners an overfull stripe spills into the spill area, never into the table.
  const bool may_stream = !g->pipeline || n_src > 1 || spill_capacity_records != 0;
  return insert_stripes(g, b, about, g->pipeline ? 2 : 0, may_stream, false, stream_of(g, cuda_stream));
}

// ------------------------------------------------------------------ queries
int lasv_graph_find(lasv_graph *g, const uint64_t *keys, uint64_t n, uint32_t *covg, uint8_t *edges,
                    uint8_t *found) {
  if (int e = use(g)) return e;
  if (!n) return LASV_OK;
  CUDA_TRY(cudaDeviceSynchronize());
  uint64_t *d_keys = nullptr;
  uint32_t *d_covg = nullptr;
  uint8_t *d_edges = nullptr, *d_found = nullptr;
  int rc = LASV_OK;
  auto body = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_keys, n * g->N * 8));
    CUDA_TRY(cudaMalloc(&d_covg, n * 4));
    CUDA_TRY(cudaMalloc(&d_edges, n));
    CUDA_TRY(cudaMalloc(&d_found, n));
    CUDA_TRY(cudaMemcpyAsync(d_keys, keys, n * g->N * 8, cudaMemcpyHostToDevice, g->compute));
    launch_find(g->N, g->view(), d_keys, n, d_covg, d_edges, d_found, g->compute);
    if (int e = check_launch()) return e;
    CUDA_TRY(cudaMemcpyAsync(covg, d_covg, n * 4, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaMemcpyAsync(edges, d_edges, n, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaMemcpyAsync(found, d_found, n, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    return LASV_OK;
  };
  rc = body();
  cudaFree(d_keys); cudaFree(d_covg); cudaFree(d_edges); cudaFree(d_found);
  return rc;
}

int lasv_graph_summary(lasv_graph *g, lasv_table_summary *out) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  TableSummary *d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(TableSummary)));
  int rc = [&]() -> int {
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(TableSummary), g->compute));
    launch_table_summary(g->N, g->view(), d, g->compute);
    if (int e = check_launch()) return e;
    TableSummary h;
    CUDA_TRY(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    out->n_nodes = h.n_nodes;
    out->sum_covg = h.sum_covg;
    out->sum_edge_bits = h.sum_edge_bits;
    out->fingerprint = h.fingerprint;
    return LASV_OK;
  }();
  cudaFree(d);
  return rc;
}

static int finalize_impl(lasv_graph *g, int16_t *d_next) {
  launch_finalize(g->N, g->view(), d_next, g->compute);
  if (int e = check_launch()) return e;
  CUDA_TRY(cudaStreamSynchronize(g->compute));
  g->finalized = true;
  return LASV_OK;
}

int lasv_graph_finalize(lasv_graph *g) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  return finalize_impl(g, nullptr);
}

extern "C++" {
namespace {
// finalize + copy to the host.  `moved` (shards of a multi-GPU graph): the Elements that do not sit in their
// first-choice bucket are taken out first and returned there instead (extract_misplaced_kernel).
int export_table_impl(lasv_graph *g, void *elements, short *next_element, std::vector<uint8_t> *moved) {
  if (int e = use(g)) return e;
  if (!elements) return fail(LASV_ERR_ARG, "elements is NULL");
  CUDA_TRY(cudaDeviceSynchronize());
  int16_t *d_next = nullptr;
  uint8_t *d_moved = nullptr;
  unsigned long long *d_n_moved = nullptr;
  CUDA_TRY(cudaMalloc(&d_next, g->n_buckets * sizeof(int16_t)));
  int rc = [&]() -> int {
    if (int e = finalize_impl(g, d_next)) return e;
    if (moved) {
      // room for 1/64 of the slots (a W=250 bucket of an 80 %-full table overflows about once in 10^4), less if the
      // device is short of memory; more than that is an error, reported below
      uint64_t cap = std::max<uint64_t>(4096, g->n_slots / 64);
      size_t free_b = 0, total_b = 0;
      CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
      cap = std::max<uint64_t>(1, std::min<uint64_t>(cap, (uint64_t)(free_b / 2) / g->slot_bytes));
      CUDA_TRY(cudaMalloc(&d_n_moved, sizeof(unsigned long long)));
      CUDA_TRY(cudaMalloc(&d_moved, cap * g->slot_bytes));
      CUDA_TRY(cudaMemsetAsync(d_n_moved, 0, sizeof(unsigned long long), g->compute));
      launch_extract_misplaced(g->N, g->view(), d_next, d_moved, d_n_moved, cap, g->compute);
      if (int e = check_launch()) return e;
      unsigned long long n_moved = 0;
      CUDA_TRY(cudaMemcpyAsync(&n_moved, d_n_moved, sizeof n_moved, cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaStreamSynchronize(g->compute));
      if (n_moved > cap)
        return fail(LASV_ERR_CAPACITY, "%llu keys of a shard sit outside their first-choice bucket, more than the %llu the "
                    "export can move (the table is too full to be sharded)", n_moved, (unsigned long long)cap);
      const size_t old = moved->size();
      moved->resize(old + (size_t)n_moved * g->slot_bytes);
      if (n_moved)
        CUDA_TRY(cudaMemcpy(moved->data() + old, d_moved, (size_t)n_moved * g->slot_bytes, cudaMemcpyDeviceToHost));
    }
    // bucket b is W consecutive Elements at the start of its NL lines: a pitched copy
    // (DMA engine) packs the buckets into the reference's contiguous Element[]
    const size_t row = (size_t)g->W * g->slot_bytes, pitch = (size_t)(g->table_bytes / g->n_buckets);
    const uint64_t rows_per_copy = 1ull << 20;
    for (uint64_t b0 = 0; b0 < g->n_buckets; b0 += rows_per_copy) {
      const uint64_t nb = std::min(rows_per_copy, g->n_buckets - b0);
      CUDA_TRY(cudaMemcpy2D((uint8_t *)elements + b0 * row, row, g->d_lines + b0 * pitch, pitch, row,
                            (size_t)nb, cudaMemcpyDeviceToHost));
    }
    if (next_element)
      CUDA_TRY(cudaMemcpy(next_element, d_next, g->n_buckets * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return LASV_OK;
  }();
  cudaFree(d_next);
  cudaFree(d_moved);
  cudaFree(d_n_moved);
  return rc;
}
}  // namespace
}

int lasv_graph_export_table(lasv_graph *g, void *elements, short *next_element) {
  return export_table_impl(g, elements, next_element, nullptr);
}

int lasv_graph_import_table(lasv_graph *g, const void *elements) {
  if (int e = use(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  CUDA_TRY(cudaDeviceSynchronize());
  // stream the host table through a bounded device buffer, whole slots at a time
  const uint64_t slots_per_chunk = std::max<uint64_t>(1, (256ull << 20) / g->slot_bytes);
  uint8_t *d_buf = nullptr;
  CUDA_TRY(cudaMalloc(&d_buf, slots_per_chunk * g->slot_bytes));
  int rc = [&]() -> int {
    for (uint64_t s0 = 0; s0 < g->n_slots; s0 += slots_per_chunk) {
      const uint64_t ns = std::min(slots_per_chunk, g->n_slots - s0);
      CUDA_TRY(cudaMemcpyAsync(d_buf, (const uint8_t *)elements + s0 * g->slot_bytes,
                               ns * g->slot_bytes, cudaMemcpyHostToDevice, g->compute));
      launch_ingest_elements(g->N, d_buf, ns, g->view(), g->d_ctr, g->compute);
      if (int e = check_launch()) return e;
      CUDA_TRY(cudaStreamSynchronize(g->compute));
    }
    return LASV_OK;
  }();
  cudaFree(d_buf);
  return rc;
}

// Path records of all starts: one walk per start (launch_sn_paths), or -- LASV_B200_SN_RANKING=1, for graphs with long
// paths; opt-in until it has run on a device -- by list ranking (ranking.cuh).  Same records either way.
static bool ranking_wanted() {
  const char *e = getenv("LASV_B200_SN_RANKING");
  return e && atoi(e) != 0;
}

static int enumerate_paths(lasv_graph *g, const TableView &tv, const uint64_t *d_starts, uint64_t n_starts,
                           uint64_t n_interior, uint8_t *d_covered, SnPath *d_recs, SnCounters *d_c,
                           RkState *keep = nullptr) {
  if (ranking_wanted()) {
    const cudaError_t err = sn_paths_ranked(g->N, tv, g->k, d_starts, n_starts, n_interior, d_covered, d_recs, n_starts,
                                            d_c, g->compute, keep);
    if (err != cudaSuccess) return fail(LASV_ERR_CUDA, "path enumeration by list ranking failed: %s", cudaGetErrorString(err));
    return LASV_OK;
  }
  launch_sn_paths(g->N, tv, g->k, d_starts, n_starts, d_covered, d_recs, n_starts, d_c, g->compute);
  return n_starts ? check_launch() : LASV_OK;
}

// ------------------------------------------------------------- supernodes
// `--output_supernodes`: enumerate all paths on the device, replay the reference's
// sequential traversal over the path records on the host, write the sequences on
// the device (supernodes.cuh has the argument and the reference lines).
int lasv_graph_write_supernodes(lasv_graph *g, const char *fasta_path, int max_length,
                                lasv_supernode_stats *stats) {
  if (int e = use(g)) return e;
  if (!fasta_path) return fail(LASV_ERR_ARG, "fasta_path is NULL");
  CUDA_TRY(cudaDeviceSynchronize());
  const TableView tv = g->view();
  SnCounters *d_c = nullptr, h_c;
  uint64_t *d_starts = nullptr, *d_list = nullptr, *d_keys = nullptr, *d_keys_alt = nullptr;
  uint32_t *d_vals = nullptr, *d_vals_alt = nullptr;
  unsigned long long *d_n = nullptr;
  uint8_t *d_covered = nullptr;
  SnPath *d_recs = nullptr, *d_cyc = nullptr;
  SnEvent *d_events = nullptr;
  SnPrint *d_prints = nullptr;
  char *d_image = nullptr, *h_stage[2] = {nullptr, nullptr};
  void *d_tmp = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr};
  FILE *fp = nullptr;
  constexpr uint64_t STAGE_BYTES = 16ull << 20;
  std::vector<SnPrint> prints;
  std::vector<uint64_t> end_keys;
  RkState ranked;  // LASV_B200_SN_RANKING=1: the ranked elements, kept for the per-node writer
  RkPrintKey *d_print_keys = nullptr;
  uint64_t *d_end_keys = nullptr;
  SnReplayStats rs;
  uint64_t over = 0;
  // LASV_B200_SN_TRACE=1: stage times on stderr (tests/perf/supernodes_speed.py)
  const bool trace = getenv("LASV_B200_SN_TRACE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char *what) {
    if (!trace) return;
    cudaStreamSynchronize(g->compute);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[supernodes] %-36s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  auto counters = [&]() -> int {
    CUDA_TRY(cudaMemcpyAsync(&h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    return LASV_OK;
  };
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_c, sizeof(SnCounters)));
    CUDA_TRY(cudaMalloc(&d_n, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(d_c, 0, sizeof(SnCounters), g->compute));
    CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), g->compute));
    launch_sn_count(g->N, tv, d_c, g->compute);
    if (int e = check_launch()) return e;
    if (int e = counters()) return e;
    stage("count nodes and starts");
    const uint64_t n_starts = h_c.n_starts;
    // every path has two starts (twins) and keeps one record; self-twins keep theirs
    CUDA_TRY(cudaMalloc(&d_starts, std::max<uint64_t>(1, n_starts) * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(&d_recs, std::max<uint64_t>(1, n_starts) * sizeof(SnPath)));
    CUDA_TRY(cudaMalloc(&d_covered, g->n_slots));
    CUDA_TRY(cudaMemsetAsync(d_covered, 0, g->n_slots, g->compute));
    launch_sn_starts(g->N, tv, d_starts, n_starts, d_n, d_c, g->compute);
    if (int e = check_launch()) return e;
    stage("list starts");
    if (int e = enumerate_paths(g, tv, d_starts, n_starts, h_c.n_interior, d_covered, d_recs, d_c, &ranked)) return e;
    stage("walk paths");
    // interior nodes no path crossed lie on cycles: count, list, walk
    launch_sn_uncovered(g->N, tv, d_covered, nullptr, 0, d_c, g->compute);
    if (int e = check_launch()) return e;
    if (int e = counters()) return e;
    const uint64_t n_unc = h_c.n_uncovered;
    stage("find uncovered interior nodes");
    if (n_unc) {
      CUDA_TRY(cudaMalloc(&d_list, n_unc * sizeof(uint64_t)));
      CUDA_TRY(cudaMalloc(&d_cyc, n_unc * sizeof(SnPath)));
      CUDA_TRY(cudaMemsetAsync(&d_c->n_uncovered, 0, sizeof(unsigned long long), g->compute));
      launch_sn_uncovered(g->N, tv, d_covered, d_list, n_unc, d_c, g->compute);
      if (int e = check_launch()) return e;
      launch_sn_cycles(g->N, tv, g->k, d_list, n_unc, d_cyc, n_unc, d_c, g->compute);
      if (int e = check_launch()) return e;
      if (int e = counters()) return e;
      stage("walk cycles");
    }
    if (h_c.n_overflow) return fail(LASV_ERR_STATE, "supernode record buffers overflowed (internal sizing error)");
    if (h_c.n_dangling)
      return fail(LASV_ERR_STATE, "%llu walks met an edge to a node that is not in the graph (the reference "
                  "die()s: dB_graph_population.c:842-850)", (unsigned long long)h_c.n_dangling);
    cudaFree(d_starts); d_starts = nullptr;
    cudaFree(d_covered); d_covered = nullptr;
    cudaFree(d_list); d_list = nullptr;

    // replay events: (slot, path << 2 | role) pairs, sorted by slot on the device, expanded to
    // SnEvent records in that order (the path records arrive in atomic-append order; sorting the
    // events by slot makes the output independent of it)
    const uint64_t n_rec = h_c.n_paths + h_c.n_cycles;
    if (n_rec >= (1ull << 30)) return fail(LASV_ERR_CAPACITY, "more than 2^30 paths");
    uint64_t n_ev = 0;
    if (n_rec) {
      CUDA_TRY(cudaMalloc(&d_keys, 3 * n_rec * sizeof(uint64_t)));
      CUDA_TRY(cudaMalloc(&d_keys_alt, 3 * n_rec * sizeof(uint64_t)));
      CUDA_TRY(cudaMalloc(&d_vals, 3 * n_rec * sizeof(uint32_t)));
      CUDA_TRY(cudaMalloc(&d_vals_alt, 3 * n_rec * sizeof(uint32_t)));
      CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), g->compute));
      launch_sn_events(d_recs, h_c.n_paths, 0u, d_keys, d_vals, d_n, g->compute);
      if (h_c.n_paths)
        if (int e = check_launch()) return e;
      launch_sn_events(d_cyc, h_c.n_cycles, (uint32_t)h_c.n_paths, d_keys, d_vals, d_n, g->compute);
      if (h_c.n_cycles)
        if (int e = check_launch()) return e;
      unsigned long long nev = 0;
      CUDA_TRY(cudaMemcpyAsync(&nev, d_n, sizeof nev, cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaStreamSynchronize(g->compute));
      n_ev = nev;
    }
    if (n_ev) {
      int key_bits = 1;
      while (key_bits < 64 && (g->n_slots >> key_bits)) key_bits++;
      const size_t tmp_bytes = sn_sort_tmp_bytes(n_ev);
      CUDA_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
      CUDA_TRY(cudaMalloc(&d_events, n_ev * sizeof(SnEvent)));
      launch_sn_sort_events(d_keys, d_keys_alt, d_vals, d_vals_alt, n_ev, key_bits, d_tmp, tmp_bytes, d_recs,
                            h_c.n_paths, d_cyc, d_events, g->compute);
      if (int e = check_launch()) return e;
      stage("events: build, sort, expand");
    }
    cudaFree(d_recs); d_recs = nullptr;
    cudaFree(d_cyc); d_cyc = nullptr;
    cudaFree(d_keys); d_keys = nullptr;
    cudaFree(d_keys_alt); d_keys_alt = nullptr;
    cudaFree(d_vals); d_vals = nullptr;
    cudaFree(d_vals_alt); d_vals_alt = nullptr;
    cudaFree(d_tmp); d_tmp = nullptr;

    // the replay consumes the events in order, so they stream through the two pinned staging
    // buffers: the copy of piece i+1 runs under the replay of piece i
    for (int b = 0; b < 2; b++) {
      CUDA_TRY(cudaMallocHost(&h_stage[b], STAGE_BYTES));
      CUDA_TRY(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
    }
    SnReplayer replayer(n_rec, g->n_slots, g->k);
    replayer.want_end_keys = ranked.el != nullptr;
    {
      const uint64_t per = STAGE_BYTES / sizeof(SnEvent);
      const uint64_t n_pieces = (n_ev + per - 1) / per;
      auto count_of = [&](uint64_t c) { return std::min(per, n_ev - c * per); };
      for (uint64_t c = 0; c <= n_pieces; c++) {
        if (c < n_pieces) {
          CUDA_TRY(cudaMemcpyAsync(h_stage[c & 1], d_events + c * per, count_of(c) * sizeof(SnEvent),
                                   cudaMemcpyDeviceToHost, g->compute));
          CUDA_TRY(cudaEventRecord(copied[c & 1], g->compute));
        }
        if (c > 0) {
          CUDA_TRY(cudaEventSynchronize(copied[(c - 1) & 1]));
          replayer.feed(reinterpret_cast<const SnEvent *>(h_stage[(c - 1) & 1]), count_of(c - 1));
        }
      }
    }
    cudaFree(d_events); d_events = nullptr;
    rs = replayer.finish();
    prints.swap(replayer.prints);
    end_keys.swap(replayer.end_keys);
    for (const SnPrint &p : prints) over += max_length > 0 && p.len > (uint32_t)max_length;
    if (over && g->sn_strict) {
      if (stats) stats->over_max_length = over;
      return fail(LASV_ERR_LIMIT, "%llu supernodes have more than %d edges: the reference cuts its walks there (--max_var_len)",
                  (unsigned long long)over, max_length);
    }
    const uint64_t image_bytes = sn_layout_fasta(prints, g->k);
    stage("replay the traversal (host)");

    fp = fopen(fasta_path, "w");
    if (!fp) return fail(LASV_ERR_IO, "cannot open %s for writing: %s", fasta_path, strerror(errno));
    if (!prints.empty()) {
      CUDA_TRY(cudaMalloc(&d_prints, prints.size() * sizeof(SnPrint)));
      CUDA_TRY(cudaMalloc(&d_image, image_bytes));
      CUDA_TRY(cudaMemcpyAsync(d_prints, prints.data(), prints.size() * sizeof(SnPrint), cudaMemcpyHostToDevice,
                               g->compute));
      // the device writes the whole file image: header lines and sequence lines
      if (ranked.el) {  // one thread per record frame and one per interior node instead of one walk per record
        std::vector<RkPrintKey> keys;
        keys.reserve(prints.size());
        for (size_t i = 0; i < prints.size(); i++)
          if (end_keys[i] != SN_NONE) keys.push_back({end_keys[i], (uint64_t)i});
        std::sort(keys.begin(), keys.end(), [](const RkPrintKey &a, const RkPrintKey &b) { return a.key < b.key; });
        CUDA_TRY(cudaMalloc(&d_print_keys, std::max<size_t>(1, keys.size()) * sizeof(RkPrintKey)));
        CUDA_TRY(cudaMalloc(&d_end_keys, prints.size() * sizeof(uint64_t)));
        CUDA_TRY(cudaMemcpyAsync(d_print_keys, keys.data(), keys.size() * sizeof(RkPrintKey), cudaMemcpyHostToDevice, g->compute));
        CUDA_TRY(cudaMemcpyAsync(d_end_keys, end_keys.data(), prints.size() * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                 g->compute));
        rk_write_supernodes(g->N, tv, g->k, ranked, d_print_keys, keys.size(), d_prints, d_end_keys, prints.size(), d_image, d_c,
                            g->compute);
        CUDA_TRY(cudaStreamSynchronize(g->compute));  // keys / end_keys go out of scope
      } else {
        launch_sn_write(g->N, tv, g->k, d_prints, prints.size(), d_image, d_c, g->compute);
      }
      if (int e = check_launch()) return e;
      if (int e = counters()) return e;
      if (h_c.n_dangling) return fail(LASV_ERR_STATE, "supernode sequence walk left the graph");
      stage("write the FASTA image (device)");
      // image -> file through two pinned staging buffers: the copy of chunk i+1 runs under the fwrite of chunk i
      const uint64_t chunk = STAGE_BYTES;
      const uint64_t n_chunks = (image_bytes + chunk - 1) / chunk;
      auto bytes_of = [&](uint64_t c) { return std::min(chunk, image_bytes - c * chunk); };
      for (uint64_t c = 0; c <= n_chunks; c++) {
        if (c < n_chunks) {
          CUDA_TRY(cudaMemcpyAsync(h_stage[c & 1], d_image + c * chunk, bytes_of(c), cudaMemcpyDeviceToHost, g->compute));
          CUDA_TRY(cudaEventRecord(copied[c & 1], g->compute));
        }
        if (c > 0) {
          CUDA_TRY(cudaEventSynchronize(copied[(c - 1) & 1]));
          if (fwrite(h_stage[(c - 1) & 1], 1, bytes_of(c - 1), fp) != bytes_of(c - 1))
            return fail(LASV_ERR_IO, "short write to %s", fasta_path);
        }
      }
      stage("copy back + write the file");
    }
    return LASV_OK;
  }();
  cudaFree(d_c); cudaFree(d_n); cudaFree(d_starts); cudaFree(d_list); cudaFree(d_covered);
  cudaFree(d_recs); cudaFree(d_cyc); cudaFree(d_keys); cudaFree(d_keys_alt); cudaFree(d_vals); cudaFree(d_vals_alt);
  cudaFree(d_tmp); cudaFree(d_events); cudaFree(d_prints); cudaFree(d_image); cudaFree(d_print_keys); cudaFree(d_end_keys);
  rk_release(ranked);
  for (int b = 0; b < 2; b++) {
    if (h_stage[b]) cudaFreeHost(h_stage[b]);
    if (copied[b]) cudaEventDestroy(copied[b]);
  }
  if (fp && fclose(fp) != 0 && rc == LASV_OK) rc = fail(LASV_ERR_IO, "cannot close %s: %s", fasta_path, strerror(errno));
  if (rc != LASV_OK) return rc;
  if (stats) {
    stats->n_supernodes = rs.printed;
    stats->n_nodes = h_c.n_nodes;
    stats->n_interior = h_c.n_interior;
    stats->n_paths = h_c.n_paths;
    stats->n_cycles = h_c.n_cycles;
    stats->n_not_printed = rs.never_printed;
    stats->longest_edges = rs.longest;
    stats->total_bases = rs.total_bases;
    stats->over_max_length = over;
  }
  return LASV_OK;
}

// ------------------------------------------------------------ .ctx records
// .ctx records in table order, streamed through a bounded device buffer: the table is walked in ranges
// of buckets whose records fit DUMP_CHUNK_BYTES even if every slot is occupied (a 100 GB table holds
// ~65 GB of records -- they cannot be materialised next to it); each range is counted, scanned, packed
// by dump_ctx_kernel and handed to `sink` in host memory.
constexpr uint64_t DUMP_CHUNK_BYTES = 256ull << 20;

using CtxSink = std::function<int(const uint8_t *records, uint64_t n, uint64_t n_before)>;
static int stream_ctx_records_out(lasv_graph *g, uint64_t *n_out, const CtxSink &sink) {
  const size_t rec = 8 * (size_t)g->N + 5;
  uint64_t chunk_bytes = DUMP_CHUNK_BYTES;
  if (const char *e = getenv("LASV_B200_DUMP_CHUNK_BYTES")) {  // tests: many ranges on small tables
    const long long v = atoll(e);
    if (v >= 64 && (uint64_t)v <= DUMP_CHUNK_BYTES) chunk_bytes = (uint64_t)v;
  }
  const uint64_t per_range = std::max<uint64_t>(1, chunk_bytes / ((uint64_t)g->W * rec));
  const uint64_t range = std::min<uint64_t>(g->n_buckets, per_range);
  uint32_t *d_counts = nullptr;
  uint64_t *d_offsets = nullptr;
  void *d_tmp = nullptr;
  uint8_t *d_out = nullptr, *h_out = nullptr;
  const size_t tmp_bytes = exclusive_scan_tmp_bytes(range);
  const size_t out_bytes = std::max<size_t>(16, range * (size_t)g->W * rec);
  uint64_t total = 0;
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_counts, range * 4));
    CUDA_TRY(cudaMalloc(&d_offsets, range * 8));
    CUDA_TRY(cudaMalloc(&d_tmp, tmp_bytes));
    CUDA_TRY(cudaMalloc(&d_out, out_bytes));
    CUDA_TRY(cudaMallocHost(&h_out, out_bytes));
    for (uint64_t b0 = 0; b0 < g->n_buckets; b0 += range) {
      const uint64_t nb = std::min(range, g->n_buckets - b0);
      launch_bucket_counts(g->N, g->view(), b0, nb, d_counts, g->compute);
      if (int e = check_launch()) return e;
      launch_exclusive_scan_u32_to_u64(d_counts, d_offsets, nb, d_tmp, tmp_bytes, g->compute);
      if (int e = check_launch()) return e;
      uint64_t last_off = 0;
      uint32_t last_cnt = 0;
      CUDA_TRY(cudaMemcpyAsync(&last_off, d_offsets + nb - 1, 8, cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaMemcpyAsync(&last_cnt, d_counts + nb - 1, 4, cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaStreamSynchronize(g->compute));
      const uint64_t n = last_off + last_cnt;
      if (!n) continue;
      launch_dump_ctx(g->N, g->view(), b0, nb, d_offsets, d_out, g->compute);
      if (int e = check_launch()) return e;
      CUDA_TRY(cudaMemcpyAsync(h_out, d_out, n * rec, cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaStreamSynchronize(g->compute));
      if (int e = sink(h_out, n, total)) return e;
      total += n;
    }
    return LASV_OK;
  }();
  cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_tmp); cudaFree(d_out);
  if (h_out) cudaFreeHost(h_out);
  if (n_out) *n_out = total;
  return rc;
}

long long lasv_graph_dump_ctx_records(lasv_graph *g, uint8_t *out, uint64_t capacity_records) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t rec = 8 * (size_t)g->N + 5;
  uint64_t n = 0;
  const int rc = stream_ctx_records_out(g, &n, [&](const uint8_t *h, uint64_t m, uint64_t done) -> int {
    if (done + m > capacity_records)
      return fail(LASV_ERR_CAPACITY, "need room for more than %llu records", (unsigned long long)capacity_records);
    memcpy(out + done * rec, h, m * rec);
    return LASV_OK;
  });
  return rc == LASV_OK ? (long long)n : rc;
}

int lasv_ctx_header(int kmer_size, uint32_t mean_read_len, uint64_t total_seq, uint8_t *out) {
  // print_binary_signature_NEW, file_reader.c:2930-2994, one colour, default
  // GraphInfo (graph_info.c): sample id "undefined", seq_err 0.01, no cleaning.
  uint8_t *p = out;
  auto put = [&](const void *src, size_t n) { memcpy(p, src, n); p += n; };
  int32_t v;
  put("CORTEX", 6);
  v = 6; put(&v, 4);
  v = kmer_size; put(&v, 4);
  v = (2 * kmer_size) / 64 + 1; put(&v, 4);
  v = 1; put(&v, 4);
  v = (int32_t)mean_read_len; put(&v, 4);
  long long ts = (long long)total_seq; put(&ts, 8);
  v = 9; put(&v, 4); put("undefined", 9);
  uint8_t ld[16];
  memset(ld, 0, sizeof ld);
  long double err = (long double)0.01;  // graph_info.c:127 assigns the double literal 0.01
  memcpy(ld, &err, 10);  // x87 extended: 10 significant bytes of the 16 stored
  put(ld, 16);
  uint8_t flags[4] = {0, 0, 0, 0};
  put(flags, 4);
  v = -1; put(&v, 4);
  v = -1; put(&v, 4);
  v = 9; put(&v, 4); put("undefined", 9);
  put("CORTEX", 6);
  return (int)(p - out);
}

int lasv_graph_write_ctx(lasv_graph *g, const char *path, uint32_t mean_read_len, uint64_t total_seq) {
  if (int e = use(g)) return e;
  CUDA_TRY(cudaDeviceSynchronize());
  FILE *fp = fopen(path, "wb");
  if (!fp) return fail(LASV_ERR_IO, "cannot open %s for writing: %s", path, strerror(errno));
  uint8_t hdr[128];
  const int hl = lasv_ctx_header(g->k, mean_read_len, total_seq, hdr);
  {  // the error-cleaning object of print_binary_signature_NEW: 4 flags, then the supernode threshold
    uint8_t *flags = hdr + 6 + 16 + 4 + 8 + 4 + 9 + 16;
    flags[0] = g->cleaned_tips ? 1 : 0;
    flags[1] = g->cleaned_sups ? 1 : 0;
    if (g->cleaned_sups) {
      const int32_t v = g->clean_threshold;
      memcpy(flags + 4, &v, 4);
    }
  }
  int rc = LASV_OK;
  if (fwrite(hdr, 1, (size_t)hl, fp) != (size_t)hl) rc = fail(LASV_ERR_IO, "short write to %s", path);
  const size_t rec = 8 * (size_t)g->N + 5;
  if (rc == LASV_OK)
    rc = stream_ctx_records_out(g, nullptr, [&](const uint8_t *h, uint64_t m, uint64_t) -> int {
      if (fwrite(h, rec, m, fp) != m) return fail(LASV_ERR_IO, "short write to %s", path);
      return LASV_OK;
    });
  if (fclose(fp) != 0 && rc == LASV_OK) rc = fail(LASV_ERR_IO, "cannot close %s: %s", path, strerror(errno));
  return rc;
}

long long lasv_graph_append_ctx_records(lasv_graph *g, const char *path) {
  if (int e = use(g)) return e;
  if (!path) return fail(LASV_ERR_ARG, "path is NULL");
  CUDA_TRY(cudaDeviceSynchronize());
  FILE *fp = fopen(path, "ab");
  if (!fp) return fail(LASV_ERR_IO, "cannot open %s for appending: %s", path, strerror(errno));
  const size_t rec = 8 * (size_t)g->N + 5;
  uint64_t n = 0;
  int rc = stream_ctx_records_out(g, &n, [&](const uint8_t *h, uint64_t m, uint64_t) -> int {
    if (fwrite(h, rec, m, fp) != m) return fail(LASV_ERR_IO, "short write to %s", path);
    return LASV_OK;
  });
  if (fclose(fp) != 0 && rc == LASV_OK) rc = fail(LASV_ERR_IO, "cannot close %s: %s", path, strerror(errno));
  return rc == LASV_OK ? (long long)n : rc;
}

int lasv_graph_load_ctx_records(lasv_graph *g, const uint8_t *records, uint64_t n_records) {
  if (int e = use(g)) return e;
  if (g->finalized) return fail(LASV_ERR_STATE, "graph is finalised; no further inserts");
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t rec = 8 * (size_t)g->N + 5;
  const uint64_t per = (256ull << 20) / rec;
  uint8_t *d_buf = nullptr;
  CUDA_TRY(cudaMalloc(&d_buf, std::max<uint64_t>(1, std::min(per, n_records)) * rec));
  int rc = [&]() -> int {
    for (uint64_t i = 0; i < n_records; i += per) {
      const uint64_t m = std::min(per, n_records - i);
      CUDA_TRY(cudaMemcpyAsync(d_buf, records + i * rec, m * rec, cudaMemcpyHostToDevice, g->compute));
      launch_ingest_ctx(g->N, d_buf, m, g->view(), g->d_ctr, g->compute);
      if (int e = check_launch()) return e;
      CUDA_TRY(cudaStreamSynchronize(g->compute));
    }
    GraphCounters c;
    CUDA_TRY(cudaMemcpy(&c, g->d_ctr, sizeof c, cudaMemcpyDeviceToHost));
    if (c.table_full) return fail(LASV_ERR_TABLE_FULL, "%s", TABLE_FULL_MSG);
    return LASV_OK;
  }();
  cudaFree(d_buf);
  return rc;
}

// packed .ctx records from the current position of `fp` to its end, into the device table
static int stream_ctx_records(lasv_graph *g, FILE *fp, unsigned long long *n_records) {
  const size_t rec = 8 * (size_t)g->N + 5;
  const uint64_t per = (64ull << 20) / rec;
  std::vector<uint8_t> buf(per * rec);
  unsigned long long n = 0;
  int rc = LASV_OK;
  // read in BYTES: a trailing partial record (truncated file) must be an error, not a silently shorter graph
  size_t carry = 0;
  for (;;) {
    const size_t got = fread(buf.data() + carry, 1, buf.size() - carry, fp);
    const size_t have = carry + got, whole = have / rec;
    if (whole) {
      rc = lasv_graph_load_ctx_records(g, buf.data(), whole);
      n += whole;
      if (rc != LASV_OK) break;
    }
    carry = have - whole * rec;
    if (carry) memmove(buf.data(), buf.data() + whole * rec, carry);
    if (got == 0 || feof(fp) || ferror(fp)) break;
  }
  if (n_records) *n_records = n;
  if (rc == LASV_OK && ferror(fp)) rc = fail(LASV_ERR_IO, "read error in the .ctx records: %s", strerror(errno));
  if (rc == LASV_OK && carry)
    rc = fail(LASV_ERR_IO, "truncated .ctx: %zu trailing bytes are not a whole %zu-byte record (after %llu records)", carry, rec, n);
  return rc;
}

int lasv_graph_load_ctx_file(lasv_graph *g, const char *path, uint32_t *mean_read_len,
                             uint64_t *total_seq) {
  if (int e = use(g)) return e;
  FILE *fp = fopen(path, "rb");
  if (!fp) return fail(LASV_ERR_IO, "cannot open %s: %s", path, strerror(errno));
  // query_binary_NEW, file_reader.c:3023-3232 (single colour)
  auto bad = [&](const char *why) {
    fclose(fp);
    return fail(LASV_ERR_IO, "%s is not a loadable .ctx: %s", path, why);
  };
  char magic[6];
  int32_t version, k, nbf, ncols;
  if (fread(magic, 1, 6, fp) != 6 || memcmp(magic, "CORTEX", 6)) return bad("missing magic number");
  if (fread(&version, 4, 1, fp) != 1 || version < 4 || version > 6) return bad("version not in [4,6]");
  if (fread(&k, 4, 1, fp) != 1 || k != g->k) return bad("kmer size differs from the graph's");
  if (fread(&nbf, 4, 1, fp) != 1 || nbf != g->N) return bad("number of bitfields differs");
  if (fread(&ncols, 4, 1, fp) != 1 || ncols != 1) return bad("only single-colour binaries are supported");
  int32_t mrl;
  long long ts;
  if (fread(&mrl, 4, 1, fp) != 1 || fread(&ts, 8, 1, fp) != 1) return bad("truncated header");
  if (version > 5) {
    int32_t idlen;
    if (fread(&idlen, 4, 1, fp) != 1 || idlen < 0 || idlen > 100000) return bad("bad sample id");
    if (fseek(fp, idlen + 16 + 4, SEEK_CUR)) return bad("truncated header");
    int32_t t1, t2, namelen;
    if (fread(&t1, 4, 1, fp) != 1 || fread(&t2, 4, 1, fp) != 1 || fread(&namelen, 4, 1, fp) != 1 ||
        namelen < 0 || namelen > 100000 || fseek(fp, namelen, SEEK_CUR))
      return bad("truncated header");
  }
  if (fread(magic, 1, 6, fp) != 6 || memcmp(magic, "CORTEX", 6)) return bad("missing end-of-header magic");
  if (mean_read_len) *mean_read_len = (uint32_t)mrl;
  if (total_seq) *total_seq = (uint64_t)ts;
  const int rc = stream_ctx_records(g, fp, nullptr);
  fclose(fp);
  return rc;
}


// ------------------------------------------------------------ one process, several GPUs
// lasv_multi_graph: the hash-sharded build of SURVEY 8e driven by ONE host thread, for callers that are one process
// (the reference's main(): graph_operations.c:746-765 calls the loader once and expects its HashTable filled).
// Shard o of n (a power of two) lives on devices[o] and owns the buckets whose index has high bits o
// (lasv_b200/sharded.py is the same scheme with one process per GPU).  Host batches go to the shards in turn:
//   copy + bin kernel on the batch's device (records by (owner, region) into that device's SEND stripes, appended
//   until the stripes hold `round_bytes` of reads)
//   exchange: block (sender d -> owner o) by cudaMemcpyPeerAsync into slot d of o's RECEIVE stripes -- copy engines
//   over NVLink when the devices are peers
//   insert: every owner inserts the received stripes of all senders (streamed when dense, lasv_graph_insert_bins_device)
struct lasv_multi_graph {
  int n = 0, k = 0, L = 0, W = 0, max_rehash = 15;
  std::vector<int> devices;
  std::vector<lasv_graph *> shard;
  struct Dev {
    void *send = nullptr, *send_counts = nullptr, *recv = nullptr, *recv_counts = nullptr;
    cudaStream_t comm = nullptr;
    cudaEvent_t binned = nullptr, sent = nullptr, received = nullptr, inserted = nullptr;
    lasv_graph::ShardSink sink;
  };
  std::vector<Dev> dev;
  uint32_t bins_per_owner = 0, cap = 0, spill = 0, rec_bytes = 0, count_stride = 0;
  uint64_t records_per_owner = 0, counters_per_owner = 0, round_bound = 0;
  int turn = 0;
  bool any_binned = false;
};

extern "C++" {
namespace {
void multi_release(lasv_multi_graph *m) {
  for (int d = 0; d < (int)m->dev.size(); d++) {
    cudaSetDevice(m->devices[d]);
    cudaDeviceSynchronize();
    auto &x = m->dev[d];
    cudaFree(x.send); cudaFree(x.send_counts); cudaFree(x.recv); cudaFree(x.recv_counts);
    if (x.comm) cudaStreamDestroy(x.comm);
    for (cudaEvent_t e : {x.binned, x.sent, x.received, x.inserted}) if (e) cudaEventDestroy(e);
  }
  for (auto &g : m->shard) if (g) { g->sink = nullptr; g->multi = nullptr; lasv_graph_free(&g); }
  delete m;
}

// Exchange what the send stripes hold and insert it on the owners.  Everything is queued; nothing waits.
int multi_ship(lasv_multi_graph *m) {
  if (!m->any_binned) return LASV_OK;
  const int n = m->n;
  const size_t block_bytes = (size_t)m->records_per_owner * m->rec_bytes;
  const size_t count_bytes = (size_t)m->counters_per_owner * m->count_stride;
  for (int d = 0; d < n; d++) {
    CUDA_TRY(cudaSetDevice(m->devices[d]));
    CUDA_TRY(cudaEventRecord(m->dev[d].binned, m->shard[d]->compute));
  }
  for (int d = 0; d < n; d++) {  // sender d pushes its n blocks
    auto &x = m->dev[d];
    CUDA_TRY(cudaSetDevice(m->devices[d]));
    CUDA_TRY(cudaStreamWaitEvent(x.comm, x.binned, 0));
    for (int i = 0; i < n; i++) {
      const int o = (d + i) % n;  // start with the local block, spread the peers
      CUDA_TRY(cudaStreamWaitEvent(x.comm, m->dev[o].inserted, 0));  // o's receive stripes are free again
      if (x.sink.fresh) {  // this device binned nothing in this round: its counters on the owners must say so
        CUDA_TRY(cudaMemsetAsync((uint8_t *)m->dev[o].recv_counts + (size_t)d * count_bytes, 0, count_bytes, x.comm));
        continue;
      }
      CUDA_TRY(cudaMemcpyPeerAsync((uint8_t *)m->dev[o].recv + (size_t)d * block_bytes, m->devices[o],
                                   (uint8_t *)x.send + (size_t)o * block_bytes, m->devices[d], block_bytes, x.comm));
      CUDA_TRY(cudaMemcpyPeerAsync((uint8_t *)m->dev[o].recv_counts + (size_t)d * count_bytes, m->devices[o],
                                   (uint8_t *)x.send_counts + (size_t)o * count_bytes, m->devices[d], count_bytes, x.comm));
    }
    CUDA_TRY(cudaEventRecord(x.sent, x.comm));
    CUDA_TRY(cudaStreamWaitEvent(m->shard[d]->compute, x.sent, 0));  // the send stripes may be refilled after this
    x.sink.fresh = true;
    x.sink.bound = 0;
  }
  for (int o = 0; o < n; o++) {  // owner o inserts the blocks of all senders
    CUDA_TRY(cudaSetDevice(m->devices[o]));
    lasv_graph *g = m->shard[o];
    for (int d = 0; d < n; d++) CUDA_TRY(cudaStreamWaitEvent(g->compute, m->dev[d].sent, 0));
    g->sink = nullptr;  // (an ordinary entry point of the shard)
    const int rc = lasv_graph_insert_bins_device(g, m->dev[o].recv, m->dev[o].recv_counts, n, m->bins_per_owner, m->cap,
                                                 m->spill, g->compute);
    g->sink = &m->dev[o].sink;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(m->dev[o].inserted, g->compute));
  }
  m->any_binned = false;
  return LASV_OK;
}

int multi_load_host(lasv_multi_graph *m, const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes,
                    const uint64_t *offsets, uint64_t n_reads, int cutoff) {
  if (!n_reads) return LASV_OK;
  const int d = m->turn;
  lasv_graph *g = m->shard[d];
  const uint64_t bound = kmer_bound(g, n_bytes, n_reads);
  if (bound > m->round_bound) return fail(LASV_ERR_ARG, "a host batch of %llu bytes exceeds the exchange stripes of the multi-GPU graph",
                                          (unsigned long long)n_bytes);
  if (m->dev[d].sink.bound + bound > m->round_bound)
    if (int e = multi_ship(m)) return e;
  m->turn = (d + 1) % m->n;
  m->any_binned = true;
  return load_reads_host_impl(g, seq, qual, n_bytes, offsets, n_reads, cutoff, nullptr, true, nullptr);
}

// A chunk of FASTQ text to the next device in turn (device parse there); 1 = not plain four-line FASTQ, nothing loaded.
int multi_load_raw(lasv_multi_graph *m, const uint8_t *text, uint64_t n_bytes, int cutoff, uint64_t *n_reads) {
  if (!n_bytes) return LASV_OK;
  const int d = m->turn;
  lasv_graph *g = m->shard[d];
  const uint64_t bound = n_bytes / 2;  // the sequence lines are less than half of the text
  if (bound > m->round_bound) return fail(LASV_ERR_ARG, "a FASTQ chunk of %llu bytes exceeds the exchange stripes of the multi-GPU graph",
                                          (unsigned long long)n_bytes);
  if (m->dev[d].sink.bound + bound > m->round_bound)
    if (int e = multi_ship(m)) return e;
  const int rc = load_raw_fastq_host(g, text, n_bytes, cutoff, n_reads);
  if (rc == LASV_OK && *n_reads) {
    m->turn = (d + 1) % m->n;
    m->any_binned = true;
  }
  return rc;
}

int multi_sync(lasv_multi_graph *m) {
  if (int e = multi_ship(m)) return e;
  for (int d = 0; d < m->n; d++) {
    CUDA_TRY(cudaSetDevice(m->devices[d]));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  return LASV_OK;
}

int multi_stats(lasv_multi_graph *m, lasv_load_stats *out, uint64_t *hist, uint64_t hist_size) {
  if (int e = multi_sync(m)) return e;
  lasv_load_stats acc;
  memset(&acc, 0, sizeof acc);
  std::vector<uint64_t> h(hist_size);
  if (hist) for (uint64_t i = 0; i < hist_size; i++) hist[i] = 0;
  for (int d = 0; d < m->n; d++) {
    lasv_load_stats s;
    lasv_graph *g = m->shard[d];
    g->sink = nullptr;
    const int rc = lasv_graph_stats(g, &s, hist ? h.data() : nullptr, hist ? hist_size : 0, nullptr);
    g->sink = &m->dev[d].sink;
    if (rc) return rc;
    acc.n_reads += s.n_reads;
    acc.bad_reads += s.bad_reads;
    acc.bases_read += s.bases_read;
    acc.bases_loaded += s.bases_loaded;
    acc.n_contigs += s.n_contigs;
    acc.kmers_inserted += s.kmers_inserted;
    acc.unique_kmers += s.unique_kmers;
    if (hist) for (uint64_t i = 0; i < hist_size; i++) hist[i] += h[i];
  }
  if (out) *out = acc;
  return LASV_OK;
}
}  // namespace
}

lasv_multi_graph *lasv_multi_graph_new(int kmer_size, int number_bits, int bucket_size, int max_rehash_tries,
                                       const int *devices, int n_devices, uint64_t round_bytes) {
  if (!devices || n_devices < 1 || n_devices > 16 || (n_devices & (n_devices - 1))) {
    fail(LASV_ERR_ARG, "the number of devices must be a power of two <= 16, got %d", n_devices);
    return nullptr;
  }
  const int bits = log2_u32((uint32_t)n_devices);
  if (number_bits - bits < 1) {
    fail(LASV_ERR_ARG, "too few buckets for %d shards", n_devices);
    return nullptr;
  }
  lasv_multi_graph *m = new lasv_multi_graph();
  m->n = n_devices;
  m->k = kmer_size;
  m->L = number_bits;
  m->W = bucket_size;
  m->max_rehash = max_rehash_tries;
  m->devices.assign(devices, devices + n_devices);
  m->shard.assign((size_t)n_devices, nullptr);
  m->dev.resize((size_t)n_devices);
  auto bail = [&](const char *what) -> lasv_multi_graph * {
    if (what) fail(LASV_ERR_CUDA, "multi-GPU graph: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    multi_release(m);
    return nullptr;
  };
  for (int d = 0; d < n_devices; d++) {
    m->shard[d] = lasv_graph_new(kmer_size, number_bits - bits, bucket_size, max_rehash_tries, devices[d]);
    if (!m->shard[d]) return bail(nullptr);
  }
  for (int d = 0; d < n_devices; d++) {  // peers where the hardware allows it (else the copies are staged by the driver)
    cudaSetDevice(devices[d]);
    for (int o = 0; o < n_devices; o++) {
      if (devices[o] == devices[d]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, devices[d], devices[o]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(devices[o], 0);
      cudaGetLastError();  // "already enabled" is fine
    }
  }
  // exchange stripes for `round_bytes` of reads per device and round: at most a quarter of what is free next to the
  // smallest shard (send + receive), 2 GB of reads by default
  uint64_t want = round_bytes ? round_bytes : (2ull << 30);
  {
    size_t min_free = ~(size_t)0;
    for (int d = 0; d < n_devices; d++) {
      size_t f = 0, t = 0;
      cudaSetDevice(devices[d]);
      if (cudaMemGetInfo(&f, &t) != cudaSuccess) return bail("cudaMemGetInfo");
      int sharing = 0;  // shards on the same device (tests) share its memory
      for (int o = 0; o < n_devices; o++) sharing += devices[o] == devices[d];
      min_free = std::min(min_free, f / (size_t)sharing);
    }
    const uint64_t per_byte = 2ull * 17ull;  // send + receive, <= 1 record of 16 (32: below) bytes per stream byte + slack
    const uint64_t rec = 8ull * bin_record_words(m->shard[0]->N, kmer_size);
    const uint64_t fit = (uint64_t)(min_free / 4) / (per_byte * rec / 16);
    want = std::max<uint64_t>(std::min(want, fit), 1ull << 20);
  }
  uint32_t rb = 0, cs = 0;
  if (lasv_graph_bin_geometry(m->shard[0], n_devices, want, 0, &m->bins_per_owner, &m->cap, &m->spill, &rb, &cs) != LASV_OK)
    return bail(nullptr);
  m->rec_bytes = rb;
  m->count_stride = cs;
  m->records_per_owner = (uint64_t)m->bins_per_owner * m->cap + m->spill;
  m->counters_per_owner = (uint64_t)m->bins_per_owner + 1;
  m->round_bound = want;  // (no read count assumed: a stream byte yields at most one k-mer)
  const size_t block = (size_t)m->records_per_owner * rb * (size_t)n_devices;
  const size_t cblock = (size_t)m->counters_per_owner * cs * (size_t)n_devices;
  for (int d = 0; d < n_devices; d++) {
    auto &x = m->dev[d];
    cudaSetDevice(devices[d]);
    if (cudaMalloc(&x.send, block) != cudaSuccess || cudaMalloc(&x.recv, block) != cudaSuccess ||
        cudaMalloc(&x.send_counts, cblock) != cudaSuccess || cudaMalloc(&x.recv_counts, cblock) != cudaSuccess ||
        cudaMemset(x.send_counts, 0, cblock) != cudaSuccess || cudaMemset(x.recv_counts, 0, cblock) != cudaSuccess ||
        cudaStreamCreateWithFlags(&x.comm, cudaStreamNonBlocking) != cudaSuccess)
      return bail("exchange stripes");
    for (cudaEvent_t *e : {&x.binned, &x.sent, &x.received, &x.inserted})
      if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return bail("events");
    x.sink.n_owners = n_devices;
    x.sink.bins_per_owner = m->bins_per_owner;
    x.sink.cap = m->cap;
    x.sink.spill = m->spill;
    x.sink.records = x.send;
    x.sink.counts = x.send_counts;
    m->shard[d]->sink = &x.sink;
  }
  return m;
}

void lasv_multi_graph_free(lasv_multi_graph **mp) {
  if (!mp || !*mp) return;
  multi_release(*mp);
  *mp = nullptr;
}

int lasv_multi_graph_n_shards(lasv_multi_graph *m) { return m ? m->n : -1; }

int lasv_multi_graph_load_reads_host(lasv_multi_graph *m, const uint8_t *seq, const uint8_t *qual, uint64_t n_bytes,
                                     const uint64_t *offsets, uint64_t n_reads, int quality_cutoff) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  return multi_load_host(m, seq, qual, n_bytes, offsets, n_reads, quality_cutoff);
}

int lasv_multi_graph_flush(lasv_multi_graph *m) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  return multi_sync(m);
}

int lasv_multi_graph_stats(lasv_multi_graph *m, lasv_load_stats *out, uint64_t *contig_hist, uint64_t hist_size) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  return multi_stats(m, out, contig_hist, hist_size);
}

int lasv_multi_graph_rehash_histogram(lasv_multi_graph *m, long long out16[16]) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  if (int e = multi_sync(m)) return e;
  for (int r = 0; r < 16; r++) out16[r] = 0;
  for (int d = 0; d < m->n; d++) {
    long long h[16];
    lasv_graph *g = m->shard[d];
    g->sink = nullptr;
    const int rc = lasv_graph_rehash_histogram(g, h);
    g->sink = &m->dev[d].sink;
    if (rc) return rc;
    for (int r = 0; r < 16; r++) out16[r] += h[r];
  }
  return LASV_OK;
}

int lasv_multi_graph_summary(lasv_multi_graph *m, lasv_table_summary *out) {
  if (!m || !out) return fail(LASV_ERR_ARG, "NULL argument");
  if (int e = multi_sync(m)) return e;
  memset(out, 0, sizeof *out);
  for (int d = 0; d < m->n; d++) {
    lasv_table_summary s;
    lasv_graph *g = m->shard[d];
    g->sink = nullptr;
    const int rc = lasv_graph_summary(g, &s);
    g->sink = &m->dev[d].sink;
    if (rc) return rc;
    out->n_nodes += s.n_nodes;
    out->sum_covg += s.sum_covg;
    out->sum_edge_bits += s.sum_edge_bits;
    out->fingerprint += s.fingerprint;
  }
  return LASV_OK;
}

// The whole sharded graph as ONE reference-layout HashTable (hash_table.h:40-50).  Owner = high bits of the bucket
// index, so the exported shards laid end to end are the reference's table for every key that sits in its first-choice
// bucket.  A key that left its first-choice bucket (full) was rehashed INSIDE its shard; the reference's chain
// (hash_table.c:69-122,270-327: bucket(key, r) over ALL 2^L buckets) leads elsewhere, so those keys are taken out
// of the exported shards and inserted again by the reference's own rule on the merged table -- bucket r+1 only if
// bucket r holds W keys.  (At the fill the reference recommends, < 80 %, a W=250 bucket overflows about once in 10^4.)
int lasv_multi_graph_export_table(lasv_multi_graph *m, void *elements, short *next_element, long long *unique_kmers) {
  if (!m || !elements) return fail(LASV_ERR_ARG, "NULL argument");
  if (int e = multi_sync(m)) return e;
  lasv_graph *g0 = m->shard[0];
  const size_t slot = g0->slot_bytes;
  const uint64_t shard_buckets = g0->n_buckets, shard_slots = g0->n_slots, n_buckets = shard_buckets * (uint64_t)m->n;
  std::vector<short> next_local;
  short *next = next_element;
  if (!next) {
    next_local.resize(n_buckets);
    next = next_local.data();
  }
  long long uniq = 0;
  std::vector<uint8_t> moved;  // Elements outside their first-choice bucket, back to back
  for (int d = 0; d < m->n; d++) {
    lasv_graph *g = m->shard[d];
    g->sink = nullptr;
    uniq += lasv_graph_unique_kmers(g);
    const int rc = export_table_impl(g, (uint8_t *)elements + (size_t)d * shard_slots * slot, next + (size_t)d * shard_buckets,
                                     m->n > 1 ? &moved : nullptr);
    g->sink = &m->dev[d].sink;
    if (rc) return rc;
  }
  if (unique_kmers) *unique_kmers = uniq;
  if (m->n == 1) return LASV_OK;
  // back in by the reference's rule: bucket(key, r) over all 2^L buckets, r+1 only if bucket r holds W keys
  const int N = g0->N;
  const uint32_t mask = (uint32_t)(n_buckets - 1), W = (uint32_t)m->W;
  auto bucket_of = [&](const uint64_t *key, uint32_t r) -> uint32_t {
    if (N == 1) { uint64_t k1[1] = {key[0]}; return lookup3_key<1>(k1, r).c & mask; }
    if (N == 2) { uint64_t k2[2] = {key[0], key[1]}; return lookup3_key<2>(k2, r).c & mask; }
    uint64_t k3[3] = {key[0], key[1], key[2]};
    return lookup3_key<3>(k3, r).c & mask;
  };
  uint8_t *tab = (uint8_t *)elements;
  for (size_t i = 0; i < moved.size(); i += slot) {
    const uint64_t *key = reinterpret_cast<const uint64_t *>(&moved[i]);
    bool placed = false;
    for (int r = 0; r <= m->max_rehash && !placed; r++) {
      const uint32_t b = bucket_of(key, (uint32_t)r);
      if ((uint32_t)next[b] < W) {
        memcpy(tab + ((uint64_t)b * W + (uint32_t)next[b]) * slot, &moved[i], slot);
        next[b]++;
        placed = true;
      }
    }
    if (!placed) return fail(LASV_ERR_TABLE_FULL, "%s", TABLE_FULL_MSG);
  }
  return LASV_OK;
}

// ---------------------------------------------------------- sequence files
namespace {

constexpr uint64_t LOADER_BATCH_BYTES = 32ull << 20;  // x 2 slots x parser threads, pinned, kept by the graph
constexpr uint64_t LOADER_BATCH_READS = 4ull << 20;

// A raw chunk (FASTQ text) through the device parse; *n_reads receives its read count.  Falls back to the host
// parser (the reference's full grammar, pageable buffers: rare) if the chunk is not plain four-line FASTQ.
int flush_raw_batch(lasv_graph *g, HostBatch &b, int cutoff, uint64_t *n_reads) {
  int rc = g->multi ? multi_load_raw(g->multi, b.seq, b.n_bytes, cutoff, n_reads) : load_raw_fastq_host(g, b.seq, b.n_bytes, cutoff, n_reads);
  if (rc != 1) return rc;
  std::vector<uint8_t> sq(b.n_bytes + 16), ql(b.n_bytes + 16);
  std::vector<uint64_t> off(b.n_bytes / 4 + 2);
  HostBatch t;
  t.seq = sq.data();
  t.qual = ql.data();
  t.offsets = off.data();
  t.cap_bytes = b.n_bytes + 8;
  t.cap_reads = off.size() - 1;
  t.reset();
  SeqReader rd;
  rd.open_memory(b.seq, b.n_bytes, "FASTQ chunk");
  std::string err;
  uint64_t n = rd.fill_batch_fast(t, true);
  while (rd.next(err)) {
    if (!rd.append_to(t, true)) return fail(LASV_ERR_IO, "a FASTQ chunk does not fit its own size when parsed");
    n++;
    n += rd.fill_batch_fast(t, true);
  }
  if (!err.empty()) return fail(LASV_ERR_IO, "%s", err.c_str());
  *n_reads = n;
  if (!t.n_reads) return LASV_OK;
  if (g->multi) return multi_load_host(g->multi, t.seq, t.qual, t.n_bytes, t.offsets, t.n_reads, cutoff);
  return lasv_graph_load_reads_host(g, t.seq, t.qual, t.n_bytes, t.offsets, t.n_reads, cutoff, nullptr);
}

int flush_batch(lasv_graph *g, HostBatch &b, bool with_qual, int cutoff, uint64_t *raw_reads) {
  if (b.raw) {
    uint64_t n = 0;
    const int rc = b.n_bytes ? flush_raw_batch(g, b, cutoff, &n) : LASV_OK;
    if (raw_reads) *raw_reads += n;
    b.reset();
    return rc;
  }
  if (!b.n_reads) return LASV_OK;
  int rc;
  if (g->multi)  // shard 0 of a multi-GPU graph stands in for the whole graph: the batch goes to the next device in turn
    rc = multi_load_host(g->multi, b.seq, with_qual ? b.qual : nullptr, b.n_bytes, b.offsets, b.n_reads, cutoff);
  else
    rc = lasv_graph_load_reads_host(g, b.seq, with_qual ? b.qual : nullptr, b.n_bytes, b.offsets, b.n_reads, cutoff, nullptr);
  b.reset();
  return rc;
}
// the loaders' statistics: of the graph, or of the multi-GPU graph it stands in for
int loader_stats(lasv_graph *g, lasv_load_stats *s) {
  if (g->multi) return multi_stats(g->multi, s, nullptr, 0);
  return lasv_graph_stats(g, s, nullptr, 0, nullptr);
}

size_t find_fastq_record(const unsigned char *d, size_t n, size_t from);

// One parser thread per file: it fills two pinned batches in turn while the caller's thread hands the
// full ones to the GPU.  (The text parse is the slow end of the file loaders by two orders of magnitude;
// the two mates of a pair need not be interleaved -- the order of reads is irrelevant to the graph,
// SURVEY 8a-B -- so each file streams on its own and only the read counts are compared.)
struct FileProducer {
  SeqReader reader;
  std::string err;
  bool with_qual = false;
  HostBatch *slot[2] = {nullptr, nullptr};
  int full[2] = {0, 0};  // guarded by m
  bool done = false;     // guarded by m: no further batches will be produced
  bool too_long = false;
  std::atomic<uint64_t> reads{0};  // (raw mode: counted by the consumer, else by the parser thread)
  std::mutex m;
  std::condition_variable cv;
  std::thread th;
  // raw mode (a slice of a mapped plain four-line FASTQ file): the thread only copies whole records of text into the
  // pinned batches, the device takes them apart (fastq_parse.cu); `reads` is then counted by the consumer
  const unsigned char *raw_data = nullptr;
  size_t raw_size = 0;
  // raw mode on a BGZF file: windows of inflated text (seq_reader.cpp: BgzfStream, several inflate threads), cut at
  // the last record boundary; the tail is carried into the next batch
  bool raw_bgzf = false;
  std::vector<unsigned char> raw_carry;

  // end of the longest run of whole records in raw_data[pos, ...) that fits `cap` bytes; 0 if none is found
  size_t raw_cut(size_t pos, size_t cap) const {
    if (raw_size - pos <= cap) return raw_size;
    const size_t back = std::min<size_t>(cap / 2, 1u << 16);  // a record boundary within the last 64 KB of the window
    const size_t c = find_fastq_record(raw_data, raw_size, pos + cap - back);
    return (c == SIZE_MAX || c <= pos || c - pos > cap) ? 0 : c;
  }

  void run() {
    int k = 0;
    bool have = false;  // a parsed read waits for room
    size_t raw_pos = 0;
    if (raw_bgzf) {
      if (slot[0]->cap_bytes < (20u << 20)) raw_bgzf = false;  // small batches (tests): the reader parses on the host
      else reader.windows()->rewind();                            // (open() had a look at the first window)
    }
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return !full[k]; });
      }
      HostBatch &b = *slot[k];
      b.reset();
      bool eof = false;
      if (raw_bgzf) {
        const size_t cap = (size_t)std::min<uint64_t>(b.cap_bytes - 1, (32ull << 20) - 1);
        // text = carry + windows, up to the batch's capacity (a window is ~8 MB)
        size_t n = raw_carry.size();
        bool more = true;
        if (n) memcpy(b.seq, raw_carry.data(), n);
        raw_carry.clear();
        std::string berr;
        while (more && n + (9u << 20) <= cap) {  // (inflated straight into the pinned batch)
          size_t got = 0;
          more = reader.windows()->next_into(b.seq + n, cap - n, &got, berr);
          if (!more) break;
          n += got;
        }
        if (!berr.empty()) {
          err = berr + " [file: " + reader.path() + "]";
          eof = true;
          more = false;
          n = 0;
        }
        size_t end = n;
        if (more && n) {  // not the end of the file: cut at a record boundary within the last 64 KB
          const size_t back = std::min<size_t>(n / 2, 1u << 16);
          const size_t c = find_fastq_record(b.seq, n, n - back);
          end = (c == SIZE_MAX || c == 0 || c > n) ? 0 : c;
        }
        if (n && !end) {
          // no record boundary where one should be (multi-line records ...): the host parser takes over from here
          raw_bgzf = false;
          reader.seed(std::vector<unsigned char>(b.seq, b.seq + n));
        } else {
          if (end < n) raw_carry.assign(b.seq + end, b.seq + n);
          b.n_bytes = end;
          if (end && b.seq[end - 1] != '\n') b.seq[b.n_bytes++] = '\n';  // the file's last line
          b.raw = true;
          if (!more) eof = true;
          {
            std::lock_guard<std::mutex> lk(m);
            if (b.n_bytes) full[k] = 1;
            if (eof) done = true;
          }
          cv.notify_all();
          if (eof) return;
          if (b.n_bytes) k ^= 1;
          continue;
        }
      }
      if (raw_data) {
        const size_t cap = (size_t)std::min<uint64_t>(b.cap_bytes - 1, (32ull << 20) - 1);
        const size_t end = raw_cut(raw_pos, cap);
        if (end) {
          const size_t n = end - raw_pos;
          memcpy(b.seq, raw_data + raw_pos, n);
          b.n_bytes = n;
          if (b.seq[n - 1] != '\n') b.seq[b.n_bytes++] = '\n';  // the file's last line
          b.raw = true;
          raw_pos = end;
          eof = raw_pos >= raw_size;
          {
            std::lock_guard<std::mutex> lk(m);
            full[k] = 1;
            if (eof) done = true;
          }
          cv.notify_all();
          if (eof) return;
          k ^= 1;
          continue;
        }
        // no record boundary where one should be: the rest of the slice goes through the host parser
        reader.open_memory(raw_data + raw_pos, raw_size - raw_pos, reader.path());
        raw_data = nullptr;
      }
      for (;;) {
        if (!have) {
          reads += reader.fill_batch_fast(b, with_qual);  // plain four-line records of a mapped slice
          if (!reader.next(err)) { eof = true; break; }
          have = true;
        }
        if (!reader.append_to(b, with_qual)) {
          if (b.n_reads == 0) { too_long = true; eof = true; }
          break;  // batch full: the read stays pending
        }
        have = false;
        reads++;
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (b.n_reads) full[k] = 1;
        if (eof) done = true;
      }
      cv.notify_all();
      if (eof) return;
      k ^= 1;
    }
  }
};

// An uncompressed FASTQ file mapped read-only, for the parser threads that share it.
struct MappedFile {
  int fd = -1;
  const unsigned char *data = nullptr;
  size_t size = 0;
  bool map(const char *path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) return false;
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) return false;
    data = (const unsigned char *)p;
    size = (size_t)st.st_size;
    madvise(p, size, MADV_SEQUENTIAL);
    return true;
  }
  ~MappedFile() {
    if (data) munmap((void *)data, size);
    if (fd >= 0) ::close(fd);
  }
};

// Start of the first FASTQ record at or after the line that follows position `from` (from == 0: the file
// start), recognised by the shape of the record that begins there: '@' line, sequence line, '+' line,
// quality line of the sequence's length, and then end of data or another '@'.  A quality line may start
// with '@' too, but then the line after the next one is a sequence, not a '+' line.  Multi-line records
// never match: the caller falls back to one sequential parser.  SIZE_MAX if nothing matches nearby.
size_t find_fastq_record(const unsigned char *d, size_t n, size_t from) {
  auto line_end = [&](size_t p) -> size_t {  // index of the '\n' ending the line at p, or n
    const void *nl = p < n ? memchr(d + p, '\n', n - p) : nullptr;
    return nl ? (size_t)((const unsigned char *)nl - d) : n;
  };
  auto content = [&](size_t b, size_t e) -> size_t {  // line length without a trailing '\r'
    return (e > b && d[e - 1] == '\r') ? e - b - 1 : e - b;
  };
  size_t p = from;
  if (p != 0) {
    p = line_end(p);
    if (p >= n) return SIZE_MAX;
    p++;
  }
  for (int tries = 0; tries < 64 && p < n; tries++) {
    const size_t e1 = line_end(p);
    if (d[p] == '@' && e1 < n) {
      const size_t e2 = line_end(e1 + 1);
      if (e2 < n && e2 + 1 < n && d[e2 + 1] == '+') {
        const size_t e3 = line_end(e2 + 1);
        if (e3 < n) {
          const size_t e4 = line_end(e3 + 1);
          if (content(e3 + 1, e4) == content(e1 + 1, e2) && content(e1 + 1, e2) > 0 &&
              (e4 + 1 >= n || d[e4 + 1] == '@'))
            return p;
        }
      }
    }
    if (e1 >= n) return SIZE_MAX;
    p = e1 + 1;
  }
  return SIZE_MAX;
}

// Parser threads of one or two sequence files: one per file, or -- for a large uncompressed FASTQ file --
// several, each on its own slice of the mapped file (the text parse is the slow end).
struct ParseJob {
  std::vector<std::unique_ptr<FileProducer>> prod;
  std::vector<int> file_of;  // producer -> file index (entry e: files 2e and 2e+1)
  std::vector<std::unique_ptr<MappedFile>> maps;
  std::string err;

  // entries: (file 1, file 2 or "") -- every file of every entry is parsed concurrently; `threads` parser
  // threads per large uncompressed FASTQ file (>= 1); `alloc` hands out the two batches of every producer
  // raw_slices: the slices of a mapped plain FASTQ file are handed over as TEXT (HostBatch::raw) for the device parse
  bool open(const std::vector<std::pair<std::string, std::string>> &entries, int threads, size_t min_slice_bytes,
            const std::function<HostBatch *()> &alloc, bool raw_slices = false) {
    for (size_t e = 0; e < entries.size(); e++) {
      const std::string paths[2] = {entries[e].first, entries[e].second};
      const int n_files = paths[1].empty() ? 1 : 2;
      std::vector<std::unique_ptr<FileProducer>> heads;
      for (int f = 0; f < n_files; f++) {
        heads.emplace_back(new FileProducer());
        if (!heads.back()->reader.open(paths[f], err)) return false;
      }
      // quality filtering applies to an entry if either of its files carries qualities (file_reader.c:389)
      const bool with_qual = heads[0]->reader.has_quality() || (n_files == 2 && heads[1]->reader.has_quality());
      for (int f = 0; f < n_files; f++) {
        std::vector<size_t> cuts;
        SeqReader &r = heads[f]->reader;
        maps.emplace_back(new MappedFile());
        MappedFile &m = *maps.back();
        if (threads > 1 && r.type() == SeqReader::FASTQ && r.is_plain_file() && m.map(paths[f].c_str()) &&
            m.size >= (size_t)threads * min_slice_bytes) {
          cuts.push_back(find_fastq_record(m.data, m.size, 0));
          for (int w = 1; w < threads && cuts.back() != SIZE_MAX; w++) {
            const size_t c = find_fastq_record(m.data, m.size, m.size / (size_t)threads * (size_t)w);
            if (c == SIZE_MAX || c <= cuts.back()) { cuts.assign(1, SIZE_MAX); break; }
            cuts.push_back(c);
          }
          if (cuts.back() == SIZE_MAX || cuts[0] != 0) cuts.clear();  // not the 4-line shape: sequential
        }
        const size_t first = prod.size();
        if (cuts.size() > 1) {
          cuts.push_back(m.size);
          const char *dp = getenv("LASV_B200_DEVICE_PARSE");  // 0: the parser threads take the slices apart on the host
          const bool device_parse = raw_slices && (!dp || atoi(dp) != 0);
          for (size_t w = 0; w + 1 < cuts.size(); w++) {
            prod.emplace_back(new FileProducer());
            prod.back()->reader.open_memory(m.data + cuts[w], cuts[w + 1] - cuts[w], paths[f]);
            if (device_parse && with_qual) {  // the text goes to the device as it is (fastq_parse.cu)
              prod.back()->raw_data = m.data + cuts[w];
              prod.back()->raw_size = cuts[w + 1] - cuts[w];
            }
          }
        } else {
          prod.emplace_back(std::move(heads[f]));
          const char *dp = getenv("LASV_B200_DEVICE_PARSE");
          FileProducer &hp = *prod.back();
          if (raw_slices && (!dp || atoi(dp) != 0) && with_qual && hp.reader.type() == SeqReader::FASTQ && hp.reader.windows()) {
            hp.raw_bgzf = true;  // (decided for good when the thread starts: the batches must hold a window)
          }
        }
        for (size_t i = first; i < prod.size(); i++) {
          file_of.push_back((int)(2 * e) + f);
          prod[i]->with_qual = with_qual;
        }
      }
    }
    for (auto &p : prod) {
      for (int k = 0; k < 2; k++) {
        p->slot[k] = alloc();
        if (!p->slot[k]) { err = "cannot allocate batch buffers"; return false; }
      }
    }
    return true;
  }

  // runs the producers; `consume(batch)` is called from this thread for every full batch
  void run(const std::function<void(HostBatch &, FileProducer &)> &consume) {
    for (auto &p : prod) {
      FileProducer *fp = p.get();
      fp->th = std::thread([fp] { fp->run(); });
    }
    const size_t n = prod.size();
    std::vector<int> next_slot(n, 0);
    std::vector<char> finished(n, 0);
    size_t n_finished = 0;
    while (n_finished < n) {
      bool progressed = false;
      for (size_t i = 0; i < n; i++) {
        if (finished[i]) continue;
        FileProducer &p = *prod[i];
        const int k = next_slot[i];
        bool ready = false;
        {
          std::unique_lock<std::mutex> lk(p.m);
          // never sleep long on one producer while another may have a batch ready
          if (n - n_finished == 1) p.cv.wait(lk, [&] { return p.full[k] || p.done; });
          else p.cv.wait_for(lk, std::chrono::microseconds(200), [&] { return p.full[k] || p.done; });
          ready = p.full[k] != 0;
          if (!ready && p.done) {
            finished[i] = 1;
            n_finished++;
          }
        }
        if (!ready) continue;
        consume(*p.slot[k], p);
        {
          std::lock_guard<std::mutex> lk(p.m);
          p.full[k] = 0;
        }
        p.cv.notify_all();
        next_slot[i] ^= 1;
        progressed = true;
      }
      (void)progressed;
    }
    for (auto &p : prod) p->th.join();
  }

  uint64_t reads_of_file(int f) const {
    uint64_t n = 0;
    for (size_t i = 0; i < prod.size(); i++)
      if (file_of[i] == f) n += prod[i]->reads;
    return n;
  }
};

// Parser threads per file and the smallest slice worth a thread of its own (every thread brings two pinned
// batches whose allocation costs more than parsing a small file).  LASV_B200_PARSE_THREADS overrides both.
int parse_threads_per_file(int n_files, size_t *min_slice_bytes) {
  *min_slice_bytes = 32u << 20;
  if (const char *e = getenv("LASV_B200_PARSE_THREADS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 64) {
      *min_slice_bytes = 1u << 20;
      return v;
    }
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const int per = hw ? (int)hw / (2 * n_files) : 1;  // leave half of the cores to the rest of the process
  return std::max(1, std::min(4, per));
}

void add_stats(lasv_load_stats *acc, const lasv_load_stats &s);

constexpr size_t LOADER_GROUP = 4;  // entries (files or file pairs) of a filelist parsed at the same time

// Common body of the SE and PE loaders (file_reader.c:475-773) for a group of files / file pairs that
// are parsed concurrently (gzip input is bound by inflate: one file, one thread -- but a filelist usually
// names several files).
int load_file_group(lasv_graph *g, const std::vector<std::pair<std::string, std::string>> &entries, int cutoff,
                    lasv_load_stats *stats) {
  if (int e = use(g)) return e;
  if (entries.empty()) {
    if (stats) memset(stats, 0, sizeof *stats);
    return LASV_OK;
  }
  size_t n_files = 0;
  for (auto &e : entries) n_files += e.second.empty() ? 1 : 2;
  size_t min_slice = 0;
  const int threads = parse_threads_per_file((int)n_files, &min_slice);
  uint64_t batch_bytes = LOADER_BATCH_BYTES;
  if (const char *e = getenv("LASV_B200_LOADER_BATCH_BYTES")) {  // tests: many small batches per file
    const long long v = atoll(e);
    if (v >= 1024 && (uint64_t)v <= LOADER_BATCH_BYTES) batch_bytes = (uint64_t)v;
  }
  if (!g->host_pool.empty() && g->host_pool[0]->cap_bytes != batch_bytes) g->host_pool.clear();
  size_t pool_next = 0;
  ParseJob job;
  if (!job.open(entries, threads, min_slice, [&]() -> HostBatch * {
        if (pool_next == g->host_pool.size()) {
          // reads of >= 15 bases fill the bytes first; shorter ones fill the read slots first -- either way a batch
          g->host_pool.emplace_back(new PinnedBatch(batch_bytes, std::min<uint64_t>(LOADER_BATCH_READS, batch_bytes / 16 + 16)));
          if (!g->host_pool.back()->ok) {
            g->host_pool.pop_back();
            return nullptr;
          }
        }
        return g->host_pool[pool_next++].get();
      }, true))
    return fail(job.err.find("batch buffers") != std::string::npos ? LASV_ERR_CUDA : LASV_ERR_IO, "%s",
                job.err.c_str());
  lasv_load_stats before;
  if (int e = loader_stats(g, &before)) return e;
  // Large tables: the batches of the whole group accumulate in the graph's stripes and are inserted when those are
  // full and at the end of the group (the streamed insert pays from about a record per table line on); sized from
  // the files' bytes, which bound the stream bytes.
  bool group_accumulates = false;
  if (!g->multi && !g->pipeline && g->acc_limit == 0) {
    uint64_t file_bytes = 0;
    for (auto &e : entries)
      for (const std::string *pth : {&e.first, &e.second}) {
        struct stat sb;
        if (!pth->empty() && stat(pth->c_str(), &sb) == 0) file_bytes += (uint64_t)sb.st_size;
      }
    uint32_t nb = 0;
    if (file_bytes && want_binned(g, file_bytes, &nb) && g->table_bytes >= (256ull << 20)) {
      if (int e = lasv_graph_set_accumulation(g, file_bytes, nullptr)) return e;
      group_accumulates = true;
    }
  }
  int rc = LASV_OK;
  job.run([&](HostBatch &b, FileProducer &p) {
    uint64_t raw_reads = 0;
    if (rc == LASV_OK) rc = flush_batch(g, b, p.with_qual, cutoff, &raw_reads);  // after an error: just drain
    else b.reset();
    p.reads += raw_reads;
  });
  if (group_accumulates) {
    const int e2 = lasv_graph_set_accumulation(g, 0, nullptr);  // inserts what the stripes hold, accumulation off again
    if (rc == LASV_OK) rc = e2;
  }
  if (rc != LASV_OK) return rc;
  for (size_t i = 0; i < job.prod.size(); i++) {
    const auto &e = entries[(size_t)job.file_of[i] / 2];
    const std::string &path = (job.file_of[i] & 1) ? e.second : e.first;
    if (!job.prod[i]->err.empty()) return fail(LASV_ERR_IO, "%s", job.prod[i]->err.c_str());
    if (job.prod[i]->too_long) return fail(LASV_ERR_ARG, "a read of %s exceeds the loader batch", path.c_str());
  }
  for (size_t e = 0; e < entries.size(); e++)
    if (!entries[e].second.empty() && job.reads_of_file((int)(2 * e)) != job.reads_of_file((int)(2 * e + 1)))
      return fail(LASV_ERR_IO, "Paired-end files don't have the same number of reads\nfile1: %s\nfile2: %s",
                  entries[e].first.c_str(), entries[e].second.c_str());
  lasv_load_stats after;
  if (int e = loader_stats(g, &after)) return e;
  if (stats) {
    stats->n_reads = after.n_reads - before.n_reads;
    stats->bad_reads = after.bad_reads - before.bad_reads;
    stats->bases_read = after.bases_read - before.bases_read;
    stats->bases_loaded = after.bases_loaded - before.bases_loaded;
    stats->n_contigs = after.n_contigs - before.n_contigs;
    stats->kmers_inserted = after.kmers_inserted - before.kmers_inserted;
    stats->unique_kmers = after.unique_kmers;
  }
  return LASV_OK;
}

int load_files(lasv_graph *g, const char *path1, const char *path2, int cutoff, lasv_load_stats *stats) {
  return load_file_group(g, {{std::string(path1), std::string(path2 ? path2 : "")}}, cutoff, stats);
}

// the entries of a filelist, LOADER_GROUP at a time
int load_entries(lasv_graph *g, const std::vector<std::pair<std::string, std::string>> &entries, int cutoff,
                 lasv_load_stats *acc) {
  for (size_t i = 0; i < entries.size(); i += LOADER_GROUP) {
    const std::vector<std::pair<std::string, std::string>> group(
        entries.begin() + (long)i, entries.begin() + (long)std::min(entries.size(), i + LOADER_GROUP));
    lasv_load_stats s;
    if (int e = load_file_group(g, group, cutoff, &s)) return e;
    add_stats(acc, s);
  }
  return LASV_OK;
}

bool read_list_line(FILE *f, std::string &line) {
  line.clear();
  int c;
  bool any = false;
  while ((c = fgetc(f)) != EOF) {
    any = true;
    if (c == '\n') break;
    line.push_back((char)c);
  }
  while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
  return any;
}

std::string dir_of(const std::string &abs) {
  std::vector<char> tmp(abs.begin(), abs.end());
  tmp.push_back('\0');
  return std::string(dirname(tmp.data())) + "/";
}

void add_stats(lasv_load_stats *acc, const lasv_load_stats &s) {
  acc->n_reads += s.n_reads;
  acc->bad_reads += s.bad_reads;
  acc->bases_read += s.bases_read;
  acc->bases_loaded += s.bases_loaded;
  acc->n_contigs += s.n_contigs;
  acc->kmers_inserted += s.kmers_inserted;
  acc->unique_kmers = s.unique_kmers;
}

}  // namespace

int lasv_graph_load_se_file(lasv_graph *g, const char *path, int quality_cutoff,
                            lasv_load_stats *stats) {
  return load_files(g, path, nullptr, quality_cutoff, stats);
}

int lasv_graph_load_pe_files(lasv_graph *g, const char *path1, const char *path2, int quality_cutoff,
                             lasv_load_stats *stats) {
  return load_files(g, path1, path2, quality_cutoff, stats);
}

int lasv_graph_load_se_filelist(lasv_graph *g, const char *list, int qual_thresh, int ascii_fq_offset,
                                unsigned int *files_loaded, lasv_load_stats *stats) {
  // file_reader.c:789-907
  const int cutoff = (int)(signed char)(qual_thresh + ascii_fq_offset);
  char abs[PATH_MAX + 1];
  if (!realpath(list, abs)) return fail(LASV_ERR_IO, "Cannot get absolute path to filelist of SE files: %s", list);
  FILE *f = fopen(abs, "r");
  if (!f) return fail(LASV_ERR_IO, "Cannot open filelist of SE files: %s", abs);
  const std::string dir = dir_of(abs);
  lasv_load_stats acc;
  memset(&acc, 0, sizeof acc);
  std::string line;
  std::vector<std::pair<std::string, std::string>> entries;
  int rc = LASV_OK;
  while (rc == LASV_OK && read_list_line(f, line)) {
    if (line.empty()) continue;
    if (line[0] != '/') line = dir + line;
    char p[PATH_MAX + 1];
    if (!realpath(line.c_str(), p)) { rc = fail(LASV_ERR_IO, "Cannot find sequence file: %s", line.c_str()); break; }
    entries.emplace_back(std::string(p), std::string());
  }
  fclose(f);
  if (rc == LASV_OK) rc = load_entries(g, entries, cutoff, &acc);
  if (rc != LASV_OK) return rc;
  const unsigned int n = (unsigned int)entries.size();
  if (files_loaded) *files_loaded += n;
  if (stats) *stats = acc;
  return LASV_OK;
}

int lasv_graph_load_pe_filelists(lasv_graph *g, const char *list1, const char *list2, int qual_thresh,
                                 int ascii_fq_offset, unsigned int *pairs_loaded,
                                 lasv_load_stats *stats) {
  // file_reader.c:910-1077
  const int cutoff = (int)(signed char)(qual_thresh + ascii_fq_offset);
  char abs1[PATH_MAX + 1], abs2[PATH_MAX + 1];
  if (!realpath(list1, abs1)) return fail(LASV_ERR_IO, "Cannot get absolute path to filelist of PE files: %s", list1);
  if (!realpath(list2, abs2)) return fail(LASV_ERR_IO, "Cannot get absolute path to filelist of PE files: %s", list2);
  FILE *f1 = fopen(abs1, "r");
  if (!f1) return fail(LASV_ERR_IO, "Cannot open filelist of PE files: %s", abs1);
  FILE *f2 = fopen(abs2, "r");
  if (!f2) { fclose(f1); return fail(LASV_ERR_IO, "Cannot open filelist of PE files: %s", abs2); }
  const std::string dir1 = dir_of(abs1), dir2 = dir_of(abs2);
  lasv_load_stats acc;
  memset(&acc, 0, sizeof acc);
  std::string l1, l2;
  std::vector<std::pair<std::string, std::string>> entries;
  int rc = LASV_OK;
  for (;;) {
    const bool g1 = read_list_line(f1, l1), g2 = read_list_line(f2, l2);
    if (!g1 && !g2) break;
    if (l1.empty() != l2.empty()) {
      rc = fail(LASV_ERR_IO, "Paired-end files don't have the same number of lines:\nFile 1: %s\nFile 2: %s",
                list1, list2);
      break;
    }
    if (l1.empty()) continue;
    if (l1[0] != '/') l1 = dir1 + l1;
    if (l2[0] != '/') l2 = dir2 + l2;
    char p1[PATH_MAX + 1], p2[PATH_MAX + 1];
    if (!realpath(l1.c_str(), p1)) { rc = fail(LASV_ERR_IO, "Cannot find sequence file: %s", l1.c_str()); break; }
    if (!realpath(l2.c_str(), p2)) { rc = fail(LASV_ERR_IO, "Cannot find sequence file: %s", l2.c_str()); break; }
    entries.emplace_back(std::string(p1), std::string(p2));
  }
  fclose(f1);
  fclose(f2);
  if (rc == LASV_OK) rc = load_entries(g, entries, cutoff, &acc);
  if (rc != LASV_OK) return rc;
  const unsigned int n = (unsigned int)entries.size();
  if (pairs_loaded) *pairs_loaded += n;
  if (stats) *stats = acc;
  return LASV_OK;
}

int lasv_multi_graph_load_se_filelist(lasv_multi_graph *m, const char *list, int qual_thresh, int ascii_fq_offset,
                                      unsigned int *files_loaded, lasv_load_stats *stats) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  lasv_graph *g = m->shard[0];
  g->multi = m;
  const int rc = lasv_graph_load_se_filelist(g, list, qual_thresh, ascii_fq_offset, files_loaded, stats);
  g->multi = nullptr;
  return rc;
}

int lasv_multi_graph_load_pe_filelists(lasv_multi_graph *m, const char *list1, const char *list2, int qual_thresh,
                                       int ascii_fq_offset, unsigned int *pairs_loaded, lasv_load_stats *stats) {
  if (!m) return fail(LASV_ERR_ARG, "NULL multi graph");
  lasv_graph *g = m->shard[0];
  g->multi = m;
  const int rc = lasv_graph_load_pe_filelists(g, list1, list2, qual_thresh, ascii_fq_offset, pairs_loaded, stats);
  g->multi = nullptr;
  return rc;
}

// ----------------------------------------------- (A) reference-named drop-ins
namespace {

[[noreturn]] void die_like_reference(const char *msg) {  // src/basic/global.c:44-63
  fflush(stdout);
  fprintf(stderr, "Error: %s", msg);
  const size_t n = strlen(msg);
  if (!n || msg[n - 1] != '\n') fprintf(stderr, "\n");
  exit(EXIT_FAILURE);
}

int log2_exact(long long v) {
  int l = 0;
  while ((1ll << l) < v) l++;
  return ((1ll << l) == v) ? l : -1;
}

// Build on the GPU into the caller's reference HashTable.
void dropin_load(lasv_ref_hash_table *t, const char *list1, const char *list2, int qual_thresh,
                 int homopol_limit, signed char remove_dups, char ascii_fq_offset, char is_colour_list,
                 unsigned int *files_loaded, unsigned long long *bad_reads,
                 unsigned long long *dup_reads, unsigned long long *bases_read,
                 unsigned long long *bases_loaded, unsigned long *hist, unsigned long hist_size) {
  (void)dup_reads;
  if (homopol_limit != 0) die_like_reference("lasv_b200: homopolymer cut-off is not supported by the GPU loader");
  if (remove_dups) die_like_reference("lasv_b200: PCR duplicate removal is not supported by the GPU loader");
  if (is_colour_list) die_like_reference("lasv_b200: colour lists are not supported by the GPU loader");
  if (!t || !t->table) die_like_reference("NULL table!");
  const int L = log2_exact(t->number_buckets);
  if (L < 0) die_like_reference("lasv_b200: number_buckets is not a power of two");
  int dev = 0;
  if (const char *e = getenv("LASV_B200_DEVICE")) dev = atoi(e);
  // LASV_B200_DEVICES=0,1,2,3 (or "all"): the table is hash-sharded over these GPUs (lasv_multi_graph) -- a power of
  // two of them; the caller's table must be empty (loading on top of an existing graph stays on one device)
  std::vector<int> devs;
  if (const char *e = getenv("LASV_B200_DEVICES")) {
    if (!strcmp(e, "all")) {
      for (int i = 0, n_dev = lasv_device_count(); i < n_dev; i++) devs.push_back(i);
      while (devs.size() & (devs.size() - 1)) devs.pop_back();
    } else {
      for (const char *q = e; *q;) {
        devs.push_back(atoi(q));
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
    }
  }
  lasv_load_stats s;
  memset(&s, 0, sizeof s);
  unsigned int n = 0;
  long long coll[16];
  long long uniq = 0;
  printf(list2 ? "Load paired-ended files\n" : "Load single-ended files\n");
  if (devs.size() > 1 && t->unique_kmers == 0) {
    lasv_multi_graph *m = lasv_multi_graph_new(t->kmer_size, L, t->bucket_size, t->max_rehash_tries, devs.data(), (int)devs.size(), 0);
    if (!m) die_like_reference(lasv_last_error());
    const int rc = list2 ? lasv_multi_graph_load_pe_filelists(m, list1, list2, qual_thresh, ascii_fq_offset, &n, &s)
                         : lasv_multi_graph_load_se_filelist(m, list1, qual_thresh, ascii_fq_offset, &n, &s);
    if (rc != LASV_OK) die_like_reference(lasv_last_error());
    if (hist && hist_size) {
      std::vector<uint64_t> h(hist_size);
      if (lasv_multi_graph_stats(m, nullptr, h.data(), hist_size) != LASV_OK) die_like_reference(lasv_last_error());
      for (unsigned long i = 0; i < hist_size; i++) hist[i] += (unsigned long)h[i];
    }
    if (lasv_multi_graph_rehash_histogram(m, coll) != LASV_OK) die_like_reference(lasv_last_error());
    if (lasv_multi_graph_export_table(m, t->table, t->next_element, &uniq) != LASV_OK) die_like_reference(lasv_last_error());
    lasv_multi_graph_free(&m);
  } else {
    lasv_graph *g = lasv_graph_new(t->kmer_size, L, t->bucket_size, t->max_rehash_tries, devs.size() == 1 ? devs[0] : dev);
    if (!g) die_like_reference(lasv_last_error());
    long long coll0[16] = {0};  // rehashes of the import itself: the caller's collisions[] holds those already
    if (t->unique_kmers > 0 &&
        (lasv_graph_import_table(g, t->table) != LASV_OK || lasv_graph_rehash_histogram(g, coll0) != LASV_OK))
      die_like_reference(lasv_last_error());
    const int rc = list2 ? lasv_graph_load_pe_filelists(g, list1, list2, qual_thresh, ascii_fq_offset, &n, &s)
                         : lasv_graph_load_se_filelist(g, list1, qual_thresh, ascii_fq_offset, &n, &s);
    if (rc != LASV_OK) die_like_reference(lasv_last_error());
    if (hist && hist_size) {
      std::vector<uint64_t> h(hist_size);
      if (lasv_graph_stats(g, nullptr, h.data(), hist_size, nullptr) != LASV_OK)
        die_like_reference(lasv_last_error());
      for (unsigned long i = 0; i < hist_size; i++) hist[i] += (unsigned long)h[i];
    }
    if (lasv_graph_rehash_histogram(g, coll) != LASV_OK) die_like_reference(lasv_last_error());
    for (int r = 1; r < 16; r++) {  // (entry 0 is derived from the k-mers the loaders counted, which the import does not touch)
      coll[r] -= coll0[r];
      coll[0] += coll0[r];
    }
    uniq = lasv_graph_unique_kmers(g);
    if (lasv_graph_export_table(g, t->table, t->next_element) != LASV_OK)
      die_like_reference(lasv_last_error());
    lasv_graph_free(&g);
  }
  t->unique_kmers = uniq;
  if (t->collisions)
    for (int r = 0; r < t->max_rehash_tries && r < 16; r++) t->collisions[r] += coll[r];
  *files_loaded += n;
  *bad_reads += s.bad_reads;
  *bases_read += s.bases_read;
  *bases_loaded += s.bases_loaded;
  // file_reader.c:898-906 / 1068-1076
  printf(list2 ? "\nNum PE file pairs loaded:%u\n" : "\nNum SE files loaded:%u\n", n);
  printf("\tKmers:%lld\n", uniq);
  printf("\tNumber of bad reads:%llu\n", (unsigned long long)s.bad_reads);
  printf("\tNumber of dupe reads:%llu\n", 0ull);
  printf(list2 ? "\tPE sequence parsed:%llu\n" : "\tSE sequence parsed:%llu\n",
         (unsigned long long)s.bases_read);
  printf(list2 ? "\tTotal PE sequence that passed filters:%llu\n"
               : "\tTotal SE sequence that passed filters:%llu\n",
         (unsigned long long)s.bases_loaded);
}

}  // namespace

/* The record loop of load_multicolour_binary_from_filename_into_graph (file_reader.c:1686-1703:
 * hash_table_insert + add_edges + db_node_update_coverage per record, single-threaded) on the GPU,
 * for a caller that has parsed the header with the reference's own check_binary_signature_NEW and
 * hands over the file offset of the first record.  The caller's reference-layout table is merged in
 * first if it already holds k-mers, then replaced by the result.  Returns the number of records read,
 * or a negative error code. */
// ---------------------------------------------------------------- cleaning
// One pass of the cleaning on the device table: enumerate the paths (the supernode kernels),
// summarise their coverage, copy the compressed graph to the host, let cleaning_host.h take the
// reference's decisions in slot order, apply them on the device.
static int clean_pass(lasv_graph *g, int what, uint32_t threshold, lasv_clean_stats *st) {
  const TableView tv = g->view();
  SnCounters *d_c = nullptr, h_c;
  ClCounters *d_cc = nullptr, h_cc;
  uint64_t *d_starts = nullptr, *d_list = nullptr, *d_low = nullptr, *d_anodes = nullptr;
  unsigned long long *d_n = nullptr;
  uint8_t *d_covered = nullptr;
  SnPath *d_recs = nullptr, *d_cyc = nullptr;
  ClNodeRec *d_nodes = nullptr;
  uint32_t *d_max = nullptr, *d_scov = nullptr;
  ClPathPrune *d_apaths = nullptr;
  ClEdgeReset *d_aresets = nullptr;
  auto sync_counters = [&]() -> int {
    CUDA_TRY(cudaMemcpyAsync(&h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaMemcpyAsync(&h_cc, d_cc, sizeof h_cc, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    return LASV_OK;
  };
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_c, sizeof(SnCounters)));
    CUDA_TRY(cudaMalloc(&d_cc, sizeof(ClCounters)));
    CUDA_TRY(cudaMalloc(&d_n, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(d_c, 0, sizeof(SnCounters), g->compute));
    CUDA_TRY(cudaMemsetAsync(d_cc, 0, sizeof(ClCounters), g->compute));
    CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), g->compute));
    // vertices of the compressed graph: count, then list
    launch_sn_count(g->N, tv, d_c, g->compute);
    if (int e = check_launch()) return e;
    launch_cl_nodes(g->N, tv, nullptr, 0, d_cc, g->compute);
    if (int e = check_launch()) return e;
    if (int e = sync_counters()) return e;
    const uint64_t n_starts = h_c.n_starts, n_nodes = h_cc.n_nodes;
    CUDA_TRY(cudaMalloc(&d_nodes, std::max<uint64_t>(1, n_nodes) * sizeof(ClNodeRec)));
    CUDA_TRY(cudaMemsetAsync(&d_cc->n_nodes, 0, sizeof(unsigned long long), g->compute));
    launch_cl_nodes(g->N, tv, d_nodes, n_nodes, d_cc, g->compute);
    if (int e = check_launch()) return e;
    // its edges: the paths between them, and the cycles of interior nodes
    CUDA_TRY(cudaMalloc(&d_starts, std::max<uint64_t>(1, n_starts) * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(&d_recs, std::max<uint64_t>(1, n_starts) * sizeof(SnPath)));
    CUDA_TRY(cudaMalloc(&d_covered, g->n_slots));
    CUDA_TRY(cudaMemsetAsync(d_covered, 0, g->n_slots, g->compute));
    launch_sn_starts(g->N, tv, d_starts, n_starts, d_n, d_c, g->compute);
    if (int e = check_launch()) return e;
    if (int e = enumerate_paths(g, tv, d_starts, n_starts, h_c.n_interior, d_covered, d_recs, d_c)) return e;
    launch_sn_uncovered(g->N, tv, d_covered, nullptr, 0, d_c, g->compute);
    if (int e = check_launch()) return e;
    if (int e = sync_counters()) return e;
    const uint64_t n_unc = h_c.n_uncovered;
    if (n_unc) {
      CUDA_TRY(cudaMalloc(&d_list, n_unc * sizeof(uint64_t)));
      CUDA_TRY(cudaMalloc(&d_cyc, n_unc * sizeof(SnPath)));
      CUDA_TRY(cudaMemsetAsync(&d_c->n_uncovered, 0, sizeof(unsigned long long), g->compute));
      launch_sn_uncovered(g->N, tv, d_covered, d_list, n_unc, d_c, g->compute);
      if (int e = check_launch()) return e;
      launch_sn_cycles(g->N, tv, g->k, d_list, n_unc, d_cyc, n_unc, d_c, g->compute);
      if (int e = check_launch()) return e;
      if (int e = sync_counters()) return e;
    }
    if (h_c.n_overflow) return fail(LASV_ERR_STATE, "path record buffers overflowed (internal sizing error)");
    if (h_c.n_dangling)
      return fail(LASV_ERR_STATE, "%llu walks met an edge to a node that is not in the graph (the reference die()s)",
                  (unsigned long long)h_c.n_dangling);
    const uint64_t n_paths = h_c.n_paths, n_cycles = h_c.n_cycles, n_rec = n_paths + n_cycles;
    if (n_rec >= (1ull << 31) || n_nodes >= (1ull << 31)) return fail(LASV_ERR_CAPACITY, "more than 2^31 paths");
    // coverage summaries (the path records first, the cycle records behind them)
    CUDA_TRY(cudaMalloc(&d_max, std::max<uint64_t>(1, n_rec) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d_scov, std::max<uint64_t>(1, n_rec) * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&d_low, std::max<uint64_t>(1, n_rec) * sizeof(uint64_t)));
    launch_cl_coverage(g->N, tv, g->k, d_recs, n_paths, threshold, d_max, d_low, d_scov, d_cc, g->compute);
    if (n_paths)
      if (int e = check_launch()) return e;
    launch_cl_coverage(g->N, tv, g->k, d_cyc, n_cycles, threshold, d_max + n_paths, d_low + n_paths, d_scov + n_paths,
                       d_cc, g->compute);
    if (n_cycles)
      if (int e = check_launch()) return e;
    std::vector<ClNodeRec> nodes(n_nodes);
    std::vector<SnPath> recs(n_rec);
    std::vector<uint32_t> mx(n_rec), scov(n_rec);
    std::vector<uint64_t> low(n_rec);
    if (n_nodes) CUDA_TRY(cudaMemcpyAsync(nodes.data(), d_nodes, n_nodes * sizeof(ClNodeRec), cudaMemcpyDeviceToHost, g->compute));
    if (n_paths) CUDA_TRY(cudaMemcpyAsync(recs.data(), d_recs, n_paths * sizeof(SnPath), cudaMemcpyDeviceToHost, g->compute));
    if (n_cycles)
      CUDA_TRY(cudaMemcpyAsync(recs.data() + n_paths, d_cyc, n_cycles * sizeof(SnPath), cudaMemcpyDeviceToHost, g->compute));
    if (n_rec) {
      CUDA_TRY(cudaMemcpyAsync(mx.data(), d_max, n_rec * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaMemcpyAsync(scov.data(), d_scov, n_rec * sizeof(uint32_t), cudaMemcpyDeviceToHost, g->compute));
      CUDA_TRY(cudaMemcpyAsync(low.data(), d_low, n_rec * sizeof(uint64_t), cudaMemcpyDeviceToHost, g->compute));
    }
    if (int e = sync_counters()) return e;
    if (h_cc.n_failed) return fail(LASV_ERR_STATE, "coverage walks left the graph");
    cudaFree(d_starts); d_starts = nullptr;
    cudaFree(d_recs); d_recs = nullptr;
    cudaFree(d_cyc); d_cyc = nullptr;
    cudaFree(d_covered); d_covered = nullptr;

    // the reference's decisions, in slot order, on the compressed graph
    ClGraph cg;
    cg.nodes.reserve(n_nodes);
    for (const ClNodeRec &r : nodes) cg.add_node(r.q, r.covg, r.edges);
    cg.seal_nodes(g->n_slots);
    if (!cg.add_paths(recs.data(), mx.data(), low.data(), n_paths))
      return fail(LASV_ERR_STATE, "a path ends in a node that is not in the node list");
    for (uint64_t i = n_paths; i < n_rec; i++) {
      ClCycle c;
      c.s_q = recs[i].s_q;
      c.s_on = ((recs[i].bits >> SNB_S_N) & 3u) << 1;
      c.len = recs[i].len;
      c.s_covg = scov[i];
      c.max_covg = mx[i];
      c.min_low_q = low[i];
      if (c.s_covg > 0 && c.s_covg <= threshold && c.s_q < c.min_low_q) c.min_low_q = c.s_q;
      c.both_ways = (recs[i].bits & SNB_BOTH_WAYS) ? 1u : 0u;
      cg.cycles.push_back(c);
    }
    if (what == LASV_CLEAN_SUPERNODES && g->walk_limit > 0) {
      // the reference walks a supernode only from a node with 0 < coverage <= T, for at most walk_limit edges each way
      // (dB_graph_population.c:3371-3462): a longer path with such a node inside, or at one of its ends, is where it
      // would judge a PIECE of the supernode
      uint64_t cut = 0;
      auto low_node = [&](int32_t i) { return i >= 0 && cg.nodes[(size_t)i].covg > 0 && cg.nodes[(size_t)i].covg <= threshold; };
      for (const ClPath &pth : cg.paths)
        cut += pth.len > (uint32_t)g->walk_limit && (pth.min_low_q != SN_NONE || low_node(pth.a) || low_node(pth.b));
      for (const ClCycle &c : cg.cycles) cut += c.len > (uint32_t)g->walk_limit && c.min_low_q != SN_NONE;
      if (cut)
        return fail(LASV_ERR_LIMIT, "%llu supernodes with a low-coverage node have more than %d edges: the reference cuts "
                    "its walks there (--max_var_len)", (unsigned long long)cut, g->walk_limit);
    }
    ClActions act;
    if (what == LASV_CLEAN_TIPS) cg.clip_tips(g->k, act);
    else cg.remove_low_coverage_supernodes(g->k, threshold, act);

    // apply
    const uint64_t np = act.paths.size(), nn = act.nodes.size(), nr = act.resets.size();
    if (np) {
      CUDA_TRY(cudaMalloc(&d_apaths, np * sizeof(ClPathPrune)));
      CUDA_TRY(cudaMemcpyAsync(d_apaths, act.paths.data(), np * sizeof(ClPathPrune), cudaMemcpyHostToDevice, g->compute));
    }
    if (nn) {
      CUDA_TRY(cudaMalloc(&d_anodes, nn * sizeof(uint64_t)));
      CUDA_TRY(cudaMemcpyAsync(d_anodes, act.nodes.data(), nn * sizeof(uint64_t), cudaMemcpyHostToDevice, g->compute));
    }
    if (nr) {
      CUDA_TRY(cudaMalloc(&d_aresets, nr * sizeof(ClEdgeReset)));
      CUDA_TRY(cudaMemcpyAsync(d_aresets, act.resets.data(), nr * sizeof(ClEdgeReset), cudaMemcpyHostToDevice, g->compute));
    }
    launch_cl_apply(g->N, tv, g->k, d_apaths, np, d_anodes, nn, d_aresets, nr, d_cc, g->compute);
    if (np + nn + nr)
      if (int e = check_launch()) return e;
    if (int e = sync_counters()) return e;
    if (h_cc.n_failed) return fail(LASV_ERR_STATE, "pruning walks left the graph");
    if (st) {
      st->tips_clipped += act.tips;
      st->tip_nodes_pruned += act.tip_edges;
      st->supernodes_pruned += act.supernodes_pruned;
      st->nodes_pruned += act.nodes_pruned();
    }
    return LASV_OK;
  }();
  cudaFree(d_c); cudaFree(d_cc); cudaFree(d_n); cudaFree(d_starts); cudaFree(d_list); cudaFree(d_low); cudaFree(d_anodes);
  cudaFree(d_covered); cudaFree(d_recs); cudaFree(d_cyc); cudaFree(d_nodes); cudaFree(d_max); cudaFree(d_scov);
  cudaFree(d_apaths); cudaFree(d_aresets);
  return rc;
}

// `--health_check` (db_graph_health_check, dB_graph_population.c:6274-6350): every edge leads to a node of
// the graph that holds the return arrow.  Problems are counted, not fixed or printed.
int lasv_graph_health_check(lasv_graph *g, lasv_health_stats *out) {
  if (int e = use(g)) return e;
  if (!out) return fail(LASV_ERR_ARG, "out is NULL");
  CUDA_TRY(cudaDeviceSynchronize());
  HcCounters *d = nullptr, h;
  CUDA_TRY(cudaMalloc(&d, sizeof(HcCounters)));
  int rc = [&]() -> int {
    CUDA_TRY(cudaMemsetAsync(d, 0, sizeof(HcCounters), g->compute));
    launch_hc_check(g->N, g->view(), g->k, d, g->compute);
    if (int e = check_launch()) return e;
    CUDA_TRY(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    out->nodes_checked = h.nodes;
    out->edges_checked = h.edges;
    out->missing_target = h.missing_target;
    out->missing_return_arrow = h.missing_return;
    out->zero_kmers = h.zero_kmers;
    return LASV_OK;
  }();
  cudaFree(d);
  return rc;
}

int lasv_graph_clean(lasv_graph *g, int what, int threshold, lasv_clean_stats *stats) {
  if (int e = use(g)) return e;
  if (!(what & (LASV_CLEAN_TIPS | LASV_CLEAN_SUPERNODES)) || (what & ~(LASV_CLEAN_TIPS | LASV_CLEAN_SUPERNODES)))
    return fail(LASV_ERR_ARG, "what must be LASV_CLEAN_TIPS, LASV_CLEAN_SUPERNODES or both");
  if ((what & LASV_CLEAN_SUPERNODES) && threshold < 0) return fail(LASV_ERR_ARG, "negative coverage threshold");
  CUDA_TRY(cudaDeviceSynchronize());
  if (stats) memset(stats, 0, sizeof *stats);
  if (what & LASV_CLEAN_TIPS) {
    if (int e = clean_pass(g, LASV_CLEAN_TIPS, (uint32_t)std::max(threshold, 0), stats)) return e;
    g->cleaned_tips = true;
  }
  if (what & LASV_CLEAN_SUPERNODES) {
    if (int e = clean_pass(g, LASV_CLEAN_SUPERNODES, (uint32_t)threshold, stats)) return e;
    g->cleaned_sups = true;
    g->clean_threshold = threshold;
  }
  return LASV_OK;
}

// The caller's reference-layout table <-> a device table of the same geometry in the finalised arrangement
// (bucket b = W consecutive Elements at the start of its NL lines, slack bytes zero): the inverse of
// lasv_graph_export_table, slot order preserved, so a traversal in table order is the caller's.
static lasv_graph *graph_for_reference_table(const lasv_ref_hash_table *t) {
  if (!t || !t->table) { fail(LASV_ERR_ARG, "NULL table!"); return nullptr; }
  const int L = log2_exact(t->number_buckets);
  if (L < 0) { fail(LASV_ERR_ARG, "number_buckets is not a power of two"); return nullptr; }
  int dev = 0;
  if (const char *e = getenv("LASV_B200_DEVICE")) dev = atoi(e);
  return lasv_graph_new(t->kmer_size, L, t->bucket_size, t->max_rehash_tries, dev);
}

static int upload_reference_table(lasv_graph *g, const lasv_ref_hash_table *t) {
  const size_t row = (size_t)g->W * g->slot_bytes, pitch = (size_t)(g->table_bytes / g->n_buckets);
  const uint64_t rows_per_copy = 1ull << 20;
  for (uint64_t b0 = 0; b0 < g->n_buckets; b0 += rows_per_copy) {
    const uint64_t nb = std::min(rows_per_copy, g->n_buckets - b0);
    CUDA_TRY(cudaMemcpy2D(g->d_lines + b0 * pitch, pitch, (const uint8_t *)t->table + b0 * row, row, row, (size_t)nb,
                          cudaMemcpyHostToDevice));
  }
  g->finalized = true;
  // a copy from pageable host memory may return before its last DMA has landed, and the kernels run
  // on a non-blocking stream that does not wait for the legacy stream the copy was issued on
  CUDA_TRY(cudaDeviceSynchronize());
  return LASV_OK;
}

static int download_reference_table(lasv_graph *g, lasv_ref_hash_table *t) {
  CUDA_TRY(cudaDeviceSynchronize());
  const size_t row = (size_t)g->W * g->slot_bytes, pitch = (size_t)(g->table_bytes / g->n_buckets);
  const uint64_t rows_per_copy = 1ull << 20;
  for (uint64_t b0 = 0; b0 < g->n_buckets; b0 += rows_per_copy) {
    const uint64_t nb = std::min(rows_per_copy, g->n_buckets - b0);
    CUDA_TRY(cudaMemcpy2D((uint8_t *)t->table + b0 * row, row, g->d_lines + b0 * pitch, pitch, row, (size_t)nb,
                          cudaMemcpyDeviceToHost));
  }
  return LASV_OK;
}

int lasv_ref_hash_table_clean(lasv_ref_hash_table *t, int what, int threshold, lasv_clean_stats *stats) {
  return lasv_ref_hash_table_clean_limited(t, what, threshold, 0, stats);
}

int lasv_graph_clean_limited(lasv_graph *g, int what, int threshold, int max_length, lasv_clean_stats *stats) {
  if (!g) return fail(LASV_ERR_ARG, "NULL graph");
  g->walk_limit = max_length;
  const int rc = lasv_graph_clean(g, what, threshold, stats);
  g->walk_limit = 0;
  return rc;
}

int lasv_ref_hash_table_clean_limited(lasv_ref_hash_table *t, int what, int threshold, int max_length, lasv_clean_stats *stats) {
  if (!(what & (LASV_CLEAN_TIPS | LASV_CLEAN_SUPERNODES)) || (what & ~(LASV_CLEAN_TIPS | LASV_CLEAN_SUPERNODES)))
    return fail(LASV_ERR_ARG, "what must be LASV_CLEAN_TIPS, LASV_CLEAN_SUPERNODES or both");
  if ((what & LASV_CLEAN_SUPERNODES) && threshold < 0) return fail(LASV_ERR_ARG, "negative coverage threshold");
  lasv_graph *g = graph_for_reference_table(t);
  if (!g) return last_error_code();
  g->walk_limit = max_length;
  if (stats) memset(stats, 0, sizeof *stats);
  int rc = [&]() -> int {
    if (int e = upload_reference_table(g, t)) return e;
    if (what & LASV_CLEAN_TIPS)
      if (int e = clean_pass(g, LASV_CLEAN_TIPS, (uint32_t)std::max(threshold, 0), stats)) return e;
    if (what & LASV_CLEAN_SUPERNODES)
      if (int e = clean_pass(g, LASV_CLEAN_SUPERNODES, (uint32_t)threshold, stats)) return e;
    return download_reference_table(g, t);  // pruned nodes marked `pruned`, their edges reset
  }();
  lasv_graph_free(&g);
  return rc;
}

// ------------------------------------------------------------- branches
// `--mode build`'s ./branches.fa (print_branches.c:131-199): enumerate the paths between
// non-interior nodes on the device (the supernode kernels), replay the reference's traversal with
// its `visited` marks on those records on the host (branches_host.h, O(1) per branch), write the
// records on the device (br_write_kernel), stream the image to the file.
int lasv_graph_write_branches(lasv_graph *g, const char *fasta_path, int mark_visited, lasv_branch_stats *stats) {
  if (int e = use(g)) return e;
  if (!fasta_path) return fail(LASV_ERR_ARG, "fasta_path is NULL");
  CUDA_TRY(cudaDeviceSynchronize());
  const TableView tv = g->view();
  SnCounters *d_c = nullptr, h_c;
  ClCounters *d_cc = nullptr, h_cc;
  uint64_t *d_starts = nullptr;
  unsigned long long *d_n = nullptr;
  SnPath *d_recs = nullptr;
  ClNodeRec *d_nodes = nullptr;
  BrPrint *d_prints = nullptr;
  char *d_image = nullptr, *h_stage[2] = {nullptr, nullptr};
  cudaEvent_t copied[2] = {nullptr, nullptr};
  FILE *fp = nullptr;
  constexpr uint64_t STAGE_BYTES = 16ull << 20;
  BrStats bs;
  auto sync_counters = [&]() -> int {
    CUDA_TRY(cudaMemcpyAsync(&h_c, d_c, sizeof h_c, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaMemcpyAsync(&h_cc, d_cc, sizeof h_cc, cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    return LASV_OK;
  };
  int rc = [&]() -> int {
    CUDA_TRY(cudaMalloc(&d_c, sizeof(SnCounters)));
    CUDA_TRY(cudaMalloc(&d_cc, sizeof(ClCounters)));
    CUDA_TRY(cudaMalloc(&d_n, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(d_c, 0, sizeof(SnCounters), g->compute));
    CUDA_TRY(cudaMemsetAsync(d_cc, 0, sizeof(ClCounters), g->compute));
    CUDA_TRY(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), g->compute));
    if (mark_visited) {  // print_branches.c:133
      launch_br_reset_marks(g->N, tv, g->compute);
      if (int e = check_launch()) return e;
    }
    // the compressed graph: non-interior nodes, and the paths between them (cycles of interior
    // nodes are out of reach of any branching node and are not enumerated)
    launch_sn_count(g->N, tv, d_c, g->compute);
    if (int e = check_launch()) return e;
    launch_cl_nodes(g->N, tv, nullptr, 0, d_cc, g->compute);
    if (int e = check_launch()) return e;
    if (int e = sync_counters()) return e;
    const uint64_t n_starts = h_c.n_starts, n_nodes = h_cc.n_nodes;
    CUDA_TRY(cudaMalloc(&d_nodes, std::max<uint64_t>(1, n_nodes) * sizeof(ClNodeRec)));
    CUDA_TRY(cudaMemsetAsync(&d_cc->n_nodes, 0, sizeof(unsigned long long), g->compute));
    launch_cl_nodes(g->N, tv, d_nodes, n_nodes, d_cc, g->compute);
    if (int e = check_launch()) return e;
    CUDA_TRY(cudaMalloc(&d_starts, std::max<uint64_t>(1, n_starts) * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(&d_recs, std::max<uint64_t>(1, n_starts) * sizeof(SnPath)));
    launch_sn_starts(g->N, tv, d_starts, n_starts, d_n, d_c, g->compute);
    if (int e = check_launch()) return e;
    if (int e = enumerate_paths(g, tv, d_starts, n_starts, h_c.n_interior, nullptr, d_recs, d_c)) return e;
    if (int e = sync_counters()) return e;
    if (h_c.n_overflow) return fail(LASV_ERR_STATE, "path record buffers overflowed (internal sizing error)");
    if (h_c.n_dangling)
      return fail(LASV_ERR_STATE, "%llu walks met an edge to a node that is not in the graph (the reference die()s)",
                  (unsigned long long)h_c.n_dangling);
    const uint64_t n_paths = h_c.n_paths;
    if (n_paths >= (1ull << 31) || n_nodes >= (1ull << 31)) return fail(LASV_ERR_CAPACITY, "more than 2^31 paths");
    std::vector<ClNodeRec> nodes(n_nodes);
    std::vector<SnPath> recs(n_paths);
    if (n_nodes) CUDA_TRY(cudaMemcpyAsync(nodes.data(), d_nodes, n_nodes * sizeof(ClNodeRec), cudaMemcpyDeviceToHost, g->compute));
    if (n_paths) CUDA_TRY(cudaMemcpyAsync(recs.data(), d_recs, n_paths * sizeof(SnPath), cudaMemcpyDeviceToHost, g->compute));
    CUDA_TRY(cudaStreamSynchronize(g->compute));
    cudaFree(d_starts); d_starts = nullptr;
    cudaFree(d_recs); d_recs = nullptr;
    cudaFree(d_nodes); d_nodes = nullptr;

    // the reference's traversal, in slot order, on the compressed graph
    ClGraph cg;
    cg.nodes.reserve(n_nodes);
    for (const ClNodeRec &r : nodes) cg.add_node(r.q, r.covg, r.edges);
    cg.seal_nodes(g->n_slots);
    if (!cg.add_paths(recs.data(), nullptr, nullptr, n_paths))
      return fail(LASV_ERR_STATE, "a path ends in a node that is not in the node list");
    std::vector<BrPrint> prints;
    BrReplay replay(cg, g->k);
    bs = replay.run(prints);
    if (bs.inconsistent) return fail(LASV_ERR_STATE, "%llu edges of branching nodes have no path record", (unsigned long long)bs.inconsistent);

    fp = fopen(fasta_path, "w");
    if (!fp) return fail(LASV_ERR_IO, "cannot open %s for writing: %s", fasta_path, strerror(errno));
    if (!prints.empty()) {
      const uint64_t image_bytes = bs.image_bytes;
      CUDA_TRY(cudaMalloc(&d_prints, prints.size() * sizeof(BrPrint)));
      CUDA_TRY(cudaMalloc(&d_image, image_bytes + 8));
      CUDA_TRY(cudaMemcpyAsync(d_prints, prints.data(), prints.size() * sizeof(BrPrint), cudaMemcpyHostToDevice, g->compute));
      launch_br_write(g->N, tv, g->k, d_prints, prints.size(), d_image, mark_visited ? 1 : 0, d_cc, g->compute);
      if (int e = check_launch()) return e;
      if (int e = sync_counters()) return e;
      if (h_cc.n_failed) return fail(LASV_ERR_STATE, "branch sequence walk left the graph");
      for (int b = 0; b < 2; b++) {
        CUDA_TRY(cudaMallocHost(&h_stage[b], STAGE_BYTES));
        CUDA_TRY(cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming));
      }
      // image -> file through two pinned staging buffers: the copy of chunk i+1 runs under the fwrite of chunk i
      const uint64_t n_chunks = (image_bytes + STAGE_BYTES - 1) / STAGE_BYTES;
      auto bytes_of = [&](uint64_t c) { return std::min(STAGE_BYTES, image_bytes - c * STAGE_BYTES); };
      for (uint64_t c = 0; c <= n_chunks; c++) {
        if (c < n_chunks) {
          CUDA_TRY(cudaMemcpyAsync(h_stage[c & 1], d_image + c * STAGE_BYTES, bytes_of(c), cudaMemcpyDeviceToHost, g->compute));
          CUDA_TRY(cudaEventRecord(copied[c & 1], g->compute));
        }
        if (c > 0) {
          CUDA_TRY(cudaEventSynchronize(copied[(c - 1) & 1]));
          if (fwrite(h_stage[(c - 1) & 1], 1, bytes_of(c - 1), fp) != bytes_of(c - 1))
            return fail(LASV_ERR_IO, "short write to %s", fasta_path);
        }
      }
    }
    return LASV_OK;
  }();
  cudaFree(d_c); cudaFree(d_cc); cudaFree(d_n); cudaFree(d_starts); cudaFree(d_recs); cudaFree(d_nodes);
  cudaFree(d_prints); cudaFree(d_image);
  for (int b = 0; b < 2; b++) {
    if (h_stage[b]) cudaFreeHost(h_stage[b]);
    if (copied[b]) cudaEventDestroy(copied[b]);
  }
  if (fp && fclose(fp) != 0 && rc == LASV_OK) rc = fail(LASV_ERR_IO, "cannot close %s: %s", fasta_path, strerror(errno));
  if (rc != LASV_OK) return rc;
  if (stats) {
    stats->n_branches = bs.branches;
    stats->n_branching_nodes = bs.branching_nodes;
    stats->n_nodes = bs.nodes;
    stats->n_refused = bs.refused;
    stats->n_cut = bs.cut;
    stats->bytes = bs.image_bytes;
  }
  return LASV_OK;
}

int lasv_ref_hash_table_write_branches(lasv_ref_hash_table *t, const char *fasta_path, lasv_branch_stats *stats) {
  lasv_graph *g = graph_for_reference_table(t);
  if (!g) return last_error_code();
  int rc = [&]() -> int {
    if (int e = upload_reference_table(g, t)) return e;
    if (int e = lasv_graph_write_branches(g, fasta_path, 1, stats)) return e;
    return download_reference_table(g, t);  // the marks
  }();
  lasv_graph_free(&g);
  return rc;
}

static int ref_write_supernodes(lasv_ref_hash_table *t, const char *fasta_path, int max_length,
                                lasv_supernode_stats *stats, bool strict) {
  lasv_graph *g = graph_for_reference_table(t);
  if (!g) return last_error_code();
  g->sn_strict = strict;
  int rc = upload_reference_table(g, t);
  if (rc == LASV_OK) rc = lasv_graph_write_supernodes(g, fasta_path, max_length, stats);
  lasv_graph_free(&g);
  return rc;
}

int lasv_ref_hash_table_write_supernodes(lasv_ref_hash_table *t, const char *fasta_path, int max_length,
                                         lasv_supernode_stats *stats) {
  return ref_write_supernodes(t, fasta_path, max_length, stats, false);
}

int lasv_ref_hash_table_write_supernodes_strict(lasv_ref_hash_table *t, const char *fasta_path, int max_length,
                                                lasv_supernode_stats *stats) {
  return ref_write_supernodes(t, fasta_path, max_length, stats, true);
}

int lasv_graph_write_supernodes_strict(lasv_graph *g, const char *fasta_path, int max_length, lasv_supernode_stats *stats) {
  if (!g) return fail(LASV_ERR_ARG, "NULL graph");
  g->sn_strict = true;
  const int rc = lasv_graph_write_supernodes(g, fasta_path, max_length, stats);
  g->sn_strict = false;
  return rc;
}

long long lasv_ref_hash_table_load_ctx_records(lasv_ref_hash_table *t, const char *path, long record_offset) {
  if (!t || !t->table) return fail(LASV_ERR_ARG, "NULL table!");
  const int L = log2_exact(t->number_buckets);
  if (L < 0) return fail(LASV_ERR_ARG, "number_buckets is not a power of two");
  FILE *fp = fopen(path, "rb");
  if (!fp) return fail(LASV_ERR_IO, "load_multicolour_binary_from_filename_into_graph cannot open file:%s", path);
  if (fseek(fp, record_offset, SEEK_SET)) {
    fclose(fp);
    return fail(LASV_ERR_IO, "cannot seek to the records of %s", path);
  }
  int dev = 0;
  if (const char *e = getenv("LASV_B200_DEVICE")) dev = atoi(e);
  lasv_graph *g = lasv_graph_new(t->kmer_size, L, t->bucket_size, t->max_rehash_tries, dev);
  if (!g) {
    fclose(fp);
    return LASV_ERR_CUDA;
  }
  unsigned long long n = 0;
  int rc = LASV_OK;
  long long coll0[16] = {0};  // rehashes of the import itself: the caller's collisions[] counted those when they happened
  if (t->unique_kmers > 0) {
    rc = lasv_graph_import_table(g, t->table);
    if (rc == LASV_OK) rc = lasv_graph_rehash_histogram(g, coll0);
  }
  if (rc == LASV_OK) rc = stream_ctx_records(g, fp, &n);
  fclose(fp);
  long long coll[16];
  if (rc == LASV_OK) rc = lasv_graph_rehash_histogram(g, coll);
  if (rc == LASV_OK)
    for (int r = 1; r < 16; r++) coll[r] -= coll0[r];
  const long long uniq = rc == LASV_OK ? lasv_graph_unique_kmers(g) : -1;
  if (rc == LASV_OK) rc = lasv_graph_export_table(g, t->table, t->next_element);
  if (rc == LASV_OK) {
    t->unique_kmers = uniq;
    // one hash_table_insert per record (hash_table.c:333-382 counts collisions[] like find_or_insert)
    if (t->collisions) {
      long long later = 0;
      for (int r = 1; r < t->max_rehash_tries && r < 16; r++) {
        t->collisions[r] += coll[r];
        later += coll[r];
      }
      t->collisions[0] += (long long)n - later;
    }
  }
  lasv_graph_free(&g);
  return rc == LASV_OK ? (long long)n : rc;
}

void load_se_filelist_into_graph_colour(
    char *se_filelist_path, int qual_thresh, int homopol_limit, signed char remove_dups_se,
    char ascii_fq_offset, int colour, lasv_ref_hash_table *db_graph, char is_colour_list,
    unsigned int *total_files_loaded, unsigned long long *total_bad_reads,
    unsigned long long *total_dup_reads, unsigned long long *total_bases_read,
    unsigned long long *total_bases_loaded, unsigned long *readlen_count_array,
    unsigned long readlen_count_array_size) {
  (void)colour;
  dropin_load(db_graph, se_filelist_path, nullptr, qual_thresh, homopol_limit, remove_dups_se,
              ascii_fq_offset, is_colour_list, total_files_loaded, total_bad_reads, total_dup_reads,
              total_bases_read, total_bases_loaded, readlen_count_array, readlen_count_array_size);
}

void load_pe_filelists_into_graph_colour(
    char *pe_filelist_path1, char *pe_filelist_path2, int qual_thresh, int homopol_limit,
    signed char remove_dups_pe, char ascii_fq_offset, int colour, lasv_ref_hash_table *db_graph,
    char is_colour_lists, unsigned int *total_file_pairs_loaded, unsigned long long *total_bad_reads,
    unsigned long long *total_dup_reads, unsigned long long *total_bases_read,
    unsigned long long *total_bases_loaded, unsigned long *readlen_count_array,
    unsigned long readlen_count_array_size) {
  (void)colour;
  dropin_load(db_graph, pe_filelist_path1, pe_filelist_path2, qual_thresh, homopol_limit,
              remove_dups_pe, ascii_fq_offset, is_colour_lists, total_file_pairs_loaded,
              total_bad_reads, total_dup_reads, total_bases_read, total_bases_loaded,
              readlen_count_array, readlen_count_array_size);
}

// ------------------------------------------------- bench / test utilities
static SynthParams to_synth(const lasv_synth_params *p) {
  SynthParams s;
  s.seed = p->seed; s.genome_len = p->genome_len; s.read_len = p->read_len; s.insert = p->insert;
  s.err_q20 = p->err_q20; s.n_q20 = p->n_q20; s.lowq_q20 = p->lowq_q20;
  s.err_lowq_pct = p->err_lowq_pct; s.tail_min = p->tail_min; s.tail_max = p->tail_max;
  s.tiling = p->tiling; s.tiling_step = p->tiling_step;
  return s;
}

static int check_synth(const lasv_synth_params *p) {
  if (!p || p->read_len < 1 || p->read_len > 1023) return fail(LASV_ERR_ARG, "read_len must be in [1,1023]");
  if (!p->tiling && (p->insert < p->read_len || p->genome_len < p->insert))
    return fail(LASV_ERR_ARG, "need read_len <= insert <= genome_len");
  if (p->tail_max < p->tail_min) return fail(LASV_ERR_ARG, "tail_max < tail_min");
  if (p->tiling && !p->tiling_step) return fail(LASV_ERR_ARG, "tiling_step must be > 0");
  return LASV_OK;
}

int lasv_synth_reads_device(const lasv_synth_params *p, uint64_t first_read, uint64_t n_reads,
                            uint8_t *d_seq, uint8_t *d_qual, void *cuda_stream) {
  if (int e = check_synth(p)) return e;
  launch_synth(to_synth(p), first_read, n_reads, d_seq, d_qual, (cudaStream_t)cuda_stream);
  return check_launch();
}

int lasv_synth_reads_host(const lasv_synth_params *p, uint64_t first_read, uint64_t n_reads,
                          uint8_t *seq, uint8_t *qual) {
  if (int e = check_synth(p)) return e;
  synth_host(to_synth(p), first_read, n_reads, seq, qual);
  return LASV_OK;
}

/* The parallel gzip inflate on its own (no GPU; gz_parallel.h): the text of a gzip file inflated by `threads` threads
 * in chunks of chunk_kb compressed KB.  out4 = {bytes of text, CRC-32 of the text, chunks whose guessed start was
 * confirmed, chunks dropped}.  LASV_ERR_IO if the file is not gzip or is corrupt (lasv_last_error says which). */
int lasv_seqfile_inflate_check(const char *path, int threads, uint64_t chunk_kb, uint64_t *out4) {
  if (!path || !out4) return fail(LASV_ERR_ARG, "NULL argument");
  PgzStream s;
  if (!s.open(path, threads, 0, (size_t)chunk_kb << 10)) return fail(LASV_ERR_IO, "not a gzip file of at least 64 bytes: %s", path);
  std::vector<unsigned char> w(9u << 20);
  std::string err;
  uint64_t total = 0;
  uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
  size_t got = 0;
  while (s.next_into(w.data(), w.size(), &got, err)) {
    crc = (uint32_t)crc32(crc, w.data(), (uInt)got);  // (zlib's own: the check does not lean on fast_crc32)
    total += got;
  }
  if (!err.empty()) return fail(LASV_ERR_IO, "%s [file: %s]", err.c_str(), path);
  out4[0] = total;
  out4[1] = crc;
  out4[2] = s.chunks_confirmed();
  out4[3] = s.chunks_dropped();
  return LASV_OK;
}

/* The file loaders' parser on its own (no GPU): reads, bases and an order-independent checksum of all
 * (sequence, quality) pairs of one or two files, parsed with `threads` parser threads per file exactly as
 * the loaders do (threads <= 1: one sequential parser per file).  out4 = {reads of file 1, reads of file
 * 2, bases, checksum}; producers_used reports how many parser threads ran. */
static int parse_check_impl(const char *path1, const char *path2, int threads, uint64_t batch_bytes, uint64_t *out4,
                            int *producers_used, bool raw_slices, uint64_t *raw_batches);

int lasv_seqfile_parse_check(const char *path1, const char *path2, int threads, uint64_t batch_bytes,
                             uint64_t *out4, int *producers_used) {
  return parse_check_impl(path1, path2, threads, batch_bytes, out4, producers_used, false, nullptr);
}

/* The same with the producers set up as the file loaders set them up for the DEVICE parse: batches of FASTQ text
 * (slices of mapped plain files, windows of inflated BGZF / gzip text cut at record boundaries, tails carried over).
 * Here the text batches are taken apart by the host parser, so that the producers' side of that path can be checked
 * without a GPU; *raw_batches counts them. */
int lasv_seqfile_parse_check_text(const char *path1, const char *path2, int threads, uint64_t batch_bytes,
                                  uint64_t *out4, int *producers_used, uint64_t *raw_batches) {
  return parse_check_impl(path1, path2, threads, batch_bytes, out4, producers_used, true, raw_batches);
}

static int parse_check_impl(const char *path1, const char *path2, int threads, uint64_t batch_bytes, uint64_t *out4,
                            int *producers_used, bool raw_slices, uint64_t *raw_batches) {
  if (!path1) return fail(LASV_ERR_ARG, "NULL argument");
  const bool checksum = out4 != nullptr;  // out4 == NULL: parse only (parser timing)
  if (batch_bytes < 1024) batch_bytes = 1024;
  struct PlainBatch : HostBatch {
    std::vector<uint8_t> s, q;
    std::vector<uint64_t> o;
    PlainBatch(uint64_t bytes, uint64_t reads) : s(bytes), q(bytes), o(reads + 1) {
      seq = s.data(); qual = q.data(); offsets = o.data();
      cap_bytes = bytes; cap_reads = reads;
      reset();
    }
  };
  std::vector<std::unique_ptr<PlainBatch>> batches;
  const auto t_open0 = std::chrono::steady_clock::now();
  ParseJob job;
  if (!job.open({{std::string(path1), std::string(path2 ? path2 : "")}}, threads < 1 ? 1 : threads, 1u << 20,
                [&]() -> HostBatch * {
                  batches.emplace_back(new PlainBatch(batch_bytes, batch_bytes / 16 + 16));
                  return batches.back().get();
                }, raw_slices))
    return fail(LASV_ERR_IO, "%s", job.err.c_str());
  uint64_t bases = 0, sum = 0, n_raw = 0;
  std::string raw_err;
  const auto t_run0 = std::chrono::steady_clock::now();
  if (getenv("LASV_B200_TRACE_PARSE"))
    fprintf(stderr, "parse_check: open %.3f s\n", std::chrono::duration<double>(t_run0 - t_open0).count());
  job.run([&](HostBatch &b, FileProducer &fp) {
    const bool wq = fp.with_qual;
    if (b.raw) {  // a batch of text: whole records (what fastq_parse.cu takes apart on the device)
      n_raw++;
      if (!checksum) {  // (producer timing)
        bases += b.n_bytes;
        return;
      }
      SeqReader mr;
      mr.open_memory(b.seq, b.n_bytes, fp.reader.path());
      std::string e2;
      uint64_t n = 0;
      while (mr.next(e2)) {
        uint64_t h = 0xCBF29CE484222325ull;
        for (unsigned char c : mr.seq()) h = (h ^ c) * 0x100000001B3ull;
        if (wq) for (unsigned char c : mr.qual()) h = (h ^ c) * 0x100000001B3ull;
        sum += mix64(h ^ (uint64_t)mr.seq().size());
        bases += mr.seq().size();
        n++;
      }
      if (!e2.empty() && raw_err.empty()) raw_err = e2;
      fp.reads += n;
      return;
    }
    if (!checksum) { bases += b.n_bytes - b.n_reads; return; }
    for (uint64_t r = 0; r < b.n_reads; r++) {
      const uint64_t lo = b.offsets[r], hi = b.offsets[r + 1] - 1;  // without the separator
      uint64_t h = 0xCBF29CE484222325ull;
      for (uint64_t i = lo; i < hi; i++) h = (h ^ b.seq[i]) * 0x100000001B3ull;
      if (wq) for (uint64_t i = lo; i < hi; i++) h = (h ^ b.qual[i]) * 0x100000001B3ull;
      sum += mix64(h ^ (hi - lo));
      bases += hi - lo;
    }
  });
  if (getenv("LASV_B200_TRACE_PARSE"))
    fprintf(stderr, "parse_check: run %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run0).count());
  for (auto &p : job.prod)
    if (!p->err.empty()) return fail(LASV_ERR_IO, "%s", p->err.c_str());
  if (!raw_err.empty()) return fail(LASV_ERR_IO, "%s", raw_err.c_str());
  if (raw_batches) *raw_batches = n_raw;
  if (out4) {
    out4[0] = job.reads_of_file(0);
    out4[1] = path2 ? job.reads_of_file(1) : 0;
    out4[2] = bases;
    out4[3] = sum;
  }
  if (producers_used) *producers_used = (int)job.prod.size();
  return LASV_OK;
}

long long lasv_seqfile_to_streams(const char *path, uint8_t *seq, uint8_t *qual, uint64_t cap_bytes,
                                  uint64_t *offsets, uint64_t cap_reads, int *has_qual) {
  std::string err;
  SeqReader r;
  if (!r.open(path, err)) return fail(LASV_ERR_IO, "%s", err.c_str());
  HostBatch b;
  b.seq = seq; b.qual = qual; b.offsets = offsets;
  b.cap_bytes = cap_bytes; b.cap_reads = cap_reads;
  b.reset();
  if (has_qual) *has_qual = r.has_quality() ? 1 : 0;
  while (r.next(err)) {
    if (!r.append_to(b, r.has_quality()))
      return fail(LASV_ERR_CAPACITY, "output buffers too small for %s", path);
  }
  if (!err.empty()) return fail(LASV_ERR_IO, "%s", err.c_str());
  return (long long)b.n_reads;
}

int lasv_random_access_probe(void *d_buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed,
                             void *cuda_stream) {
  return lasv_random_access_probe_mode(d_buf, n_sectors, n_ops, seed, 0, cuda_stream);
}

int lasv_random_access_probe_mode(void *d_buf, uint64_t n_sectors, uint64_t n_ops, uint64_t seed,
                                  int mode, void *cuda_stream) {
  if (!d_buf || !n_sectors || mode < 0 || mode > 3) return fail(LASV_ERR_ARG, "bad probe arguments");
  launch_random_rmw((uint64_t *)d_buf, n_sectors, n_ops, seed, mode, (cudaStream_t)cuda_stream);
  return check_launch();
}

int lasv_graph_record_buckets_device(lasv_graph *g, const uint32_t *d_records, uint64_t n_records,
                                     uint32_t *d_buckets, void *cuda_stream) {
  if (int e = use(g)) return e;
  launch_record_buckets(g->N, d_records, n_records, (uint32_t)(g->n_buckets - 1), d_buckets,
                        (cudaStream_t)cuda_stream);
  return check_launch();
}

}  // extern "C"
