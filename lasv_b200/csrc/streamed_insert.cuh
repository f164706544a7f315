// streamed_insert.cuh -- K3 as a STREAM over the table instead of one random DRAM access per k-mer
// (replaces insert_bins_kernel whenever a batch puts about one record or more on every 128-byte line).
//
// Reference behaviour reproduced: hash_table_find_or_insert / hash_table_find_in_bucket
// (src/hash_table/open_hash/hash_table.c:69-122,270-327), db_node_update_coverage (element.c:395-417) and the
// edge OR (element.c:223-228) -- same buckets, same rehash chain (a key leaves its bucket only when the bucket
// holds W keys; 16 full buckets are fatal), same multiset (key, coverage, edges).
//
// Scheme.  The partition kernel has already grouped the batch's records by table REGION (<= 64 MB, one stripe per
// region and sender).  Here, SLAB by slab (a run of regions whose groups' write frontiers fit in L2):
//   count        one pass over the slab's stripes: records per BUCKET
//   scan         exclusive prefix over the slab's buckets (cub)
//   scatter      second pass: every record to its bucket's run in the slab's sorted buffer.  Both passes rank a
//                chunk of one region's records in shared memory first ("staged"), so that the global atomics and
//                the stores are per (chunk, bucket) run, not per record
//   sg_insert    one CTA per GROUP (a few neighbouring buckets, 16-26 KB of lines), persistent: the group's lines
//                arrive in shared memory by cp.async.bulk (TMA,
//                double buffered behind two mbarriers), its records are applied there with the table protocol of
//                table.cuh on shared-memory atomics (bucket_upsert<N, SmemOps>), and the lines stream back with a
//                bulk store -- DRAM sees two sequential streams at full line width and nothing else
//   sg_drain     records whose bucket held W other keys leave the group (their rehash chain leads anywhere): they
//                were put on a spill list and now go through the ordinary table_upsert, after ALL groups of the
//                slab are back in the table
// Slot positions inside a bucket depend on the order of the claims, as they do on the random-access path and in
// the reference (insertion order); the multiset and the rehash-chain placement do not.
#pragma once
#include <functional>

#include "graph_kernels.cuh"

namespace lasv {

struct StreamedGeom {
  uint32_t unit_shift;        // buckets per unit (what one warp stages at a time) = 1 << unit_shift
  uint32_t unit_bytes;        // (NL * 128) << unit_shift
  uint32_t rshift;            // log2(buckets per region)
  uint32_t regions_per_slab;  // power of two
  uint32_t buckets_per_slab;
  uint32_t n_slabs;
  uint32_t fused;             // regions small enough for the one-CTA-per-region sort
  uint64_t slab_cap;          // records a slab can hold at most (its stripes' capacities)
};

// device scratch of the streamed insert, owned by the graph and grown on demand
struct StreamedScratch {
  uint32_t *counts = nullptr;   // [buckets_per_slab + 1] records per bucket (direct sort) / per-region prefix (fused sort)
  uint32_t *offsets = nullptr;  // [buckets_per_slab + 1] exclusive prefix; the last entry = records of the slab
  uint32_t *cursor = nullptr;   // [buckets_per_slab]
  void *scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  uint64_t *sorted = nullptr;   // [slab_cap * rec_words]
  uint32_t *spill = nullptr;    // [slab_cap] indices into `sorted`
  unsigned long long *n_spill = nullptr;
  uint64_t alloc_records = 0;
  uint32_t alloc_buckets = 0;
};

// The largest group that still lets two staging buffers of four CTAs share an SM; 0 if even one bucket is too big
// for double buffering (the streamed insert then does not apply).
bool streamed_geometry(const TableView &t, const BinView &b, StreamedGeom *out);
size_t streamed_scan_tmp_bytes(uint32_t n);
// The whole insert of one partitioned batch (all senders, all regions, then the senders' spill stripes).
// Everything is queued on `st`; nothing synchronises with the host.  `overflow` is raised if a record sits in
// a stripe of another region (never, by construction).  `span(kind, begin)`, if set, brackets the launches of
// one kind (LASV_T_SORT / _STREAM / _DRAIN of include/lasv_b200.h) so that the caller can time them.
// Returns 0, or -1 if a kernel attribute could not be set (no launch was queued then).
int launch_streamed_insert(int N, const BinView &b, const TableView &t, GraphCounters *ctr, unsigned long long *overflow,
                           const StreamedGeom &geom, const StreamedScratch &s, cudaStream_t st, uint64_t *launches,
                           const std::function<void(int, int)> &span);

}  // namespace lasv
