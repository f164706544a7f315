// groups.cuh -- PROTOTYPE of the streamed insert planned in DESIGN section 9 (not on the default path; nothing
// in the loaders or in bench.py calls it): instead of one random DRAM access per k-mer, the table is updated
// GROUP by group -- a few neighbouring buckets, tens of KB -- in shared memory:
//
//   records sorted by the group of their first-choice bucket (lookup3 & mask) / buckets_per_group
//   one CTA per group:  lines of the group  global -> shared   (streamed, full 128-byte lines)
//                       its records applied there with the protocol of table.cuh (bucket_upsert<N, SmemOps>:
//                       same tag bytes, same CAS claim, shared-memory atomics)
//                       lines  shared -> global
//   a record whose bucket is full leaves the group (its rehash chain leads anywhere): it is put on a spill
//   list and goes through the ordinary table_upsert after ALL groups have been written back.
//
// Parity: slot positions inside a bucket depend on the order of the claims, as they do on the random-access
// path; the multiset (key, coverage, edges) and the rehash-chain placement (a key leaves its bucket only when
// the bucket holds W keys) do not.  The find-or-insert of the reference (hash_table.c:270-327) is unchanged.
//
// The per-group logic is host/device code so that tests/host_emul runs it on the CPU; the kernels are in
// group_kernels.cu, the entry point is lasv_graph_insert_grouped_records_device (capi.cu).
#pragma once
#include "table.cuh"

namespace lasv {

// (2N+1)-word record of the records path: key words as (lo, hi) pairs, then {edges : 8, count : 24}
template <int N>
LASV_HD void group_load_record(const uint32_t *src, uint64_t (&key)[N], uint32_t &edge, uint32_t &count) {
#pragma unroll
  for (int i = 0; i < N; i++) key[i] = (uint64_t)src[2 * i] | ((uint64_t)src[2 * i + 1] << 32);
  edge = src[2 * N] & 0xFFu;
  count = src[2 * N] >> 8;
}

LASV_HD uint64_t group_bytes(const TableView &t, uint32_t buckets_per_group) {
  return (uint64_t)buckets_per_group * t.NL * LINE_BYTES;
}

// The table as seen from a staged copy of group g: bucket b of the group lives at
// staged + (b - g * buckets_per_group) * NL * 128, every other bucket is out of bounds.
LASV_HD TableView group_local_view(const TableView &t, uint8_t *staged, uint64_t g, uint32_t buckets_per_group) {
  TableView l = t;
  l.lines = staged - g * group_bytes(t, buckets_per_group);
  return l;
}

// One record against the staged group: 1 = new key, 0 = found, -1 = its bucket holds W other keys.
#pragma nv_exec_check_disable
template <int N, class Ops>
LASV_HD int group_upsert(const TableView &local, GraphCounters *ctr, const uint32_t *rec) {
  uint64_t key[N];
  uint32_t edge, count;
  group_load_record<N>(rec, key, edge, count);
  const Probe pr = probe_of_key<N>(key);
  return bucket_upsert<N, Ops>(local, ctr, lookup3_key<N>(key, 0).c & local.bucket_mask, pr, key, edge, count);
}

// The group of a record (what the caller sorts by).
template <int N>
LASV_HD uint64_t group_of_record(const TableView &t, const uint32_t *rec, uint32_t buckets_per_group) {
  uint64_t key[N];
  uint32_t edge, count;
  group_load_record<N>(rec, key, edge, count);
  return (uint64_t)(lookup3_key<N>(key, 0).c & t.bucket_mask) / buckets_per_group;
}

#if defined(__CUDACC__)
// launch interface (group_kernels.cu).  recs: n records sorted by group; group_offsets[g] .. [g+1]: the records of
// group g (n_groups + 1 entries); spill: room for spill_cap record indices, *n_spill counts what wanted in.
void launch_insert_groups(int N, const uint32_t *recs, const uint64_t *group_offsets, uint64_t n_groups,
                          uint32_t buckets_per_group, const TableView &t, GraphCounters *ctr, uint32_t *spill,
                          unsigned long long *n_spill, uint64_t spill_cap, cudaStream_t st);
void launch_insert_spilled(int N, const uint32_t *recs, const uint32_t *spill, uint64_t n_spill, const TableView &t,
                           GraphCounters *ctr, cudaStream_t st);
size_t insert_groups_smem_limit();
// The caller's sort, in the library: a counting sort of the records by group (histogram, exclusive scan, scatter).
// counts: n_groups uint32 of scratch; offsets: n_groups + 1 entries out; sorted: room for n records out.
void launch_group_records(int N, const uint32_t *recs, uint64_t n, uint32_t buckets_per_group, const TableView &t,
                          uint32_t *counts, uint64_t *offsets, void *scan_tmp, size_t scan_tmp_bytes, uint32_t *sorted,
                          cudaStream_t st);
#endif

}  // namespace lasv
