// supernode_kernels.cu -- sm_100a kernels of the supernode (maximal unambiguous
// contig) extraction; the arithmetic and the reference lines it restates are in
// supernodes.cuh, the sequential replay in supernodes_host.h.
//
//   sn_count_kernel      per slot: occupied / interior / number of path starts     (streams the meta words)
//   sn_starts_kernel     per slot: append (slot, orientation, nucleotide) of every
//                        edge of a non-interior node to the start list
//   sn_paths_kernel      one THREAD per start: walk to the next non-interior node
//                        (one hash_table_find per step, a dependent chain of random
//                        DRAM reads), mark the interior nodes crossed, keep the record
//                        if this start is the smaller twin
//   sn_uncovered_kernel  per slot: interior nodes no walk crossed (cycles) -> list
//   sn_cycles_kernel     one thread per such node: walk the cycle, lowest slot keeps it
//   sn_events_kernel     per path record: its replay events as (slot, path << 2 | role) pairs
//   cub radix sort       the events by slot; sn_expand_events_kernel: SnEvent records in that order
//   sn_write_kernel      one thread per printed supernode: header line, first k-mer, one base per edge
//
// All of them are latency-bound pointer chases except the per-slot sweeps, which
// stream the table once; grids are sized to the resident-CTA limit of the device.
#include <stdint.h>

#include <cub/device/device_radix_sort.cuh>
#include <type_traits>

#include "graph_kernels.cuh"
#include "supernodes.cuh"

namespace lasv {

namespace {

constexpr int SN_THREADS = 256;

template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_count_kernel(const TableView t, SnCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  unsigned long long nodes = 0, interior = 0, starts = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * SN_THREADS) {
    const uint32_t v = sn_classify<N, DeviceOps>(t, q);
    nodes += v & 1u;
    interior += (v >> 1) & 1u;
    starts += v >> 2;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    nodes += __shfl_xor_sync(0xFFFFFFFFu, nodes, o);
    interior += __shfl_xor_sync(0xFFFFFFFFu, interior, o);
    starts += __shfl_xor_sync(0xFFFFFFFFu, starts, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nodes) atomicAdd(&c->n_nodes, nodes);
    if (interior) atomicAdd(&c->n_interior, interior);
    if (starts) atomicAdd(&c->n_starts, starts);
  }
}

// start = slot << 3 | orientation << 2 | nucleotide
template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_starts_kernel(const TableView t, uint64_t *starts,
                                                              const uint64_t cap, unsigned long long *n_out,
                                                              SnCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * SN_THREADS) {
    const uint32_t bucket = (uint32_t)(q / t.W);
    const uint64_t m = DeviceOps::load(slot_ptr<N>(t, bucket, (uint32_t)(q - (uint64_t)bucket * t.W)) + N);
    if (meta_status(m) == ST_EMPTY) continue;
    const uint32_t e = meta_edges(m);
    if (sn_interior(e) || e == 0) continue;
    const uint32_t n = popc4(e & 15u) + popc4(e >> 4);
    uint64_t at = atomicAdd(n_out, (unsigned long long)n);
    for (uint32_t bit = 0; bit < 8; bit++) {
      if (!((e >> bit) & 1u)) continue;
      if (at < cap) starts[at] = (q << 3) | bit;  // bit = orientation << 2 | nucleotide
      else atomicAdd(&c->n_overflow, 1ull);
      at++;
    }
  }
}

template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_paths_kernel(const TableView t, const int k,
                                                             const uint64_t *__restrict__ starts,
                                                             const uint64_t n_starts, uint8_t *covered,
                                                             SnPath *recs, const uint64_t cap, SnCounters *c) {
  const uint64_t max_steps = 2 * (((uint64_t)t.bucket_mask + 1) * t.W) + 2;
  for (uint64_t i = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; i < n_starts;
       i += (uint64_t)gridDim.x * SN_THREADS) {
    const uint64_t s = starts[i];
    const uint64_t q = s >> 3;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, DeviceOps>(t, q, key);
    SnPath rec;
    bool dangling = false;
    const bool keep = sn_walk_path<N, DeviceOps>(t, k, q, key, meta_edges(m), (uint32_t)(s >> 2) & 1u,
                                                 (uint32_t)s & 3u, max_steps, covered, rec, dangling);
    if (dangling) atomicAdd(&c->n_dangling, 1ull);
    if (keep) {
      const uint64_t at = atomicAdd(&c->n_paths, 1ull);
      if (at < cap) recs[at] = rec;
      else atomicAdd(&c->n_overflow, 1ull);
    }
  }
}

template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_uncovered_kernel(const TableView t, const uint8_t *covered,
                                                                 uint64_t *list, const uint64_t cap,
                                                                 SnCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * SN_THREADS) {
    if (covered[q]) continue;
    if ((sn_classify<N, DeviceOps>(t, q) & 3u) != 3u) continue;
    const uint64_t at = atomicAdd(&c->n_uncovered, 1ull);
    if (list && at < cap) list[at] = q;
  }
}

template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_cycles_kernel(const TableView t, const int k,
                                                              const uint64_t *__restrict__ list, const uint64_t n,
                                                              SnPath *recs, const uint64_t cap, SnCounters *c) {
  const uint64_t max_steps = 2 * (((uint64_t)t.bucket_mask + 1) * t.W) + 2;
  for (uint64_t i = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * SN_THREADS) {
    const uint64_t q = list[i];
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, DeviceOps>(t, q, key);
    SnPath rec;
    bool dangling = false;
    const bool keep = sn_walk_cycle<N, DeviceOps>(t, k, q, key, meta_edges(m), max_steps, rec, dangling);
    if (dangling) atomicAdd(&c->n_dangling, 1ull);
    if (keep) {
      const uint64_t at = atomicAdd(&c->n_cycles, 1ull);
      if (at < cap) recs[at] = rec;
      else atomicAdd(&c->n_overflow, 1ull);
    }
  }
}

template <int N>
__global__ void __launch_bounds__(SN_THREADS) sn_write_kernel(const TableView t, const int k,
                                                             const SnPrint *__restrict__ prints, const uint64_t n,
                                                             char *image, SnCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * SN_THREADS) {
    const SnPrint p = prints[i];
    if (!sn_write_record<N, DeviceOps>(t, k, p, i, image)) atomicAdd(&c->n_dangling, 1ull);
  }
}

__global__ void __launch_bounds__(SN_THREADS) sn_events_kernel(const SnPath *__restrict__ recs, const uint64_t n,
                                                              const uint32_t index_base, uint64_t *keys,
                                                              uint32_t *vals, unsigned long long *n_events) {
  for (uint64_t i = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * SN_THREADS) {
    uint64_t pos[3];
    uint32_t role[3];
    const int m = sn_path_events(recs[i], pos, role);
    if (!m) continue;
    const uint64_t at = atomicAdd(n_events, (unsigned long long)m);
    for (int j = 0; j < m; j++) {
      keys[at + j] = pos[j];
      vals[at + j] = ((index_base + (uint32_t)i) << 2) | role[j];
    }
  }
}

__global__ void __launch_bounds__(SN_THREADS) sn_expand_events_kernel(const uint32_t *__restrict__ vals,
                                                                     const uint64_t n_events,
                                                                     const SnPath *__restrict__ paths,
                                                                     const uint64_t n_paths,
                                                                     const SnPath *__restrict__ cycles,
                                                                     SnEvent *events) {
  for (uint64_t i = (uint64_t)blockIdx.x * SN_THREADS + threadIdx.x; i < n_events;
       i += (uint64_t)gridDim.x * SN_THREADS) {
    const uint32_t v = vals[i], idx = v >> 2;
    const SnPath p = idx < n_paths ? paths[idx] : cycles[idx - n_paths];
    SnEvent e;
    e.s_q = p.s_q;
    e.t_q = p.t_q;
    e.len = p.len;
    e.bits = p.bits;
    e.path = idx;
    e.role = v & 3u;
    events[i] = e;
  }
}

template <typename F>
void sn_dispatch(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

template <typename K>
unsigned sn_grid(K kernel, uint64_t items) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SN_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const uint64_t cap = (uint64_t)per_sm * (uint64_t)sm_count();
  uint64_t grid = (items + SN_THREADS - 1) / SN_THREADS;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return (unsigned)grid;
}

}  // namespace

void launch_sn_count(int N, const TableView &t, SnCounters *c, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_count_kernel<NN><<<sn_grid(sn_count_kernel<NN>, n_slots), SN_THREADS, 0, st>>>(t, c);
  });
}

void launch_sn_starts(int N, const TableView &t, uint64_t *starts, uint64_t cap, unsigned long long *n_out,
                      SnCounters *c, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_starts_kernel<NN><<<sn_grid(sn_starts_kernel<NN>, n_slots), SN_THREADS, 0, st>>>(t, starts, cap, n_out, c);
  });
}

void launch_sn_paths(int N, const TableView &t, int k, const uint64_t *starts, uint64_t n_starts,
                     uint8_t *covered, SnPath *recs, uint64_t cap, SnCounters *c, cudaStream_t st) {
  if (!n_starts) return;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_paths_kernel<NN><<<sn_grid(sn_paths_kernel<NN>, n_starts), SN_THREADS, 0, st>>>(t, k, starts, n_starts,
                                                                                       covered, recs, cap, c);
  });
}

void launch_sn_uncovered(int N, const TableView &t, const uint8_t *covered, uint64_t *list, uint64_t cap,
                         SnCounters *c, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_uncovered_kernel<NN><<<sn_grid(sn_uncovered_kernel<NN>, n_slots), SN_THREADS, 0, st>>>(t, covered, list,
                                                                                              cap, c);
  });
}

void launch_sn_cycles(int N, const TableView &t, int k, const uint64_t *list, uint64_t n, SnPath *recs,
                      uint64_t cap, SnCounters *c, cudaStream_t st) {
  if (!n) return;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_cycles_kernel<NN><<<sn_grid(sn_cycles_kernel<NN>, n), SN_THREADS, 0, st>>>(t, k, list, n, recs, cap, c);
  });
}

void launch_sn_write(int N, const TableView &t, int k, const SnPrint *prints, uint64_t n, char *image,
                     SnCounters *c, cudaStream_t st) {
  if (!n) return;
  sn_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    sn_write_kernel<NN><<<sn_grid(sn_write_kernel<NN>, n), SN_THREADS, 0, st>>>(t, k, prints, n, image, c);
  });
}

void launch_sn_events(const SnPath *recs, uint64_t n, uint32_t index_base, uint64_t *keys, uint32_t *vals,
                      unsigned long long *n_events, cudaStream_t st) {
  if (!n) return;
  sn_events_kernel<<<sn_grid(sn_events_kernel, n), SN_THREADS, 0, st>>>(recs, n, index_base, keys, vals, n_events);
}

size_t sn_sort_tmp_bytes(uint64_t n_events) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (long long)n_events, 0, 64);
  return bytes;
}

void launch_sn_sort_events(uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint64_t n_events,
                           int key_bits, void *tmp, size_t tmp_bytes, const SnPath *paths, uint64_t n_paths,
                           const SnPath *cycles, SnEvent *events, cudaStream_t st) {
  if (!n_events) return;
  // slots are distinct per event, so any order of equal keys would do; the sort is stable anyway
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t *)keys, keys_alt, (const uint32_t *)vals, vals_alt,
                                  (long long)n_events, 0, key_bits, st);
  sn_expand_events_kernel<<<sn_grid(sn_expand_events_kernel, n_events), SN_THREADS, 0, st>>>(vals_alt, n_events, paths,
                                                                                           n_paths, cycles, events);
}

}  // namespace lasv
