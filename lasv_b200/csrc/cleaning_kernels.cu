// cleaning_kernels.cu -- sm_100a kernels of the graph cleaning (tip clipping, removal
// of low-coverage supernodes); the per-path arithmetic and the reference lines are in
// cleaning.cuh, the sequential decisions in cleaning_host.h, the path enumeration in
// supernode_kernels.cu.
//
//   cl_nodes_kernel      per slot: the non-interior nodes (slot, coverage, edges) -- the
//                        vertices of the compressed graph
//   cl_coverage_kernel   per path record: largest interior coverage, lowest low-coverage
//                        interior slot (one walk per record)
//   cl_apply_*_kernel    the decisions: mark the interior of a path as pruned (one walk),
//                        prune a node, clear edge bits of a node (atomic AND: several
//                        paths may end in one node)
//   cl_settle_kernel     per slot: nodes marked by the walks lose their edges
//   br_write_kernel      branch printing (branches.cuh): one thread per printed branch writes its
//                        FASTA record into the file image and, if asked, leaves the `visited` marks
#include <stdint.h>

#include <type_traits>

#include "branches.cuh"
#include "cleaning.cuh"
#include "graph_kernels.cuh"

namespace lasv {

namespace {

constexpr int CL_THREADS = 256;

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_nodes_kernel(const TableView t, ClNodeRec *out, const uint64_t cap,
                                                             ClCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * CL_THREADS) {
    const uint64_t m = DeviceOps::load(cl_meta_ptr<N, DeviceOps>(t, q));
    const uint32_t st = meta_status(m), e = meta_edges(m);
    if (st == ST_EMPTY || st == ST_PRUNED || e == 0 || sn_interior(e)) continue;
    const uint64_t at = atomicAdd(&c->n_nodes, 1ull);
    if (out && at < cap) {
      ClNodeRec r;
      r.q = q;
      r.covg = meta_covg(m);
      r.edges = e;
      out[at] = r;
    }
  }
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_coverage_kernel(const TableView t, const int k,
                                                                const SnPath *__restrict__ recs, const uint64_t n,
                                                                const uint32_t threshold, uint32_t *max_covg,
                                                                uint64_t *min_low_q, uint32_t *s_covg, ClCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * CL_THREADS) {
    const SnPath p = recs[i];
    uint32_t mx = 0;
    uint64_t lo = SN_NONE;
    if (!cl_path_coverage<N, DeviceOps>(t, k, p, threshold, mx, lo)) atomicAdd(&c->n_failed, 1ull);
    max_covg[i] = mx;
    min_low_q[i] = lo;
    s_covg[i] = meta_covg(DeviceOps::load(cl_meta_ptr<N, DeviceOps>(t, p.s_q)));
  }
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_apply_paths_kernel(const TableView t, const int k,
                                                                   const ClPathPrune *__restrict__ paths,
                                                                   const uint64_t n, ClCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * CL_THREADS) {
    const ClPathPrune p = paths[i];
    if (!cl_prune_path<N, DeviceOps>(t, k, p.q, p.on & 1u, p.on >> 1, p.len, p.keep_q)) atomicAdd(&c->n_failed, 1ull);
  }
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_apply_nodes_kernel(const TableView t, const uint64_t *__restrict__ nodes,
                                                                   const uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * CL_THREADS)
    cl_prune_node<N, DeviceOps>(t, nodes[i]);
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_apply_resets_kernel(const TableView t,
                                                                    const ClEdgeReset *__restrict__ resets,
                                                                    const uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * CL_THREADS)
    cl_reset_edges<N, DeviceOps>(t, resets[i].q, resets[i].mask);
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) cl_settle_kernel(const TableView t) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * CL_THREADS)
    cl_settle_node<N, DeviceOps>(t, q);
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) br_write_kernel(const TableView t, const int k,
                                                             const BrPrint *__restrict__ prints, const uint64_t n,
                                                             char *image, const int mark, ClCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * CL_THREADS)
    if (!br_write_record<N, DeviceOps>(t, k, prints[i], image, mark)) atomicAdd(&c->n_failed, 1ull);
}

template <int N>
__global__ void __launch_bounds__(CL_THREADS) br_reset_marks_kernel(const TableView t) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * CL_THREADS)
    br_reset_mark<N, DeviceOps>(t, q);
}

// `--health_check`: one thread per slot, every edge followed once (a random table read per edge)
template <int N>
__global__ void __launch_bounds__(CL_THREADS) hc_check_kernel(const TableView t, const int k, HcCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  unsigned long long nodes = 0, edges = 0, mt = 0, mr = 0, zk = 0;
  for (uint64_t q = (uint64_t)blockIdx.x * CL_THREADS + threadIdx.x; q < n_slots;
       q += (uint64_t)gridDim.x * CL_THREADS) {
    if (meta_status(DeviceOps::load(cl_meta_ptr<N, DeviceOps>(t, q))) == ST_EMPTY) continue;
    uint32_t a = 0, b = 0;
    bool z = false;
    edges += hc_check_node<N, DeviceOps>(t, k, q, a, b, z);
    nodes++;
    mt += a;
    mr += b;
    zk += z;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    nodes += __shfl_xor_sync(0xFFFFFFFFu, nodes, o);
    edges += __shfl_xor_sync(0xFFFFFFFFu, edges, o);
    mt += __shfl_xor_sync(0xFFFFFFFFu, mt, o);
    mr += __shfl_xor_sync(0xFFFFFFFFu, mr, o);
    zk += __shfl_xor_sync(0xFFFFFFFFu, zk, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nodes) atomicAdd(&c->nodes, nodes);
    if (edges) atomicAdd(&c->edges, edges);
    if (mt) atomicAdd(&c->missing_target, mt);
    if (mr) atomicAdd(&c->missing_return, mr);
    if (zk) atomicAdd(&c->zero_kmers, zk);
  }
}

template <typename F>
void cl_dispatch(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

template <typename K>
unsigned cl_grid(K kernel, uint64_t items) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CL_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const uint64_t cap = (uint64_t)per_sm * (uint64_t)sm_count();
  uint64_t grid = (items + CL_THREADS - 1) / CL_THREADS;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  return (unsigned)grid;
}

}  // namespace

void launch_cl_nodes(int N, const TableView &t, ClNodeRec *out, uint64_t cap, ClCounters *c, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    cl_nodes_kernel<NN><<<cl_grid(cl_nodes_kernel<NN>, n_slots), CL_THREADS, 0, st>>>(t, out, cap, c);
  });
}

void launch_cl_coverage(int N, const TableView &t, int k, const SnPath *recs, uint64_t n, uint32_t threshold,
                        uint32_t *max_covg, uint64_t *min_low_q, uint32_t *s_covg, ClCounters *c, cudaStream_t st) {
  if (!n) return;
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    cl_coverage_kernel<NN><<<cl_grid(cl_coverage_kernel<NN>, n), CL_THREADS, 0, st>>>(t, k, recs, n, threshold, max_covg,
                                                                                     min_low_q, s_covg, c);
  });
}

// path walks (status only), then the sweep that clears the edges of what they marked, then nodes, then edge bits
void launch_cl_apply(int N, const TableView &t, int k, const ClPathPrune *paths, uint64_t n_paths, const uint64_t *nodes,
                     uint64_t n_nodes, const ClEdgeReset *resets, uint64_t n_resets, ClCounters *c, cudaStream_t st) {
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    if (n_paths) {
      const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
      cl_apply_paths_kernel<NN><<<cl_grid(cl_apply_paths_kernel<NN>, n_paths), CL_THREADS, 0, st>>>(t, k, paths, n_paths, c);
      cl_settle_kernel<NN><<<cl_grid(cl_settle_kernel<NN>, n_slots), CL_THREADS, 0, st>>>(t);
    }
    if (n_nodes)
      cl_apply_nodes_kernel<NN><<<cl_grid(cl_apply_nodes_kernel<NN>, n_nodes), CL_THREADS, 0, st>>>(t, nodes, n_nodes);
    if (n_resets)
      cl_apply_resets_kernel<NN><<<cl_grid(cl_apply_resets_kernel<NN>, n_resets), CL_THREADS, 0, st>>>(t, resets, n_resets);
  });
}

void launch_br_write(int N, const TableView &t, int k, const BrPrint *prints, uint64_t n, char *image, int mark,
                     ClCounters *c, cudaStream_t st) {
  if (!n) return;
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    br_write_kernel<NN><<<cl_grid(br_write_kernel<NN>, n), CL_THREADS, 0, st>>>(t, k, prints, n, image, mark, c);
  });
}

void launch_hc_check(int N, const TableView &t, int k, HcCounters *c, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    hc_check_kernel<NN><<<cl_grid(hc_check_kernel<NN>, n_slots), CL_THREADS, 0, st>>>(t, k, c);
  });
}

void launch_br_reset_marks(int N, const TableView &t, cudaStream_t st) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  cl_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    br_reset_marks_kernel<NN><<<cl_grid(br_reset_marks_kernel<NN>, n_slots), CL_THREADS, 0, st>>>(t);
  });
}

}  // namespace lasv
