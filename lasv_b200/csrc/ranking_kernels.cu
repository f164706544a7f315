// ranking_kernels.cu -- sm_100a kernels of the path enumeration by list ranking (ranking.cuh has the scheme;
// opt-in through LASV_B200_SN_RANKING=1).  All kernels are grid-stride sweeps without atomics on the data: an element
// is written by one thread per round (double buffering), the only atomics are the counters.
//
//   rk_flags_kernel     per slot: one bit per interior node
//   rk_popc_kernel      per 64-slot word: its number of interior nodes (then an exclusive scan: the ranks)
//   rk_init_kernel      per interior node: its two elements (one table lookup each)
//   rk_jump_kernel      per element: one round of pointer jumping, counts the elements still on their way
//   rk_paths_kernel     per start of a non-interior node: the path record, read off the first element
//   rk_covered_kernel   per interior node: crossed by some path (its element is done) or left to the cycle walker
#include <stdint.h>

#include <type_traits>

#include "graph_kernels.cuh"
#include "ranking.cuh"

namespace lasv {

namespace {

constexpr int RK_THREADS = 256;

template <int N>
__global__ void __launch_bounds__(RK_THREADS) rk_flags_kernel(const TableView t, unsigned long long *bits) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; q < n_slots; q += (uint64_t)gridDim.x * RK_THREADS)
    if ((sn_classify<N, DeviceOps>(t, q) & 3u) == 3u) atomicOr(&bits[q >> 6], 1ull << (q & 63));
}

__global__ void __launch_bounds__(RK_THREADS) rk_popc_kernel(const uint64_t *__restrict__ bits, const uint64_t n_words,
                                                            uint32_t *counts) {
  for (uint64_t w = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * RK_THREADS)
    counts[w] = (uint32_t)__popcll(bits[w]);
}

template <int N>
__global__ void __launch_bounds__(RK_THREADS) rk_init_kernel(const TableView t, const int k, const RkMap map, RkElem *el,
                                                            SnCounters *c) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; q < n_slots; q += (uint64_t)gridDim.x * RK_THREADS) {
    if (!((map.bits[q >> 6] >> (q & 63)) & 1ull)) continue;
    const uint32_t cid = rk_cid(map, q);
    for (uint32_t o = 0; o < 2; o++) {
      RkElem e;
      if (!rk_init<N, DeviceOps>(t, k, map, q, o, e)) atomicAdd(&c->n_dangling, 1ull);
      el[2u * cid + o] = e;
    }
  }
}

__global__ void __launch_bounds__(RK_THREADS) rk_jump_kernel(const RkElem *__restrict__ cur, RkElem *nxt, const uint64_t n,
                                                            unsigned long long *active) {
  unsigned long long mine = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * RK_THREADS) {
    RkElem e;
    mine += rk_jump(cur, (uint32_t)i, e) ? 1ull : 0ull;
    nxt[i] = e;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(active, mine);
}

template <int N>
__global__ void __launch_bounds__(RK_THREADS) rk_paths_kernel(const TableView t, const int k, const RkMap map,
                                                             const RkElem *__restrict__ el, const uint64_t *__restrict__ starts,
                                                             const uint64_t n_starts, SnPath *recs, const uint64_t cap,
                                                             SnCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n_starts; i += (uint64_t)gridDim.x * RK_THREADS) {
    const uint64_t s = starts[i], q = s >> 3;
    uint64_t key[N];
    const uint64_t m = sn_load_node<N, DeviceOps>(t, q, key);
    SnPath rec;
    bool dangling = false, unfinished = false;
    const bool keep = rk_path_of_start<N, DeviceOps>(t, k, map, el, q, key, meta_edges(m), (uint32_t)(s >> 2) & 1u,
                                                     (uint32_t)s & 3u, rec, dangling, unfinished);
    if (dangling || unfinished) atomicAdd(&c->n_dangling, 1ull);
    if (keep) {
      const uint64_t at = atomicAdd(&c->n_paths, 1ull);
      if (at < cap) recs[at] = rec;
      else atomicAdd(&c->n_overflow, 1ull);
    }
  }
}

__global__ void __launch_bounds__(RK_THREADS) rk_covered_kernel(const RkMap map, const RkElem *__restrict__ el,
                                                               const uint64_t n_slots, uint8_t *covered) {
  for (uint64_t q = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; q < n_slots; q += (uint64_t)gridDim.x * RK_THREADS)
    if ((map.bits[q >> 6] >> (q & 63)) & 1ull) covered[q] = el[2u * rk_cid(map, q)].succ == RK_DONE ? 1 : 0;
}

template <int N>
__global__ void __launch_bounds__(RK_THREADS) rk_write_frames_kernel(const TableView t, const int k,
                                                                    const SnPrint *__restrict__ prints,
                                                                    const uint64_t *__restrict__ end_keys, const uint64_t n,
                                                                    char *image, SnCounters *c) {
  for (uint64_t i = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * RK_THREADS)
    if (!rk_write_frame<N, DeviceOps>(t, k, prints[i], end_keys[i], i, image)) atomicAdd(&c->n_dangling, 1ull);
}

template <int N>
__global__ void __launch_bounds__(RK_THREADS) rk_write_nodes_kernel(const TableView t, const int k, const RkMap map,
                                                                   const RkElem *__restrict__ el,
                                                                   const RkPrintKey *__restrict__ keys, const uint64_t n_keys,
                                                                   const SnPrint *__restrict__ prints, char *image) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  for (uint64_t q = (uint64_t)blockIdx.x * RK_THREADS + threadIdx.x; q < n_slots; q += (uint64_t)gridDim.x * RK_THREADS)
    if ((map.bits[q >> 6] >> (q & 63)) & 1ull) rk_write_node<N, DeviceOps>(t, k, map, el, keys, n_keys, prints, q, image);
}

template <typename F>
void rk_dispatch(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

unsigned rk_grid(uint64_t items) {
  uint64_t grid = (items + RK_THREADS - 1) / RK_THREADS;
  const uint64_t cap = (uint64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  return (unsigned)(grid < 1 ? 1 : grid);
}

}  // namespace

#define RK_TRY(x)                       \
  do {                                  \
    err = (x);                          \
    if (err != cudaSuccess) goto done;  \
  } while (0)

cudaError_t sn_paths_ranked(int N, const TableView &t, int k, const uint64_t *starts, uint64_t n_starts,
                            uint64_t n_interior, uint8_t *covered, SnPath *recs, uint64_t cap, SnCounters *c,
                            cudaStream_t st, RkState *keep) {
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W, n_words = (n_slots + 63) / 64;
  const uint64_t n_el = 2 * n_interior;
  if (n_el >= RK_DONE) return cudaErrorInvalidValue;  // element indices are 32-bit
  cudaError_t err = cudaSuccess;
  unsigned long long *d_bits = nullptr, *d_active = nullptr, h_active = 0, prev_active = ~0ull;
  uint64_t *d_rank = nullptr;
  uint32_t *d_counts = nullptr;
  void *d_tmp = nullptr;
  RkElem *d_el[2] = {nullptr, nullptr};
  int cur = 0;
  const size_t tmp_bytes = exclusive_scan_tmp_bytes(n_words);
  RkMap map;
  RK_TRY(cudaMalloc(&d_bits, n_words * 8));
  RK_TRY(cudaMalloc(&d_rank, n_words * 8));
  RK_TRY(cudaMalloc(&d_counts, n_words * 4));
  RK_TRY(cudaMalloc(&d_tmp, tmp_bytes > 16 ? tmp_bytes : 16));
  RK_TRY(cudaMalloc(&d_active, 8));
  RK_TRY(cudaMalloc(&d_el[0], (n_el ? n_el : 1) * sizeof(RkElem)));
  RK_TRY(cudaMalloc(&d_el[1], (n_el ? n_el : 1) * sizeof(RkElem)));
  RK_TRY(cudaMemsetAsync(d_bits, 0, n_words * 8, st));
  rk_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    rk_flags_kernel<NN><<<rk_grid(n_slots), RK_THREADS, 0, st>>>(t, d_bits);
  });
  rk_popc_kernel<<<rk_grid(n_words), RK_THREADS, 0, st>>>(reinterpret_cast<const uint64_t *>(d_bits), n_words, d_counts);
  launch_exclusive_scan_u32_to_u64(d_counts, d_rank, n_words, d_tmp, tmp_bytes, st);
  map.bits = reinterpret_cast<const uint64_t *>(d_bits);
  map.rank = d_rank;
  rk_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    rk_init_kernel<NN><<<rk_grid(n_slots), RK_THREADS, 0, st>>>(t, k, map, d_el[0], c);
  });
  RK_TRY(cudaGetLastError());
  // rounds: every round finishes at least one element of every chain that has a terminal; what is still on its
  // way when a round changes nothing lies on cycles
  for (int round = 0; round < 64 && n_el; round++) {
    RK_TRY(cudaMemsetAsync(d_active, 0, 8, st));
    rk_jump_kernel<<<rk_grid(n_el), RK_THREADS, 0, st>>>(d_el[cur], d_el[cur ^ 1], n_el, d_active);
    cur ^= 1;
    RK_TRY(cudaMemcpyAsync(&h_active, d_active, 8, cudaMemcpyDeviceToHost, st));
    RK_TRY(cudaStreamSynchronize(st));
    if (h_active == 0 || h_active == prev_active) break;
    prev_active = h_active;
  }
  if (n_starts)
    rk_dispatch(N, [&](auto nn) {
      constexpr int NN = decltype(nn)::value;
      rk_paths_kernel<NN><<<rk_grid(n_starts), RK_THREADS, 0, st>>>(t, k, map, d_el[cur], starts, n_starts, recs, cap, c);
    });
  if (covered) rk_covered_kernel<<<rk_grid(n_slots), RK_THREADS, 0, st>>>(map, d_el[cur], n_slots, covered);
  RK_TRY(cudaGetLastError());
  RK_TRY(cudaStreamSynchronize(st));
done:
  if (keep && err == cudaSuccess) {  // the final elements and the interior-node map stay for the per-node writer
    keep->bits = d_bits;
    keep->rank = d_rank;
    keep->el = d_el[cur];
    d_bits = nullptr;
    d_rank = nullptr;
    d_el[cur] = nullptr;
  }
  cudaFree(d_bits); cudaFree(d_rank); cudaFree(d_counts); cudaFree(d_tmp); cudaFree(d_active);
  cudaFree(d_el[0]); cudaFree(d_el[1]);
  return err;
}

void rk_release(RkState &s) {
  cudaFree(s.bits);
  cudaFree(s.rank);
  cudaFree(s.el);
  s = RkState();
}

void rk_write_supernodes(int N, const TableView &t, int k, const RkState &s, const RkPrintKey *keys, uint64_t n_keys,
                         const SnPrint *prints, const uint64_t *end_keys, uint64_t n_prints, char *image, SnCounters *c,
                         cudaStream_t st) {
  if (!n_prints) return;
  const uint64_t n_slots = ((uint64_t)t.bucket_mask + 1) * t.W;
  RkMap map;
  map.bits = reinterpret_cast<const uint64_t *>(s.bits);
  map.rank = s.rank;
  rk_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    rk_write_frames_kernel<NN><<<rk_grid(n_prints), RK_THREADS, 0, st>>>(t, k, prints, end_keys, n_prints, image, c);
    if (n_keys) rk_write_nodes_kernel<NN><<<rk_grid(n_slots), RK_THREADS, 0, st>>>(t, k, map, s.el, keys, n_keys, prints, image);
  });
}

}  // namespace lasv
