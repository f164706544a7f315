// group_kernels.cu -- sm_100a kernels of the streamed-insert PROTOTYPE (groups.cuh has the scheme and
// its status: not on the default path).
//
//   insert_groups_kernel   persistent CTAs; per group: lines global -> shared (128-bit loads, whole lines,
//                          coalesced), records applied in shared memory, lines shared -> global
//   insert_groups_bulk_kernel  the same with cp.async.bulk staging, double buffered (LASV_B200_GROUPS_BULK=1)
//   insert_spilled_kernel  the records whose bucket was full, through the ordinary table_upsert
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "graph_kernels.cuh"
#include "groups.cuh"

namespace lasv {

namespace {

constexpr int GR_THREADS = 256;

// The table protocol on shared memory: same loads / CAS / reductions as DeviceOps, CTA scope.
struct SmemOps {
  static __device__ __forceinline__ uint64_t load(const uint64_t *p) { return *reinterpret_cast<const volatile uint64_t *>(p); }
  static __device__ __forceinline__ void load16(const uint64_t *p, uint64_t &a, uint64_t &b) {
    a = *reinterpret_cast<const volatile uint64_t *>(p);
    b = *reinterpret_cast<const volatile uint64_t *>(p + 1);
  }
  static __device__ __forceinline__ void store(uint64_t *p, uint64_t v) { *reinterpret_cast<volatile uint64_t *>(p) = v; }
  static __device__ __forceinline__ void store_tag(uint8_t *p, uint32_t v) { *reinterpret_cast<volatile uint8_t *>(p) = (uint8_t)v; }
  static __device__ __forceinline__ uint64_t cas(uint64_t *p, uint64_t expect, uint64_t want) {
    return atomicCAS(reinterpret_cast<unsigned long long *>(p), (unsigned long long)expect, (unsigned long long)want);
  }
  static __device__ __forceinline__ void xor32(uint32_t *p, uint32_t v) { atomicXor(p, v); }
  static __device__ __forceinline__ void add32(uint32_t *p, uint32_t v) { atomicAdd(p, v); }
  static __device__ __forceinline__ void or32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
  static __device__ __forceinline__ void and32(uint32_t *p, uint32_t v) { atomicAnd(p, v); }
  static __device__ __forceinline__ void fence() { __threadfence_block(); }
  static __device__ __forceinline__ void backoff() { __nanosleep(20); }
  static __device__ __forceinline__ void count(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }  // global counters
  static __device__ __forceinline__ void flag(unsigned long long *p) { atomicExch(p, 1ull); }
};

template <int N>
__global__ void __launch_bounds__(GR_THREADS) insert_groups_kernel(const uint32_t *__restrict__ recs,
                                                                  const uint64_t *__restrict__ group_offsets,
                                                                  const uint64_t n_groups, const uint32_t buckets_per_group,
                                                                  const TableView t, GraphCounters *ctr, uint32_t *spill,
                                                                  unsigned long long *n_spill, const uint64_t spill_cap) {
  constexpr int RECW = 2 * N + 1;
  extern __shared__ __align__(128) unsigned char staged[];
  const uint64_t gbytes = group_bytes(t, buckets_per_group);
  const uint32_t n_vec = (uint32_t)(gbytes / 16);
  uint4 *sm = reinterpret_cast<uint4 *>(staged);
  unsigned long long my_new = 0;
  for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const uint64_t lo = group_offsets[g], hi = group_offsets[g + 1];
    if (lo == hi) continue;  // the same for every thread of the CTA: no line of this group is touched
    uint4 *lines = reinterpret_cast<uint4 *>(t.lines + g * gbytes);
    for (uint32_t i = threadIdx.x; i < n_vec; i += GR_THREADS) sm[i] = lines[i];
    __syncthreads();
    const TableView local = group_local_view(t, staged, g, buckets_per_group);
    for (uint64_t i = lo + threadIdx.x; i < hi; i += GR_THREADS) {
      const int r = group_upsert<N, SmemOps>(local, ctr, recs + i * RECW);
      my_new += (r == 1);
      if (r < 0) {
        const unsigned long long at = atomicAdd(n_spill, 1ull);
        if (at < spill_cap) spill[at] = (uint32_t)i;
      }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_vec; i += GR_THREADS) lines[i] = sm[i];
    __syncthreads();  // the staging buffer is free for the next group
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, o);
  if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// The same with the copy engines of the SM doing the staging (cp.async.bulk, 1-D TMA copies), double buffered:
// while the records of group i are applied in one buffer, group i+1 (the CTA's next group that has records)
// arrives in the other and group i-1 is on its way back.  One thread issues the copies; an mbarrier per buffer
// counts the arriving bytes; the write-back of a buffer must have READ it before the buffer is refilled
// (wait_group.read), and the updates made through the generic proxy are fenced before the async proxy reads them.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void __launch_bounds__(GR_THREADS) insert_groups_bulk_kernel(const uint32_t *__restrict__ recs,
                                                                       const uint64_t *__restrict__ group_offsets,
                                                                       const uint64_t n_groups, const uint32_t buckets_per_group,
                                                                       const TableView t, GraphCounters *ctr, uint32_t *spill,
                                                                       unsigned long long *n_spill, const uint64_t spill_cap) {
  constexpr int RECW = 2 * N + 1;
  extern __shared__ __align__(128) unsigned char staged[];
  __shared__ __align__(8) uint64_t mbar[2];
  const uint32_t gbytes = (uint32_t)group_bytes(t, buckets_per_group);
  unsigned char *bufs[2] = {staged, staged + gbytes};
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto next_with_records = [&](uint64_t g) {  // the same value in every thread
    for (; g < n_groups; g += gridDim.x)
      if (group_offsets[g] != group_offsets[g + 1]) return g;
    return n_groups;
  };
  auto load = [&](uint64_t g, int b) {  // one thread
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[b])), "r"(gbytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(bufs[b])),
                 "l"(t.lines + g * (uint64_t)gbytes), "r"(gbytes), "r"(smem_addr(&mbar[b]))
                 : "memory");
  };
  unsigned long long my_new = 0;
  uint64_t g = next_with_records(blockIdx.x);
  if (threadIdx.x == 0 && g < n_groups) load(g, 0);
  for (uint32_t it = 0; g < n_groups; it++) {
    const int b = it & 1;
    const uint64_t gn = next_with_records(g + gridDim.x);
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the other buffer's write-back has read it
      if (gn < n_groups) load(gn, b ^ 1);
    }
    const uint32_t parity = (it >> 1) & 1u;
    uint32_t done = 0;
    for (uint32_t tries = 0; !done; tries++) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_addr(&mbar[b])), "r"(parity) : "memory");
      if (tries > (1u << 24)) __trap();  // a copy that never lands must end the kernel with an error, not hang the GPU
    }
    const TableView local = group_local_view(t, bufs[b], g, buckets_per_group);
    for (uint64_t i = group_offsets[g] + threadIdx.x, hi = group_offsets[g + 1]; i < hi; i += GR_THREADS) {
      const int r = group_upsert<N, SmemOps>(local, ctr, recs + i * RECW);
      my_new += (r == 1);
      if (r < 0) {
        const unsigned long long at = atomicAdd(n_spill, 1ull);
        if (at < spill_cap) spill[at] = (uint32_t)i;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my updates -> visible to the bulk copy
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(t.lines + g * (uint64_t)gbytes),
                   "r"(smem_addr(bufs[b])), "r"(gbytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    g = gn;
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, o);
  if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

template <int N>
__global__ void __launch_bounds__(GR_THREADS) insert_spilled_kernel(const uint32_t *__restrict__ recs,
                                                                   const uint32_t *__restrict__ spill, const uint64_t n,
                                                                   const TableView t, GraphCounters *ctr) {
  constexpr int RECW = 2 * N + 1;
  unsigned long long my_new = 0;
  for (uint64_t j = (uint64_t)blockIdx.x * GR_THREADS + threadIdx.x; j < n; j += (uint64_t)gridDim.x * GR_THREADS) {
    uint64_t key[N];
    uint32_t edge, count;
    group_load_record<N>(recs + (uint64_t)spill[j] * RECW, key, edge, count);
    my_new += (table_upsert<N, DeviceOps>(t, ctr, key, edge, count) == 1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) my_new += __shfl_xor_sync(0xFFFFFFFFu, my_new, o);
  if ((threadIdx.x & 31) == 0 && my_new) atomicAdd(&ctr->unique_kmers, my_new);
}

// counting sort of the records by group: histogram ...
template <int N>
__global__ void __launch_bounds__(GR_THREADS) group_count_kernel(const uint32_t *__restrict__ recs, const uint64_t n,
                                                                const uint32_t buckets_per_group, const TableView t,
                                                                uint32_t *counts) {
  constexpr int RECW = 2 * N + 1;
  for (uint64_t i = (uint64_t)blockIdx.x * GR_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * GR_THREADS)
    atomicAdd(&counts[group_of_record<N>(t, recs + i * RECW, buckets_per_group)], 1u);
}

// ... and scatter (the order inside a group is whatever the atomics make it: the graph does not depend on it)
template <int N>
__global__ void __launch_bounds__(GR_THREADS) group_scatter_kernel(const uint32_t *__restrict__ recs, const uint64_t n,
                                                                  const uint32_t buckets_per_group, const TableView t,
                                                                  const uint64_t *__restrict__ offsets, uint32_t *cursor,
                                                                  uint32_t *sorted) {
  constexpr int RECW = 2 * N + 1;
  for (uint64_t i = (uint64_t)blockIdx.x * GR_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * GR_THREADS) {
    const uint64_t g = group_of_record<N>(t, recs + i * RECW, buckets_per_group);
    const uint64_t at = offsets[g] + atomicAdd(&cursor[g], 1u);
#pragma unroll
    for (int w = 0; w < RECW; w++) sorted[at * RECW + w] = recs[i * RECW + w];
  }
}

__global__ void group_set_last_offset_kernel(uint64_t *offsets, uint64_t n_groups, uint64_t n) { offsets[n_groups] = n; }

template <typename F>
void gr_dispatch(int N, F &&f) {
  if (N == 1) f(std::integral_constant<int, 1>());
  else if (N == 2) f(std::integral_constant<int, 2>());
  else f(std::integral_constant<int, 3>());
}

}  // namespace

size_t insert_groups_smem_limit() { return 200u * 1024u; }  // of the 227 KB a CTA may have on sm_100a

void launch_insert_groups(int N, const uint32_t *recs, const uint64_t *group_offsets, uint64_t n_groups,
                          uint32_t buckets_per_group, const TableView &t, GraphCounters *ctr, uint32_t *spill,
                          unsigned long long *n_spill, uint64_t spill_cap, cudaStream_t st) {
  if (!n_groups) return;
  // LASV_B200_GROUPS_BULK=1: staging by cp.async.bulk, double buffered (two buffers per CTA)
  const char *e = getenv("LASV_B200_GROUPS_BULK");
  const bool bulk = e && atoi(e) != 0 && 2 * group_bytes(t, buckets_per_group) <= insert_groups_smem_limit();
  const size_t smem = (size_t)group_bytes(t, buckets_per_group) * (bulk ? 2 : 1);
  gr_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    auto kernel = bulk ? insert_groups_bulk_kernel<NN> : insert_groups_kernel<NN>;
    int per_sm = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, GR_THREADS, smem) != cudaSuccess || per_sm < 1) {
      fprintf(stderr, "lasv_b200: grouped-insert prototype: %zu bytes of shared memory per CTA cannot be set up: %s\n", smem,
              cudaGetErrorString(cudaPeekAtLastError()));
      return;  // nothing is launched; the caller's check of the launch reports the error left in the runtime
    }
    uint64_t grid = (uint64_t)per_sm * (uint64_t)sm_count();  // persistent: resident CTAs x SMs
    if (grid > n_groups) grid = n_groups;
    kernel<<<(unsigned)grid, GR_THREADS, smem, st>>>(recs, group_offsets, n_groups, buckets_per_group, t, ctr, spill, n_spill,
                                                    spill_cap);
  });
}

void launch_group_records(int N, const uint32_t *recs, uint64_t n, uint32_t buckets_per_group, const TableView &t,
                          uint32_t *counts, uint64_t *offsets, void *scan_tmp, size_t scan_tmp_bytes, uint32_t *sorted,
                          cudaStream_t st) {
  const uint64_t n_groups = ((uint64_t)t.bucket_mask + 1) / buckets_per_group;
  uint64_t grid = (n + GR_THREADS - 1) / GR_THREADS;
  const uint64_t cap = (uint64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaMemsetAsync(counts, 0, n_groups * sizeof(uint32_t), st);
  gr_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    if (n) group_count_kernel<NN><<<(unsigned)grid, GR_THREADS, 0, st>>>(recs, n, buckets_per_group, t, counts);
  });
  launch_exclusive_scan_u32_to_u64(counts, offsets, n_groups, scan_tmp, scan_tmp_bytes, st);
  group_set_last_offset_kernel<<<1, 1, 0, st>>>(offsets, n_groups, n);
  cudaMemsetAsync(counts, 0, n_groups * sizeof(uint32_t), st);  // now the per-group cursors
  gr_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    if (n) group_scatter_kernel<NN><<<(unsigned)grid, GR_THREADS, 0, st>>>(recs, n, buckets_per_group, t, offsets, counts, sorted);
  });
}

void launch_insert_spilled(int N, const uint32_t *recs, const uint32_t *spill, uint64_t n_spill, const TableView &t,
                           GraphCounters *ctr, cudaStream_t st) {
  if (!n_spill) return;
  gr_dispatch(N, [&](auto nn) {
    constexpr int NN = decltype(nn)::value;
    uint64_t grid = (n_spill + GR_THREADS - 1) / GR_THREADS;
    const uint64_t cap = (uint64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    insert_spilled_kernel<NN><<<(unsigned)grid, GR_THREADS, 0, st>>>(recs, spill, n_spill, t, ctr);
  });
}

}  // namespace lasv
