// ubench.cu -- microbenchmarks behind the design of the partitioned insert
// (tools only; not part of the product library).
//   A. locality: random 8-byte read + 32-bit reduction inside a WINDOW of the table;
//      the window is global (all CTAs share it and it slides), per SM, or per CTA.
//   B. scatter: N records appended to n_bins stripes, one global atomic each, for
//      several record sizes / store shapes; atomics alone; stores alone.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned smid() { unsigned r; asm("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// scope: 0 global sliding window, 1 per-SM window, 2 per-CTA window, 3 per-warp window
__global__ void __launch_bounds__(256) locality_kernel(uint64_t *buf, uint64_t total_bytes, uint64_t win_bytes,
                                                       int scope, int iters, uint64_t seed, int ops_per_window) {
  const uint64_t n_win = total_bytes / win_bytes;
  const uint64_t sect_per_win = win_bytes / 32;
  uint64_t acc = 0;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; it++) {
    uint64_t wsel;
    const uint64_t epoch = it / ops_per_window;
    if (scope == 0) wsel = mix64(seed + epoch);
    else if (scope == 1) wsel = mix64(seed + epoch * 1000003ull + smid());
    else if (scope == 2) wsel = mix64(seed + epoch * 1000003ull + blockIdx.x);
    else wsel = mix64(seed + epoch * 1000003ull + (tid >> 5));
    const uint64_t w = wsel % n_win;
    const uint64_t s = mix64(seed ^ (tid * 0x100000001B3ull + it)) % sect_per_win;
    uint64_t *p = buf + (w * sect_per_win + s) * 4;
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v != 0x123456789ull) atomicAdd(reinterpret_cast<unsigned *>(p + 1), 1u);
    acc += v;
  }
  if (acc == 0x9999) buf[0] = acc;
}

// mode: 0 ATOM + 3x st.u64 (24 B rec), 1 ATOM + st.v2.u64 + st.u64 (32 B stride), 2 ATOM + st.v4.u64 (32 B),
//       3 ATOM only, 4 RED only, 5 stores only (3x u64, 24 B, position from hash), 6 stores only v4.u64 32 B,
//       7 ATOM + 2x st.v2.u64 (32 B)
__global__ void __launch_bounds__(256) scatter_kernel(uint64_t *recs, unsigned *count, uint64_t n, uint32_t n_bins,
                                                      uint32_t cap, int mode, uint64_t seed) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = mix64(seed + i);
    const uint32_t bin = (uint32_t)(h >> 40) % n_bins;
    uint32_t idx;
    if (mode == 4) { atomicAdd(&count[bin * 32], 1u); continue; }
    if (mode == 5 || mode == 6) idx = (uint32_t)(i / n_bins) % cap;   // unique-ish position, no atomic
    else idx = atomicAdd(&count[bin * 32], 1u);
    if (mode == 3) { if (idx == 0xFFFFFFFFu) recs[0] = 1; continue; }
    if (idx >= cap) continue;
    const uint64_t at = (uint64_t)bin * cap + idx;
    if (mode == 0 || mode == 5) {
      uint64_t *d = recs + at * 3;
      d[0] = h; d[1] = h + 1; d[2] = h + 2;
    } else if (mode == 1) {
      uint64_t *d = recs + at * 4;
      asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(d), "l"(h), "l"(h + 1) : "memory");
      d[2] = h + 2;
    } else if (mode == 7) {
      uint64_t *d = recs + at * 4;
      asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(d), "l"(h), "l"(h + 1) : "memory");
      asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(d + 2), "l"(h + 2), "l"(h + 3) : "memory");
    } else {
      uint64_t *d = recs + at * 4;
      asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(d), "l"(h), "l"(h + 1), "l"(h + 2), "l"(h + 3) : "memory");
    }
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char **argv) {
  const char *what = argc > 1 ? argv[1] : "all";
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  if (!strcmp(what, "all") || !strcmp(what, "locality")) {
    const uint64_t total = 100ull << 30;
    uint64_t *buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 0, total));
    const int iters = 256;
    const uint64_t wins[] = {2ull << 20, 8ull << 20, 32ull << 20, 128ull << 20, 512ull << 20, 2048ull << 20, 16384ull << 20, 100ull << 30};
    for (int scope = 0; scope < 4; scope++)
      for (uint64_t w : wins)
        for (int opw : {16, 256}) {
          if (w == (100ull << 30) && (scope > 0 || opw != 256)) continue;
          const int grid = sms * 6;
          float best = 1e30f;
          for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            locality_kernel<<<grid, 256>>>(buf, total, w, scope, iters, 1234 + rep, opw);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
          }
          const double ops = (double)grid * 256 * iters;
          printf("{\"bench\":\"locality\",\"scope\":%d,\"window_MB\":%.0f,\"ops_per_window\":%d,\"Gops\":%.2f}\n", scope,
                 (double)w / (1 << 20), opw, ops / best / 1e6);
          fflush(stdout);
        }
    CK(cudaFree(buf));
  }
  if (!strcmp(what, "all") || !strcmp(what, "scatter")) {
    const uint64_t n = 180000000ull;
    for (uint32_t n_bins : {64u, 1024u, 8192u, 65536u}) {
      const uint32_t cap = (uint32_t)(n / n_bins * 1.25) + 1024;
      uint64_t *recs; unsigned *count;
      CK(cudaMalloc(&recs, (uint64_t)n_bins * cap * 32));
      CK(cudaMalloc(&count, (size_t)n_bins * 128));
      for (int mode = 0; mode < 8; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaMemset(count, 0, (size_t)n_bins * 128));
          CK(cudaEventRecord(e0));
          scatter_kernel<<<sms * 8, 256>>>(recs, count, n, n_bins, cap, mode, 99 + rep);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"bench\":\"scatter\",\"n_bins\":%u,\"mode\":%d,\"ms\":%.3f,\"Grec_per_s\":%.2f}\n", n_bins, mode, best,
               n / best / 1e6);
        fflush(stdout);
      }
      CK(cudaFree(recs)); CK(cudaFree(count));
    }
  }
  return 0;
}
