// ubench.cu -- microbenchmarks behind the design of the partitioned insert
// (tools only; not part of the product library).
//   A. locality: random 8-byte read + 32-bit reduction inside a WINDOW of the table;
//      the window is global (all CTAs share it and it slides), per SM, or per CTA.
//   C. groups: stream groups of lines through shared memory, update them there (the planned insert).
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

// DRAM fetch granularity: random lines of a big buffer, a few accesses per line (run under ncu)
// pat 0: ld s0 | 1: ld s0,s1 | 2: ld s0,s2 | 3: ld s0,s3 | 4: ld s0 + red s1 | 5: ld s0 + red s0 | 6: ld s1,s2 (straddle atoms)
// 7: ld s0 (L1-cached plain ld) | 8: red s0 only | 9: ld s0, then dependent ld s3
__global__ void __launch_bounds__(256) gran_kernel(uint64_t *buf, uint64_t n_lines, uint64_t n_ops, int pat, uint64_t seed) {
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t *line = buf + (mix64(seed + i) % n_lines) * 16;
    auto ld = [&](int sector) { uint64_t v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(line + 4 * sector) : "memory"); return v; };
    if (pat == 0) acc += ld(0);
    else if (pat == 1) acc += ld(0) + ld(1);
    else if (pat == 2) acc += ld(0) + ld(2);
    else if (pat == 3) acc += ld(0) + ld(3);
    else if (pat == 4) { uint64_t v = ld(0); if (v != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(line + 4), 1u); }
    else if (pat == 5) { uint64_t v = ld(0); if (v != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(line + 1), 1u); }
    else if (pat == 6) acc += ld(1) + ld(2);
    else if (pat == 7) acc += line[0];
    else if (pat == 8) atomicAdd(reinterpret_cast<unsigned *>(line + 1), 1u);
    else if (pat == 9) { uint64_t v = ld(0); acc += line[12 + (v & 1)]; }
    else if (pat == 10) { uint64_t a, b; asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(line) : "memory"); acc += a + b; }
    else if (pat == 11) { uint64_t a, b, c, d; asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(line) : "memory"); acc += a + b + c + d; }
    else if (pat == 12) { acc += line[0] + line[4]; }
    else if (pat == 13) { uint64_t v = ld(0); uint64_t *s = line + 4 + (v & 1); uint64_t a = ld(1), b, c;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(s + 1) : "memory");
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(s + 2) : "memory");
      if (a + b + c != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(s), 1u); }
    else if (pat == 14) { uint64_t v = ld(0); uint64_t *s = line + 4 + 4 * (v & 1); uint64_t a, b, c, d;
      asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(s) : "memory");
      if (a + b + c + d != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(s), 1u); }
    else if (pat == 15) { uint64_t a, b, c, d; asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(line) : "memory");
      if (a + b + c + d != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(line + 1), 1u); }
    else if (pat == 16) { // two independent ops per iteration (MLP 2): ld s0 each, then red each
      uint64_t *l2 = buf + (mix64(seed + i + 0x5555555ull) % n_lines) * 16;
      uint64_t v = ld(0), w; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(l2) : "memory");
      if (v != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(line + 1), 1u);
      if (w != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(l2 + 1), 1u); }
    else if (pat == 18) { uint64_t v; asm volatile("ld.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 19) { uint64_t v; asm volatile("ld.global.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 20) { uint64_t v; asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 21) { uint64_t v; asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 22) { uint64_t v; asm volatile("ld.global.lu.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 23) { uint64_t v; asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(line) : "memory"); acc += v; }
    else if (pat == 24) { unsigned r = atomicAdd(reinterpret_cast<unsigned *>(line + 1), 0u); acc += r; }
    else { // 17: whole line as 4 x v4 loads issued together
      uint64_t t = 0;
      for (int q = 0; q < 4; q++) { uint64_t a, b, c, d; asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(line + 4 * q)); t += a + b + c + d; }
      acc += t; }
  }
  if (acc == 0x9999) buf[0] = acc;
}

// "sorted insert" access: ops sweep the table region by region (all CTAs inside the same region at a time),
// inside a region each op goes to a random 128-byte line, ~one op per 5 lines.
template <int PAT, bool L64>
__device__ __forceinline__ uint64_t sweep_op(uint64_t *line) {
  auto ld8 = [&](uint64_t *p) { uint64_t v; if (L64) asm volatile("ld.relaxed.gpu.global.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; };
  auto ld16 = [&](uint64_t *p) { uint64_t a, b; if (L64) asm volatile("ld.relaxed.gpu.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); return a + b; };
  auto ld32 = [&](uint64_t *p) { uint64_t a, b, c, d; if (L64) asm volatile("ld.relaxed.gpu.global.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory"); else asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory"); return a + b + c + d; };
  if (PAT == 0) return ld8(line);
  if (PAT == 1) { uint64_t v = ld8(line); return v + ld16(line + 2 + 2 * (v & 1)); }
  if (PAT == 2) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(s + 2), 1u); return x; }
  if (PAT == 3) { uint64_t v = ld8(line); uint64_t *s = line + 4 + 4 * (v & 1); uint64_t x = ld32(s); if (x != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(s + 2), 1u); return x; }
  if (PAT == 4) { uint64_t x = ld32(line); if (x != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(line + 2), 1u); return x; }
  if (PAT == 5) { atomicAdd(reinterpret_cast<unsigned *>(line + 2), 1u); return 0; }
  if (PAT == 6) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); return x; }
  if (PAT == 7) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(s + 2), "l"(x + 1) : "memory"); return x; }
  if (PAT == 8) { uint64_t v = ld8(line); uint64_t *s = line + 4 + 4 * (v & 1); uint64_t x = ld32(s); if (x != 0x1234567ull) asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(s), "l"(x), "l"(x + 1), "l"(x + 2), "l"(x + 3) : "memory"); return x; }
  if (PAT == 9) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned long long *>(s + 2), 1ull); return x; }
  if (PAT == 10) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) x += atomicAdd(reinterpret_cast<unsigned *>(s + 2), 1u); return x; }
  if (PAT == 12) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) { atomicAdd(reinterpret_cast<unsigned *>(s + 2), 1u); atomicAdd(reinterpret_cast<unsigned *>(reinterpret_cast<uintptr_t>(s + 2) ^ 32), 0u); } return x; }
  if (PAT == 13) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) { unsigned *q = reinterpret_cast<unsigned *>(reinterpret_cast<uintptr_t>(s + 2) & ~(uintptr_t)127); atomicAdd(q + 1, 1u); atomicAdd(q + 9, 0u); atomicAdd(q + 17, 0u); atomicAdd(q + 25, 0u); } return x; }
  if (PAT == 14) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(s + 2), "r"(1u), "l"(0x12F0000000000000ull) : "memory"); return x; }
  if (PAT == 11) { uint64_t v = ld8(line); uint64_t *s = line + 2 + 2 * (v & 1); uint64_t x = ld16(s) + ld16(s + 2); if (x != 0x1234567ull) asm volatile("st.global.u64 [%0], %1;" ::"l"(s + 2), "l"(x + 1) : "memory"); return x; }
  return 0;
}
template <int PAT, bool L64, int U>
__global__ void __launch_bounds__(256) sweep_kernel(uint64_t *buf, uint64_t region_lines, uint64_t ops_per_region, uint64_t n_ops, uint64_t seed) {
  uint64_t acc = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < n_ops; g0 += stride * U) {
    uint64_t *line[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t g = (g0 + u * stride) % n_ops;
      const uint64_t region = g / ops_per_region;
      line[u] = buf + (region * region_lines + mix64(seed + g) % region_lines) * 16;
    }
    if (U == 1) acc += sweep_op<PAT, L64>(line[0]);
    else {
      // U independent ops with their stages interleaved by hand for PAT 2 (hdr, slot pieces, red)
      uint64_t v[U], x[U]; uint64_t *s[U];
#pragma unroll
      for (int u = 0; u < U; u++) { if (L64) asm volatile("ld.relaxed.gpu.global.L2::64B.u64 %0, [%1];" : "=l"(v[u]) : "l"(line[u]) : "memory"); else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[u]) : "l"(line[u]) : "memory"); }
#pragma unroll
      for (int u = 0; u < U; u++) { s[u] = line[u] + 2 + 2 * (v[u] & 1); uint64_t a, b, c, d;
        if (L64) { asm volatile("ld.relaxed.gpu.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(s[u]) : "memory"); asm volatile("ld.relaxed.gpu.global.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(s[u] + 2) : "memory"); }
        else { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(s[u]) : "memory"); asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(s[u] + 2) : "memory"); }
        x[u] = a + b + c + d; }
#pragma unroll
      for (int u = 0; u < U; u++) { if (x[u] != 0x1234567ull) atomicAdd(reinterpret_cast<unsigned *>(s[u] + 2), 1u); acc += x[u]; }
    }
  }
  if (acc == 0x9999) buf[0] = acc;
}

// C. groups: the access pattern of the planned streamed insert: the table is cut into GROUPS of
//    lines (a few buckets, tens of KB); a CTA brings one group into shared memory, applies `ops` random
//    header-read + slot-read + shared-memory atomic updates there, and streams the group back.
//    V = 0: LDG.128 -> STS, LDS -> STG.128 through registers.  V = 1: cp.async.bulk (TMA 1-D copies) both ways,
//    double buffered: group i+1 is loaded and group i-1 written back while group i is updated.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void group_ops(uint64_t *lines, uint32_t n_lines, uint32_t ops, uint64_t seed) {
  uint64_t acc = 0;
  for (uint32_t i = threadIdx.x; i < ops; i += blockDim.x) {
    uint64_t *line = lines + (uint64_t)(mix64(seed + i) % n_lines) * 16;
    const uint64_t hdr = line[0];
    uint64_t *slot = line + 1 + 3 * (hdr & 3);  // a 24-byte slot behind the 8-byte header
    acc += slot[0] + slot[1];
    atomicAdd(reinterpret_cast<unsigned *>(slot + 2), 1u);
  }
  if (acc == 0x1234567ull) lines[0] = acc;
}
template <int V>
__global__ void __launch_bounds__(256) groups_kernel(uint64_t *buf, uint64_t n_groups, uint32_t group_bytes, uint32_t ops,
                                                    uint64_t seed) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t n_lines = group_bytes / 128, n_vec = group_bytes / 16;
  if (V == 0) {
    uint4 *sm = reinterpret_cast<uint4 *>(smem_raw);
    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
      uint4 *src = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(buf) + g * group_bytes);
      for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x) sm[i] = src[i];
      __syncthreads();
      group_ops(reinterpret_cast<uint64_t *>(sm), n_lines, ops, seed + g * 7919);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x) src[i] = sm[i];
      __syncthreads();
    }
  } else {
    __shared__ __align__(8) uint64_t mbar[2];
    unsigned char *bufs[2] = {smem_raw, smem_raw + group_bytes};
    if (threadIdx.x == 0) {
      for (int b = 0; b < 2; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[b])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto load = [&](uint64_t g, int b) {  // one thread
      const unsigned char *src = reinterpret_cast<unsigned char *>(buf) + g * group_bytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[b])), "r"(group_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(bufs[b])),
                   "l"(src), "r"(group_bytes), "r"(smem_u32(&mbar[b]))
                   : "memory");
    };
    if (threadIdx.x == 0 && blockIdx.x < n_groups) load(blockIdx.x, 0);
    uint32_t it = 0;
    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x, it++) {
      const int b = it & 1;
      if (threadIdx.x == 0) {
        // the write-back of the group that used the other buffer must have read it before it is refilled
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (g + gridDim.x < n_groups) load(g + gridDim.x, b ^ 1);
      }
      const uint32_t parity = (it >> 1) & 1u;
      uint32_t done = 0;
      for (uint32_t tries = 0; !done; tries++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(&mbar[b])), "r"(parity) : "memory");
        if (tries > (1u << 24)) __trap();  // never hang the GPU box
      }
      group_ops(reinterpret_cast<uint64_t *>(bufs[b]), n_lines, ops, seed + g * 7919);
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned char *dst = reinterpret_cast<unsigned char *>(buf) + g * group_bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(bufs[b])), "r"(group_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
  if (argc > 2) { size_t g = (size_t)atoi(argv[2]); cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("{\"l2fetch_req\":%zu,\"rc\":%d,\"got\":%zu}\n", g, (int)e, got); }
  if (!strcmp(what, "sweep")) {
    const uint64_t total = 100ull << 30;
    uint64_t *buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 0, total));
    const uint64_t n_ops = 180000000ull;
    for (uint64_t region_mb : {64ull}) {
      const uint64_t region_lines = (region_mb << 20) / 128;
      const uint64_t n_regions = total / (region_mb << 20);
      const uint64_t opr = n_ops / n_regions;
      auto run = [&](const char *name, auto kern, int ctas_per_sm) {
        float best = 1e30f;
        for (int rep = 0; rep < 2; rep++) {
          CK(cudaEventRecord(e0));
          kern<<<sms * ctas_per_sm, 256>>>(buf, region_lines, opr, opr * n_regions, 777 + rep);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"bench\":\"sweep\",\"region_MB\":%llu,\"kernel\":\"%s\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"Gops\":%.2f}\n",
               (unsigned long long)region_mb, name, ctas_per_sm, best, n_ops / best / 1e6);
        fflush(stdout);
      };
      run("hdr", sweep_kernel<0, false, 1>, 8);
      run("hdr+2v2 (no red)", sweep_kernel<6, false, 1>, 8);
      for (int c : {2, 3, 4, 6, 8}) run("hdr+2v2+red", sweep_kernel<2, false, 1>, c);
      for (int c : {2, 4, 8}) run("hdr+2v2+red_64B", sweep_kernel<2, true, 1>, c);
      for (int c : {2, 4}) run("hdr+2v2+red U2", sweep_kernel<2, false, 2>, c);
      for (int c : {1, 2}) run("hdr+2v2+red U4", sweep_kernel<2, false, 4>, c);
      for (int c : {2, 4}) run("hdr+2v2+red+red0(other sector of atom)", sweep_kernel<12, false, 1>, c);
      for (int c : {2, 4}) run("hdr+2v2+red all 4 sectors", sweep_kernel<13, false, 1>, c);
      for (int c : {4, 8}) run("hdr+2v2+st8 strong", sweep_kernel<7, false, 1>, c);
      for (int c : {4, 8}) run("hdr+2v2+st8 weak", sweep_kernel<11, false, 1>, c);
      for (int c : {4, 8}) run("hdr+v4+st32", sweep_kernel<8, false, 1>, c);
      for (int c : {4, 8}) run("hdr+2v2+red64", sweep_kernel<9, false, 1>, c);
      for (int c : {4, 8}) run("hdr+2v2+atom32", sweep_kernel<10, false, 1>, c);
      for (int c : {4, 8}) run("hdr+v4+red", sweep_kernel<3, false, 1>, c);
      for (int c : {4, 8}) run("v4+red", sweep_kernel<4, false, 1>, c);
      for (int c : {4, 8}) run("red only", sweep_kernel<5, false, 1>, c);
    }
    CK(cudaFree(buf));
  }
  if (!strcmp(what, "groups")) {
    // ./ubench groups [table_GB] : default 12.5 GB (one GPU's share of the human-scale table at 8 GPUs)
    const uint64_t total = (uint64_t)((argc > 2 ? atof(argv[2]) : 12.5) * (1ull << 30));
    uint64_t *buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 0, total));
    const uint64_t n_ops = 180000000ull;
    for (uint32_t group_bytes : {6400u, 12800u, 25600u, 51200u}) {  // 1, 2, 4, 8 buckets of 50 lines
      const uint64_t n_groups = total / group_bytes;
      const uint32_t ops = (uint32_t)(n_ops / n_groups) + 1;
      auto run = [&](const char *name, auto kern, int ctas_per_sm, uint32_t smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return; }
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
          CK(cudaEventRecord(e0));
          kern<<<sms * ctas_per_sm, 256, smem>>>(buf, n_groups, group_bytes, ops, 4242 + rep);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          if (cudaGetLastError() != cudaSuccess) return;
          best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"bench\":\"groups\",\"table_GB\":%.1f,\"group_bytes\":%u,\"ops_per_group\":%u,\"kernel\":\"%s\",\"ctas_per_sm\":%d,"
               "\"ms\":%.3f,\"GBps_in_plus_out\":%.0f,\"Gops\":%.2f}\n", (double)total / (1ull << 30), group_bytes, ops, name,
               ctas_per_sm, best, 2.0 * (double)n_groups * group_bytes / best / 1e6, (double)n_groups * ops / best / 1e6);
        fflush(stdout);
      };
      for (int c : {2, 4, 8}) if ((uint64_t)c * group_bytes <= 200u * 1024u) run("ldg/stg", groups_kernel<0>, c, group_bytes);
      for (int c : {1, 2, 4}) if ((uint64_t)c * 2 * group_bytes <= 200u * 1024u) run("bulk copy, double buffered", groups_kernel<1>, c, 2 * group_bytes);
    }
    CK(cudaFree(buf));
  }
  if (!strcmp(what, "gran")) {
    const uint64_t total = 32ull << 30;
    uint64_t *buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 0, total));
    const uint64_t n_ops = 64ull << 20;
    for (int pat = 0; pat < 25; pat++) {
      if (argc > 3 && pat != 0 && pat < 17) continue;
      CK(cudaEventRecord(e0));
      gran_kernel<<<sms * 8, 256>>>(buf, total / 128, n_ops, pat, 4242 + pat);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      printf("{\"bench\":\"gran\",\"pat\":%d,\"ms\":%.3f,\"Gops\":%.2f}\n", pat, time_ms(e0, e1), n_ops / time_ms(e0, e1) / 1e6);
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
