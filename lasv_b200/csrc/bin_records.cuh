// bin_records.cuh -- device-side access to the bin records of the partitioned insert (graph_kernels.cuh, BinView):
// one vector store / load per record.  Shared by the partition kernel, the random-access insert (graph_kernels.cu)
// and the streamed insert (streamed_insert.cu).
#pragma once
#include "graph_kernels.cuh"

namespace lasv {

template <int N>
__device__ __forceinline__ void store_bin_record(uint64_t *dst, uint32_t rec_words, const uint64_t (&key)[N],
                                                 uint32_t c, uint32_t edge, uint32_t count) {
  if (N == 1) {
    asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(key[0]), "l"(bin_aux(c, edge, count)) : "memory");
  } else if (N == 2 && rec_words == 2) {  // k <= 56: edges and count ride in the top key word (count <= 32 per round)
    const uint64_t w0 = key[0] | ((uint64_t)(edge & 0xFFu) << 48) | ((uint64_t)(count & 0xFFu) << 56);
    asm volatile("st.global.v2.u64 [%0], {%1,%2};" ::"l"(dst), "l"(w0), "l"(key[1 % N]) : "memory");
  } else {
    const uint64_t aux = bin_aux(c, edge, count);
    const uint64_t w2 = (N == 2) ? aux : key[2 % N], w3 = (N == 2) ? 0ull : aux;
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "l"(key[0]), "l"(key[1 % N]), "l"(w2), "l"(w3)
                 : "memory");
  }
}
template <int N>
__device__ __forceinline__ void load_bin_record(const uint64_t *src, uint32_t rec_words, uint64_t (&key)[N],
                                                uint32_t &c, uint32_t &edge, uint32_t &count) {
  uint64_t a, b, x = 0, y = 0;
  if (N == 1 || rec_words == 2) {
    asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src));
    if (N == 1) {
      key[0] = a;
      c = (uint32_t)b;
      edge = (uint32_t)(b >> 32) & 0xFFu;
      count = (uint32_t)(b >> 40);
    } else {
      key[0] = a & 0x0000FFFFFFFFFFFFull;
      key[1 % N] = b;
      edge = (uint32_t)(a >> 48) & 0xFFu;
      count = (uint32_t)(a >> 56);
      c = lookup3_key<N>(key, 0).c;
    }
    return;
  }
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(x), "=l"(y) : "l"(src));
  key[0] = a;
  key[1 % N] = b;
  if (N == 3) key[2 % N] = x;
  const uint64_t aux = (N == 2) ? x : y;
  c = (uint32_t)aux;
  edge = (uint32_t)(aux >> 32) & 0xFFu;
  count = (uint32_t)(aux >> 40);
}


}  // namespace lasv
