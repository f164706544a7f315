// fastq_parse.cu -- plain four-line FASTQ text -> the two byte streams of the hot path, on the device.
//
// Replaces, for large uncompressed FASTQ files, the per-read text parse of the reference (libs/seq_file/seq_fastq.c:24-150
// behind seq_file.c's readers: header line, sequence line, '+' line, quality line; one read per call, one byte at a
// time for the qualities) and the host threads that did it in round 1 (seq_reader.cpp: fill_batch_fast).  The host
// threads now only copy raw text into pinned memory; the chunk (cut at record boundaries) goes to the device as it
// is and is taken apart there:
//   fq_newlines       positions of all '\n' of the chunk (cub::DeviceSelect over a counting iterator)
//   fq_records_kernel one thread per record i (lines 4i .. 4i+3): checks the SHAPE the host fast path accepts --
//                     '@' header, non-empty sequence line without '\r', a '+' line, a quality line of the sequence's
//                     length -- and writes the record's length; anything else raises `bad`
//   (exclusive scan of length + 1 -> offsets of the reads in the streams)
//   fq_copy_kernel    one warp per record: sequence and quality bytes, verbatim, each followed by '\n'
// A chunk that is not plain four-line FASTQ throughout (multi-line records, CRLF, blank lines, truncation) is handed
// back to the host parser, which implements the reference's full grammar.  Integer/byte work, HBM streaming.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <stdint.h>

#include "fastq_parse.cuh"

namespace lasv {

namespace {

struct IsNewline {
  const uint8_t *text;
  __host__ __device__ bool operator()(const uint32_t &i) const { return text[i] == (uint8_t)'\n'; }
};

constexpr int FQ_THREADS = 256;

// lens[i] = sequence length + 1 of record i (0 beyond the records); res = {n_records, bad, total stream bytes (by
// the scan's consumer)}.  The chunk ends with '\n' (the host appends one if the file does not).
__global__ void __launch_bounds__(FQ_THREADS) fq_records_kernel(const uint8_t *__restrict__ text, const uint32_t n_bytes,
                                                               const uint32_t *__restrict__ nl, const uint32_t *__restrict__ n_nl_p,
                                                               const uint32_t nl_cap, const uint32_t rec_cap,
                                                               uint64_t *__restrict__ lens, FqResult *res) {
  const uint32_t n_nl = *n_nl_p;
  const uint32_t n_rec = n_nl / 4;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    res->n_records = n_rec;
    // every line must end inside the chunk, in whole records, and the index arrays must have held them all
    if (n_nl > nl_cap || (n_nl & 3u) || n_rec > rec_cap || n_nl == 0 || nl[n_nl <= nl_cap ? n_nl - 1 : 0] != n_bytes - 1)
      atomicExch(&res->bad, 1u);
  }
  if (n_nl > nl_cap || n_rec > rec_cap) return;
  for (uint32_t i = blockIdx.x * FQ_THREADS + threadIdx.x; i < rec_cap + 1; i += gridDim.x * FQ_THREADS) {
    if (i >= n_rec) {
      lens[i] = 0;
      continue;
    }
    const uint32_t h0 = i ? nl[4 * i - 1] + 1 : 0;
    const uint32_t e0 = nl[4 * i], e1 = nl[4 * i + 1], e2 = nl[4 * i + 2], e3 = nl[4 * i + 3];
    const uint32_t s0 = e0 + 1, p0 = e1 + 1, q0 = e2 + 1;
    const uint32_t len = e1 - s0, qlen = e3 - q0;
    bool ok = text[h0] == (uint8_t)'@' && len > 0 && text[e1 - 1] != (uint8_t)'\r' && text[p0] == (uint8_t)'+' && qlen == len &&
              text[e3 - 1] != (uint8_t)'\r';
    if (!ok) atomicExch(&res->bad, 1u);
    lens[i] = (uint64_t)len + 1;
  }
}

// offsets = exclusive scan of lens (offsets[n_rec] = total); one warp per record
__global__ void __launch_bounds__(FQ_THREADS) fq_copy_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ nl,
                                                            const uint64_t *__restrict__ offsets, FqResult *res,
                                                            uint8_t *__restrict__ seq, uint8_t *__restrict__ qual,
                                                            const uint64_t out_cap) {
  const uint32_t n_rec = res->n_records;
  if (res->bad) return;
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * FQ_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * FQ_THREADS) >> 5;
  if (warp == 0 && lane == 0) {
    res->n_bytes = offsets[n_rec];
    if (offsets[n_rec] > out_cap) atomicExch(&res->bad, 1u);
  }
  if (offsets[n_rec] > out_cap) return;
  for (uint32_t i = warp; i < n_rec; i += n_warps) {
    const uint32_t s0 = nl[4 * i] + 1, q0 = nl[4 * i + 2] + 1;
    const uint64_t o = offsets[i];
    const uint32_t len = (uint32_t)(offsets[i + 1] - o) - 1;
    for (uint32_t x = lane; x < len; x += 32) {
      seq[o + x] = text[s0 + x];
      if (qual) qual[o + x] = text[q0 + x];
    }
    if (lane == 0) {
      seq[o + len] = (uint8_t)'\n';
      if (qual) qual[o + len] = (uint8_t)'\n';
    }
  }
}

}  // namespace

size_t fq_temp_bytes(uint32_t max_bytes, uint32_t rec_cap) {
  size_t a = 0, b = 0;
  cub::CountingInputIterator<uint32_t> it(0);
  cub::DeviceSelect::If(nullptr, a, it, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)max_bytes, IsNewline{nullptr});
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)(rec_cap + 1));
  return a > b ? a : b;
}

cudaError_t fq_parse_chunk(const uint8_t *d_text, uint32_t n_bytes, const FqScratch &s, uint8_t *d_seq, uint8_t *d_qual,
                           uint64_t *d_offsets, uint64_t out_cap, cudaStream_t st) {
  cudaMemsetAsync(s.result, 0, sizeof(FqResult), st);
  size_t tmp = s.temp_bytes;
  cub::CountingInputIterator<uint32_t> it(0);
  // (cub writes every selected position: nl must hold one entry per byte of the chunk, whatever the text is)
  if (n_bytes > s.nl_cap) return cudaErrorInvalidValue;
  cudaError_t e = cub::DeviceSelect::If(s.temp, tmp, it, s.nl, s.n_nl, (int)n_bytes, IsNewline{d_text}, st);
  if (e != cudaSuccess) return e;
  const int grid = 148 * 4;
  fq_records_kernel<<<grid, FQ_THREADS, 0, st>>>(d_text, n_bytes, s.nl, s.n_nl, s.nl_cap, s.rec_cap, s.lens, s.result);
  tmp = s.temp_bytes;
  e = cub::DeviceScan::ExclusiveSum(s.temp, tmp, s.lens, d_offsets, (int)(s.rec_cap + 1), st);
  if (e != cudaSuccess) return e;
  fq_copy_kernel<<<grid, FQ_THREADS, 0, st>>>(d_text, s.nl, d_offsets, s.result, d_seq, d_qual, out_cap);
  return cudaGetLastError();
}

}  // namespace lasv
