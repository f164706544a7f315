// fastq_parse.cuh -- device-side parse of plain four-line FASTQ text (fastq_parse.cu has the scheme).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lasv {

struct FqResult {
  uint32_t n_records;
  uint32_t bad;         // the chunk is not plain four-line FASTQ throughout (or does not fit): the host parser takes it
  uint64_t n_bytes;     // bytes of each output stream (sum of length + 1)
};

// device scratch of one parse in flight
struct FqScratch {
  uint32_t *nl = nullptr;      // [nl_cap] positions of the '\n' bytes; nl_cap >= bytes of the largest chunk
  uint32_t *n_nl = nullptr;
  uint64_t *lens = nullptr;    // [rec_cap + 1]
  FqResult *result = nullptr;
  void *temp = nullptr;
  size_t temp_bytes = 0;
  uint32_t nl_cap = 0, rec_cap = 0;
};

size_t fq_temp_bytes(uint32_t max_bytes, uint32_t rec_cap);
// Queues the parse of d_text[0, n_bytes) (ends with '\n') on st: d_seq / d_qual (d_qual may be NULL) receive the two
// streams, d_offsets[0 .. rec_cap] the exclusive scan of length + 1 (entry n_records = total), s.result the verdict.
cudaError_t fq_parse_chunk(const uint8_t *d_text, uint32_t n_bytes, const FqScratch &s, uint8_t *d_seq, uint8_t *d_qual,
                           uint64_t *d_offsets, uint64_t out_cap, cudaStream_t st);

}  // namespace lasv
