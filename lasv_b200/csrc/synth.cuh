// synth.cuh -- deterministic synthetic paired-end reads (counter-based, so the
// device kernel, the host loop and any subsample produce identical bytes).
// Bench / test INPUT only; nothing here models reference behaviour.
//
// Genome: iid uniform bases, base i = 2 bits of mix64(seed_g + i/32).
// Pair p (reads 2p, 2p+1): fragment start s uniform, strand bit; the mate that
// reads the forward strand is genome[s, s+R), the other is the reverse
// complement of genome[s+insert-R, s+insert).  Per base: substitution with
// probability err, N with probability n, isolated low quality (Phred 2..5) with
// probability lowq, a fraction err_lowq_pct of substituted bases forced to low
// quality, optional Phred-2 tail at the 3' end; all other qualities drawn from
// {30,35,38,40}.  Qualities are Phred+33.
// Layout: fixed stride read_len+1, a '\n' after every read in both streams.
#pragma once
#include "kmer_core.cuh"
#include "graph_kernels.cuh"

namespace lasv {

LASV_HD uint32_t synth_genome_base(uint64_t seed, uint64_t i) {
  uint64_t w = mix64((seed ^ 0x5851F42D4C957F2Dull) + (i >> 5));
  return (uint32_t)(w >> (2 * (i & 31))) & 3u;
}

// One byte of the stream: global read index r, position i in [0, read_len].
LASV_HD void synth_byte(const SynthParams &p, uint64_t r, uint32_t i, uint8_t &s, uint8_t &q) {
  if (i == p.read_len) { s = '\n'; q = '\n'; return; }
  uint32_t base;
  if (p.tiling) {
    uint64_t g = r * (uint64_t)p.tiling_step + i;
    if (g >= p.genome_len) { s = 'N'; q = 'I'; return; }
    s = (uint8_t)"ACGT"[synth_genome_base(p.seed, g)];
    q = 'I';
    return;
  }
  const uint64_t pair = r >> 1;
  const uint32_t mate = (uint32_t)(r & 1);
  const uint64_t hp = mix64((p.seed ^ 0xA0761D6478BD642Full) + pair * 2654435761ull);
  const uint64_t start = (hp >> 1) % (p.genome_len - p.insert + 1);
  const uint32_t strand = (uint32_t)(hp & 1);
  if (strand == mate) base = synth_genome_base(p.seed, start + i);
  else base = 3u - synth_genome_base(p.seed, start + p.insert - 1 - i);
  const uint64_t h = mix64((p.seed ^ 0xE7037ED1A0B428DBull) + r * 1024ull + i);
  const uint64_t h2 = mix64(h);
  bool lowq = false;
  if ((uint32_t)(h & 0xFFFFFu) < p.err_q20) {
    base = (base + 1u + ((uint32_t)(h >> 20) & 3u) % 3u) & 3u;
    if ((uint32_t)((h2 >> 8) % 100ull) < p.err_lowq_pct) lowq = true;
  }
  if ((uint32_t)((h >> 42) & 0xFFFFFu) < p.lowq_q20) lowq = true;
  uint32_t phred;
  if (lowq) {
    phred = 2u + (uint32_t)(h2 & 3u);
  } else {
    const uint32_t qs[4] = {30u, 35u, 38u, 40u};
    phred = qs[(uint32_t)(h >> 62) & 3u];
  }
  if (p.tail_max) {
    const uint64_t hr = mix64((p.seed ^ 0x8EBC6AF09C88C6E3ull) + r);
    const uint32_t tail = p.tail_min + (uint32_t)(hr % (uint64_t)(p.tail_max - p.tail_min + 1u));
    if (i + tail >= p.read_len) phred = 2u;
  }
  s = (uint8_t)"ACGT"[base];
  if ((uint32_t)((h >> 22) & 0xFFFFFu) < p.n_q20) s = 'N';
  q = (uint8_t)(33u + phred);
}

}  // namespace lasv
