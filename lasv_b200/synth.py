"""Deterministic synthetic paired-end reads (bench / test input).

The bytes come from lasv_synth_reads_{host,device} in the CUDA library, which
share one counter-based generator (lasv_b200/csrc/synth.cuh), so a subsample
written to FASTQ for the CPU reference is byte-identical to what the GPU saw.
Presets follow SURVEY.md section 8d.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, asdict

import numpy as np

from . import _capi
from ._capi import SynthParams, check


def q20(prob: float) -> int:
    return int(round(prob * (1 << 20)))


@dataclass
class Workload:
    name: str
    kmer_size: int
    number_bits: int       # L  (--mem_height)
    bucket_size: int       # W  (--mem_width)
    q_thresh: int          # -s / --quality_score_threshold
    genome_len: int
    read_len: int
    coverage: float
    insert: int = 400
    err: float = 0.002
    n_rate: float = 0.0005
    lowq: float = 0.003
    err_lowq_pct: int = 0
    tail_min: int = 0
    tail_max: int = 0
    seed: int = 20

    @property
    def n_pairs(self) -> int:
        return int(self.genome_len * self.coverage / (2 * self.read_len))

    @property
    def cutoff(self) -> int:
        return self.q_thresh + 33  # file_reader.c:919 with cmd_line.c:363

    def params(self, tiling: bool = False) -> SynthParams:
        return SynthParams(self.seed, self.genome_len, self.read_len, self.insert, q20(self.err),
                           q20(self.n_rate), q20(self.lowq), self.err_lowq_pct, self.tail_min,
                           self.tail_max, 1 if tiling else 0,
                           self.read_len - self.kmer_size + 1 if tiling else 0)

    def describe(self) -> dict:
        return asdict(self)


# BASELINE.json configs.  cfg3 is the supernode run on the cfg0 graph.
CFG0 = Workload("cfg0", 53, 22, 100, 5, 63_025_520, 100, 30.0)
CFG1 = Workload("cfg1", 31, 22, 100, 5, 63_025_520, 100, 30.0)
CFG2 = Workload("cfg2", 95, 22, 100, 5, 63_025_520, 150, 30.0, insert=500, err=0.02, lowq=0.003,
                err_lowq_pct=92, tail_min=10, tail_max=40)
CFG4 = Workload("cfg4", 53, 24, 250, 5, 3_100_000_000, 150, 30.0, insert=500, err=0.0005,
                n_rate=0.0001, lowq=0.002, err_lowq_pct=80)
WORKLOADS = {w.name: w for w in (CFG0, CFG1, CFG2, CFG4)}


def scaled(w: Workload, genome_len: int, number_bits: int, name: str | None = None) -> Workload:
    """Same read model on a smaller genome/table (parity-test sizes)."""
    d = asdict(w)
    d.update(genome_len=genome_len, number_bits=number_bits, name=name or f"{w.name}-small")
    return Workload(**d)


def stream_bytes(read_len: int, n_reads: int) -> int:
    return (read_len + 1) * n_reads


def offsets_fixed(read_len: int, n_reads: int) -> np.ndarray:
    return np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len + 1)


def reads_host(p: SynthParams, first_read: int, n_reads: int):
    """(seq, qual, offsets) as numpy arrays, generated on the host."""
    n = stream_bytes(p.read_len, n_reads)
    seq = np.empty(n + 16, dtype=np.uint8)
    qual = np.empty(n + 16, dtype=np.uint8)
    seq[n:] = 10
    qual[n:] = 10
    check(_capi.lib().lasv_synth_reads_host(C.byref(p), first_read, n_reads, seq.ctypes.data,
                                            qual.ctypes.data))
    return seq[:n], qual[:n], offsets_fixed(p.read_len, n_reads)


def reads_device(p: SynthParams, first_read: int, n_reads: int, d_seq, d_qual, stream=None):
    """Fill two device byte buffers (torch uint8 tensors or raw pointers)."""
    ps = d_seq.data_ptr() if hasattr(d_seq, "data_ptr") else d_seq
    pq = d_qual.data_ptr() if hasattr(d_qual, "data_ptr") else d_qual
    check(_capi.lib().lasv_synth_reads_device(C.byref(p), first_read, n_reads, ps, pq, stream))
