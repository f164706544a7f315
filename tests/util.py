"""Shared helpers for the tests: golden cases, inputs, comparisons."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

from oracle import oracle as O
from lasv_b200 import synth
from lasv_b200._capi import SynthParams

GOLDEN = Path(__file__).resolve().parent / "golden"

EDGE_PE = [f"edge_pe_k{k}_q5" for k in (21, 31, 33, 53, 63, 65, 95)] + ["edge_pe_k31_q0"]
EDGE_SE = ["edge_se_k53_q20", "edge_se_gz_k31_q5", "fasta_se_k31", "fasta_se_k53"]
SYNTH = ["synth_cfg0_k53", "synth_cfg1_k31", "synth_cfg2_k95", "synth_full_k53", "synth_full_k31"]
# the reference's own reader fixtures (libs/seq_file/test), expected output from the compiled reference
REFFIX_FILES = ["multiline.fastq", "ncbi.fastq", "reads.fq", "small.fq", "reads.fa", "small.fa", "reads.txt",
                "small.txt", "single.txt"]
REFFIX = ["reffix_" + f.replace(".", "_") + "_k15" for f in REFFIX_FILES]
ALL_CASES = EDGE_PE + EDGE_SE + SYNTH


def case_meta(name: str) -> dict:
    return json.loads((GOLDEN / f"{name}.json").read_text())


def case_records(name: str) -> np.ndarray:
    return np.load(GOLDEN / f"{name}.records.npy")


def case_files(name: str):
    """(file1, file2 or None) of a file-based golden case."""
    if name.startswith("edge_pe"):
        return GOLDEN / "edge_1.fq", GOLDEN / "edge_2.fq"
    if name.startswith("edge_se_gz"):
        return GOLDEN / "edge_1.fq.gz", None
    if name.startswith("edge_se"):
        return GOLDEN / "edge_1.fq", None
    if name.startswith("fasta"):
        return GOLDEN / "contigs.fa", None
    if name.startswith("reffix_"):
        return GOLDEN / "reffix" / REFFIX_FILES[REFFIX.index(name)], None
    raise KeyError(name)


def synth_params(meta: dict) -> SynthParams:
    return SynthParams(**meta["synth"])


def case_streams(name: str):
    """(seq, qual-or-None, offsets, has_qual) in loader order for any golden case:
    mate 1 then mate 2 of every pair (file_reader.c:636-765)."""
    meta = case_meta(name)
    if "synth" in meta:
        p = synth_params(meta)
        seq, qual, offs = synth.reads_host(p, 0, 2 * meta["n_pairs"])
        return seq, qual, offs
    f1, f2 = case_files(name)
    if str(f1).endswith(".fa"):
        reads = [s.upper() if False else s for s in _fasta_reads(f1)]
        seq, _, offs = O.reads_to_streams(reads)
        return seq, None, offs
    r1, q1 = O.parse_fastq_python(f1)
    if f2 is None:
        return O.reads_to_streams(r1, q1)
    r2, q2 = O.parse_fastq_python(f2)
    reads, quals = [], []
    for a, b, c, d in zip(r1, r2, q1, q2):
        reads += [a, b]
        quals += [c, d]
    return O.reads_to_streams(reads, quals)


def _fasta_reads(path):
    reads, cur = [], None
    for line in Path(path).read_text().splitlines():
        if line.startswith(">"):
            if cur is not None:
                reads.append(cur)
            cur = ""
        elif cur is not None:
            cur += line.strip()
    if cur is not None:
        reads.append(cur)
    return reads


def cutoff_of(meta: dict) -> int:
    return meta["q_thresh"] + 33


def assert_same_records(a: np.ndarray, b: np.ndarray, what: str = ""):
    a, b = O.sort_records(a), O.sort_records(b)
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} nodes"
    assert np.array_equal(a["kmer"], b["kmer"]), f"{what}: key sets differ"
    bad = np.nonzero(a["covg"] != b["covg"])[0]
    assert bad.size == 0, f"{what}: coverage differs at {bad[:5]}"
    bad = np.nonzero(a["edges"] != b["edges"])[0]
    assert bad.size == 0, f"{what}: edges differ at {bad[:5]}: {a['edges'][bad[:5]]} vs {b['edges'][bad[:5]]}"


def oracle_build(k, L, W, seq, qual, offs, cutoff, paired=False):
    g = O.OracleGraph(k, L, W)
    ok = g.load_reads(seq, qual, offs, cutoff, paired=paired)
    return g, ok
