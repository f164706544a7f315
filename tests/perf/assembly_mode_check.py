#!/usr/bin/env python
"""The reference's `--mode assembly` (read-span linkage table + BFS contig assembly: build_linkage_from_files.c:411-499,
linkage_operations.c:167-541) is a CONSUMER of the graph: it runs unchanged -- the reference's own main() and code --
on the table the GPU loader fills (integration/_build/dB_graph_gpu_*: only the two loader symbols come from
liblasv_b200.so).  This script runs it next to the all-CPU reference binary on the same reads and read.span file and
compares contigs.fa.

    python tests/perf/assembly_mode_check.py [genome_len] [n_reads] [seed] [--branching]
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from lasv_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def make_inputs(tmp: Path, genome_len: int, n_reads: int, seed: int, L: int = 12):
    w = synth.scaled(synth.CFG0, genome_len, L)
    w.seed = seed
    p = w.params()
    seq, qual, _ = synth.reads_host(p, 0, n_reads)
    rl, k = w.read_len, w.kmer_size
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp / "l").write_text(f"{tmp / 'a.fq'}\n")
    (tmp / "r").write_text(f"{tmp / 'b.fq'}\n")
    # read.span: one line per pair, k-mers of both mates back to back (what the pipeline's perl script writes for the
    # branching nodes a pair spans; any k-mers of the graph exercise the same code)
    with open(tmp / "read.span", "w") as f:
        for i in range(0, len(s2), 2):
            parts = [bytes(s2[i][:k]).decode(), bytes(s2[i][rl - k:rl]).decode(), bytes(s2[i + 1][:k]).decode()]
            if all(set(x) <= set("ACGT") for x in parts):
                f.write("".join(parts) + "\n")
    return w


def make_branching_inputs(tmp: Path, genome_len: int, n_pairs: int, seed: int, L: int = 12):
    """A diploid genome with SNPs (bubbles) and a duplicated segment (a repeat): the graph keeps branching nodes after
    the cleaning, so the BFS has to consult the linkage table.  Error-free 2x100 bp pairs, insert 400."""
    import numpy as np
    rng = np.random.default_rng(seed)
    w = synth.scaled(synth.CFG0, genome_len, L)
    rl, k, ins = w.read_len, w.kmer_size, 400
    g = rng.integers(0, 4, genome_len)
    a, b = genome_len // 5, 3 * genome_len // 5
    g[b: b + 300] = g[a: a + 300]                      # the repeat
    h2 = g.copy()
    for pos in range(350, genome_len - 350, 701):      # heterozygous SNPs
        if not (a - 60 <= pos < a + 360 or b - 60 <= pos < b + 360):
            h2[pos] = (h2[pos] + 1 + rng.integers(0, 3)) % 4
    haps = ["".join("ACGT"[x] for x in g), "".join("ACGT"[x] for x in h2)]
    comp = str.maketrans("ACGT", "TGCA")
    with open(tmp / "a.fq", "w") as fa, open(tmp / "b.fq", "w") as fb, open(tmp / "read.span", "w") as sp:
        for i in range(n_pairs):
            hp = haps[int(rng.integers(0, 2))]
            st = int(rng.integers(0, genome_len - ins))
            m1, m2 = hp[st: st + rl], hp[st + ins - rl: st + ins].translate(comp)[::-1]
            if rng.integers(0, 2):
                m1, m2 = m2, m1
            fa.write(f"@r{i}/1\n{m1}\n+\n{'I' * rl}\n")
            fb.write(f"@r{i}/2\n{m2}\n+\n{'I' * rl}\n")
            sp.write(m1[:k] + m1[rl - k:] + m2[:k] + m2[rl - k:] + "\n")
    (tmp / "l").write_text(f"{tmp / 'a.fq'}\n")
    (tmp / "r").write_text(f"{tmp / 'b.fq'}\n")
    return w


def run(exe, tmp: Path, w, sub: str, env=None):
    d = tmp / sub
    d.mkdir(exist_ok=True)
    cmd = [str(exe), "--kmer_size", str(w.kmer_size), "--mem_height", str(w.number_bits), "--mem_width", str(w.bucket_size),
           "--pe_list", f"{tmp / 'l'},{tmp / 'r'}", "--quality_score_threshold", "5", "--mode", "assembly",
           "--read_span_file", str(tmp / "read.span"), "--remove_low_coverage_supernodes", "2", "--insert_size", "400",
           "--max_read_len", str(w.read_len)]
    p = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=1200, env=env)
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-1500:] + p.stdout[-1500:])
    return (d / "contigs.fa").read_bytes()


def contig_set(fa: bytes):
    return sorted(O.canonical_seq_set([l for l in fa.decode().splitlines() if l and not l.startswith(">")]))


if __name__ == "__main__":
    import os
    import tempfile
    pos = [a_ for a_ in sys.argv[1:] if not a_.startswith("--")]
    genome_len = int(pos[0]) if len(pos) > 0 else 20000
    n_reads = int(pos[1]) if len(pos) > 1 else 6000
    seed = int(pos[2]) if len(pos) > 2 else 11
    tmp = Path(tempfile.mkdtemp(prefix="asm_"))
    w = (make_branching_inputs(tmp, genome_len, n_reads // 2, seed) if "--branching" in sys.argv
         else make_inputs(tmp, genome_len, n_reads, seed))
    maxk = O.MAXK_OF_N[O.words_per_kmer(w.kmer_size)]
    ref = run(O.REF_DIR / f"dB_graph_{maxk}", tmp, w, "ref")
    out = {"genome_len": genome_len, "reads": n_reads, "seed": seed, "ref_contigs": ref.count(b">"), "ref_bytes": len(ref)}
    for name, env in (("gpu", None), ("gpu_2shards", dict(os.environ, LASV_B200_DEVICES="0,0"))):
        got = run(ROOT / "integration" / "_build" / f"dB_graph_gpu_{maxk}", tmp, w, name, env)
        out[name] = {"identical_bytes": got == ref, "same_contig_set": contig_set(got) == contig_set(ref),
                     "contigs": got.count(b">"), "bytes": len(got)}
    print(json.dumps(out))
