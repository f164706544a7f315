#!/usr/bin/env python
"""Regenerate tests/golden/*: inputs plus the outputs of the COMPILED, UNMODIFIED
reference (oracle/_ref/dB_graph_{31,63,95}, built from /root/reference by
oracle/Makefile).  Run in the build container (the GPU box has no
/root/reference, but it does receive oracle/_ref prebuilt and these fixtures).

    python tests/golden/make_golden.py

Every case stores, next to its input FASTQ/FASTA files (or the synth parameters
that regenerate them bit-exactly), the reference's
  * sorted `.ctx` records  (kmer words, coverage, edges)   -> <case>.records.npy
  * `.ctx` header bytes, loader statistics, "tries" histogram, supernode set
                                                             -> <case>.json
"""
from __future__ import annotations

import json
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402
from lasv_b200 import synth  # noqa: E402

HERE = Path(__file__).resolve().parent

# ---- handwritten edge-case reads ------------------------------------------
EDGE_1 = """@e1 plain 80bp all good
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@e2 N in the middle
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGNTCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGGCTAACGATCGATCGGCTAG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@e3 lower case and a low-quality base (Phred 5 = '&' is cut by -s 5, Phred 6 = ''' is kept)
acgttgcatgccgataggctaacgttagccatgcatcggatcgattacggctagctaggatccgatcgtagctagctaggctaacgatcgatcggctag
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII&IIIIIIIIIIIIIIIIIIIIIIIIII'IIIIIIIIII
@e4 too short for any k-mer
ACGTACGTAC
+
IIIIIIIIII
@e5 multi-line sequence and quality, quality line starting with @
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGG
ATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGG
+e5
@IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@e6 homopolymer run (same k-mer in neighbouring windows) then unique tail
AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAACGTCA
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@e7 invalid characters (treated as N) and Phred 0
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAG.ATCCGATCGTAGCTAGCTAGGCTAACGATCGATCGGCTAX
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII!IIIIIIIIIIIII
@e8 all bad
NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
"""

EDGE_2 = """@f1 reverse complement of e1 (same nodes, opposite orientation)
CCTAGCTAGCTACGATCGGATCCTAGCTAGCCGTAATCGATCCGATGCATGGCTAACGTTAGCCTATCGGCATGCAACGT
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f2 palindromic stretch
ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f3 exactly 53 bases
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTA
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f4 low quality tail
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGGCTAACGATCGATCGGCTAG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII###############
@f5 tandem repeat of period 3
ACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACGACG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f6 second copy of e2 without the N
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGGCTAACGATCGATCGGCTAG
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f7 G and T rich
GGGGTTTTGGGGTTTTGTGTGTGTGGTTGGTTGGGTTTGGGTTTGTTGTTGGTGGTGTTTTGGGGTGTGGTTTTGGGTGGGTTGTGTTGGGGTTTTGGT
+
IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII
@f8 short
ACGT
+
IIII
"""

FASTA_SE = """>c1
ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGGCTAACGATCGATC
GGCTAGCTAAGCTTACGATCGNNACGATTAGCGCGCGATATCGCGATTAGCCGATATCGGATTAGCTAGCTAGGGATCTCGATCGATATCGGAT
>c2
acgttgcatgccgataggctaacgttagccatgcatcggatcgattacggctagctagg
"""


REF_FIXTURES = ["multiline.fastq", "ncbi.fastq", "reads.fq", "small.fq", "reads.fa", "small.fa", "reads.txt",
                "small.txt", "single.txt"]


def ref_case(name, k, L, W, q, fq1, fq2=None, supernodes=False, extra=None):
    wd = Path(tempfile.mkdtemp(prefix="lasv_golden_"))
    try:
        res = O.run_reference(k, L, W, fq1, fq2, q_thresh=q, workdir=wd, want_supernodes=supernodes)
    finally:
        pass
    ctx = res["ctx"]
    recs = O.sort_records(ctx["records"])
    np.save(HERE / f"{name}.records.npy", recs)
    out = res["stdout"]

    def grab(label):
        for line in out.splitlines():
            if label in line:
                return int(line.split(":")[-1])
        return None

    tries = {}
    for line in out.splitlines():
        s = line.strip()
        if s.startswith("tries "):
            r, n = s[len("tries "):].split(":")
            tries[int(r)] = int(n)
    meta = {
        "k": k, "L": L, "W": W, "q_thresh": q,
        "n_records": int(len(recs)),
        "inserts": res["inserts"], "tries": tries,
        "bad_reads": grab("Number of bad reads"),
        "bases_parsed": grab("Total bases parsed"),
        "bases_loaded": grab("Total bases passing filters"),
        "mean_read_len_after_filters": grab("Mean read length after filters applied"),
        "ctx_mean_read_len": ctx["mean_read_len"], "ctx_total_seq": ctx["total_seq"],
        "ctx_header_hex": ctx["header"].hex(),
    }
    if supernodes:
        meta["supernodes"] = O.canonical_seq_set(res["supernodes"])
    if extra:
        meta.update(extra)
    (HERE / f"{name}.json").write_text(json.dumps(meta, indent=1) + "\n")
    shutil.rmtree(wd, ignore_errors=True)
    print(f"{name}: {len(recs)} records, {res['inserts']} inserts, bad={meta['bad_reads']}")


def synth_case(name, base, genome_len, L, W, n_pairs, k=None, supernodes=False):
    w = synth.scaled(base, genome_len, L)
    if k is not None:
        w.kmer_size = k
    w.bucket_size = W
    p = w.params()
    seq, qual, _ = synth.reads_host(p, 0, 2 * n_pairs)
    stride = w.read_len + 1
    s2 = seq.reshape(-1, stride)
    q2 = qual.reshape(-1, stride)
    tmp = Path(tempfile.mkdtemp(prefix="lasv_golden_fq_"))
    f1, f2 = tmp / "r_1.fq", tmp / "r_2.fq"
    O.write_fastq_fixed(f1, s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len, "p")
    O.write_fastq_fixed(f2, s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len, "p")
    ref_case(name, w.kmer_size, L, W, w.q_thresh, f1, f2, supernodes=supernodes,
             extra={"synth": {k_: int(getattr(p, k_)) for k_, _ in p._fields_}, "n_pairs": n_pairs,
                    "workload": w.describe()})
    shutil.rmtree(tmp, ignore_errors=True)


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref is missing: run `make -C oracle` in the container that has /root/reference")
    (HERE / "edge_1.fq").write_text(EDGE_1)
    (HERE / "edge_2.fq").write_text(EDGE_2)
    (HERE / "contigs.fa").write_text(FASTA_SE)
    import gzip
    with gzip.open(HERE / "edge_1.fq.gz", "wb") as f:
        f.write(EDGE_1.encode())

    for k in (21, 31, 33, 53, 63, 65, 95):
        ref_case(f"edge_pe_k{k}_q5", k, 10, 20, 5, HERE / "edge_1.fq", HERE / "edge_2.fq", supernodes=True)
    ref_case("edge_pe_k31_q0", 31, 10, 20, 0, HERE / "edge_1.fq", HERE / "edge_2.fq")
    ref_case("edge_se_k53_q20", 53, 10, 20, 20, HERE / "edge_1.fq")
    ref_case("edge_se_gz_k31_q5", 31, 10, 20, 5, HERE / "edge_1.fq.gz")
    ref_case("fasta_se_k31", 31, 10, 20, 5, HERE / "contigs.fa")
    ref_case("fasta_se_k53", 53, 10, 20, 0, HERE / "contigs.fa")

    # the fixtures of the reference's own (assert-free) reader smoke test, libs/seq_file/test/: multi-line
    # records, quality lines starting with '@', '+name' separator lines, lower case, FASTA, plain text
    # with empty lines.  The inputs are copied as test data; the expected output is the reference's.
    ref_fix = Path(O.REFERENCE_ROOT) / "libs" / "seq_file" / "test" if hasattr(O, "REFERENCE_ROOT") else \
        Path("/root/reference/libs/seq_file/test")
    (HERE / "reffix").mkdir(exist_ok=True)
    for fname in REF_FIXTURES:
        shutil.copyfile(ref_fix / fname, HERE / "reffix" / fname)
        ref_case("reffix_" + fname.replace(".", "_") + "_k15", 15, 8, 20, 5, HERE / "reffix" / fname)

    # synthetic PE reads (regenerated bit-exactly from the stored synth parameters)
    synth_case("synth_cfg0_k53", synth.CFG0, 4000, 10, 40, 400, supernodes=True)
    synth_case("synth_cfg1_k31", synth.CFG1, 4000, 10, 40, 400, supernodes=True)
    synth_case("synth_cfg2_k95", synth.CFG2, 4000, 10, 40, 300, supernodes=True)
    # ~90 % full table: rehash chains exercised (hash_table.c:309-317 path, not overflowing)
    synth_case("synth_full_k53", synth.CFG0, 4000, 6, 125, 400)
    synth_case("synth_full_k31", synth.CFG1, 4000, 7, 62, 400)


if __name__ == "__main__":
    main()
