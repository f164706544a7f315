#!/usr/bin/env python
"""End-to-end speed of the file loaders (FASTQ text -> graph): the parse is the slow end."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from lasv_b200 import DBGraph, synth
from oracle import oracle as O   # FASTQ writer only (test infrastructure)

w = synth.CFG4
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 22
tmp = Path(os.environ.get("TMPDIR", "/tmp"))
seq, qual, _ = synth.reads_host(w.params(), 0, 2 * n_pairs)
s2, q2 = seq.reshape(-1, w.read_len + 1), qual.reshape(-1, w.read_len + 1)
O.write_fastq_fixed(tmp / "ls_1.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len)
O.write_fastq_fixed(tmp / "ls_2.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len)
size = os.path.getsize(tmp / "ls_1.fq") + os.path.getsize(tmp / "ls_2.fq")
with DBGraph(w.kmer_size, L, 100) as g:
    for rep in range(2):
        g.clear()
        t = time.perf_counter()
        st = g.load_pe_seq_data(tmp / "ls_1.fq", tmp / "ls_2.fq", w.cutoff)
        dt = time.perf_counter() - t
    print(json.dumps({"what": "lasv_graph_load_pe_files, uncompressed FASTQ pair", "pairs": n_pairs, "fastq_bytes": size,
                      "seconds": dt, "fastq_MBps": size / dt / 1e6, "kmers_per_s": st["kmers_inserted"] / dt}))
