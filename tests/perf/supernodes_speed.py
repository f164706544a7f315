#!/usr/bin/env python
"""`--output_supernodes` on the device graph vs the compiled reference on the same `.ctx`
(SURVEY 8f rank 2): builds a cfg0-like graph (100 bp PE, 30x, 0.2 % errors, k=53) from
device-generated reads, times DBGraph.output_supernodes, then lets the reference reload the
dumped `.ctx` and print its supernodes, and compares the two FASTA files byte for byte.

    python tests/perf/supernodes_speed.py [genome_len] [L] [--no-ref] [--clean T]
"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch  # noqa: E402
from lasv_b200 import DBGraph, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (reference runner only: test infrastructure)

_argv = list(sys.argv[1:])
if "--clean" in _argv:
    del _argv[_argv.index("--clean") + 1]
args = [a for a in _argv if not a.startswith("--")]
genome_len = int(args[0]) if len(args) > 0 else 2_000_000
L = int(args[1]) if len(args) > 1 else 18
with_ref = "--no-ref" not in sys.argv
# --clean T: run the cleaning (tip clipping + low-coverage supernodes, threshold T) on the device first -- a cleaned
# graph is a few very long paths, the case for LASV_B200_SN_RANKING=1 (list ranking + one writer thread per node)
clean_T = int(sys.argv[sys.argv.index("--clean") + 1]) if "--clean" in sys.argv else None

w = synth.scaled(synth.CFG0, genome_len, L)
tmp = Path(os.environ.get("TMPDIR", "/tmp")) / "sn_speed"
tmp.mkdir(parents=True, exist_ok=True)
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
p = w.params()
n_reads = 2 * w.n_pairs
batch = 1 << 20
stride = w.read_len + 1
d_seq = torch.empty(batch * stride + 16, dtype=torch.uint8, device="cuda")
d_qual = torch.empty_like(d_seq)
out = {"what": "output_supernodes", "genome_len": genome_len, "k": w.kmer_size, "L": L, "W": w.bucket_size,
       "reads": n_reads}
with DBGraph(w.kmer_size, L, w.bucket_size) as g:
    done = 0
    while done < n_reads:
        nr = min(batch, n_reads - done)
        synth.reads_device(p, done, nr, d_seq, d_qual, st)
        g.load_reads_device(d_seq, d_qual, nr * stride, None, 0, w.cutoff, st)
        done += nr
    torch.cuda.synchronize()
    out["unique_kmers"] = g.unique_kmers
    out["ranking"] = os.environ.get("LASV_B200_SN_RANKING", "0")
    if clean_T is not None:
        t = time.perf_counter()
        out["clean"] = g.clean(clean_T)
        out["clean_seconds"] = time.perf_counter() - t
    os.environ["LASV_B200_SN_TRACE"] = "1"

    for rep in range(2):
        t = time.perf_counter()
        stats = g.output_supernodes(tmp / "gpu.fa", max_length=10000)
        out["gpu_seconds"] = time.perf_counter() - t
    del os.environ["LASV_B200_SN_TRACE"]
    out["stats"] = stats
    out["gpu_nodes_per_s"] = stats["n_nodes"] / out["gpu_seconds"]
    g.dump_binary(tmp / "g.ctx", 100, 1000)
if with_ref and O.ref_available(w.kmer_size):
    exe = str(O.ref_binary(w.kmer_size))
    base = [exe, "--kmer_size", str(w.kmer_size), "--mem_height", str(L), "--mem_width", str(w.bucket_size),
            "--multicolour_bin", str(tmp / "g.ctx")]
    t = time.perf_counter()
    subprocess.run(base, cwd=tmp, check=True, capture_output=True)
    t_load = time.perf_counter() - t
    (tmp / "ref.fa").unlink(missing_ok=True)
    t = time.perf_counter()
    subprocess.run(base + ["--output_supernodes", str(tmp / "ref.fa"), "--max_var_len", str(1 << 24)], cwd=tmp,
                   check=True, capture_output=True)
    t_all = time.perf_counter() - t
    out["ref_reload_seconds"] = t_load
    out["ref_supernodes_seconds"] = t_all - t_load
    out["ref_nodes_per_s"] = stats["n_nodes"] / max(t_all - t_load, 1e-9)
    out["fasta_identical_to_reference"] = (tmp / "ref.fa").read_bytes() == (tmp / "gpu.fa").read_bytes()
    out["fasta_bytes"] = (tmp / "gpu.fa").stat().st_size
print(json.dumps(out))
