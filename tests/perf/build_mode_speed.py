#!/usr/bin/env python
"""What `dB_graph --mode build --remove_low_coverage_supernodes T` does after the load
(run_laSV.sh:178) on the GPU next to the compiled reference:
tip clipping + low-coverage supernode removal (lasv_ref_hash_table_clean), branch printing
(lasv_ref_hash_table_write_branches) and the supernodes of the cleaned graph, all on the
caller's reference-layout table, i.e. INCLUDING the host<->device copies of the table.

Builds a cfg0-like graph (100 bp PE, 30x, 0.2 % errors, k=53) from device-generated reads,
exports the reference table, times the three calls, then lets the reference reload the raw
`.ctx` (same slot order) and run the same steps, and compares branches.fa and the cleaned
`.ctx` byte for byte / record for record.

    python tests/perf/build_mode_speed.py [genome_len] [L] [threshold] [--no-ref]
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import torch  # noqa: E402
from lasv_b200 import DBGraph, _capi, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (reference runner only: test infrastructure)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
genome_len = int(args[0]) if len(args) > 0 else 2_000_000
L = int(args[1]) if len(args) > 1 else 18
thresh = int(args[2]) if len(args) > 2 else 2
with_ref = "--no-ref" not in sys.argv
w = synth.scaled(synth.CFG0, genome_len, L)
tmp = Path(os.environ.get("TMPDIR", "/tmp")) / "build_mode_speed"
tmp.mkdir(parents=True, exist_ok=True)
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
p = w.params()
n_reads = 2 * w.n_pairs
batch = 1 << 20
stride = w.read_len + 1
d_seq = torch.empty(batch * stride + 16, dtype=torch.uint8, device="cuda")
d_qual = torch.empty_like(d_seq)
out = {"what": "mode build after the load: clean + branches (+ supernodes)", "genome_len": genome_len,
       "k": w.kmer_size, "L": L, "W": w.bucket_size, "threshold": thresh, "reads": n_reads}
with DBGraph(w.kmer_size, L, w.bucket_size) as g:
    done = 0
    while done < n_reads:
        nr = min(batch, n_reads - done)
        synth.reads_device(p, done, nr, d_seq, d_qual, st)
        g.load_reads_device(d_seq, d_qual, nr * stride, None, 0, w.cutoff, st)
        done += nr
    torch.cuda.synchronize()
    out["unique_kmers"] = g.unique_kmers
    t = time.perf_counter()
    out["device_graph_branches"] = g.output_branches(tmp / "raw_branches.fa")     # on the device graph, raw
    out["device_graph_branches_seconds"] = time.perf_counter() - t
    g.dump_binary(tmp / "g.ctx", 100, 1000)
    table, nxt = g.export_table()
del d_seq, d_qual
torch.cuda.empty_cache()

lib = _capi.lib()
ht = _capi.RefHashTable(w.kmer_size, 1 << L, w.bucket_size, table.ctypes.data, nxt.ctypes.data, None,
                        int(out["unique_kmers"]), 15)
cs, bs, ss = _capi.CleanStats(), _capi.BranchStats(), _capi.SupernodeStats()
t = time.perf_counter()
_capi.check(lib.lasv_ref_hash_table_clean(C.byref(ht), _capi.CLEAN_TIPS | _capi.CLEAN_SUPERNODES, thresh, C.byref(cs)))
out["gpu_clean_seconds"] = time.perf_counter() - t
t = time.perf_counter()
_capi.check(lib.lasv_ref_hash_table_write_branches(C.byref(ht), str(tmp / "gpu_branches.fa").encode(), C.byref(bs)))
out["gpu_branches_seconds"] = time.perf_counter() - t
t = time.perf_counter()
_capi.check(lib.lasv_ref_hash_table_write_supernodes(C.byref(ht), str(tmp / "gpu_sup.fa").encode(), 10000, C.byref(ss)))
out["gpu_supernodes_seconds"] = time.perf_counter() - t
out["clean"], out["branches"], out["supernodes"] = cs.as_dict(), bs.as_dict(), ss.as_dict()
out["nodes_left"] = int(((table["status"] == 1) | (table["status"] == 2)).sum())

if with_ref and O.ref_available(w.kmer_size):
    exe = str(O.ref_binary(w.kmer_size))
    base = [exe, "--kmer_size", str(w.kmer_size), "--mem_height", str(L), "--mem_width", str(w.bucket_size),
            "--multicolour_bin", str(tmp / "g.ctx")]

    def timed(extra):
        t0 = time.perf_counter()
        subprocess.run(base + extra, cwd=tmp, check=True, capture_output=True)
        return time.perf_counter() - t0

    t_load = timed([])
    (tmp / "ref_clean.ctx").unlink(missing_ok=True)
    t_clean = timed(["--remove_low_coverage_supernodes", str(thresh)])
    t_build = timed(["--remove_low_coverage_supernodes", str(thresh), "--mode", "build", "--dump_binary",
                     str(tmp / "ref_clean.ctx")])
    out["ref_reload_seconds"] = t_load
    out["ref_clean_seconds"] = t_clean - t_load
    out["ref_branches_and_dump_seconds"] = t_build - t_clean
    out["branches_identical_to_reference"] = (tmp / "branches.fa").read_bytes() == (tmp / "gpu_branches.fa").read_bytes()
    ref = O.parse_ctx(tmp / "ref_clean.ctx")["records"]
    occ = table[(table["status"] == 1) | (table["status"] == 2)]
    out["cleaned_records_identical_to_reference"] = bool(
        len(ref) == len(occ) and np.array_equal(ref["kmer"], occ["kmer"]) and np.array_equal(ref["covg"], occ["coverage"])
        and np.array_equal(ref["edges"], occ["edges"]))
print(json.dumps(out))
