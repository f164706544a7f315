#!/usr/bin/env python
"""Experiment: how fast does insert_records_kernel run when the records of a batch
are grouped by table region (bucket >> shift) before insertion?  Scaffolding for
the partitioned-insert design: grouping is done with torch.sort here."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from lasv_b200 import DBGraph, _capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # steps of records per insert call
w = synth.WORKLOADS[wl]
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
lib = _capi.lib()
pairs = 1 << 20
reads = 2 * pairs
stride = w.read_len + 1
nb = reads * stride
g = DBGraph(w.kmer_size, w.number_bits, w.bucket_size)
seq = torch.empty(nb + 16, dtype=torch.uint8, device="cuda")
qual = torch.empty_like(seq)
if wl == "cfg4":
    pt = w.params(tiling=True)
    n_t = (w.genome_len + pt.tiling_step - 1) // pt.tiling_step
    done = 0
    while done < n_t:
        nr = min(reads, n_t - done)
        synth.reads_device(pt, done, nr, seq, qual, st)
        g.load_reads_device(seq, qual, nr * stride, None, 0, w.cutoff, st)
        done += nr
    torch.cuda.synchronize()
p = w.params()
kpr = w.read_len - w.kmer_size + 1
cap = R * reads * kpr + 8
rw = 2 * g.N + 1
recs = torch.empty((cap, rw), dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
batch0 = 0
for rnd, shifts in enumerate([[None, 0, 4, 8, 12, 16, 20], [None, 8, 12]]):
    cnt.zero_()
    for s in range(R):
        synth.reads_device(p, (batch0 + s) * reads, reads, seq, qual, st)
        # extract appends: advance the output pointer by the running count
        n_so_far = int(cnt.item())
        c2 = torch.zeros(1, dtype=torch.int64, device="cuda")
        g.extract_records_device(seq, qual, nb, w.cutoff, recs[n_so_far:], cap - n_so_far, c2, st)
        torch.cuda.synchronize()
        cnt += c2
    batch0 += R
    n = int(cnt.item())
    buckets = torch.empty(n, dtype=torch.int32, device="cuda")
    _capi.check(lib.lasv_graph_record_buckets_device(g._h, recs.data_ptr(), n, buckets.data_ptr(), st))
    for sh in shifts:
        if sh is None:
            cur = recs[:n]
            t_sort = 0.0
        else:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            key = (buckets.to(torch.int64) & 0xFFFFFFFF) >> sh
            perm = torch.argsort(key)
            cur = recs[:n].index_select(0, perm).contiguous()
            torch.cuda.synchronize()
            t_sort = time.perf_counter() - t0
            del key, perm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.insert_records_device(cur, n, st)
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        region = None if sh is None else (1 << sh) * w.bucket_size * g.element_bytes
        print(json.dumps({"workload": wl, "steps_per_insert": R, "round": rnd, "records": n, "group_shift": sh,
                          "region_bytes": region, "insert_ms": dt * 1e3, "inserts_per_s": n / dt,
                          "torch_sort_ms": t_sort * 1e3}), flush=True)
        del cur
g.stats()
g.free()
