#!/usr/bin/env python
"""Streamed-insert prototype (groups.cuh) against the random-access insert on the same records:
one GPU's share of the human-scale table at 8 GPUs (L=21, W=250: 13.4 GB) pre-filled with the genome's k-mers,
then one batch of reads turned into records (the older variable-length records path), inserted once through
insert_records_kernel and once -- sorted by group on the device with torch -- through insert_groups_kernel.
Only the insert calls are timed (CUDA events); the sort is reported separately (the real design partitions inside
the bins instead).  Both graphs must have the same fingerprint.

    python tools/grouped_insert_speed.py [pairs_per_batch] [buckets_per_group] [L]
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from lasv_b200 import DBGraph, _capi, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
bpg = int(sys.argv[2]) if len(sys.argv) > 2 else 4
L = int(sys.argv[3]) if len(sys.argv) > 3 else 21
w = synth.scaled(synth.CFG4, synth.CFG4.genome_len >> (24 - L), L)      # same fill as the full table
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
p, pt = w.params(), w.params(tiling=True)
stride = w.read_len + 1
N = 2 * w.kmer_size // 64 + 1
batch = 1 << 20
d_seq = torch.empty(2 * max(pairs, batch) * stride + 16, dtype=torch.uint8, device="cuda")
d_qual = torch.empty_like(d_seq)
out = {"what": "insert_groups_kernel (prototype) vs insert_records_kernel", "k": w.kmer_size, "L": L, "W": w.bucket_size,
       "pairs_per_batch": pairs, "buckets_per_group": bpg,
       "staging": "cp.async.bulk, double buffered" if os.environ.get("LASV_B200_GROUPS_BULK", "0") != "0" else "ldg/stg"}


def event_ms(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), r


def prefilled():
    g = DBGraph(w.kmer_size, L, w.bucket_size)
    n_tiles = (w.genome_len + pt.tiling_step - 1) // pt.tiling_step
    done = 0
    while done < n_tiles:                                  # the genome's own k-mers, once each
        nr = min(2 * batch, n_tiles - done)
        synth.reads_device(pt, done, nr, d_seq, d_qual, st)
        g.load_reads_device(d_seq, d_qual, nr * stride, None, 0, w.cutoff, st)
        done += nr
    torch.cuda.synchronize()
    return g


n_reads = 2 * pairs
recs = torch.empty((n_reads * (w.read_len - w.kmer_size + 1) + 8, 2 * N + 1), dtype=torch.int32, device="cuda")
cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
fingerprints = []
for mode in ("records", "grouped"):
    g = prefilled()
    out["fill_before"] = g.unique_kmers / g.capacity
    synth.reads_device(p, 0, n_reads, d_seq, d_qual, st)
    cnt.zero_()
    g.extract_records_device(d_seq, d_qual, n_reads * stride, w.cutoff, recs, recs.shape[0], cnt, st)
    torch.cuda.synchronize()
    n = int(cnt.item())
    out["records"] = n
    if mode == "records":
        ms, _ = event_ms(lambda: g.insert_records_device(recs, n, st))
        out["insert_records_ms"] = ms
    else:
        def sort_by_group():
            buckets = torch.empty(n, dtype=torch.int32, device="cuda")
            _capi.check(_capi.lib().lasv_graph_record_buckets_device(g._h, recs.data_ptr(), n, buckets.data_ptr(), st))
            groups = (buckets.to(torch.int64) & 0xFFFFFFFF) // bpg
            order = torch.argsort(groups)
            offs = torch.searchsorted(groups[order].contiguous(),
                                      torch.arange((1 << L) // bpg + 1, device="cuda", dtype=torch.int64))
            return recs[:n][order].contiguous(), offs.contiguous()
        def sort_in_library():
            srt = torch.empty((n, 2 * N + 1), dtype=torch.int32, device="cuda")
            offs_ = torch.empty((1 << L) // bpg + 1, dtype=torch.int64, device="cuda")
            g.group_records_device(recs, n, bpg, srt, offs_, st)
            return srt, offs_
        ms_lib, _ = event_ms(sort_in_library)
        out["group_records_ms_library"] = ms_lib
        ms_sort, (srecs, offs) = event_ms(sort_by_group)
        ms, spilled = event_ms(lambda: g.insert_grouped_records_device(srecs, n, offs, bpg, st))
        out["sort_by_group_ms_torch"] = ms_sort
        out["insert_grouped_ms"] = ms
        out["spilled"] = spilled
        out["records_per_line"] = n / (g.capacity / (5 if N == 2 else 7 if N == 1 else 3))
        out["table_GB"] = (1 << L) * 50 * 128 / 1e9 if (N == 2 and w.bucket_size == 250) else None
    s = g.summary()
    fingerprints.append((s["n_nodes"], s["sum_covg"], s["fingerprint"]))
    g.free()
out["same_graph"] = fingerprints[0] == fingerprints[1]
out["fingerprint"] = fingerprints[0][2]
print(json.dumps(out))
