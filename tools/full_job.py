#!/usr/bin/env python
"""BASELINE.json configs[4] from start to finish: the complete 30x build of the synthetic 3.1 Gbp genome
(310 M read pairs of 2x150 bp, ~60 G k-mer insertions, k=53, L=24 W=250) from an EMPTY table, on 1..8 GPUs.
The read set is the same for every GPU count (global read index space split across the ranks), so the
order-independent fingerprint of the graph must be identical for N = 1, 2, 4, 8.

    python tools/full_job.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
           --master-port 29531 tools/full_job.py --gpus 8

Reads are generated on the device, batch by batch, inside the timed region (the generator is a
streaming kernel, ~1.5 ms per 2^20 pairs); one JSON line on rank 0.
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist
from lasv_b200 import synth
from lasv_b200.sharded import ShardedDBGraph

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--pairs-per-batch", type=int, default=1 << 20)
ap.add_argument("--fraction", type=float, default=1.0, help="part of the 30x read set to load")
ap.add_argument("--accumulate-gb", type=float, default=0.0,
                help="N=1: lasv_graph_set_accumulation with this many GB of read stream per insert (-1: all that fits)")
ap.add_argument("--insert", choices=["auto", "random", "streamed"], default="auto")
ap.add_argument("--prefill", action="store_true",
                help="untimed: every genomic k-mer once first (error-free tiling reads), so that the timed batches run "
                     "in the steady state of a 30x build (bench.py's pre-fill)")
ap.add_argument("--timing", action="store_true", help="per-kind kernel times (CUDA events in the library)")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
w = synth.CFG4
p = w.params()
stride = w.read_len + 1
reads_per_batch = 2 * args.pairs_per_batch
n_bytes = reads_per_batch * stride
total_pairs = int(w.n_pairs * args.fraction)
n_batches_global = (total_pairs + args.pairs_per_batch - 1) // args.pairs_per_batch
n_batches_global = (n_batches_global + world - 1) // world * world       # every rank takes part in every exchange
graph = ShardedDBGraph(w.kmer_size, w.number_bits, w.bucket_size, rank, world, local,
                       max_kmers_per_batch=reads_per_batch * (w.read_len - w.kmer_size + 1),
                       max_bytes_per_batch=n_bytes, max_reads_per_batch=reads_per_batch, pipeline=world > 1)
bufs = [(torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev), torch.empty(n_bytes + 16, dtype=torch.uint8, device=dev))
        for _ in range(2)]
offs = torch.arange(reads_per_batch + 1, dtype=torch.int64, device=dev) * stride
stream = torch.cuda.current_stream().cuda_stream
from lasv_b200 import _capi
graph.graph.set_insert_mode({"auto": _capi.INSERT_AUTO, "random": _capi.INSERT_RANDOM, "streamed": _capi.INSERT_STREAMED}[args.insert])
granted = 0
if world == 1 and args.accumulate_gb:
    granted = graph.graph.set_accumulation((1 << 64) - 1 if args.accumulate_gb < 0 else int(args.accumulate_gb * 1e9))
if args.prefill:
    pt = w.params(tiling=True)
    n_tiles = (w.genome_len + pt.tiling_step - 1) // pt.tiling_step
    per_rank = (n_tiles + world - 1) // world
    done = 0
    while done < per_rank:
        nr = min(reads_per_batch, per_rank - done)
        synth.reads_device(pt, rank * per_rank + done, nr, bufs[0][0], bufs[0][1], stream)
        if nr < reads_per_batch:
            bufs[0][0][nr * stride: n_bytes].fill_(10)
        graph.load_reads_device(bufs[0][0], bufs[0][1], n_bytes, w.cutoff)
        done += reads_per_batch
    graph.flush()
    torch.cuda.synchronize()
    pre = graph.stats()["kmers_inserted"]
else:
    pre = 0
if args.timing:
    graph.graph.enable_kernel_timing(True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for i, b in enumerate(range(rank, n_batches_global, world)):
    s, q = bufs[i % 2]
    synth.reads_device(p, b * reads_per_batch, reads_per_batch, s, q, stream)
    graph.load_reads_device(s, q, n_bytes, w.cutoff, offs, reads_per_batch)
graph.flush()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
kinds = graph.graph.kernel_times_by_kind() if args.timing else None
st = graph.stats()
sm = graph.summary()
if rank == 0:
    print(json.dumps({"what": "cfg4 complete build from an empty table", "n_gpus": world, "read_pairs": n_batches_global * args.pairs_per_batch,
                      "seconds": dt, "kmers_inserted": st["kmers_inserted"] - pre, "kmers_per_s": (st["kmers_inserted"] - pre) / dt, "prefilled_kmers": pre,
                      "unique_kmers": st["unique_kmers"], "fill": st["unique_kmers"] / (graph.graph.capacity * world),
                      "sum_covg": sm["sum_covg"], "covg_equals_inserts": sm["sum_covg"] == st["kmers_inserted"],
                      "bad_reads": st["bad_reads"], "fingerprint": f"{sm['fingerprint']:016x}",
                      "accumulate_bytes_granted": granted, "inserts": graph.graph.insert_counts(), "insert_mode": args.insert,
                      "kernel_ms_by_kind": kinds}), flush=True)
graph.free()
if world > 1:
    dist.destroy_process_group()
