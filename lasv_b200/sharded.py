"""Hash-sharded graph build across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The graph of
2^L buckets x W slots is split into `world` shards of 2^(L-log2 world) buckets;
a k-mer occurrence is owned by rank  (lookup3(key) >> local_bits) & (world-1),
i.e. by the HIGH bits of its reference bucket index (hash_value.c:767-773), and
is inserted there with the LOW bits as the local bucket -- so the shards laid
end to end in rank order are the reference's table for every key that sits in
its first-choice bucket.  Node updates commute (coverage +, edges |), so the
only communication is one personalised all-to-all of (2N+1)-word records per
batch:

    reads --route kernel--> per-owner bins --all_to_all--> insert kernel --> shard

The exchange is one grouped send/recv (`torch.distributed.batch_isend_irecv`, a
single ncclGroup of ncclSend/ncclRecv under NCCL) on views of the bins -- no
packing copy; counts travel first in a tiny all_to_all_single.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from .graph import DBGraph, words_per_kmer


def is_pow2(x: int) -> bool:
    return x > 0 and (x & (x - 1)) == 0


def exchange_records(bins: torch.Tensor, send_counts: torch.Tensor, recv_buf: torch.Tensor,
                     group=None):
    """Personalised all-to-all of variable-length record blocks.

    bins        [world, bin_capacity, rec_words] int32  (bin d = records owned by rank d)
    send_counts [world] int64 on the host: records in each bin
    recv_buf    [recv_capacity, rec_words] int32
    Returns (n_received, recv_counts[world] host tensor).  Works with NCCL (CUDA
    tensors) and gloo (CPU tensors: the world_size-2 tests)."""
    world = bins.shape[0]
    sc = send_counts.to(torch.int64).cpu()
    if int(sc.max()) > bins.shape[1]:
        raise OverflowError(f"exchange bin overflow: {int(sc.max())} records > capacity {bins.shape[1]}")
    sc_dev = sc.to(bins.device)
    rc_dev = torch.empty_like(sc_dev)
    dist.all_to_all_single(rc_dev, sc_dev, group=group)
    rc = rc_dev.cpu()
    total = int(rc.sum())
    if total > recv_buf.shape[0]:
        raise OverflowError(f"receive buffer too small: {total} > {recv_buf.shape[0]}")
    starts = torch.zeros(world + 1, dtype=torch.int64)
    starts[1:] = torch.cumsum(rc, 0)
    rank = dist.get_rank(group)
    ops = []
    for d in range(world):
        src = bins[d, : int(sc[d])]
        dst = recv_buf[int(starts[d]): int(starts[d + 1])]
        if d == rank:
            dst.copy_(src)                     # own records never leave the GPU
            continue
        if src.numel():
            ops.append(dist.P2POp(dist.isend, src, d if group is None else dist.get_global_rank(group, d), group))
        if dst.numel():
            ops.append(dist.P2POp(dist.irecv, dst, d if group is None else dist.get_global_rank(group, d), group))
    if ops:
        # one grouped ncclSend/ncclRecv launch under NCCL; plain isend/irecv under gloo
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return total, rc


class ShardedDBGraph:
    """The rank-local part of a graph sharded over `world` GPUs."""

    def __init__(self, kmer_size: int, number_bits: int, bucket_size: int, rank: int, world: int,
                 device: int, max_kmers_per_batch: int, group=None, slack: float = 1.08):
        if not is_pow2(world) or world > 16:
            raise ValueError("world size must be a power of two <= 16")
        self.rank, self.world, self.group = rank, world, group
        self.N = words_per_kmer(kmer_size)
        self.rec_words = 2 * self.N + 1
        self.local_bits = number_bits - int(math.log2(world))
        if self.local_bits < 1:
            raise ValueError("too few buckets for this many ranks")
        self.graph = DBGraph(kmer_size, self.local_bits, bucket_size, device=device)
        self.device = torch.device("cuda", device)
        if world > 1:
            # a rank sends ~1/world of its k-mers to each owner and receives about as many as it sent
            self.bin_capacity = int(max_kmers_per_batch / world * slack) + 4096
            self.bins = torch.empty((world, self.bin_capacity, self.rec_words), dtype=torch.int32,
                                    device=self.device)
            self.recv = torch.empty((int(max_kmers_per_batch * slack) + 4096 * world, self.rec_words),
                                    dtype=torch.int32, device=self.device)
            self.counts = torch.zeros(world, dtype=torch.int64, device=self.device)

    def free(self):
        self.graph.free()

    def load_reads_device(self, d_seq, d_qual, n_bytes: int, quality_cutoff: int, d_offsets=None,
                          n_reads: int = 0):
        """One batch: every rank passes ITS reads; returns after the local inserts
        are queued on the current stream."""
        stream = torch.cuda.current_stream().cuda_stream
        g = self.graph
        if self.world == 1:
            g.load_reads_device(d_seq, d_qual, n_bytes, d_offsets, n_reads, quality_cutoff, stream)
            return
        self.counts.zero_()
        g.route_reads_device(d_seq, d_qual, n_bytes, quality_cutoff, self.world, self.bins,
                             self.bin_capacity, self.counts, stream)
        if d_offsets is not None and n_reads:
            g.read_stats_device(d_seq, d_qual, d_offsets, n_reads, quality_cutoff, stream)
        total, _ = exchange_records(self.bins, self.counts, self.recv, self.group)
        g.insert_records_device(self.recv, total, stream)

    def stats(self) -> dict:
        """Job-wide statistics (sum over ranks)."""
        st = self.graph.stats(stream=torch.cuda.current_stream().cuda_stream)
        keys = sorted(st)
        t = torch.tensor([st[k] for k in keys], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return dict(zip(keys, (int(x) for x in t.cpu())))

    def summary(self) -> dict:
        s = self.graph.summary()
        keys = sorted(s)
        # fingerprints add modulo 2^64: reduce as two 32-bit halves to stay inside int64
        lo = torch.tensor([s[k] & 0xFFFFFFFF for k in keys], dtype=torch.int64, device=self.device)
        hi = torch.tensor([s[k] >> 32 for k in keys], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(lo, group=self.group)
            dist.all_reduce(hi, group=self.group)
        out = {}
        for k, l, h in zip(keys, lo.cpu().tolist(), hi.cpu().tolist()):
            out[k] = (l + (h << 32)) & 0xFFFFFFFFFFFFFFFF
        return out
