"""Hash-sharded graph build across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  The graph of
2^L buckets x W slots is split into `world` shards of 2^(L-log2 world) buckets;
a k-mer occurrence is owned by rank  (lookup3(key) >> local_bits) & (world-1),
i.e. by the HIGH bits of its reference bucket index (hash_value.c:767-773), and
is inserted there with the LOW bits as the local bucket -- so the shards laid
end to end in rank order are the reference's table for every key that sits in
its first-choice bucket.  Node updates commute (coverage +, edges |), so the
only communication is one personalised all-to-all of k-mer records per batch:

    reads --bin kernel--> stripes[owner][region] --all_to_all--> insert kernel --> shard

The bin kernel writes every k-mer occurrence into the stripe of the 64 MB table
region it will touch ON ITS OWNER; the stripes of one owner are one contiguous,
fixed-size block, so the exchange is a plain `all_to_all_single` with equal
splits (NCCL over NVLink) -- no packing copy and no host round trip for sizes;
the per-stripe counters travel the same way.  By default the blocks do not even go
through NCCL: the receive buffers live in symmetric memory (every rank maps every
rank's allocation, torch.distributed._symmetric_memory) and each rank PUSHES its
blocks into its peers' buffers with peer-to-peer DMA copies -- copy engines over
NVLink, no SM of either side involved -- bracketed by two stream-ordered barriers
(`LASV_B200_EXCHANGE=nccl` selects the all_to_all_single path; it is also the
fallback when peer access is unavailable).  The insert kernel then sweeps the
received stripes region by region, all senders of a region together.  With
`pipeline=True` batch i+1 is binned and exchanged while batch i is inserted
(two sets of buffers, side streams).

`exchange_records` / route / insert_records is the older variable-length path
(kept: small tables, the world_size-2 gloo test).
"""
from __future__ import annotations

import math
import os
import sys

import torch
import torch.distributed as dist

from . import _capi
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


def exchange_stripes(send: torch.Tensor, recv: torch.Tensor, group=None):
    """Equal-split personalised all-to-all of the stripe blocks: block d of `send`
    goes to rank d, block s of `recv` comes from rank s.  NCCL: all_to_all_single;
    gloo (CPU tests): list all_to_all, or send/recv pairs if that is missing."""
    world = send.shape[0]
    if send.is_cuda:
        dist.all_to_all_single(recv.view(world, -1), send.view(world, -1), group=group)
        return
    try:
        dist.all_to_all(list(recv.view(world, -1).unbind(0)), list(send.view(world, -1).unbind(0)), group=group)
    except (RuntimeError, NotImplementedError):
        rank = dist.get_rank(group)
        ops = []
        for d in range(world):
            if d == rank:
                recv[d].copy_(send[d])
            else:
                ops.append(dist.P2POp(dist.isend, send[d].contiguous(), d, group))
                ops.append(dist.P2POp(dist.irecv, recv[d], d, group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def write_sharded_ctx(path, header: bytes, append_records, group=None) -> int:
    """One `.ctx` from a hash-sharded graph: rank 0 writes the header, then every rank in turn
    appends its shard's records with `append_records(path) -> count` (the ranks share the file system of
    the node).  Owner = high bits of the reference bucket index, so the shards laid end to end are in the
    reference's table order.  Returns the job-wide record count on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = 0
    for r in range(world):
        if r == rank:
            if rank == 0:
                with open(path, "wb") as f:
                    f.write(header)
            mine = int(append_records(path))
        dist.barrier(group=group)        # rank r's bytes are in the file before rank r+1 opens it
    total = torch.tensor([mine], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        total = total.cuda()
    dist.all_reduce(total, group=group)
    return int(total.item())


class _StripeBuffers:
    def __init__(self, world, geo, device, with_send=True):
        # (no send stripes in the fused mode: the partition kernel stores into the owners' receive stripes)
        self.send = torch.empty((world, geo["records_per_owner"], geo["record_bytes"] // 8), dtype=torch.int64,
                                device=device) if with_send else None
        self.recv = None                     # allocated by the exchange set-up: symmetric memory or plain
        self.send_counts = torch.zeros((world, geo["counters_per_owner"], geo["count_stride_bytes"] // 4),
                                       dtype=torch.int32, device=device)
        self.recv_counts = None
        self.inserted = None     # event: the insert kernel that read `recv` is done
        self.exchanged = None    # event: `send` may be overwritten, `recv` is complete


class ShardedDBGraph:
    """The rank-local part of a graph sharded over `world` GPUs."""

    def __init__(self, kmer_size: int, number_bits: int, bucket_size: int, rank: int, world: int,
                 device: int, max_kmers_per_batch: int, group=None, slack: float = 1.08,
                 max_bytes_per_batch: int = 0, max_reads_per_batch: int = 0, pipeline: bool = True,
                 accumulate_batches: int = 1):
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
        self.pipeline = pipeline
        if pipeline:
            self.graph.set_pipeline(True)     # kernels sized to share the SMs with the neighbouring stage
        self.buffers = []
        self.turn = 0
        # accumulate_batches = A: a rank bins A batches into the SAME stripes (sized for all of them) before the
        # exchange, and the owners insert A batches' records at once -- the streamed insert's cost per record falls
        # with the records per table line (streamed_insert.cuh), and the exchange moves few large blocks
        self.accumulate = max(1, int(accumulate_batches))
        self.pending = 0                      # batches binned into the current buffer set since its last exchange
        self.trace = None                     # set to [] to collect per-stage CUDA events (timeline())
        if world > 1:
            if not max_bytes_per_batch:       # k-mers <= bytes: a safe stand-in for callers that only know k-mers
                max_bytes_per_batch = int(max_kmers_per_batch * 1.7)
            self.max_bytes_per_batch = max_bytes_per_batch
            kk = kmer_size
            self.kmer_bound = (max_bytes_per_batch - max_reads_per_batch * kk
                               if max_reads_per_batch and max_reads_per_batch * kk < max_bytes_per_batch else max_bytes_per_batch)
            geo = self.graph.bin_geometry(world, self.accumulate * max_bytes_per_batch,
                                          self.accumulate * max_reads_per_batch)
            self.bins_per_owner, self.cap, self.spill = geo["bins_per_owner"], geo["stripe_capacity"], geo["spill_capacity"]
            self.rec_bytes = geo["record_bytes"]
            # LASV_B200_EXCHANGE = p2p (default): blocks pushed by peer-to-peer DMA copies; stores: FUSED producer ->
            # exchange, the partition kernel stores every record straight into its owner's receive stripes over
            # peer memory (no send stripes, no block copies; lasv_graph_bin_reads_peers_device); nccl: all_to_all_single
            want = os.environ.get("LASV_B200_EXCHANGE", "p2p")
            for _ in range(2 if self.pipeline else 1):
                self.buffers.append(_StripeBuffers(world, geo, self.device, with_send=(want != "stores")))
            self.comm_stream = torch.cuda.Stream(device=self.device)
            self.insert_stream = torch.cuda.Stream(device=self.device)
            self.exchange = "nccl"
            if want in ("p2p", "stores"):
                try:
                    self._setup_p2p(geo)
                    self.exchange = want
                except Exception as ex:           # no peer access / no symmetric memory: NCCL does the job
                    print(f"[lasv_b200] peer-memory exchange unavailable ({type(ex).__name__}: {ex}); using NCCL",
                          file=sys.stderr)
            if self.exchange == "nccl":
                for buf in self.buffers:
                    if buf.send is None:
                        buf.send = torch.empty((world, geo["records_per_owner"], geo["record_bytes"] // 8),
                                               dtype=torch.int64, device=self.device)
                    buf.recv = torch.empty_like(buf.send)
                    buf.recv_counts = torch.zeros_like(buf.send_counts)

    def _setup_p2p(self, geo):
        """Receive buffers in symmetric memory (every rank maps every rank's allocation): the exchange
        becomes `world` peer-to-peer DMA copies per batch -- copy engines over NVLink, no SM of the
        sender or the receiver involved -- bracketed by two stream-ordered barriers."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        world, rw, cw = self.world, geo["record_bytes"] // 8, geo["count_stride_bytes"] // 4
        rec_words = geo["records_per_owner"] * rw                       # int64 words per owner block
        cnt_words = geo["counters_per_owner"] * cw // 2                 # int32 counters viewed as int64 words
        set_words = world * (rec_words + cnt_words)
        self.symm = symm_mem.empty(len(self.buffers) * set_words, dtype=torch.int64, device=self.device)
        self.symm_hdl = symm_mem.rendezvous(self.symm, group)
        for k, buf in enumerate(self.buffers):
            base = k * set_words
            buf.recv = self.symm[base: base + world * rec_words].view(world, geo["records_per_owner"], rw)
            cbase = base + world * rec_words
            buf.recv_counts = self.symm[cbase: cbase + world * cnt_words].view(torch.int32).view(
                world, geo["counters_per_owner"], cw)
            # where MY block lands on every peer: slot `rank` of the peer's receive buffer
            buf.peer_recv = [self.symm_hdl.get_buffer(d, (geo["records_per_owner"], rw), torch.int64,
                                                      base + self.rank * rec_words) for d in range(world)]
            buf.peer_counts = [self.symm_hdl.get_buffer(d, (cnt_words,), torch.int64,
                                                        cbase + self.rank * cnt_words) for d in range(world)]
            buf.send_counts64 = buf.send_counts.view(torch.int64).view(world, cnt_words)
        self.symm_hdl.barrier(0)
        torch.cuda.synchronize()

    def _exchange(self, buf, k):
        """On the current (communication) stream: blocks and counters of `buf.send*` to their owners."""
        if self.exchange == "stores":
            self.symm_hdl.barrier(k)                 # every rank's partition kernels of this round are complete:
            for i in range(self.world):              # all records sit in their owners' receive stripes already
                d = (self.rank + i) % self.world
                buf.peer_counts[d].copy_(buf.send_counts64[d], non_blocking=True)
            self.symm_hdl.barrier(k)                 # every sender's counters have landed here
        elif self.exchange == "p2p":
            self.symm_hdl.barrier(k)                 # every rank's receive buffer of this set is free
            for i in range(self.world):
                d = (self.rank + i) % self.world     # start with the local block, spread the peers
                buf.peer_recv[d].copy_(buf.send[d], non_blocking=True)
                buf.peer_counts[d].copy_(buf.send_counts64[d], non_blocking=True)
            self.symm_hdl.barrier(k)                 # every block addressed to this rank has landed
        else:
            exchange_stripes(buf.send, buf.recv, self.group)
            exchange_stripes(buf.send_counts, buf.recv_counts, self.group)

    def free(self):
        self.graph.free()

    def load_reads_device(self, d_seq, d_qual, n_bytes: int, quality_cutoff: int, d_offsets=None,
                          n_reads: int = 0):
        """One batch: every rank passes ITS reads.  Returns once the batch's work is
        queued; with the pipeline on, the inserts of this batch may still be running
        on a side stream when the caller's stream is done -- `flush()` joins them
        (stats/summary/free do it themselves)."""
        cur = torch.cuda.current_stream()
        stream = cur.cuda_stream
        g = self.graph
        if self.world == 1:
            g.load_reads_device(d_seq, d_qual, n_bytes, d_offsets, n_reads, quality_cutoff, stream)
            return
        if n_bytes > self.max_bytes_per_batch:
            raise ValueError(f"batch of {n_bytes} stream bytes exceeds the {self.max_bytes_per_batch} the exchange "
                             "buffers were sized for (max_bytes_per_batch)")
        # the stripes hold what a batch of max_bytes in max_reads reads can yield (k-mers <= bytes - reads * k);
        # a batch of fewer, longer reads yields more per byte and must not be dropped on the floor
        k = self.graph.kmer_size
        bound = n_bytes - n_reads * k if (d_offsets is not None and n_reads and n_reads * k < n_bytes) else n_bytes
        if bound > self.kmer_bound:
            raise ValueError(f"batch may hold {bound} k-mers, the exchange stripes were sized for {self.kmer_bound} "
                             "(max_bytes_per_batch / max_reads_per_batch: pass fewer bytes or state fewer reads)")
        buf = self.buffers[self.turn % len(self.buffers)]
        first = self.pending == 0
        # the previous user of this buffer set must be done with it
        if first and buf.exchanged is not None:
            cur.wait_event(buf.exchanged)       # its send block is free again
        ev = None
        if self.trace is not None and first:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            self.trace.append(ev)
            ev[0].record(cur)
        if self.exchange == "stores":
            if first:
                # the owners must have inserted what this buffer set held before anybody stores into it again
                if buf.inserted is not None:
                    cur.wait_event(buf.inserted)
                self.symm_hdl.barrier(2 + self.turn % len(self.buffers))
            g.bin_reads_peers_device(d_seq, d_qual, n_bytes, quality_cutoff, self.world, self.bins_per_owner, self.cap,
                                     self.spill, buf.peer_recv, buf.send_counts, stream, append=not first)
        else:
            g.bin_reads_device(d_seq, d_qual, n_bytes, quality_cutoff, self.world, self.bins_per_owner, self.cap,
                               self.spill, buf.send, buf.send_counts, stream, append=not first)
        if d_offsets is not None and n_reads:
            g.read_stats_device(d_seq, d_qual, d_offsets, n_reads, quality_cutoff, stream)
        self.pending += 1
        if self.pending >= self.accumulate:
            self._ship()

    def _ship(self):
        """Exchange what the current buffer set has accumulated and insert it on its owners."""
        if not self.pending:
            return
        cur = torch.cuda.current_stream()
        stream = cur.cuda_stream
        g = self.graph
        k = self.turn % len(self.buffers)
        buf = self.buffers[k]
        self.turn += 1
        self.pending = 0
        ev = self.trace[-1] if self.trace else None
        if ev:
            ev[1].record(cur)
        if not self.pipeline:
            self._exchange(buf, 0)
            g.insert_bins_device(buf.recv, buf.recv_counts, self.world, self.bins_per_owner, self.cap, self.spill,
                                 stream)
            return
        binned = torch.cuda.Event()
        binned.record(cur)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(binned)
            if buf.inserted is not None:
                self.comm_stream.wait_event(buf.inserted)   # recv block free again
            if ev:
                ev[2].record(self.comm_stream)
            self._exchange(buf, k)
            if ev:
                ev[3].record(self.comm_stream)
            buf.exchanged = torch.cuda.Event()
            buf.exchanged.record(self.comm_stream)
        with torch.cuda.stream(self.insert_stream):
            self.insert_stream.wait_event(buf.exchanged)
            if ev:
                ev[4].record(self.insert_stream)
            g.insert_bins_device(buf.recv, buf.recv_counts, self.world, self.bins_per_owner, self.cap, self.spill,
                                 self.insert_stream.cuda_stream)
            if ev:
                ev[5].record(self.insert_stream)
            buf.inserted = torch.cuda.Event()
            buf.inserted.record(self.insert_stream)

    def stats_async(self, pinned_out):
        """Queue a device->host copy of this rank's statistics behind the inserts of the batches
        loaded so far (no host synchronisation); see DBGraph.stats_async for the layout."""
        if self.world > 1 and self.pipeline:
            with torch.cuda.stream(self.insert_stream):
                self.graph.stats_async(pinned_out, self.insert_stream.cuda_stream)
        else:
            self.graph.stats_async(pinned_out, torch.cuda.current_stream().cuda_stream)

    def timeline(self):
        """[(partition start, end, exchange start, end, insert start, end)] in ms relative to the first
        traced batch (after a device synchronisation)."""
        torch.cuda.synchronize()
        if not self.trace:
            return []
        t0 = self.trace[0][0]
        return [[t0.elapsed_time(e) for e in ev] for ev in self.trace]

    def join(self):
        """The caller's current stream waits for the exchanges and inserts queued so far; batches that were binned
        but not yet exchanged stay in the stripes (flush() ships them too)."""
        if self.world > 1 and self.pipeline:
            cur = torch.cuda.current_stream()
            for buf in self.buffers:
                if buf.inserted is not None:
                    cur.wait_event(buf.inserted)

    def flush(self):
        """Make the caller's current stream wait for everything queued on the side streams."""
        cur = torch.cuda.current_stream()
        if self.world == 1:
            self.graph.flush(cur.cuda_stream)
            return
        self._ship()                     # (collective: every rank has binned the same number of batches)
        if self.pipeline:
            for buf in self.buffers:
                if buf.inserted is not None:
                    cur.wait_event(buf.inserted)

    def load_reads_device_routed(self, d_seq, d_qual, n_bytes: int, quality_cutoff: int, d_offsets=None,
                                 n_reads: int = 0):
        """The older path: per-owner record bins of variable length, sizes via the host."""
        stream = torch.cuda.current_stream().cuda_stream
        g = self.graph
        if not hasattr(self, "bins"):
            kmers = int(n_bytes)
            self.bin_capacity = int(kmers / self.world * 1.08) + 4096
            self.bins = torch.empty((self.world, self.bin_capacity, self.rec_words), dtype=torch.int32,
                                    device=self.device)
            self.recv = torch.empty((int(kmers * 1.08) + 4096 * self.world, self.rec_words),
                                    dtype=torch.int32, device=self.device)
            self.counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        self.counts.zero_()
        g.route_reads_device(d_seq, d_qual, n_bytes, quality_cutoff, self.world, self.bins,
                             self.bin_capacity, self.counts, stream)
        if d_offsets is not None and n_reads:
            g.read_stats_device(d_seq, d_qual, d_offsets, n_reads, quality_cutoff, stream)
        total, _ = exchange_records(self.bins, self.counts, self.recv, self.group)
        g.insert_records_device(self.recv, total, stream)

    def dump_binary(self, path, mean_read_len: int = 0, total_seq: int = 0) -> int:
        """`--dump_binary path` of the whole sharded graph: BINVERSION 6 header + the records of all shards in
        rank order (db_graph_dump_single_colour_binary_of_colour0, dB_graph_population.c:3730).  Collective."""
        from .graph import ctx_header
        self.flush()
        if self.world == 1:
            self.graph.dump_binary(path, mean_read_len, total_seq)
            return self.graph.summary()["n_nodes"]
        return write_sharded_ctx(path, ctx_header(self.graph.kmer_size, mean_read_len, total_seq),
                                 self.graph.append_binary_records, self.group)

    def _local(self, call):
        """A rank-local library call that may raise (table full, stripe overflow, ...): the error is agreed on by
        all ranks BEFORE anybody raises, so that no rank is left alone in the collective that follows."""
        err = None
        try:
            out = call()
        except _capi.LasvError as ex:
            err, out = ex, None
        if self.world > 1:
            flag = torch.tensor([0 if err is None else 1], dtype=torch.int64, device=self.device)
            dist.all_reduce(flag, group=self.group)
            if int(flag.item()) and err is None:
                raise _capi.LasvError(_capi.LASV_ERR_STATE, f"{int(flag.item())} other rank(s) of the sharded graph failed "
                                                              "(table full or exchange stripe overflow there)")
        if err is not None:
            raise err
        return out

    def stats(self) -> dict:
        """Job-wide statistics (sum over ranks)."""
        self.flush()
        st = self._local(lambda: self.graph.stats(stream=torch.cuda.current_stream().cuda_stream))
        keys = sorted(st)
        t = torch.tensor([st[k] for k in keys], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(t, group=self.group)
        return dict(zip(keys, (int(x) for x in t.cpu())))

    def summary(self) -> dict:
        self.flush()
        s = self._local(self.graph.summary)
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
