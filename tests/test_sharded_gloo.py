"""world_size-2 gloo test (CPU) of the sharded build's host logic: the
personalised all-to-all of variable-length record blocks
(lasv_b200/sharded.py::exchange_records) carrying records produced and consumed
by the kernels' host emulation (tests/host_emul).  The union of the two shards
must equal the reference golden graph."""
import ctypes as C
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lasv_b200 import element_dtype, words_per_kmer
    from lasv_b200.sharded import exchange_records
    from tests import util
    emul = C.CDLL(str(ROOT / "tests" / "host_emul" / "libhost_emul.so"))
    emul.emul_extract_records.restype = C.c_uint64
    emul.emul_extract_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_uint64]
    emul.emul_insert_records.argtypes = [C.c_void_p, C.c_uint64] + [C.c_int] * 4 + [C.c_void_p] * 2
    emul.emul_table_bytes.restype = C.c_uint64
    emul.emul_table_bytes.argtypes = [C.c_int] * 3
    emul.emul_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    emul.emul_hash_c.restype = C.c_uint32
    emul.emul_hash_c.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = words_per_kmer(k)
    rw = 2 * N + 1
    seq, qual, offs = util.case_streams(name)
    # this rank's slice of the reads (cut at a read boundary)
    n_reads = len(offs) - 1
    lo, hi = int(offs[n_reads * rank // world]), int(offs[n_reads * (rank + 1) // world])
    s, q = np.ascontiguousarray(seq[lo:hi]), np.ascontiguousarray(qual[lo:hi])
    cap = meta["inserts"] + 8
    recs = np.zeros((cap, rw), dtype=np.uint32)
    n = emul.emul_extract_records(s.ctypes.data, q.ctypes.data, len(s), k, util.cutoff_of(meta), 384,
                                  recs.ctypes.data, cap)
    recs = recs[:n]
    bits = int(np.log2(world))
    owner = np.zeros(n, dtype=np.int64)
    for i in range(n):
        key = np.zeros(N, dtype=np.uint64)
        for w in range(N):
            key[w] = np.uint64(recs[i, 2 * w]) | (np.uint64(recs[i, 2 * w + 1]) << np.uint64(32))
        owner[i] = (emul.emul_hash_c(key.ctypes.data, N, 0) >> (L - bits)) & (world - 1)
    bins = torch.zeros((world, cap, rw), dtype=torch.int32)
    counts = torch.zeros(world, dtype=torch.int64)
    for d in range(world):
        mine = recs[owner == d].view(np.int32)
        bins[d, : len(mine)] = torch.from_numpy(mine.copy())
        counts[d] = len(mine)
    recv = torch.zeros((cap * world, rw), dtype=torch.int32)
    total, rc = exchange_records(bins, counts, recv)
    assert total == int(rc.sum())
    got = np.ascontiguousarray(recv[:total].numpy().view(np.uint32))
    raw = np.zeros(int(emul.emul_table_bytes(k, L - bits, W)), dtype=np.uint8)   # device line layout
    ctr = np.zeros(20, dtype=np.uint64)
    assert emul.emul_insert_records(got.ctypes.data, total, k, L - bits, W, 15, raw.ctypes.data,
                                    ctr.ctypes.data) == 0
    table = np.zeros((1 << (L - bits)) * W, dtype=element_dtype(N))
    emul.emul_gather(raw.ctypes.data, k, L - bits, W, 0, table.ctypes.data)
    np.save(Path(outdir) / f"shard{rank}.npy", table)
    # one `.ctx` from all shards, as ShardedDBGraph.dump_binary writes it: rank 0's header, then every rank's
    # records in rank order
    from lasv_b200 import ctx_header, record_dtype
    from lasv_b200.sharded import write_sharded_ctx

    def append_records(path):
        occ = table[table["status"] != 0]
        out = np.zeros(len(occ), dtype=record_dtype(N))
        out["kmer"], out["covg"], out["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
        with open(path, "ab") as f:
            f.write(out.tobytes())
        return len(out)

    n_all = write_sharded_ctx(Path(outdir) / "all.ctx", ctx_header(k, 100, 1000), append_records)
    assert n_all == meta["n_records"]
    # job-wide insert count through a collective, as ShardedDBGraph.stats does
    t = torch.tensor([n], dtype=torch.int64)
    dist.all_reduce(t)
    assert int(t.item()) == meta["inserts"]
    # an overflowing bin must be reported, not truncated
    try:
        exchange_records(bins[:, :1], counts.clamp(min=2), recv)
        raise AssertionError("overflow not detected")
    except OverflowError:
        pass
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["synth_cfg0_k53", "synth_cfg1_k31"])
def test_two_rank_exchange_builds_the_reference_graph(name, tmp_path):
    import subprocess
    so = ROOT / "tests" / "host_emul" / "libhost_emul.so"
    if not so.exists():
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I/usr/local/cuda/include", "-o", str(so), str(so.with_name("host_emul.cpp"))], check=True)
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    from lasv_b200 import record_dtype
    from tests import util
    recs = []
    for r in range(world):
        t = np.load(tmp_path / f"shard{r}.npy")
        occ = t[t["status"] != 0]
        out = np.zeros(len(occ), dtype=record_dtype(t["kmer"].shape[1]))
        out["kmer"], out["covg"], out["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
        recs.append(out)
    assert all(len(r) > 0 for r in recs)
    util.assert_same_records(np.concatenate(recs), util.case_records(name), name)
    # the sharded dump: a valid BINVERSION 6 file, shard 0's records first, then shard 1's
    from oracle import oracle as O
    ctx = O.parse_ctx(tmp_path / "all.ctx")
    meta = util.case_meta(name)
    assert bytes(ctx["header"]) == O.ctx_header(meta["k"], 100, 1000)
    assert np.array_equal(ctx["records"], np.concatenate(recs))


def _stripe_worker(rank, world, port, name, outdir):
    """The production exchange on the CPU: every rank bins its records by (owner, region) into the
    fixed-size per-owner blocks the partition kernel writes (stripe counters one per 128-byte line, a
    spill area behind every block), `exchange_stripes` moves the blocks, and the owner inserts what it
    received sender by sender -- through the kernels' host emulation."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lasv_b200 import element_dtype, words_per_kmer
    from lasv_b200.sharded import exchange_stripes
    from tests import util
    emul = C.CDLL(str(ROOT / "tests" / "host_emul" / "libhost_emul.so"))
    emul.emul_extract_records.restype = C.c_uint64
    emul.emul_extract_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_uint64]
    emul.emul_insert_records.argtypes = [C.c_void_p, C.c_uint64] + [C.c_int] * 4 + [C.c_void_p] * 2
    emul.emul_table_bytes.restype = C.c_uint64
    emul.emul_table_bytes.argtypes = [C.c_int] * 3
    emul.emul_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    emul.emul_hash_c.restype = C.c_uint32
    emul.emul_hash_c.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = words_per_kmer(k)
    rw = 2 * N + 1
    seq, qual, offs = util.case_streams(name)
    n_reads = len(offs) - 1
    lo, hi = int(offs[n_reads * rank // world]), int(offs[n_reads * (rank + 1) // world])
    s, q = np.ascontiguousarray(seq[lo:hi]), np.ascontiguousarray(qual[lo:hi])
    cap_all = meta["inserts"] + 8
    recs = np.zeros((cap_all, rw), dtype=np.uint32)
    n = emul.emul_extract_records(s.ctypes.data, q.ctypes.data, len(s), k, util.cutoff_of(meta), 384,
                                  recs.ctypes.data, cap_all)
    recs = recs[:n]
    bits = int(np.log2(world))
    B, cap, spill, cstride = 4, 64, cap_all, 32        # tiny stripes: the spill areas carry real traffic
    per_owner = B * cap + spill
    send = torch.zeros((world, per_owner, rw), dtype=torch.int32)
    send_counts = torch.zeros((world, B + 1, cstride), dtype=torch.int32)
    region_shift = L - bits - 2                        # B = 4 regions per owner
    # two batches, the second APPENDED to the stripes of the first (ShardedDBGraph(accumulate_batches=2):
    # lasv_graph_bin_reads_append_device keeps the counters), then ONE exchange and one insert per owner
    batches = [range(0, n // 2), range(n // 2, n)]
    counts_after_first = None
    for bi, batch in enumerate(batches):
      if bi == 1:
        counts_after_first = send_counts.clone()
      for i in batch:
          key = np.zeros(N, dtype=np.uint64)
          for w in range(N):
              key[w] = np.uint64(recs[i, 2 * w]) | (np.uint64(recs[i, 2 * w + 1]) << np.uint64(32))
          c = emul.emul_hash_c(key.ctypes.data, N, 0) & ((1 << L) - 1)
          owner, region = c >> (L - bits), (c >> region_shift) & (B - 1)
          idx = int(send_counts[owner, region, 0])
          send_counts[owner, region, 0] += 1
          if idx < cap:
              send[owner, region * cap + idx] = torch.from_numpy(recs[i].view(np.int32).copy())
          else:
              si = int(send_counts[owner, B, 0])
              send_counts[owner, B, 0] += 1
              assert si < spill
              send[owner, B * cap + si] = torch.from_numpy(recs[i].view(np.int32).copy())
    # (every record asks its stripe's counter for a place; the ones that found it full also ask the spill counter)
    assert counts_after_first is not None and int((send_counts - counts_after_first)[:, :B, 0].sum()) == n - n // 2
    recv, recv_counts = torch.zeros_like(send), torch.zeros_like(send_counts)
    exchange_stripes(send, recv, None)
    exchange_stripes(send_counts, recv_counts, None)
    raw = np.zeros(int(emul.emul_table_bytes(k, L - bits, W)), dtype=np.uint8)
    ctr = np.zeros(20, dtype=np.uint64)
    total = 0
    for region in list(range(B)) + [B]:               # regions in order, every sender; then the spill areas
        for src in range(world):
            cnt = int(recv_counts[src, region, 0])
            if region < B:
                cnt = min(cnt, cap)
                blk = recv[src, region * cap: region * cap + cnt]
            else:
                blk = recv[src, B * cap: B * cap + cnt]
            got = np.ascontiguousarray(blk.numpy().view(np.uint32))
            assert emul.emul_insert_records(got.ctypes.data, cnt, k, L - bits, W, 15, raw.ctypes.data,
                                            ctr.ctypes.data) == 0
            total += cnt
    t = torch.tensor([total, n], dtype=torch.int64)
    dist.all_reduce(t)
    assert int(t[0]) == int(t[1]) == meta["inserts"]
    assert int(recv_counts[:, B, 0].sum()) > 0
    table = np.zeros((1 << (L - bits)) * W, dtype=element_dtype(N))
    emul.emul_gather(raw.ctypes.data, k, L - bits, W, 0, table.ctypes.data)
    np.save(Path(outdir) / f"shard{rank}.npy", table)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stripe_exchange_builds_the_reference_graph(tmp_path):
    name = "synth_cfg0_k53"
    world = 2
    mp.spawn(_stripe_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    from lasv_b200 import record_dtype
    from tests import util
    recs = []
    for r in range(world):
        t = np.load(tmp_path / f"shard{r}.npy")
        occ = t[t["status"] != 0]
        out = np.zeros(len(occ), dtype=record_dtype(t["kmer"].shape[1]))
        out["kmer"], out["covg"], out["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
        recs.append(out)
    util.assert_same_records(np.concatenate(recs), util.case_records(name), name)
