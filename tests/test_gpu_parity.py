"""GPU parity tests (run with -m gpu on the B200 box): every call goes through
the C ABI of liblasv_b200.so.  The CUDA graph build is compared bit-exactly
(sorted multiset of (canonical k-mer, coverage, edges), loader statistics, .ctx
bytes) with the golden outputs of the compiled reference, with the oracle on
seeded synthetic reads, and -- at BASELINE.json's full config-0 size -- through
size-independent properties."""
import ctypes as C
import dataclasses
import os
import subprocess

import numpy as np
import pytest

import lasv_b200
from lasv_b200 import DBGraph, _capi, synth
from lasv_b200.graph import mean_contig_length
from oracle import oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _have_gpu():
    try:
        return _capi.lib().lasv_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    if not _have_gpu():
        pytest.fail("no CUDA device: the GPU tests need one (the product has no CPU fallback)")


def torch_cuda():
    import torch
    torch.cuda.set_device(0)
    return torch


# ---------------------------------------------------------------- golden cases
@pytest.mark.parametrize("name", util.ALL_CASES)
def test_host_buffer_path_matches_reference_golden(name):
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        st = g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        util.assert_same_records(g.records(), util.case_records(name), name)
        assert st["kmers_inserted"] == meta["inserts"]
        assert st["unique_kmers"] == meta["n_records"] == g.unique_kmers
        assert st["bad_reads"] == meta["bad_reads"]
        assert st["bases_read"] == meta["bases_parsed"]
        assert st["bases_loaded"] == meta["bases_loaded"]
        _, hist = g.stats(hist_size=20001)
        assert mean_contig_length(hist) == meta["ctx_mean_read_len"]
        assert int(g.rehash_histogram().sum()) == meta["inserts"]


@pytest.mark.parametrize("name", util.EDGE_PE + util.EDGE_SE + util.REFFIX)
def test_sequence_file_loaders_match_reference_golden(name, tmp_path):
    meta = util.case_meta(name)
    f1, f2 = util.case_files(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        if f2 is None:
            lst = tmp_path / "se.list"
            lst.write_text(f"{f1}\n")
            st = g.load_se_filelist(lst, meta["q_thresh"], 33)
        else:
            (tmp_path / "l").write_text(f"{f1}\n")
            (tmp_path / "r").write_text(f"{f2}\n")
            st = g.load_pe_filelists(tmp_path / "l", tmp_path / "r", meta["q_thresh"], 33)
        assert st["files_loaded"] == 1
        util.assert_same_records(g.records(), util.case_records(name), name)
        assert st["kmers_inserted"] == meta["inserts"] and st["bad_reads"] == meta["bad_reads"]
        assert st["bases_read"] == meta["bases_parsed"] and st["bases_loaded"] == meta["bases_loaded"]
        # .ctx file: header bytes and record multiset equal the reference's own dump
        _, hist = g.stats(hist_size=20001)
        out = tmp_path / "g.ctx"
        g.dump_binary(out, mean_contig_length(hist), st["bases_loaded"])
        ctx = O.parse_ctx(out)
        assert ctx["header"].hex() == meta["ctx_header_hex"]
        util.assert_same_records(ctx["records"], util.case_records(name), name)


@pytest.mark.parametrize("name", ["synth_cfg0_k53", "synth_cfg2_k95"])
@pytest.mark.parametrize("batch_bytes", [4096, 50_000])
def test_file_loaders_with_many_small_batches(name, batch_bytes, tmp_path, monkeypatch):
    """The parser threads hand their pinned batches to the GPU in turn: with batches of a few KB every
    file spans hundreds of them (slot hand-off, the read that did not fit, the last partial batch)."""
    monkeypatch.setenv("LASV_B200_LOADER_BATCH_BYTES", str(batch_bytes))
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        st = g.load_pe_seq_data(tmp_path / "a.fq", tmp_path / "b.fq", util.cutoff_of(meta))
        assert st["n_reads"] == len(s2) and st["kmers_inserted"] == meta["inserts"]
        assert st["bad_reads"] == meta["bad_reads"] and st["bases_loaded"] == meta["bases_loaded"]
        util.assert_same_records(g.records(), util.case_records(name), name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:      # single-ended: one parser thread
        st = g.load_se_seq_data(tmp_path / "a.fq", util.cutoff_of(meta))
        assert st["n_reads"] == len(s2) // 2


@pytest.mark.parametrize("batch_bytes,spoil", [(0, False), (150_000, False), (150_000, True)])
def test_file_loaders_parse_fastq_text_on_the_device(batch_bytes, spoil, tmp_path, monkeypatch):
    """Large uncompressed FASTQ files are mapped, cut into slices at record boundaries, and the slices go to the
    device as TEXT (HostBatch::raw): newline index, record shape check, offsets, stream copy (fastq_parse.cu).
    Same graph and statistics as the host parse of the same files; a chunk that is not plain four-line FASTQ
    (one multi-line record in the middle of a file) falls back to the host parser for that chunk only."""
    monkeypatch.setenv("LASV_B200_PARSE_THREADS", "2")           # slices from 1 MB on
    if batch_bytes:
        monkeypatch.setenv("LASV_B200_LOADER_BATCH_BYTES", str(batch_bytes))
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    reps = 32                                                    # 32 x 84 KB per file: two slices each
    for f, sel in (("a.fq", slice(0, None, 2)), ("b.fq", slice(1, None, 2))):
        O.write_fastq_fixed(tmp_path / "one.fq", s2[sel].reshape(-1), q2[sel].reshape(-1), rl)
        text = (tmp_path / "one.fq").read_bytes()
        if spoil and f == "a.fq":
            # the 200th record with its sequence and quality wrapped over two lines: legal FASTQ, not four-line
            lines = text.split(b"\n")
            i = 4 * 200
            lines[i + 1: i + 2] = [lines[i + 1][:40], lines[i + 1][40:]]
            lines[i + 4: i + 5] = [lines[i + 4][:55], lines[i + 4][55:]]
            text = b"\n".join(lines)
        one = (tmp_path / "one.fq").read_bytes()
        (tmp_path / f).write_bytes(one * 9 + text + one * (reps - 10))      # (the wrapped record sits in the first slice)
    gold = util.case_records(name).copy()
    gold["covg"] *= reps
    results = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LASV_B200_DEVICE_PARSE", mode)
        with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
            l0 = _capi.lib().lasv_kernel_launches()
            st = g.load_pe_seq_data(tmp_path / "a.fq", tmp_path / "b.fq", util.cutoff_of(meta))
            results[mode] = (st, _capi.lib().lasv_kernel_launches() - l0)
            assert st["n_reads"] == reps * len(s2) and st["kmers_inserted"] == reps * meta["inserts"]
            assert st["bad_reads"] == reps * meta["bad_reads"] and st["bases_loaded"] == reps * meta["bases_loaded"]
            util.assert_same_records(g.records(), gold, name)
    assert results["1"][0] == results["0"][0]
    assert results["1"][1] > results["0"][1]                     # the parse kernels ran


@pytest.mark.parametrize("kind", ["bgzf", "gzip"])
def test_bgzf_fastq_goes_to_the_device_parse(kind, tmp_path, monkeypatch):
    """Compressed FASTQ: several threads inflate windows of text -- blocks of a BGZF file (BgzfStream), chunks of an
    ordinary gzip stream (PgzStream, gz_parallel.cpp) -- and the text goes to the device parse like the slices of plain
    files (cut at the last record boundary of a window, the tail carried over).  Same graph and statistics as the
    uncompressed files; a multi-line record hands the rest of that file to the host parser."""
    from tests.test_capi_cpu import _bgzf, _gz
    if kind == "gzip":
        monkeypatch.setenv("LASV_B200_PGZ_MIN_KB", "0")          # (the files are ~5 MB compressed)
        monkeypatch.setenv("LASV_B200_PGZ_CHUNK_KB", "512")
        _bgzf = lambda data: _gz(data, 1)                        # noqa: E731
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    reps = 260                                                   # 22 MB of text per file: three windows
    gold = util.case_records(name).copy()
    gold["covg"] *= reps
    for f, sel in (("a", slice(0, None, 2)), ("b", slice(1, None, 2))):
        O.write_fastq_fixed(tmp_path / "one.fq", s2[sel].reshape(-1), q2[sel].reshape(-1), rl)
        one = (tmp_path / "one.fq").read_bytes()
        (tmp_path / f"{f}.fq.gz").write_bytes(_bgzf(one * reps))
        if f == "a":                                             # the same reads with one record wrapped, late in the file
            lines = one.split(b"\n")
            i = 4 * 100
            lines[i + 1: i + 2] = [lines[i + 1][:30], lines[i + 1][30:]]
            lines[i + 3: i + 4] = [lines[i + 3][:70], lines[i + 3][70:]]
            (tmp_path / "w.fq.gz").write_bytes(_bgzf(one * 200 + b"\n".join(lines) + one * (reps - 201)))
    for first in ("a.fq.gz", "w.fq.gz"):
        with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
            st = g.load_pe_seq_data(tmp_path / first, tmp_path / "b.fq.gz", util.cutoff_of(meta))
            assert st["n_reads"] == reps * len(s2) and st["kmers_inserted"] == reps * meta["inserts"]
            assert st["bad_reads"] == reps * meta["bad_reads"] and st["bases_loaded"] == reps * meta["bases_loaded"]
            util.assert_same_records(g.records(), gold, first)


def _merge_records(*arrays):
    acc = {}
    for a in arrays:
        for r in a:
            key = tuple(int(x) for x in r["kmer"])
            c, e = acc.get(key, (0, 0))
            acc[key] = (c + int(r["covg"]), e | int(r["edges"]))
    out = np.zeros(len(acc), dtype=arrays[0].dtype)
    for i, (key, (c, e)) in enumerate(acc.items()):
        out[i]["kmer"], out[i]["covg"], out[i]["edges"] = key, c, e
    return out


def test_filelists_with_several_files_are_parsed_concurrently(tmp_path):
    """A filelist names several files: they are parsed at the same time, four entries per group.  Five
    copies of a pair (two groups) give five times the coverage; a list that mixes FASTQ and FASTA applies
    the quality filter only where there are qualities."""
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text("a.fq\n" * 5)                   # relative to the list's directory
    (tmp_path / "r").write_text("b.fq\n" * 5)
    gold = util.case_records(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        st = g.load_pe_filelists(tmp_path / "l", tmp_path / "r", meta["q_thresh"], 33)
        assert st["files_loaded"] == 5 and st["kmers_inserted"] == 5 * meta["inserts"]
        assert st["bad_reads"] == 5 * meta["bad_reads"] and st["n_reads"] == 5 * len(s2)
        five = gold.copy()
        five["covg"] *= 5
        util.assert_same_records(g.records(), five, name)
    # SE list: FASTQ (quality threshold 20 applies) + FASTA (no qualities) at k=53
    mq, mf = util.case_meta("edge_se_k53_q20"), util.case_meta("fasta_se_k53")
    assert (mq["k"], mq["L"], mq["W"]) == (mf["k"], mf["L"], mf["W"])
    (tmp_path / "se").write_text(f"{util.GOLDEN / 'edge_1.fq'}\n{util.GOLDEN / 'contigs.fa'}\n")
    with DBGraph(mq["k"], mq["L"], mq["W"]) as g:
        st = g.load_se_filelist(tmp_path / "se", mq["q_thresh"], 33)
        assert st["files_loaded"] == 2 and st["kmers_inserted"] == mq["inserts"] + mf["inserts"]
        want = _merge_records(util.case_records("edge_se_k53_q20"), util.case_records("fasta_se_k53"))
        util.assert_same_records(g.records(), want, "fastq+fasta list")


def test_paired_files_of_unequal_length_are_an_error(tmp_path):
    a = tmp_path / "a.fq"
    a.write_text("@x\nACGT\n+\nIIII\n@y\nACGT\n+\nIIII\n")
    b = tmp_path / "b.fq"
    b.write_text("@x\nACGT\n+\nIIII\n")
    with DBGraph(3, 4, 4) as g:
        with pytest.raises(_capi.LasvError) as e:
            g.load_pe_seq_data(a, b, 33)
        assert "same number of reads" in str(e.value)        # file_reader.c:641-647


# ------------------------------------------------------- seeded synthetic reads
def _synth_oracle(w, n_reads, seed=None):
    if seed is not None:
        w.seed = seed
    p = w.params()
    seq, qual, offs = synth.reads_host(p, 0, n_reads)
    og = O.OracleGraph(w.kmer_size, w.number_bits, w.bucket_size)
    assert og.load_reads(seq, qual, offs, w.cutoff, paired=True)
    return p, seq, qual, offs, og


@pytest.mark.parametrize("base,genome,L,W,n_reads", [
    (synth.CFG1, 100_000, 13, 60, 30_000),    # k=31, one word
    (synth.CFG0, 100_000, 13, 60, 30_000),    # k=53, two words
    (synth.CFG2, 60_000, 12, 60, 20_000),     # k=95, three words, filter-heavy
    (synth.CFG4, 150_000, 10, 250, 20_000),   # human-scale read model, W=250
])
def test_device_resident_path_matches_oracle(base, genome, L, W, n_reads):
    torch = torch_cuda()
    w = synth.scaled(base, genome, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads)
    want = O.sort_records(og.records())
    n_bytes = len(seq)
    d_seq = torch.empty(n_bytes + 16, dtype=torch.uint8, device="cuda")
    d_qual = torch.empty_like(d_seq)
    st_ = torch.cuda.current_stream().cuda_stream
    synth.reads_device(p, 0, n_reads, d_seq, d_qual, st_)
    assert np.array_equal(d_seq[:n_bytes].cpu().numpy(), seq)
    assert np.array_equal(d_qual[:n_bytes].cpu().numpy(), qual)
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    with DBGraph(w.kmer_size, L, W) as g:
        # two batches on the caller's stream
        half = (n_reads // 2) * (w.read_len + 1)
        g.load_reads_device(d_seq, d_qual, half, d_offs, n_reads // 2, w.cutoff, st_)
        rest = d_offs[n_reads // 2:] - half
        g.load_reads_device(d_seq[half:], d_qual[half:], n_bytes - half, rest, n_reads - n_reads // 2,
                            w.cutoff, st_)
        torch.cuda.synchronize()
        st = g.stats(stream=st_)
        assert np.array_equal(O.sort_records(g.records()), want)
        assert st["kmers_inserted"] == og.stats.kmers_inserted
        assert st["unique_kmers"] == og.unique_kmers
        assert st["bad_reads"] == og.stats.bad_reads and st["bases_loaded"] == og.stats.bases_loaded
        s = g.summary()
        assert s["n_nodes"] == len(want) and s["sum_covg"] == og.stats.kmers_inserted
        # find: present and absent keys
        keys = want["kmer"][:: max(1, len(want) // 500)]
        found, covg, edges = g.find(keys)
        sel = want[:: max(1, len(want) // 500)]
        assert found.all() and np.array_equal(covg, sel["covg"]) and np.array_equal(edges, sel["edges"])
        absent = keys.copy()
        absent[:, -1] ^= np.uint64(0x2AAAA)
        present = {tuple(x) for x in want["kmer"].tolist()}
        found, _, _ = g.find(absent)
        assert found.tolist() == [tuple(x) in present for x in absent.tolist()]


@pytest.mark.parametrize("base,world", [(synth.CFG0, 2), (synth.CFG1, 4), (synth.CFG2, 8)])
def test_route_and_insert_records_match_oracle(base, world):
    """The sharded path on one GPU: ranks emulated as `world` local graphs; every
    emulated rank routes its slice of the reads, bins are handed to their owners."""
    torch = torch_cuda()
    L, W, n_reads = 13, 40, 24_000
    w = synth.scaled(base, 80_000, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads, seed=99)
    want = O.sort_records(og.records())
    N = lasv_b200.words_per_kmer(w.kmer_size)
    rw = 2 * N + 1
    bits = int(np.log2(world))
    st_ = torch.cuda.current_stream().cuda_stream
    d_seq = torch.from_numpy(np.concatenate([seq, np.full(16, 10, np.uint8)])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, np.full(16, 10, np.uint8)])).cuda()
    graphs = [DBGraph(w.kmer_size, L - bits, W) for _ in range(world)]
    try:
        cap = int(og.stats.kmers_inserted / world * 1.5) + 4096
        per = (n_reads // world) * (w.read_len + 1)
        inbox = [[] for _ in range(world)]
        total_sent = 0
        for r in range(world):
            bins = torch.zeros((world, cap, rw), dtype=torch.int32, device="cuda")
            counts = torch.zeros(world, dtype=torch.int64, device="cuda")
            lo, hi = r * per, ((r + 1) * per if r < world - 1 else len(seq))
            graphs[r].route_reads_device(d_seq[lo:], d_qual[lo:], hi - lo, w.cutoff, world, bins, cap,
                                         counts, st_)
            torch.cuda.synchronize()
            c = counts.cpu().tolist()
            assert max(c) <= cap
            total_sent += sum(c)
            for d in range(world):
                inbox[d].append(bins[d, : c[d]].clone())
        assert total_sent == og.stats.kmers_inserted
        recs = []
        for d in range(world):
            mine = torch.cat(inbox[d]).contiguous()
            graphs[d].insert_records_device(mine, mine.shape[0], st_)
            torch.cuda.synchronize()
            graphs[d].stats(stream=st_)
            recs.append(graphs[d].records())
        assert np.array_equal(O.sort_records(np.concatenate(recs)), want)
        # extract-records (single bin) + insert == fused
        with DBGraph(w.kmer_size, L, W) as g:
            out = torch.zeros((og.stats.kmers_inserted + 8, rw), dtype=torch.int32, device="cuda")
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            g.extract_records_device(d_seq, d_qual, len(seq), w.cutoff, out, out.shape[0], cnt, st_)
            torch.cuda.synchronize()
            assert int(cnt.item()) == og.stats.kmers_inserted
            g.insert_records_device(out, int(cnt.item()), st_)
            torch.cuda.synchronize()
            assert np.array_equal(O.sort_records(g.records()), want)
    finally:
        for g in graphs:
            g.free()


@pytest.mark.parametrize("base,world,bins,tiny_stripes", [(synth.CFG0, 2, 4, False), (synth.CFG1, 4, 1, False),
                                                           (synth.CFG2, 8, 8, False), (synth.CFG0, 1, 16, False),
                                                           (synth.CFG0, 2, 4, True), (synth.CFG2, 4, 2, True),
                                                           # two-word k-mers too long for the 16-byte record
                                                           (dataclasses.replace(synth.CFG0, kmer_size=63), 2, 2, False),
                                                           (dataclasses.replace(synth.CFG0, kmer_size=57), 1, 4, False)])
@pytest.mark.parametrize("insert", ["random", "streamed"])
def test_binned_exchange_matches_oracle(base, world, bins, tiny_stripes, insert, monkeypatch):
    """The production sharded path on one GPU: ranks emulated as `world` local graphs.
    Every emulated rank bins its slice of the reads by (owner, table region); the
    all-to-all is emulated by handing block d of every sender to graph d; the owners
    sweep the received stripes region by region (random access) or stream their table
    through shared memory (streamed_insert.cuh), senders' spill areas included."""
    torch = torch_cuda()
    monkeypatch.setenv("LASV_B200_BINS", str(bins))
    monkeypatch.setenv("LASV_B200_INSERT", insert)
    L, W, n_reads = 13, 40, 24_000
    w = synth.scaled(base, 80_000, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads, seed=77)
    want = O.sort_records(og.records())
    bits = int(np.log2(world))
    st_ = torch.cuda.current_stream().cuda_stream
    d_seq = torch.from_numpy(np.concatenate([seq, np.full(16, 10, np.uint8)])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, np.full(16, 10, np.uint8)])).cuda()
    graphs = [DBGraph(w.kmer_size, L - bits, W) for _ in range(world)]
    try:
        per_reads = n_reads // world
        per = per_reads * (w.read_len + 1)
        geo = graphs[0].bin_geometry(world, len(seq) - (world - 1) * per, n_reads - (world - 1) * per_reads)
        assert geo["bins_per_owner"] == bins
        B, cap, spill = geo["bins_per_owner"], geo["stripe_capacity"], geo["spill_capacity"]
        if tiny_stripes:                      # most records overflow their stripe and take the spill path
            cap, spill = 128, (int(og.stats.kmers_inserted / world / world * 1.3) + 127) // 128 * 128
        rw, cw = geo["record_bytes"] // 8, geo["count_stride_bytes"] // 4
        send = torch.zeros((world, world, B * cap + spill, rw), dtype=torch.int64, device="cuda")  # [sender][owner]
        send_counts = torch.zeros((world, world, B + 1, cw), dtype=torch.int32, device="cuda")
        for r in range(world):
            lo, hi = r * per, ((r + 1) * per if r < world - 1 else len(seq))
            # in two batches: the second one is APPENDED to the stripes of the first (accumulation before the exchange)
            mid = lo + (per_reads // 2) * (w.read_len + 1)
            if insert == "streamed" and world > 1:
                # fused producer -> exchange: the records go straight to per-owner blocks (on a multi-GPU box: the
                # owners' receive stripes over peer memory); here the blocks are this sender's rows of `send`
                blocks = [send[r, o] for o in range(world)]
                graphs[r].bin_reads_peers_device(d_seq[lo:], d_qual[lo:], mid - lo, w.cutoff, world, B, cap, spill, blocks,
                                                 send_counts[r], st_)
                graphs[r].bin_reads_peers_device(d_seq[mid:], d_qual[mid:], hi - mid, w.cutoff, world, B, cap, spill, blocks,
                                                 send_counts[r], st_, append=True)
                continue
            graphs[r].bin_reads_device(d_seq[lo:], d_qual[lo:], mid - lo, w.cutoff, world, B, cap, spill, send[r],
                                       send_counts[r], st_)
            graphs[r].bin_reads_device(d_seq[mid:], d_qual[mid:], hi - mid, w.cutoff, world, B, cap, spill, send[r],
                                       send_counts[r], st_, append=True)
        torch.cuda.synchronize()
        cnt = send_counts[..., 0].cpu().numpy().astype(np.int64)
        stored = np.minimum(cnt[..., :B], cap).sum() + np.minimum(cnt[..., B], spill).sum()
        # (a record may stand for several occurrences: identical keys of a warp round are merged)
        assert 0 < int(stored) <= og.stats.kmers_inserted
        if world > 1 or tiny_stripes:
            assert (cnt[..., B] <= spill).all()
        if tiny_stripes:
            assert cnt[..., B].sum() > 0.5 * stored
        assert sum(g.stats(stream=st_)["kmers_inserted"] for g in graphs) == og.stats.kmers_inserted
        recs = []
        for d in range(world):
            recv = send[:, d].contiguous()                        # what all_to_all_single delivers to rank d
            recv_counts = send_counts[:, d].contiguous()
            graphs[d].insert_bins_device(recv, recv_counts, world, B, cap, spill, st_)
            torch.cuda.synchronize()
            st = graphs[d].stats(stream=st_)
            recs.append(graphs[d].records())
            assert st["unique_kmers"] == len(recs[-1])
            assert graphs[d].insert_counts()[insert] == 1
        assert np.array_equal(O.sort_records(np.concatenate(recs)), want)
    finally:
        for g in graphs:
            g.free()


def test_binned_exchange_overflow_is_reported(monkeypatch):
    """A stripe that is too small must raise, never drop records silently."""
    torch = torch_cuda()
    monkeypatch.setenv("LASV_B200_BINS", "2")
    w = synth.scaled(synth.CFG0, 50_000, 10)
    seq, qual, offs = synth.reads_host(w.params(), 0, 4000)
    st_ = torch.cuda.current_stream().cuda_stream
    d_seq = torch.from_numpy(np.concatenate([seq, np.full(16, 10, np.uint8)])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, np.full(16, 10, np.uint8)])).cuda()
    with DBGraph(w.kmer_size, 9, w.bucket_size) as g:
        geo = g.bin_geometry(2, len(seq), len(offs) - 1)
        B, rw, cw = geo["bins_per_owner"], geo["record_bytes"] // 8, geo["count_stride_bytes"] // 4
        cap, spill = 128, 128                                    # far too small, spill area included
        send = torch.zeros((2, B * cap + spill, rw), dtype=torch.int64, device="cuda")
        counts = torch.zeros((2, B + 1, cw), dtype=torch.int32, device="cuda")
        g.bin_reads_device(d_seq, d_qual, len(seq), w.cutoff, 2, B, cap, spill, send, counts, st_)
        torch.cuda.synchronize()
        with pytest.raises(_capi.LasvError):
            g.stats(stream=st_)


def test_pipelined_inserts_and_async_statistics(monkeypatch):
    """lasv_graph_set_pipeline: batch i is inserted on the library's own stream while
    batch i+1 is partitioned on the caller's; the result, the asynchronous statistics
    read-back and the per-kernel timing events must match the one-stream build."""
    torch = torch_cuda()
    monkeypatch.setenv("LASV_B200_BINS", "8")
    monkeypatch.setenv("LASV_B200_INSERT", "binned")
    L, W, n_reads = 13, 40, 24_000
    w = synth.scaled(synth.CFG0, 80_000, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads, seed=5)
    want = O.sort_records(og.records())
    st_ = torch.cuda.current_stream().cuda_stream
    d_seq = torch.from_numpy(np.concatenate([seq, np.full(16, 10, np.uint8)])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, np.full(16, 10, np.uint8)])).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    n_batches = 6
    per_reads = n_reads // n_batches
    per = per_reads * (w.read_len + 1)
    with DBGraph(w.kmer_size, L, W) as g:
        g.set_pipeline(True)
        g.enable_kernel_timing(True)
        rows = torch.zeros((n_batches, 9), dtype=torch.int64).pin_memory()
        for b in range(n_batches):
            lo = b * per
            nr = per_reads if b < n_batches - 1 else n_reads - b * per_reads
            nb = nr * (w.read_len + 1)
            d_off_b = (d_offs[b * per_reads: b * per_reads + nr + 1] - lo).contiguous()
            g.load_reads_device(d_seq[lo:], d_qual[lo:], nb, d_off_b, nr, w.cutoff, st_)
            g.flush(st_)                                       # the statistics row of THIS batch
            g.stats_async(rows[b], st_)
        torch.cuda.synchronize()
        kt = g.kernel_times()
        assert kt["batches"] == n_batches and kt["partition_ms"] > 0 and kt["insert_ms"] > 0
        names = DBGraph.STATS_ASYNC_FIELDS
        last = dict(zip(names, rows[-1].tolist()))
        assert last["kmers_inserted"] == og.stats.kmers_inserted and last["n_reads"] == n_reads
        assert last["bad_reads"] == og.stats.bad_reads and last["table_full"] == 0
        ins = rows[:, names.index("kmers_inserted")].tolist()
        assert all(a < b for a, b in zip(ins, ins[1:]))        # every batch's row advanced
        st = g.stats(stream=st_)
        assert st["unique_kmers"] == len(want) == last["unique_kmers"]
        assert np.array_equal(O.sort_records(g.records()), want)


@pytest.mark.parametrize("base,L,W,bins,n_batches", [(synth.CFG0, 13, 40, 4, 5), (synth.CFG1, 12, 100, 1, 3),
                                                     (synth.CFG2, 13, 40, 16, 4), (synth.CFG0, 10, 250, 2, 1),
                                                     (dataclasses.replace(synth.CFG0, kmer_size=63), 13, 33, 8, 2)])
def test_streamed_insert_with_accumulation_matches_oracle(base, L, W, bins, n_batches):
    """lasv_graph_set_accumulation + the streamed insert (streamed_insert.cuh): load_reads_device only partitions,
    the stripes are inserted by flush / by the first call that reads the table, through sort-by-group, shared-memory
    upserts behind cp.async.bulk copies, and the drain of full buckets.  Same multiset as the oracle, and the same
    as the random-access insert."""
    torch = torch_cuda()
    n_reads = 24_000
    w = synth.scaled(base, 80_000, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads, seed=91)
    want = O.sort_records(og.records())
    st_ = torch.cuda.current_stream().cuda_stream
    d_seq = torch.from_numpy(np.concatenate([seq, np.full(16, 10, np.uint8)])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, np.full(16, 10, np.uint8)])).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    per_reads = n_reads // n_batches
    per = per_reads * (w.read_len + 1)
    os.environ["LASV_B200_BINS"] = str(bins)
    os.environ["LASV_B200_INSERT"] = "binned"
    try:
        with DBGraph(w.kmer_size, L, W) as g:
            g.set_insert_mode(_capi.INSERT_STREAMED)
            g.enable_kernel_timing(True)
            # room for two batches and a bit: the third one forces an insert by itself
            granted = g.set_accumulation(int(2.5 * per))
            assert granted == int(2.5 * per)
            for b in range(n_batches):
                lo = b * per
                nr = per_reads if b < n_batches - 1 else n_reads - b * per_reads
                d_off_b = (d_offs[b * per_reads: b * per_reads + nr + 1] - lo).contiguous()
                g.load_reads_device(d_seq[lo:], d_qual[lo:], nr * (w.read_len + 1), d_off_b, nr, w.cutoff, st_)
            before = g.insert_counts()
            assert before["random"] == 0 and before["streamed"] == (n_batches - 1) // 2
            st = g.stats(stream=st_)                   # reads the counters: inserts what is still in the stripes
            assert g.insert_counts()["streamed"] == before["streamed"] + 1
            assert st["kmers_inserted"] == og.stats.kmers_inserted and st["unique_kmers"] == len(want)
            assert st["bad_reads"] == og.stats.bad_reads
            assert np.array_equal(O.sort_records(g.records()), want)
            kt = g.kernel_times_by_kind()
            assert kt["partition"][1] == n_batches and kt["insert"][1] == before["streamed"] + 1
            assert kt["sort"][0] > 0 and kt["stream"][0] > 0 and kt["stream"][1] == kt["sort"][1] >= kt["insert"][1]
            # a second load after the flush, accumulation off again: one insert per batch
            g.set_accumulation(0)
            g.set_insert_mode(_capi.INSERT_RANDOM)
            g.load_reads_device(d_seq, d_qual, len(seq), d_offs, n_reads, w.cutoff, st_)
            recs = O.sort_records(g.records())
            assert g.insert_counts()["random"] == 1
            assert np.array_equal(recs["kmer"], want["kmer"]) and np.array_equal(recs["edges"], want["edges"])
            assert np.array_equal(recs["covg"].astype(np.int64), 2 * want["covg"].astype(np.int64))
    finally:
        os.environ.pop("LASV_B200_BINS", None)
        os.environ.pop("LASV_B200_INSERT", None)



@pytest.mark.parametrize("accumulate", [0, 2.5])
def test_asynchronous_host_calls_through_the_staging_ring(accumulate, monkeypatch):
    """lasv_graph_load_reads_host_async + lasv_graph_set_staging: the calls return when their copies have left the
    host buffers, the chunks wrap a five-leg ring many times (64 KB chunks), full stripes are inserted while later
    copies are already on their way, the per-call statistics arrive in pinned memory.  Same graph as the oracle."""
    torch = torch_cuda()
    monkeypatch.setenv("LASV_B200_BINS", "4")
    monkeypatch.setenv("LASV_B200_INSERT", "binned")
    monkeypatch.setenv("LASV_B200_CHUNK_KB", "64")
    L, W, n_reads, n_calls = 13, 40, 24_000, 4
    w = synth.scaled(synth.CFG0, 80_000, L)
    w.bucket_size = W
    p, seq, qual, offs, og = _synth_oracle(w, n_reads, seed=17)
    want = O.sort_records(og.records())
    per_reads = n_reads // n_calls
    per = per_reads * (w.read_len + 1)
    h_seq = torch.from_numpy(seq.copy()).pin_memory()
    h_qual = torch.from_numpy(qual.copy()).pin_memory()
    rows = torch.zeros((n_calls, 9), dtype=torch.int64).pin_memory()
    with DBGraph(w.kmer_size, L, W) as g:
        assert g.set_staging(5 * (128 << 20)) == 5 * (128 << 20)
        g.set_insert_mode(_capi.INSERT_STREAMED)
        if accumulate:
            g.set_accumulation(int(accumulate * per))
        torch.cuda.synchronize()
        for c in range(n_calls):
            lo = c * per
            o = (offs[c * per_reads: (c + 1) * per_reads + 1] - np.uint64(lo)).astype(np.uint64)
            g.load_reads_host_async(h_seq[lo: lo + per], h_qual[lo: lo + per], o, w.cutoff, rows[c])
            if c == 1:      # the buffers of a finished call may be overwritten at once
                h_seq[:per].fill_(ord("N"))
        st = g.stats()      # synchronous: inserts what the stripes still hold
        names = DBGraph.STATS_ASYNC_FIELDS
        reads = rows[:, names.index("n_reads")].tolist()
        assert reads == [per_reads * (c + 1) for c in range(n_calls)]
        assert st["kmers_inserted"] == og.stats.kmers_inserted and st["unique_kmers"] == len(want)
        assert st["bad_reads"] == og.stats.bad_reads and st["bases_loaded"] == og.stats.bases_loaded
        assert np.array_equal(O.sort_records(g.records()), want)
        n_ins = g.insert_counts()
        assert n_ins["random"] == 0 and n_ins["streamed"] == (n_calls if not accumulate else 2)
        # and the synchronous call still works on the same ring
        h_seq[:per].copy_(torch.from_numpy(seq[:per].copy()))
        o = offs[: per_reads + 1].astype(np.uint64)
        st2 = g.load_reads_host(h_seq[:per], h_qual[:per], o, w.cutoff)
        assert st2["n_reads"] == per_reads and st2["kmers_inserted"] > 0


def test_human_scale_table_geometry_matches_oracle():
    """BASELINE.json configs[4] at its REAL geometry: k=53, 2^24 buckets x 250 slots = 107 GB of 128-byte lines
    (64-bit line offsets, 50 lines per bucket, 65 536 regions).  A deterministic subsample of the config's read set
    (every 64th pair) goes through the streamed insert with accumulation and through the random-access insert into
    that table; every node is read back and compared with the oracle's graph of the same reads (the multiset does
    not depend on the table geometry as long as no bucket chain overflows)."""
    torch = torch_cuda()
    free_b, _ = torch.cuda.mem_get_info()
    w = synth.CFG4
    if free_b < (112 << 30):
        pytest.skip("the human-scale table needs 107 GB of device memory")
    pairs, every = 8_000, 64
    p = w.params()
    parts = [synth.reads_host(p, 2 * every * i, 2) for i in range(pairs)]
    seq = np.concatenate([a for a, _, _ in parts])
    qual = np.concatenate([b for _, b, _ in parts])
    n_reads, stride = 2 * pairs, w.read_len + 1
    offs = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(stride))
    og = O.OracleGraph(w.kmer_size, 14, w.bucket_size)
    assert og.load_reads(seq, qual, offs, w.cutoff, paired=True)
    want = O.sort_records(og.records())
    st_ = torch.cuda.current_stream().cuda_stream
    pad = np.full(16, 10, np.uint8)
    d_seq = torch.from_numpy(np.concatenate([seq, pad])).cuda()
    d_qual = torch.from_numpy(np.concatenate([qual, pad])).cuda()
    chunk = ((n_reads + 2) // 3 + 1) // 2 * 2
    d_offs = torch.arange(chunk + 1, dtype=torch.int64, device="cuda") * stride
    with DBGraph(w.kmer_size, w.number_bits, w.bucket_size) as g:
        assert g.table_bytes == (1 << 24) * 50 * 128
        for mode, kind in ((_capi.INSERT_STREAMED, "streamed"), (_capi.INSERT_RANDOM, "random")):
            g.set_accumulation(0)
            g.clear(st_)
            torch.cuda.synchronize()
            before = g.insert_counts()
            g.set_insert_mode(mode)
            g.set_accumulation(2 * chunk * stride)      # two of the three batches per insert
            for c0 in range(0, n_reads, chunk):
                nr = min(chunk, n_reads - c0)
                g.load_reads_device(d_seq[c0 * stride:], d_qual[c0 * stride:], nr * stride, d_offs[: nr + 1], nr,
                                    w.cutoff, st_)
            st = g.stats(stream=st_)
            after = g.insert_counts()
            assert after[kind] - before[kind] == 2, (kind, before, after)
            assert st["kmers_inserted"] == og.stats.kmers_inserted and st["unique_kmers"] == len(want)
            assert st["bad_reads"] == og.stats.bad_reads
            assert np.array_equal(O.sort_records(g.records()), want), kind
            s = g.summary()
            assert s["sum_covg"] == og.stats.kmers_inserted and s["n_nodes"] == len(want)

@pytest.mark.parametrize("mode", ["direct", "binned", "random", "streamed"])
def test_random_inputs_match_oracle(mode, monkeypatch):
    """Seeded random reads (bad characters, low qualities, low-complexity, empty and short reads) and
    random table geometries -- bucket sizes that are not a multiple of the slots per line, tables that
    overflow -- through the real kernels, fused and partitioned, against the oracle."""
    monkeypatch.setenv("LASV_B200_INSERT", mode)
    rng = np.random.default_rng(20261019)
    alphabet = np.array(list("ACGT" * 12 + "acgt" * 2 + "NnX-"))
    quals = np.array([chr(q) for q in (73, 73, 73, 73, 73, 60, 40, 39, 38, 37, 35, 34, 33)])
    for case in range(24):
        k = int(rng.choice([3, 5, 21, 31, 33, 53, 63, 65, 95]))
        L, W = int(rng.integers(2, 8)), int(rng.choice([1, 2, 3, 5, 7, 8, 13, 40, 100]))
        q_thresh = int(rng.choice([0, 5, 20]))
        monkeypatch.setenv("LASV_B200_BINS", str(1 << int(rng.integers(0, L + 1))))
        reads, qs = [], []
        for _ in range(int(rng.integers(0, 400))):
            n = int(rng.integers(0, 300))
            if rng.integers(0, 10) == 0:
                unit = str(rng.choice(["A", "AC", "ACG", "T", "GC"]))
                reads.append((unit * (n // len(unit) + 1))[:n])
            else:
                reads.append("".join(rng.choice(alphabet, n)))
            qs.append("".join(rng.choice(quals, n)))
        seq, qual, offs = O.reads_to_streams(reads, qs)
        og, ok = util.oracle_build(k, L + 6, W, seq, qual, offs, q_thresh + 33)       # roomy: the true node set
        assert ok
        distinct, capacity = og.unique_kmers, (1 << L) * W
        if 0.6 * capacity < distinct <= capacity:
            continue          # whether a 16-bucket rehash chain fills up first depends on the insertion order
        with DBGraph(k, L, W) as g:
            if distinct > capacity:
                with pytest.raises(_capi.TableFullError):
                    g.load_reads_host(seq, qual, offs, q_thresh + 33)
                continue
            st = g.load_reads_host(seq, qual, offs, q_thresh + 33)
            tag = f"case {case}: k={k} L={L} W={W} mode={mode}"
            assert st["kmers_inserted"] == og.stats.kmers_inserted and st["bad_reads"] == og.stats.bad_reads, tag
            assert st["bases_loaded"] == og.stats.bases_loaded, tag
            util.assert_same_records(g.records(), og.records(), tag)
            table, nxt = g.export_table()
            occ = (table["status"] != 0).reshape(1 << L, W)
            assert (occ == (np.arange(W)[None, :] < nxt[:, None])).all(), tag


@pytest.mark.parametrize("mode", [None, "streamed"])
def test_hot_kmers_and_duplicate_reads_are_counted_exactly(mode, monkeypatch):
    """Thousands of copies of the same reads (and homopolymers): every occurrence is
    an atomic on the same few slots; warp aggregation and L2 reductions (shared-memory
    atomics in the streamed insert, where all copies of a key meet in ONE group) must
    not lose or double count."""
    if mode:
        monkeypatch.setenv("LASV_B200_INSERT", mode)
        monkeypatch.setenv("LASV_B200_BINS", "4")
    rng = np.random.default_rng(3)
    base = ["".join(rng.choice(list("ACGT"), 100)) for _ in range(3)] + ["A" * 100, "AC" * 50, "ACG" * 33 + "A"]
    reads = [base[i % len(base)] for i in range(60_000)]
    quals = ["I" * 100] * len(reads)
    seq, qual, offs = O.reads_to_streams(reads, quals)
    for k in (31, 53, 95):
        og, ok = util.oracle_build(k, 8, 40, seq, qual, offs, 38)
        assert ok
        with DBGraph(k, 8, 40) as g:
            st = g.load_reads_host(seq, qual, offs, 38)
            assert st["kmers_inserted"] == og.stats.kmers_inserted
            util.assert_same_records(g.records(), og.records(), f"k={k}")


def test_table_full_is_fatal_not_silent():
    w = synth.scaled(synth.CFG0, 50_000, 6)
    w.bucket_size = 20                       # 1280 slots for ~50k distinct k-mers
    seq, qual, offs = synth.reads_host(w.params(), 0, 4000)
    with DBGraph(w.kmer_size, 6, 20) as g:
        with pytest.raises(_capi.TableFullError) as e:
            g.load_reads_host(seq, qual, offs, w.cutoff)
        assert "not allocated enough memory" in str(e.value)      # hash_table.c:312-316


def test_nearly_full_table_uses_the_rehash_chain():
    name = "synth_full_k53"
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        h = g.rehash_histogram()
        assert h[1:].sum() > 0 and h.sum() == meta["inserts"]
        util.assert_same_records(g.records(), util.case_records(name), name)
        table, nxt = g.export_table()
        _check_reference_layout(table, nxt, meta["k"], meta["L"], meta["W"])


# ---------------------------------------------- reference-layout host HashTable
def _check_reference_layout(table, nxt, k, L, W):
    """hole-free buckets; each key sits in a bucket of ITS reference rehash chain
    with every earlier bucket of the chain full (hash_table.c:69-122,234-267)."""
    N = lasv_b200.words_per_kmer(k)
    t2 = table.reshape(1 << L, W)
    occ = t2["status"] != 0
    assert np.array_equal(occ.sum(axis=1), nxt)
    assert (occ == (np.arange(W)[None, :] < nxt[:, None])).all()
    raw = table.view(np.uint8).reshape(len(table), 8 * N + 8)
    assert not raw[:, 8 * N + 6:].any() and not raw[~occ.reshape(-1)].any()
    assert (t2["status"][occ] == 1).all()
    bs, ss = np.nonzero(occ)
    step = max(1, len(bs) // 400)
    for b, s in zip(bs[::step], ss[::step]):
        chain = [O.hash_value(t2["kmer"][b, s], r, L) for r in range(16)]
        assert b in chain
        assert all(nxt[c] == W for c in chain[: chain.index(b)])


def _libref(k):
    so = O.REF_DIR / f"libref_{O.MAXK_OF_N[O.words_per_kmer(k)]}.so"
    return C.CDLL(str(so)) if so.exists() else None


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95"])
def test_export_import_and_reference_hash_table_find(name):
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    gold = util.case_records(name)
    seq, qual, offs = util.case_streams(name)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        table, nxt = g.export_table()
        _check_reference_layout(table, nxt, k, L, W)
        # finalised tables still answer find() and dump, but refuse inserts
        found, covg, edges = g.find(gold["kmer"])
        assert found.all() and np.array_equal(covg, gold["covg"]) and np.array_equal(edges, gold["edges"])
        util.assert_same_records(g.records(), gold)
        with pytest.raises(_capi.LasvError):
            g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
    # the REFERENCE's own hash_table_find (compiled from /root/reference) on our table
    ref = _libref(k)
    if ref is not None:
        ht = _capi.RefHashTable(k, 1 << L, W, table.ctypes.data, nxt.ctypes.data, None, len(gold), 15)
        ref.hash_table_find.restype = C.c_void_p
        ref.hash_table_find.argtypes = [C.c_void_p, C.POINTER(_capi.RefHashTable)]
        esz = table.dtype.itemsize
        for i in range(0, len(gold), 7):
            key = np.ascontiguousarray(gold["kmer"][i])
            ptr = ref.hash_table_find(key.ctypes.data, C.byref(ht))
            assert ptr, "reference hash_table_find misses a key in the exported table"
            e = table[(ptr - table.ctypes.data) // esz]
            assert np.array_equal(e["kmer"], key) and e["coverage"] == gold["covg"][i] and e["edges"] == gold["edges"][i]
        miss = np.ascontiguousarray(gold["kmer"][0] ^ np.uint64(0x15555))
        if tuple(miss.tolist()) not in {tuple(x) for x in gold["kmer"].tolist()}:
            assert not ref.hash_table_find(miss.ctypes.data, C.byref(ht))
    # importing a reference-layout table (an earlier load) and continuing gives the same graph
    with DBGraph(k, L, W) as g2:
        g2.import_table(table)
        util.assert_same_records(g2.records(), gold)
        g2.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        twice = gold.copy()
        twice["covg"] *= 2
        util.assert_same_records(g2.records(), twice)


def _reference_finds_every_key(table, nxt, gold, k, L, W):
    ref = _libref(k)
    if ref is None:
        return
    ht = _capi.RefHashTable(k, 1 << L, W, table.ctypes.data, nxt.ctypes.data, None, len(gold), 15)
    ref.hash_table_find.restype = C.c_void_p
    ref.hash_table_find.argtypes = [C.c_void_p, C.POINTER(_capi.RefHashTable)]
    esz = table.dtype.itemsize
    for i in range(0, len(gold), 5):
        key = np.ascontiguousarray(gold["kmer"][i])
        ptr = ref.hash_table_find(key.ctypes.data, C.byref(ht))
        assert ptr, "reference hash_table_find misses a key in the merged table"
        e = table[(ptr - table.ctypes.data) // esz]
        assert np.array_equal(e["kmer"], key) and e["coverage"] == gold["covg"][i] and e["edges"] == gold["edges"][i]


@pytest.mark.parametrize("name,shards", [("synth_full_k53", 2), ("synth_full_k53", 4), ("synth_full_k31", 2),
                                         ("synth_cfg2_k95", 4), ("synth_cfg0_k53", 8)])
def test_multi_gpu_graph_in_one_process(name, shards, monkeypatch):
    """lasv_multi_graph: the hash-sharded build driven by one host thread (the drop-in loaders with
    LASV_B200_DEVICES).  With one GPU all shards live on device 0 -- same code, the peer copies become local ones.
    The 90 %-full cases fill buckets, so keys are rehashed INSIDE their shard; the export must put them back on
    the reference's own chain over all buckets: hole-free buckets, every key behind full buckets only, and the
    compiled reference's hash_table_find finds every key with its coverage and edges."""
    monkeypatch.setenv("LASV_B200_BINS", "2")
    monkeypatch.setenv("LASV_B200_CHUNK_KB", "256")
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    gold = util.case_records(name)
    seq, qual, offs = util.case_streams(name)
    n_dev = _capi.lib().lasv_device_count()
    devices = [i % n_dev for i in range(shards)]
    n_reads = len(offs) - 1
    import torch
    h_seq, h_qual = torch.from_numpy(seq.copy()).pin_memory(), torch.from_numpy(qual.copy()).pin_memory()
    with lasv_b200.MultiDBGraph(k, L, W, devices, round_bytes=len(seq) // 3) as mg:
        assert mg.n_shards == shards
        pieces = 7
        for c in range(pieces):
            r0, r1 = n_reads * c // pieces, n_reads * (c + 1) // pieces
            lo, hi = int(offs[r0]), int(offs[r1])
            mg.load_reads_host(h_seq[lo:hi], h_qual[lo:hi], (offs[r0: r1 + 1] - offs[r0]).astype(np.uint64),
                               util.cutoff_of(meta))
        st = mg.stats()
        assert st["kmers_inserted"] == meta["inserts"] and st["unique_kmers"] == len(gold)
        assert st["bad_reads"] == meta["bad_reads"] and st["bases_loaded"] == meta["bases_loaded"]
        assert int(mg.rehash_histogram().sum()) == meta["inserts"]
        s = mg.summary()
        assert s["n_nodes"] == len(gold) and s["sum_covg"] == meta["inserts"]
        table, nxt, uniq = mg.export_table()
    assert uniq == len(gold)
    _check_reference_layout(table, nxt, k, L, W)
    occ = table["status"] != 0
    got = np.zeros(int(occ.sum()), dtype=gold.dtype)
    got["kmer"], got["covg"], got["edges"] = table["kmer"][occ], table["coverage"][occ], table["edges"][occ]
    util.assert_same_records(got, gold, name)
    _reference_finds_every_key(table, nxt, gold, k, L, W)


def test_import_keeps_pruned_nodes_pruned():
    """A reference-layout table that went through the cleaning holds `pruned` Elements (edges reset, not part of the
    graph: element.h:57-75).  Importing it -- what the drop-ins do when more reads or a .ctx are loaded onto a graph
    that is not empty -- must keep them pruned: out of the dumps, and not counted as rehashes a second time."""
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        raw = len(g.records())
        g.clean(2)
        cleaned = O.sort_records(g.records())
        table, nxt = g.export_table()
    assert (table["status"] == 3).sum() == raw - len(cleaned) > 0
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g2:
        g2.import_table(table)
        assert np.array_equal(O.sort_records(g2.records()), cleaned)
        t2, _ = g2.export_table()
        assert (t2["status"] == 3).sum() == (table["status"] == 3).sum()


@pytest.mark.parametrize("devices", [None, "0,0", "all"])
def test_dropin_loader_fills_a_reference_hash_table(devices, tmp_path, monkeypatch):
    """(A) of include/lasv_b200.h: the reference-named loader on a calloc'ed
    reference HashTable, as graph_operations.c:757-765 calls it -- on one GPU, and hash-sharded over several
    (LASV_B200_DEVICES: two shards on device 0, or every GPU of the box)."""
    if devices:
        monkeypatch.setenv("LASV_B200_DEVICES", devices)
    name = "edge_pe_k53_q5"
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = lasv_b200.words_per_kmer(k)
    table = np.zeros((1 << L) * W, dtype=lasv_b200.element_dtype(N))
    nxt = np.zeros((1 << L) * 2, dtype=np.int16)             # calloc(number_buckets, sizeof(int))
    coll = np.zeros(15, dtype=np.int64)
    ht = _capi.RefHashTable(k, 1 << L, W, table.ctypes.data, nxt.ctypes.data, coll.ctypes.data, 0, 15)
    f1, f2 = util.case_files(name)
    (tmp_path / "l").write_text(f"{f1}\n")
    (tmp_path / "r").write_text(f"{f2}\n")
    files = C.c_uint(0)
    bad, dup, parsed, loaded = (C.c_ulonglong(0) for _ in range(4))
    hist = np.zeros(20001, dtype=np.uint64)
    _capi.lib().load_pe_filelists_into_graph_colour(
        str(tmp_path / "l").encode(), str(tmp_path / "r").encode(), meta["q_thresh"], 0, 0, bytes([33]), 0,
        C.byref(ht), b"\0", C.byref(files), C.byref(bad), C.byref(dup), C.byref(parsed), C.byref(loaded),
        hist.ctypes.data, len(hist))
    assert files.value == 1 and bad.value == meta["bad_reads"]
    assert parsed.value == meta["bases_parsed"] and loaded.value == meta["bases_loaded"]
    assert ht.unique_kmers == meta["n_records"] and coll.sum() == meta["inserts"]
    assert mean_contig_length(hist) == meta["ctx_mean_read_len"]
    _check_reference_layout(table, nxt[: 1 << L], k, L, W)
    occ = table[table["status"] != 0]
    got = np.zeros(len(occ), dtype=lasv_b200.record_dtype(N))
    got["kmer"], got["covg"], got["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
    util.assert_same_records(got, util.case_records(name))


# ----------------------------------------------------- .ctx and the consumers
@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95"])
def test_ctx_roundtrip_and_reference_consumer(name, tmp_path):
    """Config 3: the reference's own supernode walk (dB_graph --multicolour_bin ...
    --output_supernodes) on OUR dumped graph gives the reference's contig set."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    gold = util.case_records(name)
    seq, qual, offs = util.case_streams(name)
    out = tmp_path / "gpu.ctx"
    with DBGraph(k, L, W) as g:
        st = g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        _, hist = g.stats(hist_size=20001)
        g.dump_binary(out, mean_contig_length(hist), st["bases_loaded"])
    ctx = O.parse_ctx(out)
    assert ctx["header"].hex() == meta["ctx_header_hex"]
    util.assert_same_records(ctx["records"], gold)
    with DBGraph(k, L, W) as g:                                # load_multicolour_binary_from_filename_into_graph
        mrl, ts = g.load_binary(out)
        assert (mrl, ts) == (meta["ctx_mean_read_len"], meta["ctx_total_seq"])
        util.assert_same_records(g.records(), gold)
        g.load_records(gold)                                   # merging a second binary adds covg, ORs edges
        twice = gold.copy()
        twice["covg"] *= 2
        util.assert_same_records(g.records(), twice)
    if O.ref_available(k):
        res = O.run_reference(k, L, W, None, ctx_in=out, workdir=tmp_path, want_supernodes=True)
        util.assert_same_records(res["ctx"]["records"], gold)  # reloaded + re-dumped by the reference
        assert O.canonical_seq_set(res["supernodes"]) == meta["supernodes"]
    # a truncated file is an error, not a silently smaller graph
    cut = tmp_path / "cut.ctx"
    cut.write_bytes(out.read_bytes()[:-3])
    with DBGraph(k, L, W) as g:
        with pytest.raises(_capi.LasvError, match="truncated"):
            g.load_binary(cut)


def _oracle_supernodes_in_table_order(k, L, W, recs, path):
    """The reference's sequential supernode walk (oracle restatement) over the given record order."""
    og = O.OracleGraph(k, L, W)
    assert og.load_records(recs)
    return og.print_supernodes(path, max_length=1 << 20)


@pytest.mark.parametrize("name", [n for n in util.ALL_CASES if "supernodes" in util.case_meta(n)])
def test_gpu_supernodes_match_reference(name, tmp_path):
    """SURVEY 8f rank 2, `--output_supernodes` on the device graph: the reference's contig set
    (golden), and byte for byte the FASTA the reference's traversal prints for OUR slot order --
    before and after finalisation, against the oracle's restatement and the compiled reference."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    seq, qual, offs = util.case_streams(name)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        st = g.output_supernodes(tmp_path / "gpu.fa")
        recs = g.records()
        g.dump_binary(tmp_path / "gpu.ctx", 100, 1000)
        seqs = O.read_fasta_seqs(tmp_path / "gpu.fa")
        assert st["n_supernodes"] == len(seqs) and st["n_nodes"] == meta["n_records"]
        assert O.canonical_seq_set(seqs) == meta["supernodes"]
        ost = _oracle_supernodes_in_table_order(k, L, W, recs, tmp_path / "oracle.fa")
        gpu = (tmp_path / "gpu.fa").read_bytes()
        assert gpu == (tmp_path / "oracle.fa").read_bytes()
        assert ost["supernodes"] == st["n_supernodes"]
        g.finalize()                                            # reference layout: same order, same output
        st2 = g.output_supernodes(tmp_path / "gpu_final.fa")
        assert (tmp_path / "gpu_final.fa").read_bytes() == gpu and st2 == st
    if O.ref_available(k):
        O.run_reference(k, L, W, None, ctx_in=tmp_path / "gpu.ctx", workdir=tmp_path, want_supernodes=True,
                        extra_args=["--max_var_len", str(1 << 20)])
        assert (tmp_path / "sup.fa").read_bytes() == gpu


@pytest.mark.parametrize("k,err_q20,n_reads,genome,L", [(21, 30000, 400, 3000, 10), (15, 80000, 400, 3000, 10),
                                                        (53, 30000, 400, 3000, 10), (95, 20000, 400, 3000, 10),
                                                        (31, 20000, 40000, 200000, 17)])
def test_gpu_supernodes_order_dependent_cases(k, err_q20, n_reads, genome, L, tmp_path):
    """Dense errors put junction next to junction: whether a one-edge supernode is printed depends on
    which end node the traversal visited first.  The device enumeration + replay must print exactly
    what the reference's sequential walk prints over the same slot order."""
    base = synth.CFG1 if k <= 31 else synth.CFG0 if k <= 63 else synth.CFG2
    w = synth.scaled(base, genome, L)
    w.seed = 100 + k
    w.kmer_size = k
    p = w.params()
    p.err_q20 = err_q20
    seq, qual, offs = synth.reads_host(p, 0, n_reads)
    with DBGraph(k, L, 40) as g:
        g.load_reads_host(seq, qual, offs, 38)
        st = g.output_supernodes(tmp_path / "gpu.fa")
        recs = g.records()
        g.dump_binary(tmp_path / "gpu.ctx", 100, 1000)
    gpu = (tmp_path / "gpu.fa").read_bytes()
    _oracle_supernodes_in_table_order(k, L, 40, recs, tmp_path / "oracle.fa")
    assert gpu == (tmp_path / "oracle.fa").read_bytes()
    assert st["n_not_printed"] > 0 or k == 95
    assert st["n_nodes"] == len(recs) and st["total_bases"] == sum(len(s) for s in O.read_fasta_seqs(tmp_path / "gpu.fa"))
    if O.ref_available(k) and n_reads <= 1000:
        O.run_reference(k, L, 40, None, ctx_in=tmp_path / "gpu.ctx", workdir=tmp_path, want_supernodes=True,
                        extra_args=["--max_var_len", str(1 << 20)])
        assert (tmp_path / "sup.fa").read_bytes() == gpu


def test_gpu_supernodes_cycles_and_long_paths(tmp_path):
    """A circular genome read without errors is one cycle of interior nodes (no node ends a path);
    an error-free linear genome is one long path (thousands of edges walked by one thread)."""
    rng = np.random.default_rng(5)
    k = 31
    circ = "".join(rng.choice(list("ACGT"), size=500))
    lin = "".join(rng.choice(list("ACGT"), size=6000))
    reads = [(circ + circ)[i:i + 100] for i in range(0, 500, 7)] + [lin[i:i + 150] for i in range(0, 5851, 50)]
    seq, _, offs = O.reads_to_streams(reads)
    with DBGraph(k, 10, 40) as g:
        g.load_reads_host(seq, None, offs, 0)
        st = g.output_supernodes(tmp_path / "gpu.fa", max_length=1000)
        recs = g.records()
    _oracle_supernodes_in_table_order(k, 10, 40, recs, tmp_path / "oracle.fa")
    assert (tmp_path / "gpu.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
    assert st["n_cycles"] == 1 and st["n_supernodes"] == 2
    assert st["longest_edges"] == 6000 - k and st["over_max_length"] == 1
    seqs = sorted(O.read_fasta_seqs(tmp_path / "gpu.fa"), key=len)
    assert len(seqs[0]) == 500 + k and len(seqs[1]) == 6000


def test_ctx_dump_streams_through_a_bounded_buffer(tmp_path, monkeypatch):
    """A table larger than the spare device memory is dumped range of buckets by range of buckets
    (a 100 GB table holds ~65 GB of records): force hundreds of ranges on a small table."""
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    gold = util.case_records(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        whole = g.records()
        monkeypatch.setenv("LASV_B200_DUMP_CHUNK_BYTES", str(meta["W"] * 21 * 3))     # three buckets per range
        parts = g.records()
        assert np.array_equal(parts, whole)                      # same records, same (table) order
        util.assert_same_records(parts, gold, name)
        g.dump_binary(tmp_path / "x.ctx", 7, 11)
        ctx = O.parse_ctx(tmp_path / "x.ctx")
        assert np.array_equal(ctx["records"], whole)


# ------------------------------------ the reference's own main() on the GPU loader
@pytest.mark.parametrize("name,devices", [("synth_cfg1_k31", None), ("synth_cfg0_k53", None), ("synth_cfg2_k95", None),
                                          ("synth_cfg0_k53", "0,0,0,0"), ("synth_cfg1_k31", "all")])
def test_reference_cli_with_gpu_loader(name, devices, tmp_path, monkeypatch):
    """integration/_build/dB_graph_gpu_*: reference main(), cmd-line parsing, cleaning,
    supernode walk and .ctx dump compiled from the reference sources, with ONLY the
    two loader symbols coming from liblasv_b200.so (INTEGRATION.md level A).  With LASV_B200_DEVICES the same
    main() builds on a hash-sharded multi-GPU graph (four shards on device 0 / every GPU of the box)."""
    from pathlib import Path
    if devices:
        monkeypatch.setenv("LASV_B200_DEVICES", devices)
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / f"dB_graph_gpu_{O.MAXK_OF_N[O.words_per_kmer(k)]}"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text(f"{tmp_path / 'a.fq'}\n")
    (tmp_path / "r").write_text(f"{tmp_path / 'b.fq'}\n")
    cmd = [str(exe), "--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W),
           "--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold", str(meta["q_thresh"]),
           "--dump_binary", str(tmp_path / "out.ctx"), "--output_supernodes", str(tmp_path / "sup.fa")]
    p = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    ctx = O.parse_ctx(tmp_path / "out.ctx")
    assert ctx["header"].hex() == meta["ctx_header_hex"]                  # incl. mean read length, total seq
    util.assert_same_records(ctx["records"], util.case_records(name), name)
    assert O.canonical_seq_set(O.read_fasta_seqs(tmp_path / "sup.fa")) == meta["supernodes"]
    assert f"Number of bad reads:{meta['bad_reads']}" in p.stdout
    assert f"Total bases passing filters and loaded into graph:{meta['bases_loaded']}" in p.stdout
    tries = sum(int(l.split(":")[1]) for l in p.stdout.splitlines() if l.strip().startswith("tries "))
    assert tries == meta["inserts"]


@pytest.mark.parametrize("devices", [None, "0,0"])
def test_reference_assembly_mode_runs_unchanged_on_the_gpu_built_table(devices, tmp_path, monkeypatch):
    """SURVEY 8f rank 3, second half: the read-span linkage table and the linkage-guided BFS contig assembly
    (`--mode assembly --read_span_file`: build_linkage_from_files.c:411-499, linkage_operations.c:167-541,
    little_hash_for_linkage.c:227) are consumers of the graph -- sequential, order-dependent host code that runs
    UNCHANGED (the reference's own main() and functions) on the HashTable the GPU loader fills.  A diploid genome
    with SNPs and a duplicated segment keeps branching nodes after the cleaning, so the BFS consults the linkage
    table: contigs.fa must hold the same contigs as the all-CPU reference binary's."""
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("assembly_mode_check", root / "tests" / "perf" / "assembly_mode_check.py")
    A = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(A)
    exe = root / "integration" / "_build" / "dB_graph_gpu_63"
    if not exe.exists() or not O.ref_available(53):
        pytest.skip("integration/_build or oracle/_ref not built (both need /root/reference at build time)")
    w = A.make_branching_inputs(tmp_path, 8000, 2000, 7)
    ref = A.run(O.REF_DIR / "dB_graph_63", tmp_path, w, "ref")
    env = dict(os.environ, LASV_B200_DEVICES=devices) if devices else None
    got = A.run(exe, tmp_path, w, "gpu", env)
    assert ref.count(b">") > 100                      # the BFS branched
    assert A.contig_set(got) == A.contig_set(ref)


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95"])
@pytest.mark.parametrize("clean", [False, True])
def test_reference_cli_prints_supernodes_through_the_gpu(name, clean, tmp_path):
    """the reference's main() with `--output_supernodes` bound to the GPU
    (integration/supernode_shim.c + lasv_ref_hash_table_write_supernodes), on the table the GPU
    loader left in host memory -- raw, and after the reference's own tip clipping and
    low-coverage supernode removal (pruned nodes stay in the table, isolated).  The FASTA must be
    what the reference's sequential walk prints for that table: the oracle's restatement walks the
    `.ctx` the same run dumped (same slot order) and must produce the same bytes."""
    from pathlib import Path
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / f"dB_graph_gpusn_{O.MAXK_OF_N[O.words_per_kmer(k)]}"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text(f"{tmp_path / 'a.fq'}\n")
    (tmp_path / "r").write_text(f"{tmp_path / 'b.fq'}\n")
    cmd = [str(exe), "--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W),
           "--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold", str(meta["q_thresh"]),
           "--dump_binary", str(tmp_path / "out.ctx"), "--output_supernodes", str(tmp_path / "sup.fa")]
    if clean:
        cmd += ["--remove_low_coverage_supernodes", "2"]
    p = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "nodes visted" in p.stdout
    ctx = O.parse_ctx(tmp_path / "out.ctx")
    seqs = O.read_fasta_seqs(tmp_path / "sup.fa")
    if not clean:
        util.assert_same_records(ctx["records"], util.case_records(name), name)
        assert O.canonical_seq_set(seqs) == meta["supernodes"]
    else:
        assert 0 < len(ctx["records"]) < meta["n_records"]
    _oracle_supernodes_in_table_order(k, L, W, ctx["records"], tmp_path / "oracle.fa")
    assert (tmp_path / "sup.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()


def test_walk_limit_is_reported_not_ignored(tmp_path):
    """--max_var_len (default 10 000 edges): the reference cuts its supernode walks there and prints / judges
    overlapping pieces (dB_graph_population.c:2739-2767, 3371-3462).  The *_strict / *_limited calls refuse such
    a graph (LASV_ERR_LIMIT, nothing written or changed) instead of silently printing the supernode whole; with
    a limit nothing exceeds they are the plain calls."""
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    with DBGraph(meta["k"], meta["L"], meta["W"]) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        whole = g.output_supernodes(tmp_path / "whole.fa", max_length=50)
        assert whole["over_max_length"] > 0 and whole["longest_edges"] > 50
        with pytest.raises(_capi.LasvError) as e:
            g.output_supernodes(tmp_path / "strict.fa", max_length=50, strict=True)
        assert e.value.code == _capi.LASV_ERR_LIMIT and not (tmp_path / "strict.fa").exists()
        ok = g.output_supernodes(tmp_path / "ok.fa", max_length=int(whole["longest_edges"]), strict=True)
        assert ok["over_max_length"] == 0
        assert (tmp_path / "ok.fa").read_bytes() == (tmp_path / "whole.fa").read_bytes()
        before = O.sort_records(g.records())
        g.clean(2, what=_capi.CLEAN_TIPS)
        tipped = O.sort_records(g.records())
        with pytest.raises(_capi.LasvError) as e:
            g.clean(2, what=_capi.CLEAN_SUPERNODES, max_length=5)     # error paths are ~k edges long
        assert e.value.code == _capi.LASV_ERR_LIMIT
        assert np.array_equal(O.sort_records(g.records()), tipped) and len(tipped) <= len(before)
        st = g.clean(2, what=_capi.CLEAN_SUPERNODES, max_length=10000)
        assert st["supernodes_pruned"] > 0 and len(g.records()) < len(tipped)


def test_reference_cli_supernodes_beyond_the_walk_limit(tmp_path):
    """The reference's main() with the GPU supernode printer and --max_var_len 50: supernodes of this graph have
    up to ~440 edges, so the GPU call refuses (LASV_ERR_LIMIT) and integration/supernode_shim.c lets the reference's
    own printer take over, so whatever is printed beyond the limit is the reference's doing.  With the default
    limit the GPU prints, byte for byte what the all-CPU reference prints for the same table (it reloads the .ctx
    the same run dumped: same slot order)."""
    from pathlib import Path
    name = "synth_cfg0_k53"
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / "dB_graph_gpusn_63"
    if not exe.exists() or not O.ref_available(k):
        pytest.skip("integration/_build or oracle/_ref not built (both need /root/reference at build time)")
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text(f"{tmp_path / 'a.fq'}\n")
    (tmp_path / "r").write_text(f"{tmp_path / 'b.fq'}\n")
    geo = ["--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W)]
    p = subprocess.run([str(exe)] + geo + ["--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold",
                                           str(meta["q_thresh"]), "--dump_binary", str(tmp_path / "out.ctx"),
                                           "--output_supernodes", str(tmp_path / "gpu.fa"), "--max_var_len", "50"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    # (the reference's printer overruns its own arrays beyond the limit, see below: our binary, which runs that very
    # code after the hand-over, may die in free() just like the all-CPU binary)
    assert p.returncode <= 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "the reference's own printer takes over" in p.stdout and "contig length equals max length" in p.stdout
    q = subprocess.run([str(O.ref_binary(k))] + geo + ["--multicolour_bin", str(tmp_path / "out.ctx"), "--output_supernodes",
                                                       str(tmp_path / "ref.fa"), "--max_var_len", "50"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    # Beyond the limit the reference itself is on thin ice: its path arrays hold max_length entries and the walk
    # writes entry [max_length] (dB_graph_population.c:2739-2742 vs :864) -- with glibc the all-CPU binary prints
    # its pieces and then dies in free().  If it survives, the bytes must agree; either way it was the reference's
    # own code that produced them, not a silent GPU approximation.
    if q.returncode == 0 and p.returncode == 0:
        assert (tmp_path / "gpu.fa").read_bytes() == (tmp_path / "ref.fa").read_bytes()
    # and with the default limit the GPU prints: same bytes as the reference
    for f in ("out.ctx",):
        (tmp_path / f).unlink()
    p = subprocess.run([str(exe)] + geo + ["--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold",
                                           str(meta["q_thresh"]), "--dump_binary", str(tmp_path / "out.ctx"),
                                           "--output_supernodes", str(tmp_path / "gpu2.fa")],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "takes over" not in p.stdout
    q = subprocess.run([str(O.ref_binary(k))] + geo + ["--multicolour_bin", str(tmp_path / "out.ctx"), "--output_supernodes",
                                                       str(tmp_path / "ref2.fa")],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert q.returncode == 0
    assert (tmp_path / "gpu2.fa").read_bytes() == (tmp_path / "ref2.fa").read_bytes()


def _table_records(table):
    occ = table[table["status"] == 1]
    out = np.zeros(len(occ), dtype=lasv_b200.record_dtype(table["kmer"].shape[1]))
    out["kmer"], out["covg"], out["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
    return out


def _clean_case(name, thresh, tmp_path, dense=None):
    """Build on the GPU, export the reference-layout table, clean it through
    lasv_ref_hash_table_clean and compare with the reference's sequential cleaning (oracle
    restatement, pinned on the compiled reference) of the same slot order."""
    if dense is None:
        meta = util.case_meta(name)
        k, L, W = meta["k"], meta["L"], meta["W"]
        seq, qual, offs = util.case_streams(name)
        cutoff = util.cutoff_of(meta)
    else:
        k, err_q20, n_reads, genome, L = dense
        W, cutoff = 40, 38
        base = synth.CFG1 if k <= 31 else synth.CFG0 if k <= 63 else synth.CFG2
        w = synth.scaled(base, genome, L)
        w.seed, w.kmer_size = 300 + k, k
        p = w.params()
        p.err_q20 = err_q20
        seq, qual, offs = synth.reads_host(p, 0, n_reads)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, cutoff)
        table, nxt = g.export_table()
    raw = _table_records(table)
    og = O.OracleGraph(k, L, W)
    assert og.load_records(raw)
    ost = og.clean(thresh)
    ht = _capi.RefHashTable(k, 1 << L, W, table.ctypes.data, nxt.ctypes.data, None, len(raw), 15)
    st = _capi.CleanStats()
    _capi.check(_capi.lib().lasv_ref_hash_table_clean(C.byref(ht), _capi.CLEAN_TIPS | _capi.CLEAN_SUPERNODES, thresh,
                                                    C.byref(st)))
    st = st.as_dict()
    assert np.array_equal(_table_records(table), og.records())          # same nodes, coverage, edges, same order
    assert int((table["status"] == 3).sum()) == st["nodes_pruned"] == len(raw) - ost["nodes_left"]
    assert not table["edges"][table["status"] == 3].any()
    assert st["tips_clipped"] == ost["tips"] and st["tip_nodes_pruned"] == ost["tip_nodes"]
    assert st["supernodes_pruned"] == ost["supernodes_pruned"]
    # the supernodes of the cleaned table, through the reference-table entry point as well
    sst = _capi.SupernodeStats()
    _capi.check(_capi.lib().lasv_ref_hash_table_write_supernodes(C.byref(ht), str(tmp_path / "gpu.fa").encode(), 10000,
                                                               C.byref(sst)))
    og.print_supernodes(tmp_path / "oracle.fa", max_length=1 << 20)
    assert (tmp_path / "gpu.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
    return st


@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", 2), ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 2),
                                         ("synth_cfg0_k53", 1), ("edge_pe_k21_q5", 1)])
def test_ref_table_clean_matches_reference_walk(name, thresh, tmp_path):
    """`--remove_low_coverage_supernodes T` (tip clipping + low-coverage supernode
    removal, both order-dependent) on the device, exact for the caller's slot order."""
    st = _clean_case(name, thresh, tmp_path)
    if name.startswith("synth"):
        assert st["tips_clipped"] > 0 and st["supernodes_pruned"] > 0


# dense1 once failed on the device with "path record buffers overflowed": a race between the pageable table upload
# and the first kernel on the non-blocking stream (the count pass saw fewer nodes than the list pass).  Fixed by a
# device synchronisation after the upload; all cases below green on a B200 since (profiles/r1/cleaning_gpu_pytest_*.log).


@pytest.mark.parametrize("dense", [(21, 40000, 600, 3000, 11), (53, 20000, 600, 3000, 11), (95, 20000, 600, 3000, 11),
                                   (31, 20000, 40000, 200000, 17)])
def test_ref_table_clean_dense_errors(dense, tmp_path):
    """Junction next to junction, several tips on one node, supernodes that merge once a
    neighbour is pruned: where the order of the traversal decides."""
    st = _clean_case(None, 2, tmp_path, dense=dense)
    assert st["tips_clipped"] > 0 and st["supernodes_pruned"] > 0


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95"])
def test_reference_cli_cleans_through_the_gpu(name, tmp_path):
    """The reference's main() with loaders, cleaning and supernode printing all bound to the GPU
    (integration/_build/dB_graph_gpuall_*): `--remove_low_coverage_supernodes 2` leaves a `.ctx`
    whose header carries the cleaning flags and whose records the oracle's supernode walk turns
    into the FASTA the same run printed."""
    from pathlib import Path
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / f"dB_graph_gpuall_{O.MAXK_OF_N[O.words_per_kmer(k)]}"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text(f"{tmp_path / 'a.fq'}\n")
    (tmp_path / "r").write_text(f"{tmp_path / 'b.fq'}\n")
    cmd = [str(exe), "--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W),
           "--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold", str(meta["q_thresh"]),
           "--remove_low_coverage_supernodes", "2", "--dump_binary", str(tmp_path / "out.ctx"),
           "--output_supernodes", str(tmp_path / "sup.fa")]
    p = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "tips clipped" in p.stdout
    ctx = O.parse_ctx(tmp_path / "out.ctx")
    want = bytearray(O.ctx_header(k, meta["ctx_mean_read_len"], meta["ctx_total_seq"], cleaned_threshold=2))
    got = bytearray(ctx["header"])
    pad = slice(6 + 16 + 12 + 4 + 9 + 10, 6 + 16 + 12 + 4 + 9 + 16)
    want[pad] = got[pad] = b"\0" * 6
    assert want == got
    assert 0 < len(ctx["records"]) < meta["n_records"]
    gold = util.case_records(name)                       # cleaning removes nodes and edges, never adds any
    keys = {tuple(r) for r in gold["kmer"].tolist()}
    assert all(tuple(r) in keys for r in ctx["records"]["kmer"].tolist())
    _oracle_supernodes_in_table_order(k, L, W, ctx["records"], tmp_path / "oracle.fa")
    assert (tmp_path / "sup.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "synth_full_k53"])
def test_reference_cli_reloads_ctx_through_the_gpu(name, tmp_path):
    """`dB_graph --multicolour_bin x.ctx` with the record loop of
    load_multicolour_binary_from_filename_into_graph replaced by the GPU loader
    (integration/ctx_loader_shim.c + lasv_ref_hash_table_load_ctx_records): the
    reference's own consumers (supernode walk, binary dump) must see exactly the graph
    the reference's single-threaded hash_table_insert loop would have built."""
    from pathlib import Path
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / f"dB_graph_gpu_{O.MAXK_OF_N[O.words_per_kmer(k)]}"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    gold = util.case_records(name)
    src = tmp_path / "in.ctx"
    src.write_bytes(bytes.fromhex(meta["ctx_header_hex"]) + gold.tobytes())      # a reference-format binary of the golden graph
    cmd = [str(exe), "--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W),
           "--multicolour_bin", str(src), "--dump_binary", str(tmp_path / "out.ctx"),
           "--output_supernodes", str(tmp_path / "sup.fa")]
    p = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert f"got {len(gold)} kmers" in p.stdout                                   # graph_operations.c:834
    ctx = O.parse_ctx(tmp_path / "out.ctx")
    assert ctx["header"].hex() == meta["ctx_header_hex"]                          # GraphInfo carried over by the reference's parser
    util.assert_same_records(ctx["records"], gold, name)
    if "supernodes" in meta:                                  # (the 90 %-full rehash-chain cases carry none)
        assert O.canonical_seq_set(O.read_fasta_seqs(tmp_path / "sup.fa")) == meta["supernodes"]
    # a file that is not a binary dies like the reference does
    bad = tmp_path / "bad.ctx"
    bad.write_bytes(b"NOTCORTEX" + bytes(100))
    p = subprocess.run(cmd[:7] + ["--multicolour_bin", str(bad), "--dump_binary", str(tmp_path / "out2.ctx")],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "not compatible" in (p.stderr + p.stdout)       # cmd_line.c checks the header first


# --------------------------------------------------- BASELINE-size properties
@pytest.mark.skipif(os.environ.get("LASV_SKIP_FULLSIZE") == "1", reason="LASV_SKIP_FULLSIZE=1")
def test_config0_full_size_properties():
    """BASELINE.json config 0 at full size (63 Mbp, 30x, 2x100 bp, k=53, L=22 W=100;
    ~0.9 G k-mer inserts into a 10 GB table).  Too big for the CPU oracle, so:
      * sum of coverages == number of inserts, nodes == unique counter;
      * the order-independent fingerprint is identical when the same reads are
        loaded in a different batch split and order, and through the
        records path (extract + insert) instead of the fused kernel;
      * a 1/256 subsample of the pairs built on the GPU equals the oracle."""
    torch = torch_cuda()
    w = synth.CFG0
    p = w.params()
    n_pairs = w.n_pairs
    n_reads = 2 * n_pairs
    stride = w.read_len + 1
    st_ = torch.cuda.current_stream().cuda_stream
    batch = 1 << 21                                            # reads per batch
    d_seq = torch.empty(batch * stride + 16, dtype=torch.uint8, device="cuda")
    d_qual = torch.empty_like(d_seq)

    def run(order, records_path=False):
        with DBGraph(w.kmer_size, w.number_bits, w.bucket_size) as g:
            N = g.N
            out = cnt = None
            if records_path:
                out = torch.empty((batch * 48 + 8, 2 * N + 1), dtype=torch.int32, device="cuda")
                cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            for b0 in order:
                nr = min(batch, n_reads - b0)
                synth.reads_device(p, b0, nr, d_seq, d_qual, st_)
                if records_path:
                    cnt.zero_()
                    g.extract_records_device(d_seq, d_qual, nr * stride, w.cutoff, out, out.shape[0], cnt, st_)
                    g.insert_records_device(out, int(cnt.item()), st_)
                else:
                    g.load_reads_device(d_seq, d_qual, nr * stride, None, 0, w.cutoff, st_)
            torch.cuda.synchronize()
            st = g.stats(stream=st_)
            return st, g.summary()

    starts = list(range(0, n_reads, batch))
    st1, s1 = run(starts)
    assert s1["sum_covg"] == st1["kmers_inserted"] > 700_000_000
    assert s1["n_nodes"] == st1["unique_kmers"]
    st2, s2 = run(starts[::-1])
    assert s2 == s1 and st2["kmers_inserted"] == st1["kmers_inserted"]
    # the records path counts inserts on the extract side only
    st3, s3 = run(starts[1::2] + starts[0::2], records_path=True)
    assert s3 == s1
    # subsample vs oracle: every 256th pair
    sel = np.arange(0, n_pairs, 256)
    seqs, quals = [], []
    for pr in sel:
        s, q, _ = synth.reads_host(p, 2 * int(pr), 2)
        seqs.append(s)
        quals.append(q)
    seq, qual = np.concatenate(seqs), np.concatenate(quals)
    offs = synth.offsets_fixed(w.read_len, 2 * len(sel))
    og = O.OracleGraph(w.kmer_size, 18, 100)
    assert og.load_reads(seq, qual, offs, w.cutoff, paired=True)
    with DBGraph(w.kmer_size, 18, 100) as g:
        st = g.load_reads_host(seq, qual, offs, w.cutoff)
        assert st["kmers_inserted"] == og.stats.kmers_inserted
        assert np.array_equal(O.sort_records(g.records()), O.sort_records(og.records()))


# ---- branch printing, cleaning on a graph handle, --health_check, list ranking, grouped insert -----------------------
# (first green device run: round 1's round-end suite; plain tests since round 2)


def _oracle_branches_in_table_order(k, L, W, recs, path):
    og = O.OracleGraph(k, L, W)
    assert og.load_records(recs)
    st = og.print_branches(path, keep_marks=True)
    return og, st


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "edge_pe_k21_q5"])
def test_gpu_branches_match_reference_walk(name, tmp_path):
    """The reference's branches.fa for OUR slot order, byte for byte, in both arrangements of the table."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    seq, qual, offs = util.case_streams(name)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        recs = g.records()
        st = g.output_branches(tmp_path / "gpu.fa")
        _, ost = _oracle_branches_in_table_order(k, L, W, recs, tmp_path / "oracle.fa")
        gpu = (tmp_path / "gpu.fa").read_bytes()
        assert gpu == (tmp_path / "oracle.fa").read_bytes()
        assert st["n_branches"] == ost["branches"] and st["n_branching_nodes"] == ost["branching_nodes"]
        assert st["bytes"] == len(gpu)
        if name.startswith("synth"):
            assert st["n_branches"] > 0
        assert np.array_equal(g.records(), recs)                # the graph is left as it was
        g.finalize()
        st2 = g.output_branches(tmp_path / "gpu_final.fa")
        assert (tmp_path / "gpu_final.fa").read_bytes() == gpu and st2 == st


@pytest.mark.parametrize("dense", [(21, 40000, 600, 3000, 11, None), (53, 20000, 600, 3000, 11, 1),
                                   (53, 20000, 600, 3000, 11, 2),
                                   (31, 3000, 3000, 60000, 14, None), (31, 20000, 40000, 200000, 17, 2)])
def test_ref_table_branches_and_marks(dense, tmp_path):
    """lasv_ref_hash_table_write_branches on the caller's reference-layout table, raw or after the GPU cleaning:
    the file, and the `visited` marks it leaves in the table, against the reference's sequential walk."""
    k, err_q20, n_reads, genome, L, thresh = dense
    W, cutoff = 40, 38
    base = synth.CFG1 if k <= 31 else synth.CFG0 if k <= 63 else synth.CFG2
    w = synth.scaled(base, genome, L)
    w.seed, w.kmer_size = 700 + k, k
    p = w.params()
    p.err_q20 = err_q20
    seq, qual, offs = synth.reads_host(p, 0, n_reads)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, cutoff)
        table, nxt = g.export_table()
    n_raw = int((table["status"] == 1).sum())
    ht = _capi.RefHashTable(k, 1 << L, W, table.ctypes.data, nxt.ctypes.data, None, n_raw, 15)
    if thresh is not None:
        _capi.check(_capi.lib().lasv_ref_hash_table_clean(C.byref(ht), _capi.CLEAN_TIPS | _capi.CLEAN_SUPERNODES, thresh,
                                                        None))
    before = table.copy()
    og, ost = _oracle_branches_in_table_order(k, L, W, _table_records(table), tmp_path / "oracle.fa")
    st = _capi.BranchStats()
    _capi.check(_capi.lib().lasv_ref_hash_table_write_branches(C.byref(ht), str(tmp_path / "gpu.fa").encode(), C.byref(st)))
    st = st.as_dict()
    assert (tmp_path / "gpu.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
    assert st["n_branches"] == ost["branches"]
    # (53, ..., T=2): the cleaning leaves a graph without a single branching node -- the reference prints an empty
    # file (the compiled reference agrees: tests/golden/branches.json holds such cases); everywhere else there are branches
    assert (st["n_branches"] > 0) == (dense != (53, 20000, 600, 3000, 11, 2))
    if err_q20 <= 3000:
        assert st["n_cut"] > 0
    # marks: exactly the reference's; nothing else in the table changed
    marked = {tuple(r) for r in table["kmer"][table["status"] == 2].tolist()}
    orecs = og.records()
    assert marked == {tuple(r) for r in orecs["kmer"][og.status() == 2].tolist()}
    after = table.copy()
    after["status"][after["status"] == 2] = 1
    assert np.array_equal(after, before)


@pytest.mark.parametrize("name,clean", [("synth_cfg1_k31", False), ("synth_cfg0_k53", False), ("synth_cfg0_k53", True),
                                        ("synth_cfg2_k95", True)])
def test_reference_cli_build_mode_through_the_gpu(name, clean, tmp_path):
    """run_laSV.sh:178 -- `dB_graph --mode build ... --dump_binary x.ctx` -- with loaders, cleaning and branch
    printing all bound to the GPU (integration/_build/dB_graph_gpubuild_*): ./branches.fa is what the reference's
    walk prints for the table the same run dumped."""
    from pathlib import Path
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    exe = Path(__file__).resolve().parents[1] / "integration" / "_build" / f"dB_graph_gpubuild_{O.MAXK_OF_N[O.words_per_kmer(k)]}"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    seq, qual, _ = util.case_streams(name)
    rl = meta["workload"]["read_len"]
    s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    (tmp_path / "l").write_text(f"{tmp_path / 'a.fq'}\n")
    (tmp_path / "r").write_text(f"{tmp_path / 'b.fq'}\n")
    cmd = [str(exe), "--kmer_size", str(k), "--mode", "build", "--mem_height", str(L), "--mem_width", str(W),
           "--pe_list", f"{tmp_path / 'l'},{tmp_path / 'r'}", "--quality_score_threshold", str(meta["q_thresh"]),
           "--dump_binary", str(tmp_path / "out.ctx")]
    if clean:
        cmd += ["--remove_low_coverage_supernodes", "2"]
    p = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:] + p.stdout[-2000:]
    assert "branches printed from" in p.stdout and "Branches printed" in p.stdout
    ctx = O.parse_ctx(tmp_path / "out.ctx")
    if not clean:
        util.assert_same_records(ctx["records"], util.case_records(name), name)
    _oracle_branches_in_table_order(k, L, W, ctx["records"], tmp_path / "oracle.fa")
    assert (tmp_path / "branches.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()


@pytest.mark.parametrize("name,thresh,finalized", [("synth_cfg1_k31", 2, False), ("synth_cfg0_k53", 2, False),
                                                   ("synth_cfg2_k95", 2, True), ("synth_cfg0_k53", 1, True)])
def test_gpu_build_mode_stays_on_the_device(name, thresh, finalized, tmp_path):
    """`--mode build --remove_low_coverage_supernodes T --dump_binary` (run_laSV.sh:178) on a graph handle, nothing
    but files leaving the device: lasv_graph_clean -> lasv_graph_write_branches -> lasv_graph_write_ctx (pruned
    nodes skipped, cleaning flags in the header) -> supernodes; all against the reference's sequential passes
    (oracle restatement) over the same slot order."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    seq, qual, offs = util.case_streams(name)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        raw = g.records()
        if finalized:
            g.finalize()
        og = O.OracleGraph(k, L, W)
        assert og.load_records(raw)
        ost = og.clean(thresh)
        st = g.clean(thresh)
        assert st["tips_clipped"] == ost["tips"] and st["supernodes_pruned"] == ost["supernodes_pruned"]
        assert st["nodes_pruned"] == len(raw) - ost["nodes_left"] > 0
        recs = g.records()
        assert np.array_equal(recs, og.records())              # same nodes, coverage, edges, same order
        assert g.summary()["n_nodes"] == len(recs)
        g.dump_binary(tmp_path / "gpu.ctx", 100, 1000)
        ctx = O.parse_ctx(tmp_path / "gpu.ctx")
        assert np.array_equal(ctx["records"], recs)
        assert bytes(ctx["header"]) == O.ctx_header(k, 100, 1000, cleaned_threshold=thresh)
        g.output_branches(tmp_path / "gpu_br.fa")
        og.print_branches(tmp_path / "oracle_br.fa")
        assert (tmp_path / "gpu_br.fa").read_bytes() == (tmp_path / "oracle_br.fa").read_bytes()
        g.output_supernodes(tmp_path / "gpu.fa")
        og.print_supernodes(tmp_path / "oracle.fa", max_length=1 << 20)
        assert (tmp_path / "gpu.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
        table, _ = g.export_table()                              # pruned nodes keep their place and status
        assert int((table["status"] == 3).sum()) == st["nodes_pruned"]
        assert not table["edges"][table["status"] == 3].any()
        assert np.array_equal(_table_records(table), recs)


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "synth_full_k53"])
def test_gpu_health_check(name):
    """`--health_check` on the device graph: every edge has its return arrow, raw, finalised and cleaned."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    seq, qual, offs = util.case_streams(name)
    with DBGraph(k, L, W) as g:
        g.load_reads_host(seq, qual, offs, util.cutoff_of(meta))
        recs = g.records()
        want = {"nodes_checked": len(recs), "edges_checked": int(np.unpackbits(recs["edges"]).sum()),
                "missing_target": 0, "missing_return_arrow": 0, "zero_kmers": 0}
        assert g.health_check() == want
        if name != "synth_full_k53":
            g.clean(2)
            h = g.health_check()
            assert h["missing_target"] == h["missing_return_arrow"] == 0 and h["nodes_checked"] == len(recs)
            assert h["edges_checked"] == int(np.unpackbits(g.records()["edges"]).sum()) < want["edges_checked"]
        g.finalize()
        h = g.health_check()
        assert h["missing_target"] == h["missing_return_arrow"] == 0


def test_gpu_health_check_at_full_config0_size():
    """Size-independent invariant at BASELINE configs[0] scale (63 Mbp, 30x, k=53, L=22 W=100): the graph of the
    full build has a return arrow for every edge, and edges_checked equals the summary's edge-bit count."""
    import torch
    w = synth.WORKLOADS["cfg0"]
    p = w.params()
    n_reads, batch, stride = 2 * w.n_pairs, 1 << 20, w.read_len + 1
    d_seq = torch.empty(batch * stride + 16, dtype=torch.uint8, device="cuda")
    d_qual = torch.empty_like(d_seq)
    st_ = torch.cuda.current_stream().cuda_stream
    with DBGraph(w.kmer_size, w.number_bits, w.bucket_size) as g:
        for b0 in range(0, n_reads, batch):
            nr = min(batch, n_reads - b0)
            synth.reads_device(p, b0, nr, d_seq, d_qual, st_)
            g.load_reads_device(d_seq, d_qual, nr * stride, None, 0, w.cutoff, st_)
        torch.cuda.synchronize()
        s = g.summary()
        h = g.health_check()
        assert h["nodes_checked"] == s["n_nodes"] > 100_000_000
        assert h["edges_checked"] == s["sum_edge_bits"]
        assert h["missing_target"] == h["missing_return_arrow"] == 0


@pytest.mark.parametrize("bulk", [0, 1])
@pytest.mark.parametrize("name,bpg", [("synth_cfg0_k53", 4), ("synth_cfg1_k31", 2), ("synth_cfg2_k95", 4),
                                      ("synth_full_k53", 4), ("synth_full_k53", 1)])
def test_gpu_grouped_insert_prototype(name, bpg, bulk, monkeypatch):
    """The streamed-insert prototype (groups.cuh): records sorted by group, every group updated in
    shared memory and written back, full buckets spill to the ordinary path -- same graph as the reference, in one
    batch and in two (the second meets the keys of the first)."""
    import torch
    monkeypatch.setenv("LASV_B200_GROUPS_BULK", str(bulk))      # 1: cp.async.bulk staging, double buffered
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = lasv_b200.words_per_kmer(k)
    seq, qual, _ = util.case_streams(name)
    d_seq, d_qual = torch.from_numpy(seq).cuda(), torch.from_numpy(qual).cuda()
    st_ = torch.cuda.current_stream().cuda_stream
    n_groups = (1 << L) // bpg

    def grouped(g, recs):
        n = recs.shape[0]
        if bpg == 4:                                # the library's own counting sort by group
            srt = torch.empty_like(recs)
            offsets = torch.empty(n_groups + 1, dtype=torch.int64, device="cuda")
            g.group_records_device(recs, n, bpg, srt, offsets, st_)
            torch.cuda.synchronize()
            assert int(offsets[0].item()) == 0 and int(offsets[-1].item()) == n
            assert bool((offsets[1:] >= offsets[:-1]).all())
            return g.insert_grouped_records_device(srt, n, offsets, bpg, st_)
        buckets = torch.empty(n, dtype=torch.int32, device="cuda")
        _capi.check(_capi.lib().lasv_graph_record_buckets_device(g._h, recs.data_ptr(), n, buckets.data_ptr(), st_))
        groups = (buckets.to(torch.int64) & 0xFFFFFFFF) // bpg
        order = torch.argsort(groups, stable=True)
        offsets = torch.searchsorted(groups[order].contiguous(),
                                     torch.arange(n_groups + 1, device="cuda", dtype=torch.int64)).to(torch.int64)
        torch.cuda.synchronize()
        return g.insert_grouped_records_device(recs[order].contiguous(), n, offsets.contiguous(), bpg, st_)

    for pieces in (1, 2):
        with DBGraph(k, L, W) as g:
            out = torch.zeros((meta["inserts"] + 8, 2 * N + 1), dtype=torch.int32, device="cuda")
            cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
            g.extract_records_device(d_seq, d_qual, len(seq), util.cutoff_of(meta), out, out.shape[0], cnt, st_)
            torch.cuda.synchronize()
            n = int(cnt.item())
            assert n == meta["inserts"]
            spilled = 0
            for part in torch.chunk(out[:n], pieces):
                spilled += grouped(g, part.contiguous())
            torch.cuda.synchronize()
            util.assert_same_records(g.records(), util.case_records(name), name)
            assert g.unique_kmers == meta["n_records"]
            assert (spilled > 0) == (name == "synth_full_k53")


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "edge_pe_k21_q5"])
def test_gpu_ranked_path_enumeration(name, tmp_path, monkeypatch):
    """LASV_B200_SN_RANKING=1: the path records come from list ranking (ranking.cuh) instead of one walk per start;
    supernodes, cleaning and branches built on them must be what the verified walks give."""
    monkeypatch.setenv("LASV_B200_SN_RANKING", "1")
    test_gpu_supernodes_match_reference(name, tmp_path)
    if name.startswith("synth"):
        (tmp_path / "c").mkdir()
        test_ref_table_clean_matches_reference_walk(name, 2, tmp_path / "c")
        (tmp_path / "b").mkdir()
        test_gpu_branches_match_reference_walk(name, tmp_path / "b")


def test_gpu_ranked_cycles_long_paths_and_dense_errors(tmp_path, monkeypatch):
    monkeypatch.setenv("LASV_B200_SN_RANKING", "1")
    test_gpu_supernodes_cycles_and_long_paths(tmp_path)
    (tmp_path / "d").mkdir()
    test_ref_table_clean_dense_errors((31, 20000, 40000, 200000, 17), tmp_path / "d")
