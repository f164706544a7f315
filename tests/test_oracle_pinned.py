"""Pins the oracle (oracle/dbg_oracle.c) to the reference: SURVEY 8c known-answer
vectors, golden `.ctx` outputs of the compiled unmodified reference
(tests/golden, generator make_golden.py) and, where oracle/_ref is present,
live runs of the reference binary."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import util

S = ("ACGTTGCATGCCGATAGGCTAACGTTAGCCATGCATCGGATCGATTACGGCTAGCTAGGATCCGATCGTAGCTAGCTAGGCTAACGATCGA"
     "TCGGCTAGCTAAGCTT")

# SURVEY.md section 8c, extracted from the compiled reference
KAT = {
    31: dict(fw=[0x06f9396329c1bc94], rc=[0x3a706f25cda4e41b], h=[(43, 451), (106539, 62915), (16228395, 10679747)]),
    53: dict(fw=[0x0000006f9396329c, 0x1bc94e4da363c69c], rc=[0x00000325b0d8d639, 0x3a706f25cda4e41b],
             h=[(366, 991), (10606, 41951), (16132462, 6071263)]),
    95: dict(fw=[0x06f9396329c1bc94, 0xe4da363c69c9ca35, 0x8db272729c18d8da],
             rc=[0x163636f25c9c9c63, 0x68d72725b0d8d639, 0x3a706f25cda4e41b],
             h=[(787, 538), (127763, 21018), (4846355, 16405018)]),
}


@pytest.mark.parametrize("k", [31, 53, 95])
def test_known_answers(k):
    v = KAT[k]
    fw = O.seq_to_kmer(S[:k], k)
    assert [int(x) for x in fw] == v["fw"]
    rc = O.revcomp(fw, k)
    assert [int(x) for x in rc] == v["rc"]
    key = O.get_key(fw, k)
    assert np.array_equal(key, fw)                      # "= fwd" in the table
    assert np.array_equal(O.get_key(rc, k), fw)         # and from the other strand
    assert O.kmer_to_seq(fw, k) == S[:k]
    for L, (h0, h1) in zip((10, 17, 24), v["h"]):
        assert O.hash_value(key, 0, L) == h0
        assert O.hash_value(key, 1, L) == h1


@pytest.mark.parametrize("name", util.ALL_CASES)
def test_oracle_matches_reference_golden(name):
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    paired = name.startswith("edge_pe") or name.startswith("synth")
    g, ok = util.oracle_build(meta["k"], meta["L"], meta["W"], seq, qual, offs, util.cutoff_of(meta),
                              paired=paired)
    assert ok
    util.assert_same_records(g.records(), util.case_records(name), name)
    assert g.stats.kmers_inserted == meta["inserts"]
    assert g.stats.bad_reads == meta["bad_reads"]
    assert g.stats.bases_read == meta["bases_parsed"]
    assert g.stats.bases_loaded == meta["bases_loaded"]
    # "tries r: n" (hash_table.c:384-394) -- the oracle keeps the reference's slot layout
    coll = g.collisions()
    for r, n in meta["tries"].items():
        assert int(coll[int(r)]) == n
    # mean contig length that lands in the .ctx header (graph_operations.c:769-774)
    tot = int(g.hist.sum())
    mean = int((g.hist.astype(object) * np.arange(len(g.hist), dtype=object)).sum() // tot) if tot else 0
    assert mean == meta["ctx_mean_read_len"] == meta["mean_read_len_after_filters"]
    # header bytes: all but the 6 padding bytes of the x87 long double
    hdr = bytearray(O.ctx_header(meta["k"], mean, meta["bases_loaded"]))
    ref = bytearray(bytes.fromhex(meta["ctx_header_hex"]))
    assert len(hdr) == len(ref) == 94
    pad = slice(6 + 16 + 12 + 4 + 9 + 10, 6 + 16 + 12 + 4 + 9 + 16)
    hdr[pad] = ref[pad] = b"\0" * 6
    assert hdr == ref


def test_oracle_table_full_is_reported():
    meta = util.case_meta("synth_cfg0_k53")
    seq, qual, offs = util.case_streams("synth_cfg0_k53")
    g, ok = util.oracle_build(53, 4, 20, seq, qual, offs, util.cutoff_of(meta))   # 320 slots << 7208 keys
    assert not ok and g.table_full


def test_oracle_reload_records_roundtrip():
    name = "synth_cfg0_k53"
    recs = util.case_records(name)
    g = O.OracleGraph(53, 10, 40)
    assert g.load_records(recs)
    util.assert_same_records(g.records(), recs)
    key = recs["kmer"][123]
    assert g.lookup(key) == (int(recs["covg"][123]), int(recs["edges"][123]))


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,L,W", [(31, 9, 30), (53, 9, 30), (95, 9, 30)])
def test_oracle_matches_live_reference(tmp_path, k, L, W):
    from lasv_b200 import synth
    base = {31: synth.CFG1, 53: synth.CFG0, 95: synth.CFG2}[k]
    w = synth.scaled(base, 3000, L)
    w.seed = 7 + k
    p = w.params()
    n_pairs = 150
    seq, qual, offs = synth.reads_host(p, 0, 2 * n_pairs)
    stride = w.read_len + 1
    s2, q2 = seq.reshape(-1, stride), qual.reshape(-1, stride)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len)
    ref = O.run_reference(k, L, W, tmp_path / "a.fq", tmp_path / "b.fq", q_thresh=5, workdir=tmp_path)
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, 38, paired=True)
    assert ok
    util.assert_same_records(g.records(), ref["ctx"]["records"])
    assert g.stats.kmers_inserted == ref["inserts"]
    # the oracle keeps the reference's slot order too: .ctx record order is identical
    assert np.array_equal(g.records(), ref["ctx"]["records"])


# ---- supernodes (config 3): the oracle's restatement of the reference's sequential walk ----
SUPERNODE_CASES = [n for n in util.ALL_CASES if "supernodes" in util.case_meta(n)]


@pytest.mark.parametrize("name", SUPERNODE_CASES)
def test_oracle_supernodes_match_reference_golden(name, tmp_path):
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    g, ok = util.oracle_build(meta["k"], meta["L"], meta["W"], seq, qual, offs, util.cutoff_of(meta), paired=True)
    assert ok
    st = g.print_supernodes(tmp_path / "sup.fa")
    seqs = O.read_fasta_seqs(tmp_path / "sup.fa")
    assert st["supernodes"] == len(seqs)
    assert O.canonical_seq_set(seqs) == meta["supernodes"]
    # the walk leaves the graph as it found it (graph_operations.c:1290)
    st2 = g.print_supernodes(tmp_path / "sup2.fa")
    assert st2 == st and (tmp_path / "sup2.fa").read_bytes() == (tmp_path / "sup.fa").read_bytes()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,L,W,err", [(31, 9, 30, None), (53, 9, 30, None), (95, 9, 30, None), (21, 10, 30, 30000)])
def test_oracle_supernodes_match_live_reference_bytes(tmp_path, k, L, W, err):
    """Same table order (the oracle keeps the reference's slot layout), so the FASTA the
    reference prints -- names, order, strand -- must come out byte for byte; the k=21 case
    has 3 % errors: junction next to junction, where what gets printed depends on the order."""
    from lasv_b200 import synth
    base = {21: synth.CFG1, 31: synth.CFG1, 53: synth.CFG0, 95: synth.CFG2}[k]
    w = synth.scaled(base, 3000, L)
    w.seed = 11 + k
    w.kmer_size = k
    p = w.params()
    if err is not None:
        p.err_q20 = err
    n_pairs = 150
    seq, qual, offs = synth.reads_host(p, 0, 2 * n_pairs)
    stride = w.read_len + 1
    s2, q2 = seq.reshape(-1, stride), qual.reshape(-1, stride)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len)
    O.run_reference(k, L, W, tmp_path / "a.fq", tmp_path / "b.fq", q_thresh=5, workdir=tmp_path,
                    want_supernodes=True)
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, 38, paired=True)
    assert ok
    g.print_supernodes(tmp_path / "ours.fa")
    assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "sup.fa").read_bytes()


# ---- cleaning (`--remove_low_coverage_supernodes T`, always on in laSV: run_laSV.sh) ----
@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", 2), ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 2),
                                         ("synth_cfg0_k53", 1), ("synth_cfg1_k31", 5), ("edge_pe_k21_q5", 1)])
def test_oracle_cleaning_matches_live_reference(name, thresh, tmp_path):
    """Tip clipping + low-coverage supernode removal restated from the reference (they mutate the
    graph as they traverse it): same slot order in, same `.ctx` out -- records in the same order,
    header with the cleaning flags -- and the same supernodes printed afterwards, byte for byte."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    f1, f2 = (tmp_path / "a.fq", tmp_path / "b.fq")
    seq, qual, offs = util.case_streams(name)
    if "synth" in meta:
        rl = meta["workload"]["read_len"]
        s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
        O.write_fastq_fixed(f1, s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
        O.write_fastq_fixed(f2, s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    else:
        f1, f2 = util.case_files(name)
    ref = O.run_reference(k, L, W, f1, f2, q_thresh=meta["q_thresh"], workdir=tmp_path, want_supernodes=True,
                          extra_args=["--remove_low_coverage_supernodes", str(thresh)])
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, util.cutoff_of(meta), paired=True)
    assert ok
    st = g.clean(thresh)
    recs = g.records()
    assert st["nodes_left"] == len(recs) == len(ref["ctx"]["records"])
    assert np.array_equal(recs, ref["ctx"]["records"])          # same nodes, coverage, edges, in the same order
    if name.startswith("synth"):
        assert st["tips"] > 0 and st["supernodes_pruned"] > 0 and len(recs) < meta["n_records"]
    hdr = bytearray(O.ctx_header(k, meta["ctx_mean_read_len"], meta["ctx_total_seq"], cleaned_threshold=thresh))
    got = bytearray(ref["ctx"]["header"])
    pad = slice(6 + 16 + 12 + 4 + 9 + 10, 6 + 16 + 12 + 4 + 9 + 16)
    hdr[pad] = got[pad] = b"\0" * 6
    assert hdr == got
    g.print_supernodes(tmp_path / "oracle.fa")
    assert (tmp_path / "oracle.fa").read_bytes() == (tmp_path / "sup.fa").read_bytes()


# ---- branch printing (`--mode build` writes ./branches.fa after the cleaning: graph_operations.c:1145-1152) ----
@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", None), ("synth_cfg0_k53", None), ("synth_cfg2_k95", None),
                                         ("synth_cfg1_k31", 2), ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 2),
                                         ("edge_pe_k21_q5", None), ("edge_pe_k21_q5", 1)])
def test_oracle_branches_match_live_reference(name, thresh, tmp_path):
    """print_branches.c restated: same slot order in, the same branches.fa out byte for byte, raw and
    after the cleaning; the `visited` marks it leaves behind are pinned through the supernodes the same
    run prints afterwards (the printer skips marked nodes)."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    f1, f2 = (tmp_path / "a.fq", tmp_path / "b.fq")
    seq, qual, offs = util.case_streams(name)
    if "synth" in meta:
        rl = meta["workload"]["read_len"]
        s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
        O.write_fastq_fixed(f1, s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
        O.write_fastq_fixed(f2, s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
    else:
        f1, f2 = util.case_files(name)
    extra = ["--mode", "build"] + (["--remove_low_coverage_supernodes", str(thresh)] if thresh is not None else [])
    ref = O.run_reference(k, L, W, f1, f2, q_thresh=meta["q_thresh"], workdir=tmp_path, want_supernodes=True,
                          extra_args=extra)
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, util.cutoff_of(meta), paired=True)
    assert ok
    if thresh is not None:
        g.clean(thresh)
    st = g.print_branches(tmp_path / "oracle_branches.fa", keep_marks=True)
    want = (tmp_path / "branches.fa").read_bytes()
    assert (tmp_path / "oracle_branches.fa").read_bytes() == want
    assert st["branches"] == want.count(b">b_")
    if name.startswith("synth") and thresh is None:     # the cleaning may leave no junction at all
        assert st["branches"] > 0 and st["visited"] >= st["branches"]
    assert np.array_equal(g.records(), ref["ctx"]["records"])     # marks do not change what is dumped
    g.print_supernodes(tmp_path / "oracle.fa")
    assert (tmp_path / "oracle.fa").read_bytes() == (tmp_path / "sup.fa").read_bytes()


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,L,W,err,genome,pairs", [(21, 10, 30, 30000, 3000, 150), (31, 12, 40, 20000, 20000, 1500),
                                                    (53, 11, 40, 600, 1500, 400), (95, 11, 40, 20000, 3000, 300)])
def test_oracle_branches_dense_errors_live_reference(tmp_path, k, L, W, err, genome, pairs):
    """Junction next to junction and long error-free stretches (walks cut at 300 nodes, walks that
    run into marks left from the other end)."""
    from lasv_b200 import synth
    base = {21: synth.CFG1, 31: synth.CFG1, 53: synth.CFG0, 95: synth.CFG2}[k]
    w = synth.scaled(base, genome, L)
    w.seed, w.kmer_size = 500 + k, k
    p = w.params()
    p.err_q20 = err
    seq, qual, offs = synth.reads_host(p, 0, 2 * pairs)
    stride = w.read_len + 1
    s2, q2 = seq.reshape(-1, stride), qual.reshape(-1, stride)
    O.write_fastq_fixed(tmp_path / "a.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len)
    O.write_fastq_fixed(tmp_path / "b.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len)
    O.run_reference(k, L, W, tmp_path / "a.fq", tmp_path / "b.fq", q_thresh=5, workdir=tmp_path,
                    extra_args=["--mode", "build"])
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, 38, paired=True)
    assert ok
    st = g.print_branches(tmp_path / "ours.fa")
    assert st["branches"] > 0
    assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "branches.fa").read_bytes()


def _golden_branches():
    import json
    return json.loads((util.GOLDEN / "branches.json").read_text())


@pytest.mark.parametrize("key", sorted(_golden_branches()))
def test_oracle_cleaning_and_branches_match_committed_reference_digests(key, tmp_path):
    """The same pinning without oracle/_ref: tests/golden/branches.json holds, from the compiled reference
    (generator: tests/golden/make_golden_branches.py), the SHA-256 of ./branches.fa and of the `.ctx` records a
    `--mode build [--remove_low_coverage_supernodes T]` run leaves for every golden case."""
    import hashlib
    gold = _golden_branches()[key]
    name, what = key.split(":")
    meta = util.case_meta(name)
    seq, qual, offs = util.case_streams(name)
    g, ok = util.oracle_build(meta["k"], meta["L"], meta["W"], seq, qual, offs, util.cutoff_of(meta), paired=True)
    assert ok
    if what != "raw":
        g.clean(int(what))
    st = g.print_branches(tmp_path / "br.fa")
    fa = (tmp_path / "br.fa").read_bytes()
    assert (len(fa), st["branches"]) == (gold["branches_bytes"], gold["branches"])
    assert hashlib.sha256(fa).hexdigest() == gold["branches_sha256"]
    recs = g.records()
    assert len(recs) == gold["ctx_records"]
    assert hashlib.sha256(recs.tobytes()).hexdigest() == gold["ctx_records_sha256"]
