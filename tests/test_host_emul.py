"""CPU check of the kernels' own arithmetic (tests/host_emul): the __host__
__device__ functions the CUDA kernels are made of, driven tile by tile on the
host, against the golden reference outputs and the oracle.  No GPU needed; the
-m gpu tests repeat the same comparisons through the real kernels."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from lasv_b200 import element_dtype, record_dtype, words_per_kmer
from tests import util

HERE = Path(__file__).resolve().parent / "host_emul"


@pytest.fixture(scope="module")
def emul():
    so = HERE / "libhost_emul.so"
    src = HERE / "host_emul.cpp"
    hdrs = list((HERE.parents[1] / "lasv_b200" / "csrc").glob("*.cuh")) + list((HERE.parents[1] / "lasv_b200" / "csrc").glob("*_host.h"))
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-I/usr/local/cuda/include", "-o", str(so), str(src)], check=True)
    L = C.CDLL(str(so))
    L.emul_table_bytes.restype = C.c_uint64
    L.emul_table_bytes.argtypes = [C.c_int] * 3
    L.emul_build_fused.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_int] * 6 + [C.c_void_p] * 2
    L.emul_extract_records.restype = C.c_uint64
    L.emul_extract_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_uint64]
    L.emul_insert_records.argtypes = [C.c_void_p, C.c_uint64] + [C.c_int] * 4 + [C.c_void_p] * 2
    L.emul_finalize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.emul_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.emul_find.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_supernodes.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_char_p, C.c_void_p]
    L.emul_revcomp.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.emul_clean.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
    L.emul_insert_grouped.argtypes = [C.c_void_p, C.c_uint64] + [C.c_int] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_set_parallel_min.argtypes = [C.c_uint64]
    L.emul_set_ranking.argtypes = [C.c_int]
    L.emul_health.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
    L.emul_branches.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_char_p, C.c_int, C.c_void_p]
    L.emul_hash_c.restype = C.c_uint32
    L.emul_hash_c.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    return L


class EmulTable:
    """The device table as the kernels see it: 2^L buckets x ceil(W/G) lines of 128
    bytes (kmer_core.cuh "line layout"), in host memory."""

    def __init__(self, emul, k, L, W):
        self.emul, self.k, self.L, self.W, self.N = emul, k, L, W, words_per_kmer(k)
        self.raw = np.zeros(int(emul.emul_table_bytes(k, L, W)), dtype=np.uint8)
        self.finalized = False

    @property
    def ptr(self):
        return self.raw.ctypes.data

    def elements(self):
        """Slot s of bucket b -> out[b*W+s], as reference Elements (holes included)."""
        out = np.zeros((1 << self.L) * self.W, dtype=element_dtype(self.N))
        self.emul.emul_gather(self.ptr, self.k, self.L, self.W, int(self.finalized), out.ctypes.data)
        return out

    def finalize(self):
        nxt = np.zeros(1 << self.L, dtype=np.int16)
        self.emul.emul_finalize(self.ptr, self.k, self.L, self.W, nxt.ctypes.data)
        self.finalized = True
        return nxt

    def find(self, keys):
        n = len(keys)
        covg, edges, found = np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.emul.emul_find(self.ptr, self.k, self.L, self.W, 15, int(self.finalized), keys.ctypes.data, n,
                            covg.ctypes.data, edges.ctypes.data, found.ctypes.data)
        return found.astype(bool), covg, edges

    def records(self):
        return table_records(self.elements())


def table_records(table):
    occ = table[(table["status"] == 1) | (table["status"] == 2)]       # pruned nodes are not part of the graph
    out = np.zeros(len(occ), dtype=record_dtype(table["kmer"].shape[1]))
    out["kmer"], out["covg"], out["edges"] = occ["kmer"], occ["coverage"], occ["edges"]
    return out


def emul_build(emul, meta, seq, qual, T=896, L=None, W=None):
    k = meta["k"]
    L = L or meta["L"]
    W = W or meta["W"]
    tab = EmulTable(emul, k, L, W)
    ctr = np.zeros(20, dtype=np.uint64)
    rc = emul.emul_build_fused(seq.ctypes.data, qual.ctypes.data if qual is not None else None, len(seq), k,
                               util.cutoff_of(meta), L, W, 15, T, tab.ptr, ctr.ctypes.data)
    return rc, tab, ctr


@pytest.mark.parametrize("name", util.ALL_CASES)
@pytest.mark.parametrize("T", [896, 384, 2048])
def test_emulated_build_matches_reference_golden(emul, name, T):
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, ctr = emul_build(emul, meta, seq, qual, T)
    assert rc == 0
    table = tab.elements()
    util.assert_same_records(table_records(table), util.case_records(name), name)
    assert int(ctr[1]) == meta["inserts"]
    assert int(ctr[0]) == meta["n_records"]
    # pad bytes carry tags until finalisation, status is `none`
    occ = table[table["status"] != 0]
    assert (occ["status"] == 1).all()


@pytest.mark.parametrize("name", ["synth_cfg0_k53", "synth_cfg1_k31", "synth_cfg2_k95", "synth_full_k53",
                                  "synth_full_k31", "edge_pe_k33_q5"])
def test_emulated_finalize_gives_reference_layout(emul, name):
    """After finalisation every bucket is hole-free and every key is found by the
    reference's scan-from-slot-0 lookup along the reference's rehash chain."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = words_per_kmer(k)
    seq, qual, _ = util.case_streams(name)
    rc, tab, ctr = emul_build(emul, meta, seq, qual)
    assert rc == 0
    gold = util.case_records(name)
    keys = np.ascontiguousarray(gold["kmer"])
    n = len(keys)
    found, covg, edges = tab.find(keys)
    assert found.all() and np.array_equal(covg, gold["covg"]) and np.array_equal(edges, gold["edges"])
    # the header tag bytes mirror the occupancy exactly (0 <=> empty slot), slots fill each
    # line in order, and bytes past the line's slots stay zero
    G = (128 - 8) // (8 * N + 8)
    NL = -(-W // G)
    lines = tab.raw.reshape(1 << L, NL, 128)
    hdr = lines[:, :, :8]
    assert not hdr[:, :, G:].any()
    occ_dev = (tab.elements()["status"] != 0).reshape(1 << L, W)
    tag_occ = (hdr[:, :, :G] != 0).reshape(1 << L, NL * G)
    assert np.array_equal(tag_occ[:, :W], occ_dev) and not tag_occ[:, W:].any()
    absent0 = keys.copy()
    absent0[:, -1] ^= np.uint64(0x5555)
    found, _, _ = tab.find(absent0)
    present0 = {tuple(x) for x in keys.tolist()}
    assert [bool(f) for f in found] == [tuple(x) in present0 for x in absent0.tolist()]
    nxt = tab.finalize()
    table = tab.elements()
    # every bucket: W consecutive Elements at the start of its lines, zero slack behind
    pitched = tab.raw.reshape(1 << L, NL * 128)
    assert np.array_equal(pitched[:, : W * (8 * N + 8)].reshape(-1), table.view(np.uint8).reshape(-1))
    assert not pitched[:, W * (8 * N + 8):].any()
    raw = table.view(np.uint8).reshape(len(table), 8 * N + 8)
    assert not raw[:, 8 * N + 6:].any()                       # pad bytes zeroed
    t2 = table.reshape(1 << L, W)
    occ = t2["status"] != 0
    assert np.array_equal(occ.sum(axis=1), nxt)
    assert (occ == (np.arange(W)[None, :] < nxt[:, None])).all()    # hole-free from slot 0
    assert not raw[~occ.reshape(-1)].any()                    # empty slots are all-zero bytes
    found, covg, edges = tab.find(keys)
    assert found.all() and np.array_equal(covg, gold["covg"]) and np.array_equal(edges, gold["edges"])
    # independent check against the reference's placement rule (bucket scan + rehash chain)
    for b in range(1 << L):                                   # bucket b holds exactly the keys the
        for s in range(int(nxt[b])):                          # reference chain would look for there
            h = [O.hash_value(t2["kmer"][b, s], r, L) for r in range(16)]
            assert b in h
            r = h.index(b)
            assert all(nxt[hb] == W for hb in h[:r])          # earlier buckets of the chain are full
    util.assert_same_records(table_records(table), gold, name)
    absent = keys.copy()
    absent[:, -1] ^= np.uint64(0x5555)
    found, _, _ = tab.find(absent)
    present = {tuple(x) for x in keys.tolist()}
    assert [bool(f) for f in found] == [tuple(x) in present for x in absent.tolist()]


@pytest.mark.parametrize("name", ["synth_cfg0_k53", "synth_cfg1_k31", "synth_cfg2_k95", "edge_pe_k65_q5"])
def test_emulated_records_path(emul, name):
    """extract -> records -> insert (the split / sharded path) gives the same graph."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = words_per_kmer(k)
    seq, qual, _ = util.case_streams(name)
    cap = meta["inserts"] + 10
    recs = np.zeros((cap, 2 * N + 1), dtype=np.uint32)
    n = emul.emul_extract_records(seq.ctypes.data, qual.ctypes.data, len(seq), k, util.cutoff_of(meta), 384,
                                  recs.ctypes.data, cap)
    assert n == meta["inserts"]
    # shard the records by owner bits exactly as the ROUTE mode does and insert per shard
    g_bits = 1
    shards = [EmulTable(emul, k, L - g_bits, W) for _ in range(2)]
    owner = np.zeros(n, dtype=np.int64)
    for i in range(n):
        key = np.zeros(N, dtype=np.uint64)
        for w in range(N):
            key[w] = np.uint64(recs[i, 2 * w]) | (np.uint64(recs[i, 2 * w + 1]) << np.uint64(32))
        c = emul.emul_hash_c(key.ctypes.data, N, 0)
        assert (c & ((1 << L) - 1)) == O.hash_value(key, 0, L)     # lookup3 == reference hash_value
        owner[i] = (c >> (L - g_bits)) & 1
    all_recs = []
    for d in range(2):
        sub = np.ascontiguousarray(recs[:n][owner == d])
        ctr = np.zeros(20, dtype=np.uint64)
        rc = emul.emul_insert_records(sub.ctypes.data, len(sub), k, L - g_bits, W, 15,
                                      shards[d].ptr, ctr.ctypes.data)
        assert rc == 0
        all_recs.append(shards[d].records())
    util.assert_same_records(np.concatenate(all_recs), util.case_records(name), name)


def test_emulated_table_full(emul):
    meta = util.case_meta("synth_cfg0_k53")
    seq, qual, _ = util.case_streams("synth_cfg0_k53")
    rc, table, ctr = emul_build(emul, meta, seq, qual, L=4, W=20)
    assert rc == -3 and ctr[2] == 1


def test_emulated_ragged_and_tiny_inputs(emul):
    """Empty stream, one short read, reads of every length around k, no trailing
    newline at a 16-byte boundary: compare with the oracle."""
    rng = np.random.default_rng(5)
    for k in (21, 33, 65):
        reads, quals = [], []
        for ln in [0, 1, k - 1, k, k + 1, 2 * k, 15, 16, 17, 31, 32, 33, 200, 5000]:
            reads.append("".join(rng.choice(list("ACGT"), ln)))
            q = rng.choice([ord("I"), ord("#"), ord("&"), ord("'")], ln, p=[0.9, 0.04, 0.03, 0.03])
            quals.append(bytes(q.astype(np.uint8)).decode())
        seq, qual, offs = O.reads_to_streams(reads, quals)
        meta = {"k": k, "L": 8, "W": 60, "q_thresh": 5}
        g, ok = util.oracle_build(k, 8, 60, seq, qual, offs, 38)
        assert ok
        for T in (384, 896, 896 | (5 << 16), 384 | (15 << 16)):    # last two: misaligned stream pointer
            rc, table, ctr = emul_build(emul, meta, seq, qual, T)
            assert rc == 0
            util.assert_same_records(table.records(), g.records(), f"k={k}")
            assert int(ctr[1]) == g.stats.kmers_inserted
    rc, table, ctr = emul_build(emul, {"k": 31, "L": 4, "W": 4, "q_thresh": 0}, np.zeros(0, np.uint8),
                                np.zeros(0, np.uint8))
    assert rc == 0 and ctr[1] == 0


# --------------------------------------------------------------- property test
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

_BASES = st.sampled_from(list("ACGT") * 12 + list("acgt") * 2 + ["N", "n", "X", "-"])
_QUALS = st.sampled_from([chr(q) for q in (73, 73, 73, 73, 73, 60, 40, 39, 38, 37, 35, 34, 33)])


@st.composite
def _read(draw):
    n = draw(st.integers(min_value=0, max_value=260))
    if draw(st.integers(0, 9)) == 0:                       # a low-complexity read now and then
        unit = draw(st.sampled_from(["A", "AC", "ACG", "T", "GC"]))
        seq = (unit * (n // len(unit) + 1))[:n]
    else:
        seq = "".join(draw(st.lists(_BASES, min_size=n, max_size=n)))
    qual = "".join(draw(st.lists(_QUALS, min_size=n, max_size=n)))
    return seq, qual


@settings(max_examples=int(__import__("os").environ.get("LASV_FUZZ_EXAMPLES", "40")), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(reads=st.lists(_read(), min_size=0, max_size=40),
       k=st.sampled_from([3, 5, 21, 31, 33, 53, 63, 65, 95]),
       L=st.integers(min_value=2, max_value=6),
       W=st.sampled_from([1, 2, 3, 5, 7, 8, 13, 40, 100]),
       q_thresh=st.sampled_from([0, 5, 20]),
       T=st.sampled_from([384, 896, 896 | (7 << 16)]))
def test_emulated_build_equals_oracle_on_random_inputs(emul, reads, k, L, W, q_thresh, T):
    """Random reads (bad characters, low qualities, homopolymers, empty and short reads), random table
    geometry incl. bucket sizes that are not a multiple of the slots per line and tables that overflow:
    the kernels' arithmetic + the line-layout table == the oracle, or both report a full table."""
    seq, qual, offs = O.reads_to_streams([r[0] for r in reads], [r[1] for r in reads])
    g, ok = util.oracle_build(k, L, W, seq, qual, offs, q_thresh + 33)
    meta = {"k": k, "L": L, "W": W, "q_thresh": q_thresh}
    rc, table, ctr = emul_build(emul, meta, seq, qual, T)
    if not ok:
        assert rc == -3                                       # hash table full: fatal on both sides
        return
    assert rc == 0
    util.assert_same_records(table.records(), g.records(), f"k={k} L={L} W={W}")
    assert int(ctr[1]) == g.stats.kmers_inserted and int(ctr[0]) == g.unique_kmers
    nxt = table.finalize()
    fin = table.elements().reshape(1 << L, W)
    occ = fin["status"] != 0
    assert (occ == (np.arange(W)[None, :] < nxt[:, None])).all()
    if len(g.records()):
        found, covg, edges = table.find(np.ascontiguousarray(g.records()["kmer"]))
        assert found.all() and np.array_equal(covg, g.records()["covg"]) and np.array_equal(edges, g.records()["edges"])


# ---- supernodes: supernodes.cuh (walks) + supernodes_host.h (replay) on the CPU ----
SN_COUNTS = ["printed", "nodes", "interior", "starts", "paths", "cycles", "never_printed", "dangling"]


def emul_supernodes(emul, tab, path):
    counts = np.zeros(8, dtype=np.uint64)
    rc = emul.emul_supernodes(tab.ptr, tab.k, tab.L, tab.W, int(tab.finalized), str(path).encode(), counts.ctypes.data)
    assert rc == 0, f"emulated supernode extraction failed ({rc}): {dict(zip(SN_COUNTS, counts.tolist()))}"
    return dict(zip(SN_COUNTS, (int(c) for c in counts)))


def oracle_supernodes_in_order(tab, path):
    """The oracle's restatement of the reference's sequential walk over OUR slot order."""
    g = O.OracleGraph(tab.k, tab.L, tab.W)
    assert g.load_records(tab.records())
    return g.print_supernodes(path, max_length=1 << 20)


@pytest.mark.parametrize("k", [21, 31, 33, 53, 63, 65, 95])
def test_closed_form_revcomp_matches_oracle(emul, k):
    rng = np.random.default_rng(k)
    N = words_per_kmer(k)
    for _ in range(200):
        s = "".join(rng.choice(list("ACGT"), size=k))
        km = O.seq_to_kmer(s, k)
        out = np.zeros(N, dtype=np.uint64)
        emul.emul_revcomp(km.ctypes.data, k, out.ctypes.data)
        assert np.array_equal(out, O.revcomp(km, k)), s


@pytest.mark.parametrize("name", [n for n in util.ALL_CASES if "supernodes" in util.case_meta(n)])
@pytest.mark.parametrize("finalized", [False, True])
def test_emulated_supernodes_match_reference_golden(emul, name, finalized, tmp_path):
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    if finalized:
        tab.finalize()
    st = emul_supernodes(emul, tab, tmp_path / "ours.fa")
    seqs = O.read_fasta_seqs(tmp_path / "ours.fa")
    assert st["printed"] == len(seqs) and st["nodes"] == meta["n_records"]
    # the contig set the compiled reference printed for its own build of the same reads
    assert O.canonical_seq_set(seqs) == meta["supernodes"]
    # and, for OUR slot order, the reference's sequential walk (oracle) prints the same bytes
    ost = oracle_supernodes_in_order(tab, tmp_path / "oracle.fa")
    assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
    assert ost["supernodes"] == st["printed"] and ost["nodes"] == st["nodes"]


@pytest.mark.parametrize("k,err_q20,seed", [(21, 30000, 3), (21, 60000, 4), (31, 40000, 5), (53, 30000, 6),
                                             (95, 20000, 7), (15, 80000, 8)])
def test_emulated_supernodes_order_dependent_cases(emul, k, err_q20, seed, tmp_path):
    """Dense errors: junctions next to junctions, where whether a one-edge supernode is printed
    depends on which of its end nodes the traversal visited first.  Ours must equal the
    reference's sequential walk over the same slot order, byte for byte -- the oracle's
    restatement always, the compiled reference (reloading our .ctx) where it is available."""
    from lasv_b200 import synth
    base = synth.CFG1 if k <= 31 else synth.CFG0 if k <= 63 else synth.CFG2
    w = synth.scaled(base, 3000, 10)
    w.seed = seed
    w.kmer_size = k
    p = w.params()
    p.err_q20 = err_q20
    seq, qual, _ = synth.reads_host(p, 0, 400)
    meta = {"k": k, "L": 10, "W": 40, "q_thresh": 5}
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    st = emul_supernodes(emul, tab, tmp_path / "ours.fa")
    oracle_supernodes_in_order(tab, tmp_path / "oracle.fa")
    ours = (tmp_path / "ours.fa").read_bytes()
    assert ours == (tmp_path / "oracle.fa").read_bytes()
    assert st["never_printed"] > 0 or k == 95, "case does not exercise the order dependence"
    if O.ref_available(k):
        recs = tab.records()
        with open(tmp_path / "ours.ctx", "wb") as f:
            f.write(O.ctx_header(k, 100, 1000))
            f.write(recs.tobytes())
        O.run_reference(k, 10, 40, None, ctx_in=tmp_path / "ours.ctx", workdir=tmp_path, want_supernodes=True,
                        extra_args=["--max_var_len", str(1 << 20)])
        assert ours == (tmp_path / "sup.fa").read_bytes()


def test_emulated_supernodes_cycles_and_long_paths(emul, tmp_path):
    """A circular genome read without errors is one cycle of interior nodes; an error-free
    linear genome is one long path."""
    rng = np.random.default_rng(5)
    k = 31
    circ = "".join(rng.choice(list("ACGT"), size=500))
    lin = "".join(rng.choice(list("ACGT"), size=6000))
    reads = [(circ + circ)[i:i + 100] for i in range(0, 500, 7)] + [lin[i:i + 150] for i in range(0, 5851, 50)]
    seq, _, _ = O.reads_to_streams(reads)
    rc, tab, _ = emul_build(emul, {"k": k, "L": 10, "W": 40, "q_thresh": 0}, seq, None)
    assert rc == 0
    st = emul_supernodes(emul, tab, tmp_path / "ours.fa")
    oracle_supernodes_in_order(tab, tmp_path / "oracle.fa")
    assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()
    assert st["cycles"] == 1 and st["printed"] == 2 and st["paths"] == 1
    seqs = sorted(O.read_fasta_seqs(tmp_path / "ours.fa"), key=len)
    assert len(seqs[0]) == 500 + k and len(seqs[1]) == 6000


# ---- cleaning: cleaning.cuh (per-path work) + cleaning_host.h (the reference's decisions on the compressed graph) ----
def emul_clean(emul, tab, threshold):
    counts = np.zeros(4, dtype=np.uint64)
    rc = emul.emul_clean(tab.ptr, tab.k, tab.L, tab.W, int(tab.finalized), threshold, counts.ctypes.data)
    assert rc == 0, f"emulated cleaning failed ({rc})"
    return dict(zip(["tips", "tip_edges", "supernodes_pruned", "nodes_pruned"], (int(c) for c in counts)))


def oracle_clean_in_order(tab, threshold):
    g = O.OracleGraph(tab.k, tab.L, tab.W)
    assert g.load_records(tab.records())
    st = g.clean(threshold)
    return g, st


@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", 2), ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 2),
                                         ("synth_cfg0_k53", 1), ("synth_cfg1_k31", 5), ("edge_pe_k21_q5", 1),
                                         ("edge_pe_k31_q5", 3)])
@pytest.mark.parametrize("finalized", [False, True])
def test_emulated_cleaning_matches_reference_walk(emul, name, thresh, finalized, tmp_path):
    """`--remove_low_coverage_supernodes T` decided on the path records: the table it leaves must hold
    exactly the nodes, coverages and edges the reference's sequential cleaning (oracle restatement,
    pinned on the compiled reference) leaves for the same slot order."""
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    if finalized:
        tab.finalize()
    og, ost = oracle_clean_in_order(tab, thresh)          # before the emulation mutates the table
    st = emul_clean(emul, tab, thresh)
    assert np.array_equal(tab.records(), og.records())
    assert st["tips"] == ost["tips"] and st["tip_edges"] == ost["tip_nodes"]
    assert st["supernodes_pruned"] == ost["supernodes_pruned"]
    elems = tab.elements()
    assert int((elems["status"] == 3).sum()) == st["nodes_pruned"] and not elems["edges"][elems["status"] == 3].any()
    if name.startswith("synth"):
        assert st["tips"] > 0 and st["supernodes_pruned"] > 0
    # and the supernodes of the cleaned graph
    emul_supernodes(emul, tab, tmp_path / "ours.fa")
    og.print_supernodes(tmp_path / "oracle.fa", max_length=1 << 20)
    assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes()


@pytest.mark.parametrize("alphabet,seed", [("AC", 1), ("CG", 2), ("ACG", 3), ("ACGT", 4)])
def test_emulated_cleaning_and_supernodes_on_repetitive_graphs(emul, alphabet, seed, tmp_path):
    """Small random graphs over reduced alphabets: short cycles, tips on tips, and -- with C and G --
    paths that are their own reverse complement (the reference's walk turns round in them and runs
    back over its own nodes).  Cleaning and the supernodes printed afterwards must equal the
    reference's sequential result (oracle restatement) for the same slot order."""
    rng = np.random.default_rng(seed)
    comp = str.maketrans("ACGT", "TGCA")
    checked = cycles = 0
    for _ in range(120):
        k = int(rng.choice([9, 11, 15, 21]))
        glen = int(rng.integers(30, 400))
        genome = "".join(rng.choice(list(alphabet), size=glen))
        src = genome * 3 if rng.random() < 0.5 else genome
        reads = []
        for _ in range(int(rng.integers(1, 40))):
            s, rl = int(rng.integers(0, glen)), int(rng.integers(k, k + 60))
            r = src[s:s + rl]
            if len(r) < k:
                continue
            if rng.random() < 0.3:
                r = list(r)
                r[int(rng.integers(0, len(r)))] = str(rng.choice(list("ACGT")))
                r = "".join(r)
            reads.append(r.translate(comp)[::-1] if rng.random() < 0.5 else r)
        if not reads:
            continue
        seq, _, _ = O.reads_to_streams(reads)
        thresh = int(rng.integers(1, 4))
        rc, tab, _ = emul_build(emul, {"k": k, "L": 6, "W": 40, "q_thresh": 0}, seq, None)
        if rc != 0:
            continue
        og, _ = oracle_clean_in_order(tab, thresh)
        emul_clean(emul, tab, thresh)
        assert np.array_equal(tab.records(), og.records()), (alphabet, k, thresh, reads)
        h = emul_health(emul, tab)                        # the cleaning keeps every remaining arrow paired
        assert h["missing_target"] == h["missing_return_arrow"] == 0, (alphabet, k, thresh, reads)
        st = emul_supernodes(emul, tab, tmp_path / "ours.fa")
        og.print_supernodes(tmp_path / "oracle.fa", max_length=1 << 20)
        assert (tmp_path / "ours.fa").read_bytes() == (tmp_path / "oracle.fa").read_bytes(), (alphabet, k, thresh, reads)
        checked += 1
        cycles += st["cycles"]
    assert checked > 100 and cycles > 0


# ---- branch printing: branches.cuh (record writer) + branches_host.h (the reference's traversal on the compressed graph) ----
def emul_branches(emul, tab, path, mark=False):
    counts = np.zeros(6, dtype=np.uint64)
    rc = emul.emul_branches(tab.ptr, tab.k, tab.L, tab.W, int(tab.finalized), str(path).encode(), int(mark),
                            counts.ctypes.data)
    assert rc == 0, f"emulated branch printing failed ({rc})"
    return dict(zip(["branches", "branching_nodes", "nodes", "refused", "cut", "inconsistent"], (int(c) for c in counts)))


def check_branches(emul, tab, tmp_path, what=""):
    """branches.fa of the emulated pipeline == the reference's sequential walk (oracle restatement,
    pinned on the compiled reference) over the same slot order, and the same `visited` marks left."""
    og = O.OracleGraph(tab.k, tab.L, tab.W)
    recs = tab.records()
    assert og.load_records(recs)
    ost = og.print_branches(tmp_path / "oracle_br.fa", keep_marks=True)
    st = emul_branches(emul, tab, tmp_path / "ours_br.fa", mark=True)
    assert (tmp_path / "ours_br.fa").read_bytes() == (tmp_path / "oracle_br.fa").read_bytes(), what
    assert st["branches"] == ost["branches"] and st["branching_nodes"] == ost["branching_nodes"], what
    elems = tab.elements()
    occ = elems[(elems["status"] == 1) | (elems["status"] == 2)]
    ours = {tuple(r) for r in occ["kmer"][occ["status"] == 2].tolist()}             # the marks, node for node
    orecs = og.records()
    assert ours == {tuple(r) for r in orecs["kmer"][og.status() == 2].tolist()}, what
    assert np.array_equal(tab.records(), recs), what                                # nothing else changed
    return st


@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", None), ("synth_cfg0_k53", None), ("synth_cfg2_k95", None),
                                         ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 1), ("edge_pe_k21_q5", None),
                                         ("edge_pe_k31_q5", None), ("synth_full_k53", None)])
@pytest.mark.parametrize("finalized", [False, True])
def test_emulated_branches_match_reference_walk(emul, name, thresh, finalized, tmp_path):
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    if finalized:
        tab.finalize()
    if thresh is not None:
        emul_clean(emul, tab, thresh)
    st = check_branches(emul, tab, tmp_path, name)
    if name.startswith("synth") and thresh is None:
        assert st["branches"] > 0


@pytest.mark.parametrize("k,err_q20,n_reads,genome,L", [(21, 40000, 600, 3000, 11), (53, 20000, 600, 3000, 11),
                                                        (95, 20000, 600, 3000, 11), (31, 3000, 3000, 60000, 14),
                                                        (31, 20000, 40000, 200000, 17)])
def test_emulated_branches_dense_errors_and_long_runs(emul, k, err_q20, n_reads, genome, L, tmp_path):
    """Junction next to junction (neighbours refused because an earlier branch marked them) and long
    error-free stretches (walks cut at 300 nodes, walks that meet marks left from the other end)."""
    from lasv_b200 import synth
    base = synth.CFG1 if k <= 31 else synth.CFG0 if k <= 63 else synth.CFG2
    w = synth.scaled(base, genome, L)
    w.seed, w.kmer_size = 700 + k, k
    p = w.params()
    p.err_q20 = err_q20
    seq, qual, _ = synth.reads_host(p, 0, n_reads)
    rc, tab, _ = emul_build(emul, {"k": k, "L": L, "W": 40, "q_thresh": 5}, seq, qual)
    assert rc == 0
    st = check_branches(emul, tab, tmp_path)
    assert st["branches"] > 0
    if k <= 53:
        assert st["refused"] > 0
    if err_q20 <= 3000:
        assert st["cut"] > 0


@pytest.mark.parametrize("alphabet,seed", [("AC", 11), ("CG", 12), ("ACG", 13), ("ACGT", 14)])
def test_emulated_branches_on_repetitive_graphs(emul, alphabet, seed, tmp_path):
    """Small random graphs over reduced alphabets: loops through a branching node, paths that are
    their own reverse complement, branching nodes that are each other's neighbours; raw and cleaned."""
    rng = np.random.default_rng(seed)
    comp = str.maketrans("ACGT", "TGCA")
    checked = printed = 0
    for _ in range(150):
        k = int(rng.choice([9, 11, 15, 21]))
        glen = int(rng.integers(30, 400))
        genome = "".join(rng.choice(list(alphabet), size=glen))
        src = genome * 3 if rng.random() < 0.5 else genome
        reads = []
        for _ in range(int(rng.integers(1, 40))):
            s, rl = int(rng.integers(0, glen)), int(rng.integers(k, k + 60))
            r = src[s:s + rl]
            if len(r) < k:
                continue
            if rng.random() < 0.3:
                r = list(r)
                r[int(rng.integers(0, len(r)))] = str(rng.choice(list("ACGT")))
                r = "".join(r)
            reads.append(r.translate(comp)[::-1] if rng.random() < 0.5 else r)
        if not reads:
            continue
        seq, _, _ = O.reads_to_streams(reads)
        rc, tab, _ = emul_build(emul, {"k": k, "L": 6, "W": 40, "q_thresh": 0}, seq, None)
        if rc != 0:
            continue
        if rng.random() < 0.4:
            emul_clean(emul, tab, int(rng.integers(1, 3)))
        st = check_branches(emul, tab, tmp_path, (alphabet, k, reads))
        checked += 1
        printed += st["branches"]
    assert checked > 120 and printed > 0


def test_emulated_clean_then_finalize_keeps_pruned_nodes_pruned(emul, tmp_path):
    """lasv_graph_clean on the line layout, then finalisation (export): pruned nodes keep their place and their
    status, as in the reference's table; the records of the graph (pruned skipped) and everything printed
    from it are the same before and after."""
    name, thresh = "synth_cfg0_k53", 2
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    og, _ = oracle_clean_in_order(tab, thresh)
    st = emul_clean(emul, tab, thresh)
    recs = tab.records()
    assert np.array_equal(recs, og.records())
    emul_supernodes(emul, tab, tmp_path / "a.fa")
    tab.finalize()
    elems = tab.elements()
    assert int((elems["status"] == 3).sum()) == st["nodes_pruned"] > 0
    assert np.array_equal(tab.records(), recs)
    emul_supernodes(emul, tab, tmp_path / "b.fa")
    assert (tmp_path / "a.fa").read_bytes() == (tmp_path / "b.fa").read_bytes()
    check_branches(emul, tab, tmp_path)


# ---- `--health_check` (dB_graph_population.c:6274-6350): hc_check_node of cleaning.cuh ----
def emul_health(emul, tab):
    counts = np.zeros(5, dtype=np.uint64)
    emul.emul_health(tab.ptr, tab.k, tab.L, tab.W, int(tab.finalized), counts.ctypes.data)
    return dict(zip(["nodes_checked", "edges_checked", "missing_target", "missing_return_arrow", "zero_kmers"],
                    (int(c) for c in counts)))


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "edge_pe_k21_q5", "synth_full_k53"])
def test_emulated_health_check(emul, name):
    """A graph built from reads has a return arrow for every edge, before and after the cleaning, in both table
    arrangements; a flipped edge bit is found."""
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    rc, tab, _ = emul_build(emul, meta, seq, qual)
    assert rc == 0
    recs = tab.records()
    n_edges = int(np.unpackbits(recs["edges"]).sum())
    zero = int((recs["kmer"] == 0).all(axis=1).sum())          # the poly-A k-mer of the homopolymer reads, if any
    want = {"nodes_checked": len(recs), "edges_checked": n_edges, "missing_target": 0, "missing_return_arrow": 0,
            "zero_kmers": zero}
    assert zero <= 1 and emul_health(emul, tab) == want
    tab.finalize()
    assert emul_health(emul, tab) == want
    if name.startswith("synth") and name != "synth_full_k53":
        st = emul_clean(emul, tab, 2)
        h = emul_health(emul, tab)
        assert h["missing_target"] == h["missing_return_arrow"] == 0
        assert h["nodes_checked"] == len(recs) and st["nodes_pruned"] > 0        # pruned nodes are visited, they have no edges
        assert h["edges_checked"] == int(np.unpackbits(tab.records()["edges"]).sum()) < n_edges
    # corrupt the finalised table: drop one arrow (its twin loses the return arrow), add one that leads nowhere
    N, slot = tab.N, 8 * tab.N + 8
    bucket_bytes = len(tab.raw) >> tab.L
    elems = tab.elements()
    victim = int(np.flatnonzero((elems["status"] == 1) & (elems["edges"] != 0) & (elems["edges"] != 0xFF))[0])
    at = (victim // tab.W) * bucket_bytes + (victim % tab.W) * slot + 8 * N + 4
    before = emul_health(emul, tab)
    old = int(tab.raw[at])
    low = old & -old
    tab.raw[at] = old & ~low
    h = emul_health(emul, tab)
    assert h["missing_return_arrow"] == 1 and h["edges_checked"] == before["edges_checked"] - 1
    free = next(b for b in range(8) if not (old >> b) & 1)
    tab.raw[at] = old | (1 << free)
    h = emul_health(emul, tab)
    assert h["missing_target"] + h["missing_return_arrow"] == 1 and h["edges_checked"] == before["edges_checked"] + 1


# ---- streamed-insert prototype (groups.cuh): records grouped by first-choice bucket, groups updated in a staged copy ----
@pytest.mark.parametrize("name,bpg", [("synth_cfg0_k53", 4), ("synth_cfg1_k31", 1), ("synth_cfg2_k95", 8),
                                      ("edge_pe_k65_q5", 2), ("synth_full_k53", 4), ("synth_full_k53", 1)])
def test_emulated_grouped_insert(emul, name, bpg):
    """Same graph as the random-access insert -- also when buckets fill up and keys leave their group through
    the rehash chain (synth_full_k53: 90 % full table) -- and the same rehash histogram."""
    meta = util.case_meta(name)
    k, L, W = meta["k"], meta["L"], meta["W"]
    N = words_per_kmer(k)
    seq, qual, _ = util.case_streams(name)
    cap = meta["inserts"] + 10
    recs = np.zeros((cap, 2 * N + 1), dtype=np.uint32)
    n = emul.emul_extract_records(seq.ctypes.data, qual.ctypes.data, len(seq), k, util.cutoff_of(meta), 384,
                                  recs.ctypes.data, cap)
    assert n == meta["inserts"]
    tab = EmulTable(emul, k, L, W)
    ctr, spilled = np.zeros(20, dtype=np.uint64), C.c_uint64(0)
    rc = emul.emul_insert_grouped(recs.ctypes.data, n, k, L, W, 15, bpg, tab.ptr, ctr.ctypes.data, C.byref(spilled))
    assert rc == 0
    util.assert_same_records(tab.records(), util.case_records(name), name)
    assert int(ctr[0]) == meta["n_records"]
    if name == "synth_full_k53":
        assert spilled.value > 0                                   # keys that left their group
        # every key sits on its reference rehash chain: the reference's find locates all of them after finalisation
        tab.finalize()
        keys = np.ascontiguousarray(util.case_records(name)["kmer"])
        found, _, _ = tab.find(keys)
        assert found.all()
    else:
        assert spilled.value == 0
    # in two halves (a second batch meets the keys of the first)
    tab2 = EmulTable(emul, k, L, W)
    for part in (recs[: n // 2], recs[n // 2: n]):
        part = np.ascontiguousarray(part)
        assert emul.emul_insert_grouped(part.ctypes.data, len(part), k, L, W, 15, bpg, tab2.ptr, ctr.ctypes.data,
                                        C.byref(spilled)) == 0
    util.assert_same_records(tab2.records(), util.case_records(name), name)


def test_emulated_compressed_graph_built_by_threads(emul, tmp_path):
    """ClGraph::add_paths hands ranges of path records to threads (no two paths write the same adjacency entry):
    forced on for a small graph, cleaning and branch printing must not change."""
    name, thresh = "synth_cfg0_k53", 2
    meta = util.case_meta(name)
    seq, qual, _ = util.case_streams(name)
    emul.emul_set_parallel_min(1)
    try:
        rc, tab, _ = emul_build(emul, meta, seq, qual)
        assert rc == 0
        check_branches(emul, tab, tmp_path)
        og, ost = oracle_clean_in_order(tab, thresh)
        st = emul_clean(emul, tab, thresh)
        assert np.array_equal(tab.records(), og.records()) and st["supernodes_pruned"] == ost["supernodes_pruned"] > 0
    finally:
        emul.emul_set_parallel_min(1 << 18)


# ---- path enumeration by list ranking (ranking.cuh) instead of one walk per path start ----
@pytest.fixture
def ranked(emul):
    emul.emul_set_ranking(1)
    yield emul
    emul.emul_set_ranking(0)


@pytest.mark.parametrize("name", ["synth_cfg1_k31", "synth_cfg0_k53", "synth_cfg2_k95", "edge_pe_k21_q5"])
@pytest.mark.parametrize("finalized", [False, True])
def test_ranked_supernodes_match_reference_golden(ranked, name, finalized, tmp_path):
    test_emulated_supernodes_match_reference_golden(ranked, name, finalized, tmp_path)


@pytest.mark.parametrize("k,err_q20,seed", [(21, 30000, 3), (31, 40000, 5), (53, 30000, 6)])
def test_ranked_supernodes_order_dependent_cases(ranked, k, err_q20, seed, tmp_path):
    test_emulated_supernodes_order_dependent_cases(ranked, k, err_q20, seed, tmp_path)


def test_ranked_supernodes_cycles_and_long_paths(ranked, tmp_path):
    """A pure cycle (elements that never finish) and a 5 969-edge path (13 rounds instead of 5 969 steps)."""
    test_emulated_supernodes_cycles_and_long_paths(ranked, tmp_path)


@pytest.mark.parametrize("name,thresh", [("synth_cfg1_k31", 2), ("synth_cfg0_k53", 2), ("synth_cfg2_k95", 2),
                                         ("edge_pe_k31_q5", 3)])
def test_ranked_cleaning_and_branches(ranked, name, thresh, tmp_path):
    test_emulated_cleaning_matches_reference_walk(ranked, name, thresh, False, tmp_path)
    test_emulated_branches_match_reference_walk(ranked, name, None, True, tmp_path)


@pytest.mark.parametrize("alphabet,seed", [("CG", 22), ("ACG", 23), ("AT", 24)])
def test_ranked_on_repetitive_graphs(ranked, alphabet, seed, tmp_path):
    """Short cycles, palindromic paths (a node crossed in both orientations: the earlier crossing decides) and loops
    through a junction: cleaning, supernodes and branches from ranked records equal the reference's."""
    test_emulated_cleaning_and_supernodes_on_repetitive_graphs(ranked, alphabet, seed, tmp_path)
    test_emulated_branches_on_repetitive_graphs(ranked, alphabet, seed, tmp_path)
