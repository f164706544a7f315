"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/liboracle.so (the plain-C restatement of the
reference's graph build, oracle/dbg_oracle.c) plus helpers to run the compiled,
unmodified reference under oracle/_ref/ and to parse its `.ctx` output.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module, and only as the checker / baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
MAXK_OF_N = {1: 31, 2: 63, 3: 95}


def words_per_kmer(k: int) -> int:
    return (2 * k) // 64 + 1  # graph_operations.c:620


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", str(HERE), "-j8"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.orc_graph_new.restype = C.c_void_p
        L.orc_graph_new.argtypes = [C.c_int] * 4
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_load_reads.restype = C.c_int
        L.orc_load_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_unique_kmers.restype = C.c_uint64
        L.orc_unique_kmers.argtypes = [C.c_void_p]
        L.orc_table_full.restype = C.c_int
        L.orc_table_full.argtypes = [C.c_void_p]
        L.orc_collisions.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dump_records.restype = C.c_uint64
        L.orc_dump_records.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dump_status.restype = C.c_uint64
        L.orc_dump_status.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_load_records.restype = C.c_int
        L.orc_load_records.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_lookup.restype = C.c_int
        L.orc_lookup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hash_value.restype = C.c_uint32
        L.orc_hash_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_seq_to_kmer.restype = C.c_int
        L.orc_seq_to_kmer.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_kmer_to_seq.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.orc_revcomp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_key.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_print_supernodes.restype = C.c_int
        L.orc_print_supernodes.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_print_branches.restype = C.c_int
        L.orc_print_branches.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
        L.orc_clean.restype = C.c_int
        L.orc_clean.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_ctx_header_cleaned.restype = C.c_int
        L.orc_ctx_header_cleaned.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_ctx_header.restype = C.c_int
        L.orc_ctx_header.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


class Stats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("bad_reads", C.c_uint64), ("bases_read", C.c_uint64),
                ("bases_loaded", C.c_uint64), ("kmers_inserted", C.c_uint64)]


def record_dtype(N: int) -> np.dtype:
    """One `.ctx` record: kmer[N] | uint32 covg | uint8 edges (element.c:894-910)."""
    return np.dtype([("kmer", "<u8", (N,)), ("covg", "<u4"), ("edges", "u1")], align=False)


def sort_records(recs: np.ndarray) -> np.ndarray:
    """Sorted multiset of (key words, covg, edges): the parity object."""
    N = recs["kmer"].shape[1]
    keys = [recs["edges"], recs["covg"]] + [recs["kmer"][:, i] for i in range(N - 1, -1, -1)]
    return recs[np.lexsort(keys)]


class OracleGraph:
    """CPU graph build following the reference's control flow (dbg_oracle.c)."""

    def __init__(self, k: int, L: int, W: int, max_rehash: int = 15):
        self.k, self.L, self.W = k, L, W
        self.N = words_per_kmer(k)
        self._g = lib().orc_graph_new(k, L, W, max_rehash)
        if not self._g:
            raise MemoryError("oracle table allocation failed")
        self.stats = Stats()
        self.hist = np.zeros(20001, dtype=np.uint64)

    def close(self):
        if self._g:
            lib().orc_graph_free(self._g)
            self._g = None

    __del__ = close

    def load_reads(self, seq: np.ndarray, qual, offsets: np.ndarray, cutoff: int, sep: int = 1,
                   paired: bool = False) -> bool:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        qp = None
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
            qp = qual.ctypes.data
        ok = lib().orc_load_reads(self._g, seq.ctypes.data, qp, offsets.ctypes.data, len(offsets) - 1,
                                  sep, cutoff, 1 if paired else 0, C.byref(self.stats),
                                  self.hist.ctypes.data, len(self.hist))
        return bool(ok)

    @property
    def unique_kmers(self) -> int:
        return int(lib().orc_unique_kmers(self._g))

    @property
    def table_full(self) -> bool:
        return bool(lib().orc_table_full(self._g))

    def collisions(self) -> np.ndarray:
        out = np.zeros(16, dtype=np.uint64)
        lib().orc_collisions(self._g, out.ctypes.data)
        return out

    def records(self) -> np.ndarray:
        out = np.zeros(self.unique_kmers, dtype=record_dtype(self.N))
        n = lib().orc_dump_records(self._g, out.ctypes.data)   # pruned nodes are not dumped
        assert n <= len(out)
        return out[:n]

    def status(self) -> np.ndarray:
        """Status byte (1 none, 2 visited) of every record of records(), in that order."""
        out = np.zeros(self.unique_kmers, dtype=np.uint8)
        n = lib().orc_dump_status(self._g, out.ctypes.data)
        return out[:n]

    def clean(self, threshold: int, max_var_len: int = 10000) -> dict:
        """`--remove_low_coverage_supernodes threshold` (graph_operations.c:1069-1090): clip tips, then
        prune low-coverage supernodes, in table order."""
        counts = np.zeros(4, dtype=np.uint64)
        if lib().orc_clean(self._g, threshold, max_var_len, counts.ctypes.data) != 0:
            raise RuntimeError("oracle cleaning met an edge to a missing node (the reference dies there)")
        return {"tips": int(counts[0]), "tip_nodes": int(counts[1]), "supernodes_pruned": int(counts[2]),
                "nodes_left": int(counts[3])}

    def load_records(self, recs: np.ndarray) -> bool:
        recs = np.ascontiguousarray(recs)
        return bool(lib().orc_load_records(self._g, recs.ctypes.data, len(recs)))

    def print_supernodes(self, path, max_length: int = 10000) -> dict:
        """`--output_supernodes` (dB_graph_population.c:2696-2803) on this table, in table order."""
        counts = np.zeros(4, dtype=np.uint64)
        rc = lib().orc_print_supernodes(self._g, str(path).encode(), max_length, counts.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle supernode walk failed (I/O error or an edge to a missing node)")
        return {"supernodes": int(counts[0]), "singletons": int(counts[1]), "nodes": int(counts[2]),
                "at_max_length": int(counts[3])}

    def print_branches(self, path, keep_marks: bool = False) -> dict:
        """`--mode build`'s branches.fa (print_branches.c:131-199) on this table, in table order.
        keep_marks=True leaves the `visited` marks in the table, as the reference does."""
        counts = np.zeros(3, dtype=np.uint64)
        rc = lib().orc_print_branches(self._g, str(path).encode(), 1 if keep_marks else 0, counts.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle branch walk failed (I/O error or an edge to a missing node)")
        return {"branches": int(counts[0]), "branching_nodes": int(counts[1]), "visited": int(counts[2])}

    def lookup(self, key_words):
        key = np.asarray(key_words, dtype=np.uint64)
        c, e = C.c_uint32(), C.c_uint8()
        if lib().orc_lookup(self._g, key.ctypes.data, C.byref(c), C.byref(e)):
            return int(c.value), int(e.value)
        return None


def hash_value(key_words, rehash: int, L: int) -> int:
    key = np.asarray(key_words, dtype=np.uint64)
    return int(lib().orc_hash_value(key.ctypes.data, len(key), rehash, L))


def seq_to_kmer(s: str, k: int) -> np.ndarray:
    N = words_per_kmer(k)
    out = np.zeros(N, dtype=np.uint64)
    assert lib().orc_seq_to_kmer(s.encode(), k, N, out.ctypes.data)
    return out


def kmer_to_seq(words, k: int) -> str:
    w = np.asarray(words, dtype=np.uint64)
    buf = C.create_string_buffer(k + 1)
    lib().orc_kmer_to_seq(w.ctypes.data, k, len(w), buf)
    return buf.value.decode()


def revcomp(words, k: int) -> np.ndarray:
    w = np.asarray(words, dtype=np.uint64)
    out = np.zeros_like(w)
    lib().orc_revcomp(w.ctypes.data, k, len(w), out.ctypes.data)
    return out


def get_key(words, k: int) -> np.ndarray:
    w = np.asarray(words, dtype=np.uint64)
    out = np.zeros_like(w)
    lib().orc_get_key(w.ctypes.data, k, len(w), out.ctypes.data)
    return out


def ctx_header(k: int, mean_read_len: int, total_seq: int, cleaned_threshold=None) -> bytes:
    buf = np.zeros(128, dtype=np.uint8)
    if cleaned_threshold is None:
        n = lib().orc_ctx_header(k, words_per_kmer(k), mean_read_len, total_seq, buf.ctypes.data)
    else:
        n = lib().orc_ctx_header_cleaned(k, words_per_kmer(k), mean_read_len, total_seq, cleaned_threshold,
                                         buf.ctypes.data)
    return buf[:n].tobytes()


# --------------------------------------------------------------------------- .ctx
def parse_ctx(path) -> dict:
    """Parse a BINVERSION 4-6 single-colour `.ctx` (file_reader.c:3023-3232)."""
    data = Path(path).read_bytes()
    assert data[:6] == b"CORTEX", "missing magic"
    version, k, N, ncols = np.frombuffer(data, dtype="<i4", count=4, offset=6)
    assert ncols == 1
    off = 22
    mean_read_len = int(np.frombuffer(data, "<i4", 1, off)[0]); off += 4
    total_seq = int(np.frombuffer(data, "<i8", 1, off)[0]); off += 8
    if version > 5:
        idlen = int(np.frombuffer(data, "<i4", 1, off)[0]); off += 4 + idlen
        off += 16  # long double seq_err
        off += 4   # 4 x boolean
        off += 8   # two thresholds
        namelen = int(np.frombuffer(data, "<i4", 1, off)[0]); off += 4 + namelen
    assert data[off:off + 6] == b"CORTEX", "missing end-of-header magic"
    off += 6
    recs = np.frombuffer(data, dtype=record_dtype(int(N)), offset=off)
    return {"version": int(version), "k": int(k), "N": int(N), "mean_read_len": mean_read_len,
            "total_seq": total_seq, "header": data[:off], "records": recs}


# ---------------------------------------------------------------- FASTQ helpers
def reads_to_streams(reads, quals=None):
    """Python lists of read strings -> (seq stream, qual stream, offsets) with the
    '\\n'-separated layout the loaders consume."""
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    parts_s, parts_q = [], []
    pos = 0
    for i, r in enumerate(reads):
        parts_s.append(r.encode() + b"\n")
        if quals is not None:
            assert len(quals[i]) == len(r)
            parts_q.append(quals[i].encode() + b"\n")
        pos += len(r) + 1
        offs[i + 1] = pos
    seq = np.frombuffer(b"".join(parts_s), dtype=np.uint8).copy()
    qual = np.frombuffer(b"".join(parts_q), dtype=np.uint8).copy() if quals is not None else None
    return seq, qual, offs


def write_fastq_fixed(path, seq: np.ndarray, qual: np.ndarray, read_len: int, name: str = "r") -> None:
    """Write a fixed-stride stream (read_len bases + '\\n' per read) as 4-line FASTQ."""
    stride = read_len + 1
    n = len(seq) // stride
    s2 = seq[: n * stride].reshape(n, stride)
    q2 = qual[: n * stride].reshape(n, stride)
    with open(path, "wb") as f:
        chunk = 200000
        for i0 in range(0, n, chunk):
            i1 = min(n, i0 + chunk)
            out = bytearray()
            for i in range(i0, i1):
                out += b"@%s%d\n" % (name.encode(), i)
                out += s2[i].tobytes()           # includes the '\n'
                out += b"+\n"
                out += q2[i].tobytes()
            f.write(out)


def parse_fastq_python(path):
    """Pure-Python restatement of the FASTQ grammar of libs/seq_file/seq_fastq.c:24-150
    (multi-line sequence up to a line starting with '+'; exactly len(seq) quality
    characters over any number of lines).  Small inputs only."""
    import gzip
    raw = Path(path).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    reads, quals = [], []
    i, n = 0, len(raw)
    if n == 0 or raw[0:1] != b"@":
        raise ValueError("not FASTQ")
    i = 1
    while True:
        j = raw.find(b"\n", i)
        if j < 0:
            break
        i = j + 1                                     # name line consumed
        seq = bytearray()
        while i < n and raw[i:i + 1] != b"+":
            c = raw[i:i + 1]
            if c in (b"\r", b"\n"):
                i += 1
                continue
            j = raw.find(b"\n", i)
            j = n if j < 0 else j
            seq += raw[i:j].rstrip(b"\r\n")
            i = j + 1
        if i >= n:
            break
        j = raw.find(b"\n", i)
        i = n if j < 0 else j + 1                     # '+' line consumed
        q = bytearray()
        while len(q) < len(seq) and i < n:
            c = raw[i:i + 1]
            if c not in (b"\r", b"\n"):
                q += c
            i += 1
        reads.append(seq.decode("latin1"))
        quals.append(q.decode("latin1"))
        while i < n and raw[i:i + 1] in (b"\r", b"\n"):
            i += 1
        if i >= n or raw[i:i + 1] != b"@":
            break
        i += 1
    return reads, quals


# ------------------------------------------------------- the compiled reference
def ref_binary(k: int) -> Path:
    return REF_DIR / f"dB_graph_{MAXK_OF_N[words_per_kmer(k)]}"


def ref_available(k: int = 31) -> bool:
    return ref_binary(k).exists()


def run_reference(k: int, L: int, W: int, fq1, fq2=None, q_thresh: int = 0, workdir=None,
                  extra_args=(), want_supernodes: bool = False, ctx_in=None, timeout=3600) -> dict:
    """Run the unmodified reference CLI (graph_operations.c:main) load-only and
    return its parsed `.ctx`, its stdout, the load wall time from an external
    monotonic clock and the insert count (sum of the "tries r:" lines,
    hash_table.c:384-394)."""
    exe = ref_binary(k)
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle` where /root/reference exists")
    own = workdir is None
    wd = Path(tempfile.mkdtemp(prefix="lasv_ref_")) if own else Path(workdir)
    out_ctx = wd / "out.ctx"
    if out_ctx.exists():
        out_ctx.unlink()
    cmd = [str(exe), "--kmer_size", str(k), "--mem_height", str(L), "--mem_width", str(W),
           "--dump_binary", str(out_ctx)]
    if ctx_in is not None:
        cmd += ["--multicolour_bin", str(ctx_in)]
    elif fq2 is not None:
        l1, l2 = wd / "left.file", wd / "right.file"
        l1.write_text(str(Path(fq1).resolve()) + "\n")
        l2.write_text(str(Path(fq2).resolve()) + "\n")
        cmd += ["--pe_list", f"{l1},{l2}"]
    else:
        l1 = wd / "se.file"
        l1.write_text(str(Path(fq1).resolve()) + "\n")
        cmd += ["--se_list", str(l1)]
    if q_thresh > 0:
        cmd += ["--quality_score_threshold", str(q_thresh)]
    sup = wd / "sup.fa"
    if want_supernodes:
        if sup.exists():
            sup.unlink()
        cmd += ["--output_supernodes", str(sup)]
    cmd += list(extra_args)
    t0 = time.monotonic()
    p = subprocess.run(cmd, cwd=wd, capture_output=True, text=True, timeout=timeout)
    wall = time.monotonic() - t0
    if p.returncode != 0:
        raise RuntimeError(f"reference failed ({p.returncode}): {p.stderr[-2000:]}\n{p.stdout[-2000:]}")
    tries = 0
    for line in p.stdout.splitlines():
        s = line.strip()
        if s.startswith("tries "):
            tries += int(s.split(":")[1])
    res = {"stdout": p.stdout, "stderr": p.stderr, "wall_s": wall, "inserts": tries,
           "ctx": parse_ctx(out_ctx), "workdir": wd}
    if want_supernodes:
        res["supernodes"] = read_fasta_seqs(sup)
    return res


def read_fasta_seqs(path):
    seqs, cur = [], []
    for line in Path(path).read_text().splitlines():
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(line.strip())
    if cur:
        seqs.append("".join(cur))
    return seqs


_COMP = str.maketrans("ACGT", "TGCA")


def canonical_seq_set(seqs):
    """Set of min(seq, revcomp(seq)): the strand-independent contig set (config 3)."""
    out = []
    for s in seqs:
        rc = s.translate(_COMP)[::-1]
        out.append(min(s, rc))
    return sorted(out)
