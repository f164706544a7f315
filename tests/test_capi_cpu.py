"""CPU-side checks of the C-ABI library: it loads, exports every symbol
include/lasv_b200.h declares, refuses to compute without a GPU (no fallback),
and its host-only pieces (sequence-file reader, synthetic-read generator, .ctx
header) behave like the reference."""
import re
from pathlib import Path

import numpy as np
import pytest

import lasv_b200
from lasv_b200 import _capi, synth
from lasv_b200.graph import seqfile_to_streams
from oracle import oracle as O
from tests import util

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    hdr = (ROOT / "include" / "lasv_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:lasv|load)_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), f"liblasv_b200.so does not export {sym}"


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/lasv_b200.h must compile as C99 (no C++, no torch types)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "include/lasv_b200.h"\nint main(void) { return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", str(ROOT),
                    str(src)], check=True)


def test_product_and_tools_never_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under lasv_b200/, include/, integration/*.c or tools/ imports, links,
    loads or executes it (bench.py and __graft_entry__.smoke() use it as the checker / CPU baseline only)."""
    import subprocess
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/_ref|dbg_oracle|oracle\.py", re.M)
    offenders = []
    for sub, globs in (("lasv_b200", ("*.py", "csrc/*.cu", "csrc/*.cuh", "csrc/*.h", "csrc/*.cpp", "csrc/Makefile")),
                       ("include", ("*.h",)), ("integration", ("*.c",)), ("tools", ("*.py", "*.sh", "ubench/*.cu"))):
        for g in globs:
            for f in (ROOT / sub).glob(g):
                if pat.search(f.read_text(errors="replace")):
                    offenders.append(str(f.relative_to(ROOT)))
    assert not offenders, offenders
    out = subprocess.run(["ldd", str(ROOT / "lasv_b200" / "liblasv_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_no_cpu_fallback_without_gpu():
    lib = _capi.lib()
    if lib.lasv_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.LasvError) as e:
        lasv_b200.DBGraph(31, 10, 20)
    assert "no usable CUDA device" in str(e.value)


def test_argument_validation():
    lib = _capi.lib()
    assert not lib.lasv_graph_new(32, 10, 20, 15, 0)          # even k (cmd_line.c:805)
    assert b"odd" in lib.lasv_last_error()
    assert not lib.lasv_graph_new(97, 10, 20, 15, 0)          # beyond MAXK=95
    assert not lib.lasv_graph_new(31, 0, 20, 15, 0)


def test_reference_table_entry_points_fail_loudly():
    """The entries a reference-side shim binds (cleaning, supernodes, branches, .ctx reload) never fall back to
    the CPU: bad arguments are refused with a message, and without a GPU they report the missing device."""
    import ctypes as C
    lib = _capi.lib()
    for call in (lambda: lib.lasv_ref_hash_table_clean(None, _capi.CLEAN_TIPS, 0, None),
                 lambda: lib.lasv_ref_hash_table_write_supernodes(None, b"/dev/null", 10000, None),
                 lambda: lib.lasv_ref_hash_table_write_branches(None, b"/dev/null", None),
                 lambda: lib.lasv_graph_write_branches(None, b"/dev/null", 0, None),
                 lambda: lib.lasv_graph_clean(None, _capi.CLEAN_TIPS, 0, None)):
        assert call() < 0 and lib.lasv_last_error()
    table = np.zeros(4 * 8, dtype=lasv_b200.element_dtype(1))
    ht = _capi.RefHashTable(31, 3, 8, table.ctypes.data, None, None, 0, 15)    # 3 buckets: not a power of two
    assert lib.lasv_ref_hash_table_write_branches(C.byref(ht), b"/dev/null", None) < 0
    assert b"power of two" in lib.lasv_last_error()
    ht = _capi.RefHashTable(31, 4, 8, table.ctypes.data, None, None, 0, 15)
    assert lib.lasv_ref_hash_table_clean(C.byref(ht), 8, 0, None) < 0          # unknown pass
    if lib.lasv_device_count() == 0:
        for call in (lambda: lib.lasv_ref_hash_table_write_branches(C.byref(ht), b"/dev/null", None),
                     lambda: lib.lasv_ref_hash_table_clean(C.byref(ht), _capi.CLEAN_TIPS, 0, None)):
            assert call() < 0
            assert b"CUDA" in lib.lasv_last_error()


@pytest.mark.parametrize("fname", ["edge_1.fq", "edge_2.fq", "edge_1.fq.gz"])
def test_fastq_reader_matches_reference_grammar(fname):
    seq, qual, offs = seqfile_to_streams(util.GOLDEN / fname)
    reads, quals = O.parse_fastq_python(util.GOLDEN / fname)
    s2, q2, o2 = O.reads_to_streams(reads, quals)
    assert np.array_equal(seq, s2) and np.array_equal(qual, q2) and np.array_equal(offs, o2)


def test_fasta_and_plain_reader(tmp_path):
    seq, qual, offs = seqfile_to_streams(util.GOLDEN / "contigs.fa")
    assert qual is None and len(offs) == 3
    reads = util._fasta_reads(util.GOLDEN / "contigs.fa")
    assert seq.tobytes().decode().split("\n")[:-1] == reads
    p = tmp_path / "plain.txt"
    p.write_text("ACGT\nGGCC\n\nTT")
    seq, qual, offs = seqfile_to_streams(p)
    assert qual is None
    assert seq.tobytes().decode().split("\n")[:-1] == ["ACGT", "GGCC", "", "TT"]   # seq_plain.c: empty line = empty read
    q = tmp_path / "x.sam"
    q.write_text("@HD\n")
    with pytest.raises(_capi.LasvError):
        seqfile_to_streams(q)
    with pytest.raises(_capi.LasvError):
        seqfile_to_streams(tmp_path / "missing.fq")


def test_reader_through_oracle_matches_reference_golden():
    """The reader's bytes, fed to the oracle, reproduce the reference's output on
    the file-based golden cases: pins the parser end to end on the CPU."""
    for name in util.EDGE_PE + util.EDGE_SE + util.REFFIX:      # REFFIX: the reference's own reader fixtures
        meta = util.case_meta(name)
        f1, f2 = util.case_files(name)
        s1, q1, o1 = seqfile_to_streams(f1)
        if f2 is None:
            seq, qual, offs, paired = s1, q1, o1, False
        else:
            s2, q2, o2 = seqfile_to_streams(f2)
            r1 = s1.tobytes().decode("latin1").split("\n")[:-1]
            r2 = s2.tobytes().decode("latin1").split("\n")[:-1]
            u1 = q1.tobytes().decode("latin1")
            u2 = q2.tobytes().decode("latin1")
            qa = [u1[int(o1[i]):int(o1[i + 1]) - 1] for i in range(len(r1))]
            qb = [u2[int(o2[i]):int(o2[i + 1]) - 1] for i in range(len(r2))]
            reads, quals = [], []
            for a, b, c, d in zip(r1, r2, qa, qb):
                reads += [a, b]
                quals += [c, d]
            seq, qual, offs = O.reads_to_streams(reads, quals)
            paired = True
        g, ok = util.oracle_build(meta["k"], meta["L"], meta["W"], seq, qual, offs, util.cutoff_of(meta), paired)
        assert ok
        util.assert_same_records(g.records(), util.case_records(name), name)
        assert g.stats.kmers_inserted == meta["inserts"] and g.stats.bad_reads == meta["bad_reads"], name
        assert g.stats.bases_loaded == meta["bases_loaded"], name


def test_synth_generator_is_deterministic_and_sane():
    w = synth.scaled(synth.CFG2, 5000, 10)
    p = w.params()
    a = synth.reads_host(p, 0, 64)
    b = synth.reads_host(p, 32, 32)
    stride = w.read_len + 1
    assert np.array_equal(a[0][32 * stride:], b[0]) and np.array_equal(a[1][32 * stride:], b[1])
    s = a[0].reshape(64, stride)
    assert (s[:, -1] == 10).all() and set(np.unique(s[:, :-1])) <= set(b"ACGTN")
    q = a[1].reshape(64, stride)[:, :-1]
    assert q.min() >= 33 + 2 and q.max() <= 33 + 40
    # mates of a pair come from opposite strands of one fragment: with no errors the
    # reverse complement of mate 2 matches the genome downstream of mate 1
    w0 = synth.scaled(synth.CFG0, 5000, 10)
    w0.err = w0.n_rate = w0.lowq = 0.0
    s0 = synth.reads_host(w0.params(), 0, 2)[0].tobytes().decode().split("\n")
    tiled = synth.reads_host(w0.params(tiling=True), 0, (5000 // 48) + 2)[0].tobytes().decode().split("\n")
    genome = tiled[0] + "".join(t[52:] for t in tiled[1:-1])
    genome = genome.replace("N", "")
    comp = str.maketrans("ACGT", "TGCA")
    m1, m2 = s0[0], s0[1]
    fwd, rev = (m1, m2.translate(comp)[::-1])
    if fwd not in genome:
        fwd, rev = (m2, m1.translate(comp)[::-1])
    i = genome.find(fwd)
    assert i >= 0 and genome[i + w0.insert - w0.read_len: i + w0.insert] == rev


def test_ctx_header_matches_reference_bytes():
    meta = util.case_meta("synth_cfg0_k53")
    hdr = lasv_b200.ctx_header(53, meta["ctx_mean_read_len"], meta["ctx_total_seq"])
    assert hdr.hex() == meta["ctx_header_hex"]
    assert hdr == O.ctx_header(53, meta["ctx_mean_read_len"], meta["ctx_total_seq"])


# ------------------------------------------------ parallel parse of large FASTQ files
def _parse_check(path1, path2, threads, batch_bytes=1 << 16):
    import ctypes as C
    out = (C.c_uint64 * 4)()
    used = C.c_int(0)
    rc = _capi.lib().lasv_seqfile_parse_check(str(path1).encode(), str(path2).encode() if path2 else None, threads,
                                              batch_bytes, out, C.byref(used))
    assert rc == 0, _capi.lib().lasv_last_error()
    return tuple(int(x) for x in out), used.value


def test_parallel_fastq_parse_equals_sequential(tmp_path):
    """Large uncompressed FASTQ files are cut at record boundaries and parsed slice by slice by several
    threads; everything else (gzip, multi-line records, FASTA, small files) gets one sequential parser.
    Either way: same reads, same bases, same checksum of all (sequence, quality) pairs."""
    rng = np.random.default_rng(11)
    n = 30_000
    lens = rng.integers(1, 220, n)
    with open(tmp_path / "a.fq", "w") as fa, open(tmp_path / "crlf.fq", "w", newline="") as fc, \
            open(tmp_path / "wrapped.fq", "w") as fw:
        for i, ln in enumerate(lens):
            s = "".join(rng.choice(list("ACGTNacgt"), ln))
            # quality lines that start with '@' or '+' are the classic trap for boundary detection
            q = "".join(rng.choice(list("@+IIIIIIFF#5&'"), ln))
            fa.write(f"@r{i} x\n{s}\n+\n{q}\n")
            fc.write(f"@r{i}\r\n{s}\r\n+r{i}\r\n{q}\r\n")
            h = ln // 2
            fw.write(f"@r{i}\n{s[:h]}\n{s[h:]}\n+\n{q[:h]}\n{q[h:]}\n")        # multi-line record
    import gzip, shutil
    with open(tmp_path / "a.fq", "rb") as src, gzip.open(tmp_path / "a.fq.gz", "wb", compresslevel=1) as dst:
        shutil.copyfileobj(src, dst)
    ref, used = _parse_check(tmp_path / "a.fq", None, 1)
    assert used == 1 and ref[0] == n and ref[2] == int(lens.sum())
    size = (tmp_path / "a.fq").stat().st_size
    for threads in (2, 3, 4, 6, 64):
        got, used = _parse_check(tmp_path / "a.fq", None, threads)
        assert used == (threads if size >= threads << 20 else 1) and got == ref, (threads, got, ref)
    got, used = _parse_check(tmp_path / "crlf.fq", None, 4)
    assert used == 4 and got == ref                                   # '\r' is not part of a read
    got, used = _parse_check(tmp_path / "wrapped.fq", None, 4)
    assert used == 1 and got == ref                                   # multi-line: sequential fallback
    got, used = _parse_check(tmp_path / "a.fq.gz", None, 4)
    assert used == 1 and got == ref                                   # compressed: sequential
    got, used = _parse_check(tmp_path / "a.fq", tmp_path / "crlf.fq", 3)
    assert used == 6 and got == (n, n, 2 * ref[2], (2 * ref[3]) % (1 << 64))
    # mostly plain four-line records (taken by the slice parsers' fast path) with odd ones in between: wrapped,
    # CRLF, an empty read, blank lines between records -- the full grammar takes over for those and hands back
    with open(tmp_path / "mixed.fq", "w", newline="") as fm:
        for i, ln in enumerate(lens):
            s_ = "".join(rng.choice(list("ACGTN"), ln))
            q_ = "".join(rng.choice(list("@+IIIIFF#5"), ln))
            if i > 2000 and i % 1000 == 0 and ln > 3:
                fm.write(f"@m{i}\n{s_[:2]}\n{s_[2:]}\n+\n{q_[:1]}\n{q_[1:]}\n")
            elif i > 2000 and i % 777 == 0:
                fm.write(f"@m{i}\r\n{s_}\r\n+\r\n{q_}\r\n")
            elif i > 2000 and i % 1501 == 0:
                fm.write(f"\n\n@m{i}\n{s_}\n+m{i}\n{q_}\n")
            else:
                fm.write(f"@m{i}\n{s_}\n+\n{q_}\n")
    one, u1 = _parse_check(tmp_path / "mixed.fq", None, 1)
    assert u1 == 1 and one[0] == n and one[2] == int(lens.sum())
    for threads in (2, 4):
        got, used = _parse_check(tmp_path / "mixed.fq", None, threads)
        assert got == one, (threads, used, got, one)
    # tiny batches: many hand-offs per slice
    got, used = _parse_check(tmp_path / "a.fq", None, 4, batch_bytes=2048)
    assert got == ref
    # the golden edge-case files (too small to be cut): unchanged
    for f in ("edge_1.fq", "edge_2.fq", "edge_1.fq.gz", "contigs.fa"):
        one, u1 = _parse_check(util.GOLDEN / f, None, 1)
        four, u4 = _parse_check(util.GOLDEN / f, None, 4)
        assert one == four and u1 == u4 == 1


def _bgzf(data: bytes, block: int = 65280, eof_marker: bool = True) -> bytes:
    """BGZF (SAM spec 4.1): gzip members of <= 64 KB with a 'BC' extra field holding the member's size - 1."""
    import struct
    import zlib
    out = bytearray()
    for i in list(range(0, len(data), block)) + ([None] if eof_marker else []):
        chunk = b"" if i is None else data[i: i + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        c = co.compress(chunk) + co.flush()
        out += (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(c) + 25) + c +
                struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return bytes(out)


def test_bgzf_input_is_inflated_block_parallel(tmp_path, monkeypatch):
    """Blocked gzip (bgzip / htslib) is inflated a window of blocks at a time by several threads (seq_reader.cpp:
    BgzfStream); ordinary gzip keeps zlib's single stream.  Same reads, bases and checksum as the plain file for any
    number of threads, for odd block sizes, and through zlib (LASV_B200_GZ_THREADS=0); a corrupt or truncated block is
    an error, never a short file."""
    rng = np.random.default_rng(5)
    n = 40_000
    lens = rng.integers(1, 200, n)
    with open(tmp_path / "a.fq", "w") as fa:
        for i, ln in enumerate(lens):
            s = "".join(rng.choice(list("ACGTN"), ln))
            q = "".join(rng.choice(list("@+IIIIIIFF#5&'"), ln))
            fa.write(f"@r{i}\n{s}\n+\n{q}\n")
    text = (tmp_path / "a.fq").read_bytes()
    ref, _ = _parse_check(tmp_path / "a.fq", None, 1)
    assert ref[0] == n
    (tmp_path / "a.fq.gz").write_bytes(_bgzf(text))
    (tmp_path / "odd.fq.gz").write_bytes(_bgzf(text, block=7919, eof_marker=False))
    for thr in ("8", "3", "1", "0"):
        monkeypatch.setenv("LASV_B200_GZ_THREADS", thr)
        for f in ("a.fq.gz", "odd.fq.gz"):
            got, used = _parse_check(tmp_path / f, None, 4)
            assert used == 1 and got == ref, (thr, f, got, ref)
    monkeypatch.setenv("LASV_B200_GZ_THREADS", "4")
    import ctypes as C
    lib = _capi.lib()
    blob = bytearray(_bgzf(text))
    blob[len(blob) // 2] ^= 0x55                                  # a flipped byte in the middle of a block
    (tmp_path / "bad.fq.gz").write_bytes(bytes(blob))
    out = (C.c_uint64 * 4)()
    used = C.c_int(0)
    assert lib.lasv_seqfile_parse_check(str(tmp_path / "bad.fq.gz").encode(), None, 1, 1 << 20, out, C.byref(used)) < 0
    assert b"BGZF" in lib.lasv_last_error()
    (tmp_path / "cut.fq.gz").write_bytes(_bgzf(text)[: len(blob) // 2])   # the file ends inside a block
    assert lib.lasv_seqfile_parse_check(str(tmp_path / "cut.fq.gz").encode(), None, 1, 1 << 20, out, C.byref(used)) < 0
    # nothing but the end-of-file block: an empty file, not an error
    (tmp_path / "empty.fq.gz").write_bytes(_bgzf(b""))
    got, _ = _parse_check(tmp_path / "empty.fq.gz", None, 1)
    assert got[0] == 0 and got[2] == 0
    # another extra subfield in front of 'BC' (the gzip header allows any number of them)
    import struct
    import zlib

    def member(chunk):
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        c = co.compress(chunk) + co.flush()
        xtra = b"XY\x03\x00abc" + b"BC\x02\x00"
        xlen = len(xtra) + 2
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", xlen) + xtra +
                struct.pack("<H", 12 + xlen + len(c) + 8 - 1) + c + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    small = text[:60000]
    (tmp_path / "small.fq").write_bytes(small[: small.rfind(b"\n@") + 1])
    body = (tmp_path / "small.fq").read_bytes()
    (tmp_path / "xtra.fq.gz").write_bytes(member(body[:30000]) + member(body[30000:]) + member(b""))
    assert _parse_check(tmp_path / "xtra.fq.gz", None, 1)[0] == _parse_check(tmp_path / "small.fq", None, 1)[0]


def _fastq_text(n, seed, max_len=200):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len, n)
    out = []
    for i, ln in enumerate(lens):
        s = "".join(rng.choice(list("ACGTN"), ln))
        q = "".join(rng.choice(list("@+IIIIIIFF#5&'"), ln))
        out.append(f"@r{i} {i}/1\n{s}\n+\n{q}\n")
    return "".join(out).encode()


def _gz(data, level=6, flush_every=0, flush_mode=None, strategy=None):
    import zlib
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 8, zlib.Z_DEFAULT_STRATEGY if strategy is None else strategy)
    out = []
    if flush_every:
        for i in range(0, len(data), flush_every):
            out.append(co.compress(data[i:i + flush_every]))
            out.append(co.flush(flush_mode))
    else:
        out.append(co.compress(data))
    out.append(co.flush())
    return b"".join(out)


def _inflate_check(path, threads, chunk_kb):
    import ctypes as C
    out = (C.c_uint64 * 4)()
    rc = _capi.lib().lasv_seqfile_inflate_check(str(path).encode(), threads, chunk_kb, out)
    return rc, tuple(int(x) for x in out)


def test_ordinary_gzip_is_inflated_by_several_threads(tmp_path):
    """Single-stream gzip (gz_parallel.cpp: PgzStream): threads that start in the middle of the deflate stream find
    a block boundary, inflate into symbols without the 32 KB before them, and are resolved once the chunk in front
    has met their start.  The text is byte for byte zlib's -- for any number of threads and any chunk size, across
    compression levels and strategies, flush points (stored empty blocks), fixed-code and stored blocks, several
    members, trailing garbage, and data that is not text (nothing is confirmed there: one thread carries on)."""
    import os
    import zlib
    text = _fastq_text(60_000, 21)
    cases = {
        "l1": _gz(text, 1), "l6": _gz(text, 6), "l9": _gz(text, 9),
        "sync": _gz(text, 6, 70_000, zlib.Z_SYNC_FLUSH), "full": _gz(text, 6, 50_001, zlib.Z_FULL_FLUSH),
        "fixed": _gz(text[:2_000_000], 6, strategy=zlib.Z_FIXED), "huffman": _gz(text[:2_000_000], 6, strategy=zlib.Z_HUFFMAN_ONLY),
        "stored": _gz(text[:1_000_000], 0),
        "members": _gz(text[:3_000_001], 6) + _gz(text[3_000_001:5_000_000], 2) + _gz(b"") + _gz(text[5_000_000:], 9),
        "tail": _gz(text, 6) + b"\0" * 777,
        "binary": _gz(os.urandom(700_000) + text[:2_000_000] + os.urandom(300_000), 6),
        "zeros": _gz(bytes(40_000_000), 6),
    }
    plain = {"fixed": text[:2_000_000], "huffman": text[:2_000_000], "stored": text[:1_000_000], "zeros": bytes(40_000_000)}
    for name, blob in cases.items():
        (tmp_path / f"{name}.gz").write_bytes(blob)
        want = plain.get(name)
        if want is None:
            want = zlib.decompressobj(31)
            # (all members, as gzread does; the trailing zeros are not one)
            data, rest = b"", blob
            while rest[:2] == b"\x1f\x8b":
                d = zlib.decompressobj(31)
                data += d.decompress(rest)
                rest = d.unused_data
            want = data
        for threads, chunk_kb in ((4, 256), (8, 64), (3, 1024), (1, 512), (5, 4), (16, 32)):
            rc, got = _inflate_check(tmp_path / f"{name}.gz", threads, chunk_kb)
            assert rc == 0, (name, threads, chunk_kb, _capi.lib().lasv_last_error())
            assert got[0] == len(want) and got[1] == zlib.crc32(want), (name, threads, chunk_kb, got)
            if name in ("l1", "l6", "l9", "sync", "full", "members", "tail") and (threads, chunk_kb) in ((4, 256), (8, 64)):
                assert got[2] >= 4, (name, threads, chunk_kb, got)      # the chunks really started in mid-stream
    # a broken file is an error, never a short or a wrong text
    good = cases["l6"]
    bad = {}
    b = bytearray(good); b[len(b) // 2] ^= 0x10; bad["flip"] = bytes(b)
    b = bytearray(good); b[-6] ^= 1; bad["crc"] = bytes(b)
    b = bytearray(good); b[-2] ^= 1; bad["isize"] = bytes(b)
    bad["cut"] = good[: len(good) // 2]
    bad["cut_trailer"] = good[:-3]
    bad["second_header"] = good + b"\x1f\x8b\x07" + b"\0" * 20
    for name, blob in bad.items():
        (tmp_path / f"bad_{name}.gz").write_bytes(blob)
        for threads, chunk_kb in ((4, 256), (1, 512), (8, 16)):
            rc, _ = _inflate_check(tmp_path / f"bad_{name}.gz", threads, chunk_kb)
            assert rc < 0 and b"gzip" in _capi.lib().lasv_last_error(), (name, threads, chunk_kb)


def test_file_loaders_read_ordinary_gzip_through_the_parallel_inflate(tmp_path, monkeypatch):
    """The reader behind the file loaders takes .gz input of some size through PgzStream (LASV_B200_PGZ_MIN_KB,
    default 8 MB; LASV_B200_PGZ_CHUNK_KB) and smaller files, or everything with LASV_B200_GZ_THREADS=0, through
    zlib: the same reads, bases and checksum as the plain file either way, and a truncated file is an error."""
    text = _fastq_text(40_000, 22)
    (tmp_path / "a.fq").write_bytes(text)
    (tmp_path / "a.fq.gz").write_bytes(_gz(text, 6))
    ref, _ = _parse_check(tmp_path / "a.fq", None, 1)
    assert ref[0] == 40_000
    for env in ({"LASV_B200_PGZ_MIN_KB": "0", "LASV_B200_PGZ_CHUNK_KB": "128", "LASV_B200_GZ_THREADS": "4"},
                {"LASV_B200_PGZ_MIN_KB": "0", "LASV_B200_PGZ_CHUNK_KB": "16", "LASV_B200_GZ_THREADS": "7"},
                {"LASV_B200_PGZ_MIN_KB": "0", "LASV_B200_GZ_THREADS": "1"},
                {"LASV_B200_GZ_THREADS": "0"}, {}):
        for k in ("LASV_B200_PGZ_MIN_KB", "LASV_B200_PGZ_CHUNK_KB", "LASV_B200_GZ_THREADS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got, used = _parse_check(tmp_path / "a.fq.gz", None, 4)
        assert used == 1 and got == ref, (env, got, ref)
    monkeypatch.setenv("LASV_B200_PGZ_MIN_KB", "0")
    monkeypatch.setenv("LASV_B200_GZ_THREADS", "4")
    # FASTA (multi-line records) and one-read-per-line text take the same way: the type is told from the first window
    rng = np.random.default_rng(24)
    fa = "".join(f">c{i} len\n" + "\n".join("".join(rng.choice(list("ACGTNacgt"), 60)) for _ in range(int(rng.integers(1, 40)))) + "\n"
                 for i in range(600)).encode()
    (tmp_path / "c.fa").write_bytes(fa)
    (tmp_path / "c.fa.gz").write_bytes(_gz(fa, 6))
    plain_reads = b"".join(bytes(rng.choice(list(b"ACGT"), int(rng.integers(0, 90))).astype(np.uint8)) + b"\n" for _ in range(20_000))
    (tmp_path / "p.txt").write_bytes(plain_reads)
    (tmp_path / "p.txt.gz").write_bytes(_gz(plain_reads, 6))
    monkeypatch.setenv("LASV_B200_PGZ_CHUNK_KB", "64")
    for a, b in (("c.fa", "c.fa.gz"), ("p.txt", "p.txt.gz")):
        assert _parse_check(tmp_path / b, None, 1)[0] == _parse_check(tmp_path / a, None, 1)[0], a
    blob = (tmp_path / "a.fq.gz").read_bytes()
    (tmp_path / "cut.fq.gz").write_bytes(blob[: len(blob) // 2])
    import ctypes as C
    out = (C.c_uint64 * 4)()
    used = C.c_int(0)
    lib = _capi.lib()
    assert lib.lasv_seqfile_parse_check(str(tmp_path / "cut.fq.gz").encode(), None, 1, 1 << 20, out, C.byref(used)) < 0
    assert b"gzip" in lib.lasv_last_error()


def test_text_batches_for_the_device_parse_hold_whole_records(tmp_path, monkeypatch):
    """The producers of the device parse (capi.cu: FileProducer in raw mode) hand over batches of FASTQ text: slices of
    a mapped plain file, windows of BGZF or gzip text cut at the last record boundary with the tail carried into the
    next batch.  Parsed on the host here (lasv_seqfile_parse_check_text): the same reads, bases and checksum as the
    sequential parse of the plain file, and a multi-line record late in a compressed file hands the rest of it to
    the host parser without losing or repeating a read."""
    import ctypes as C
    text = _fastq_text(260_000, 23, max_len=160)                 # 44 MB of text: two batches of 24 MB
    (tmp_path / "a.fq").write_bytes(text)
    (tmp_path / "a.bgzf.fq.gz").write_bytes(_bgzf(text))
    (tmp_path / "a.gzip.fq.gz").write_bytes(_gz(text, 1))
    lines = text.split(b"\n")
    i = 4 * 200_000
    lines[i + 1: i + 2] = [lines[i + 1][:1], lines[i + 1][1:]] if len(lines[i + 1]) > 1 else [lines[i + 1], b""]
    wrapped = b"\n".join(lines)
    (tmp_path / "w.gzip.fq.gz").write_bytes(_gz(wrapped, 1))
    ref, _ = _parse_check(tmp_path / "a.fq", None, 1)
    assert ref[0] == 260_000
    monkeypatch.setenv("LASV_B200_PGZ_MIN_KB", "0")
    monkeypatch.setenv("LASV_B200_PGZ_CHUNK_KB", "512")
    monkeypatch.setenv("LASV_B200_GZ_THREADS", "4")
    lib = _capi.lib()

    def text_check(path, threads):
        out = (C.c_uint64 * 4)()
        used = C.c_int(0)
        raw = C.c_uint64(0)
        rc = lib.lasv_seqfile_parse_check_text(str(path).encode(), None, threads, 24 << 20, out, C.byref(used), C.byref(raw))
        assert rc == 0, lib.lasv_last_error()
        return tuple(int(x) for x in out), raw.value
    for f, min_raw in (("a.fq", 2), ("a.bgzf.fq.gz", 2), ("a.gzip.fq.gz", 2), ("w.gzip.fq.gz", 1)):
        got, raw = text_check(tmp_path / f, 2)
        assert got == ref and raw >= min_raw, (f, got, ref, raw)
    # two mates through the same path
    got, raw = text_check(tmp_path / "a.gzip.fq.gz", 2)
    out = (C.c_uint64 * 4)()
    used = C.c_int(0)
    assert lib.lasv_seqfile_parse_check_text(str(tmp_path / "a.gzip.fq.gz").encode(), str(tmp_path / "a.bgzf.fq.gz").encode(),
                                             2, 24 << 20, out, C.byref(used), None) == 0
    assert out[0] == out[1] == ref[0] and out[2] == 2 * ref[2]


def test_bench_job_geometry_is_the_same_read_set_at_every_n():
    """bench.py: the cfg-4 job is the same set of batches at N = 1, 2, 4, 8 (a multiple of 8 batches), whatever K is;
    a rank's batches of step s are consecutive, and all ranks together cover every batch exactly once."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = synth.WORKLOADS["cfg4"]
    totals = set()
    for steps in (3, 7, 20):
        for world in (1, 2, 4, 8):
            args = argparse.Namespace(steps=steps, fraction=1.0, pairs_per_batch=1 << 20)
            geo = bench.job_geometry(args, w, world)
            totals.add(geo["batches"])
            assert geo["batches"] % 8 == 0 and geo["per_rank"] * world == geo["batches"]
            assert sum(geo["step_counts"]) == geo["per_rank"] and len(geo["step_counts"]) == steps
            assert max(geo["step_counts"]) - min(geo["step_counts"]) <= 1
            seen = set()
            for rank in range(world):
                for s_ in range(steps):
                    for i in range(geo["step_counts"][s_]):
                        seen.add((geo["step_prefix"][s_] + i) * world + rank)
            assert seen == set(range(geo["batches"]))
    assert totals == {296}
