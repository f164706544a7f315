#!/usr/bin/env python
"""End-to-end speed of the file loaders (FASTQ text -> graph): the parse is the slow end."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from lasv_b200 import DBGraph, synth
from oracle import oracle as O   # FASTQ writer only (test infrastructure)

w = synth.CFG4
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n_pairs = int(_pos[0]) if len(_pos) > 0 else 1_000_000
L = int(_pos[1]) if len(_pos) > 1 else 22
tmp = Path(os.environ.get("TMPDIR", "/tmp"))
seq, qual, _ = synth.reads_host(w.params(), 0, 2 * n_pairs)
s2, q2 = seq.reshape(-1, w.read_len + 1), qual.reshape(-1, w.read_len + 1)
O.write_fastq_fixed(tmp / "ls_1.fq", s2[0::2].reshape(-1), q2[0::2].reshape(-1), w.read_len)
O.write_fastq_fixed(tmp / "ls_2.fq", s2[1::2].reshape(-1), q2[1::2].reshape(-1), w.read_len)
f1, f2 = tmp / "ls_1.fq", tmp / "ls_2.fq"
if "--bgzf" in sys.argv:   # blocked gzip (bgzip / htslib): inflated block-parallel, text taken apart on the device
    import struct, zlib
    def bgzf(data, block=65280):
        out = bytearray()
        for i in list(range(0, len(data), block)) + [None]:
            chunk = b"" if i is None else data[i: i + block]
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            c = co.compress(chunk) + co.flush()
            out += (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(c) + 25) + c +
                    struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        return bytes(out)
    for f in (f1, f2):
        f.with_suffix(".fq.gz").write_bytes(bgzf(f.read_bytes()))
    text_size = os.path.getsize(f1) + os.path.getsize(f2)
    f1, f2 = f1.with_suffix(".fq.gz"), f2.with_suffix(".fq.gz")
if "--gzip" in sys.argv:   # ordinary single-stream gzip (gzip -1): chunks of the stream inflated by several threads
    import subprocess
    procs = [subprocess.Popen(["gzip", "-1", "-k", "-f", str(f)]) for f in (f1, f2)]
    assert all(p.wait() == 0 for p in procs)
    text_size = os.path.getsize(f1) + os.path.getsize(f2)
    f1, f2 = Path(str(f1) + ".gz"), Path(str(f2) + ".gz")
size = os.path.getsize(f1) + os.path.getsize(f2)
compressed = "--bgzf" in sys.argv or "--gzip" in sys.argv
with DBGraph(w.kmer_size, L, 100) as g:
    for rep in range(2):
        g.clear()
        t = time.perf_counter()
        st = g.load_pe_seq_data(f1, f2, w.cutoff)
        dt = time.perf_counter() - t
    print(json.dumps({"what": "lasv_graph_load_pe_files, uncompressed FASTQ pair", "pairs": n_pairs, "fastq_bytes": size,
                      "seconds": dt, "fastq_MBps": size / dt / 1e6, "kmers_per_s": st["kmers_inserted"] / dt,
                      "bgzf": "--bgzf" in sys.argv, "gzip": "--gzip" in sys.argv,
                      "text_MBps": (text_size if compressed else size) / dt / 1e6,
                      "device_parse": int(os.environ.get("LASV_B200_DEVICE_PARSE", "1")),
                      "gz_threads": os.environ.get("LASV_B200_GZ_THREADS", "default"),
                      "pgz_chunk_kb": os.environ.get("LASV_B200_PGZ_CHUNK_KB", "default"), "host_cores": os.cpu_count()}))
