"""Host-side mirror of the reference's hash_table / element / file_reader surface
for the graph-build path, over the C ABI (include/lasv_b200.h).

Names follow the reference (paths under /root/reference):
  DBGraph(...)                         hash_table_new            hash_table.c:13
  .free()                              hash_table_free           hash_table.c:55
  .unique_kmers / .capacity            hash_table_get_*          hash_table.c:397-408
  .find(keys)                          hash_table_find           hash_table.c:234
  .load_se_filelist / .load_pe_filelists   file_reader.c:789,910
  .load_se_seq_data / .load_pe_seq_data    file_reader.c:475,589
  .dump_binary(path)                   db_graph_dump_single_colour_binary_of_colour0
                                                                 dB_graph_population.c:3730
  .load_binary(path)                   load_multicolour_binary_from_filename_into_graph
                                                                 file_reader.c:1652
All compute happens in the CUDA library; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import LoadStats, TableSummary, check

QUALITY_OFFSET_DEFAULT = 33  # cmd_line.c:363


def words_per_kmer(k: int) -> int:
    return (2 * k) // 64 + 1  # graph_operations.c:620


def record_dtype(N: int) -> np.dtype:
    return np.dtype([("kmer", "<u8", (N,)), ("covg", "<u4"), ("edges", "u1")], align=False)


def element_dtype(N: int) -> np.dtype:
    """Reference Element (element.h:77-82), sizeof = 8N+8."""
    return np.dtype({"names": ["kmer", "coverage", "edges", "status"],
                     "formats": [("<u8", (N,)), "<u4", "u1", "u1"],
                     "offsets": [0, 8 * N, 8 * N + 4, 8 * N + 5], "itemsize": 8 * N + 8})


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (device or pinned host memory)
        return a.data_ptr()
    return a


class DBGraph:
    """A de Bruijn graph whose hash table lives in the HBM of one GPU."""

    def __init__(self, kmer_size: int, number_bits: int, bucket_size: int,
                 max_rehash_tries: int = 15, device: int = 0):
        self._lib = _capi.lib()
        h = self._lib.lasv_graph_new(kmer_size, number_bits, bucket_size, max_rehash_tries, device)
        if not h:
            raise _capi.LasvError(-2, self._lib.lasv_last_error().decode(errors="replace"))
        self._h = C.c_void_p(h)
        self.kmer_size, self.number_bits, self.bucket_size = kmer_size, number_bits, bucket_size
        self.N = words_per_kmer(kmer_size)
        self.device = device

    # -- lifetime
    def free(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.lasv_graph_free(C.byref(self._h))

    __del__ = free

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    def clear(self, stream=None):
        check(self._lib.lasv_graph_clear(self._h, stream))

    def set_pipeline(self, on: bool = True):
        """Insert batch i on an internal stream while batch i+1 is partitioned (large tables)."""
        check(self._lib.lasv_graph_set_pipeline(self._h, int(on)))

    def enable_kernel_timing(self, on: bool = True):
        check(self._lib.lasv_graph_enable_kernel_timing(self._h, int(on)))

    def kernel_times(self) -> dict:
        """Summed CUDA-event durations of the partition / insert kernels since enable_kernel_timing."""
        a, b, n = C.c_double(0), C.c_double(0), C.c_uint64(0)
        check(self._lib.lasv_graph_kernel_times(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return {"partition_ms": a.value, "insert_ms": b.value, "batches": int(n.value)}

    def kernel_times_by_kind(self) -> dict:
        """{kind: (ms, spans)} for partition / insert / sort / stream / drain (LASV_T_*)."""
        ms = (C.c_double * len(_capi.T_KINDS))()
        n = (C.c_uint64 * len(_capi.T_KINDS))()
        check(self._lib.lasv_graph_kernel_times_by_kind(self._h, ms, n))
        return {k: (ms[i], int(n[i])) for i, k in enumerate(_capi.T_KINDS)}

    def set_insert_mode(self, mode: int, min_records_per_line: float = 0.0):
        """_capi.INSERT_AUTO / INSERT_RANDOM / INSERT_STREAMED (streamed_insert.cuh)."""
        check(self._lib.lasv_graph_set_insert_mode(self._h, int(mode), float(min_records_per_line)))

    def insert_counts(self) -> dict:
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(self._lib.lasv_graph_insert_counts(self._h, C.byref(a), C.byref(b)))
        return {"streamed": int(a.value), "random": int(b.value)}

    def set_accumulation(self, max_stream_bytes: int) -> int:
        """load_reads_device only partitions; one insert per `max_stream_bytes` of batches (or flush()).
        Returns the limit granted (free device memory may hold less)."""
        got = C.c_uint64(0)
        check(self._lib.lasv_graph_set_accumulation(self._h, int(max_stream_bytes), C.byref(got)))
        return int(got.value)

    def flush(self, stream=None):
        """`stream` waits for every insert queued so far (accumulated stripes are inserted first)."""
        check(self._lib.lasv_graph_flush(self._h, stream))

    # -- properties
    @property
    def unique_kmers(self) -> int:
        return int(self._lib.lasv_graph_unique_kmers(self._h))

    @property
    def capacity(self) -> int:
        return int(self._lib.lasv_graph_capacity(self._h))

    @property
    def table_bytes(self) -> int:
        """Bytes of the device table (lines of 128 bytes until finalize)."""
        return int(self._lib.lasv_graph_device_table_bytes(self._h))

    @property
    def element_bytes(self) -> int:
        return int(self._lib.lasv_graph_element_bytes(self._h))

    # -- the hot path
    def load_reads_device(self, d_seq, d_qual, n_bytes: int, d_offsets=None, n_reads: int = 0,
                          quality_cutoff: int = QUALITY_OFFSET_DEFAULT, stream=None):
        check(self._lib.lasv_graph_load_reads_device(self._h, _ptr(d_seq), _ptr(d_qual), n_bytes,
                                                     _ptr(d_offsets), n_reads, quality_cutoff, stream))

    def load_reads_host(self, seq, qual, offsets, quality_cutoff: int = QUALITY_OFFSET_DEFAULT) -> dict:
        n_reads = len(offsets) - 1
        n_bytes = int(offsets[n_reads])
        st = LoadStats()
        check(self._lib.lasv_graph_load_reads_host(self._h, _ptr(seq), _ptr(qual), n_bytes,
                                                   _ptr(offsets), n_reads, quality_cutoff, C.byref(st)))
        return st.as_dict()

    def load_reads_host_async(self, seq, qual, offsets, quality_cutoff: int = QUALITY_OFFSET_DEFAULT, pinned_stats=None):
        """Returns once the copies have left `seq` / `qual`; kernels (and inserts of full stripes) still run.
        `pinned_stats`: pinned uint64[9] that receives the statistics behind this call's kernels."""
        n_reads = len(offsets) - 1
        n_bytes = int(offsets[n_reads])
        check(self._lib.lasv_graph_load_reads_host_async(self._h, _ptr(seq), _ptr(qual), n_bytes, _ptr(offsets),
                                                         n_reads, quality_cutoff, _ptr(pinned_stats)))

    def set_staging(self, n_bytes: int) -> int:
        """Device staging ring of the host-buffer calls (before set_accumulation).  Returns the bytes granted."""
        got = C.c_uint64(0)
        check(self._lib.lasv_graph_set_staging(self._h, int(n_bytes), C.byref(got)))
        return int(got.value)

    def stats(self, hist_size: int = 0, stream=None):
        st = LoadStats()
        hist = np.zeros(hist_size, dtype=np.uint64) if hist_size else None
        check(self._lib.lasv_graph_stats(self._h, C.byref(st), _ptr(hist), hist_size, stream))
        return (st.as_dict(), hist) if hist_size else st.as_dict()

    STATS_ASYNC_FIELDS = ("n_reads", "bad_reads", "bases_read", "bases_loaded", "n_contigs", "unique_kmers",
                          "kmers_inserted", "table_full", "exchange_overflow")

    def stats_async(self, pinned_out, stream=None):
        """Queue a copy of the statistics (9 x uint64, STATS_ASYNC_FIELDS) into pinned host memory."""
        check(self._lib.lasv_graph_stats_async(self._h, _ptr(pinned_out), stream))

    def rehash_histogram(self) -> np.ndarray:
        out = np.zeros(16, dtype=np.int64)
        check(self._lib.lasv_graph_rehash_histogram(self._h, out.ctypes.data))
        return out

    # -- sharded pieces
    def route_reads_device(self, d_seq, d_qual, n_bytes, quality_cutoff, n_dest, d_bins, bin_capacity,
                           d_bin_counts, stream=None):
        check(self._lib.lasv_graph_route_reads_device(self._h, _ptr(d_seq), _ptr(d_qual), n_bytes,
                                                      quality_cutoff, n_dest, _ptr(d_bins), bin_capacity,
                                                      _ptr(d_bin_counts), stream))

    def extract_records_device(self, d_seq, d_qual, n_bytes, quality_cutoff, d_records, capacity,
                               d_count, stream=None):
        check(self._lib.lasv_graph_extract_records_device(self._h, _ptr(d_seq), _ptr(d_qual), n_bytes,
                                                          quality_cutoff, _ptr(d_records), capacity,
                                                          _ptr(d_count), stream))

    def insert_records_device(self, d_records, n_records, stream=None):
        check(self._lib.lasv_graph_insert_records_device(self._h, _ptr(d_records), n_records, stream))

    def group_records_device(self, d_records, n_records, buckets_per_group: int, d_sorted, d_group_offsets, stream=None):
        """Counting sort of records by group (first-choice bucket // buckets_per_group) for insert_grouped_records_device."""
        check(self._lib.lasv_graph_group_records_device(self._h, _ptr(d_records), n_records, buckets_per_group,
                                                        _ptr(d_sorted), _ptr(d_group_offsets), stream))

    def insert_grouped_records_device(self, d_sorted_records, n_records, d_group_offsets, buckets_per_group: int,
                                      stream=None) -> int:
        """PROTOTYPE of the streamed insert: records sorted by group, offsets per group; returns
        the number of records whose bucket was full (they went through the ordinary path)."""
        spilled = C.c_uint64(0)
        check(self._lib.lasv_graph_insert_grouped_records_device(self._h, _ptr(d_sorted_records), n_records,
                                                                  _ptr(d_group_offsets), buckets_per_group,
                                                                  C.byref(spilled), stream))
        return int(spilled.value)

    def bin_geometry(self, n_owners: int, n_bytes: int, n_reads: int = 0) -> dict:
        """Buffer sizes of the region-binned exchange for one batch (include/lasv_b200.h)."""
        a, b, sp, c, d = (C.c_uint32(0) for _ in range(5))
        check(self._lib.lasv_graph_bin_geometry(self._h, n_owners, n_bytes, n_reads, C.byref(a), C.byref(b),
                                                C.byref(sp), C.byref(c), C.byref(d)))
        geo = {"bins_per_owner": int(a.value), "stripe_capacity": int(b.value), "spill_capacity": int(sp.value),
               "record_bytes": int(c.value), "count_stride_bytes": int(d.value)}
        geo["records_per_owner"] = geo["bins_per_owner"] * geo["stripe_capacity"] + geo["spill_capacity"]
        geo["counters_per_owner"] = geo["bins_per_owner"] + 1
        return geo

    def bin_reads_device(self, d_seq, d_qual, n_bytes, quality_cutoff, n_owners, bins_per_owner,
                         stripe_capacity, spill_capacity, d_records, d_counts, stream=None, append=False):
        """append: the stripes keep what earlier batches put there (lasv_graph_bin_reads_append_device)."""
        fn = self._lib.lasv_graph_bin_reads_append_device if append else self._lib.lasv_graph_bin_reads_device
        check(fn(self._h, _ptr(d_seq), _ptr(d_qual), n_bytes, quality_cutoff,
                 n_owners, bins_per_owner, stripe_capacity, spill_capacity, _ptr(d_records), _ptr(d_counts), stream))

    def bin_reads_peers_device(self, d_seq, d_qual, n_bytes, quality_cutoff, n_owners, bins_per_owner,
                               stripe_capacity, spill_capacity, owner_blocks, d_counts, stream=None, append=False):
        """Fused producer -> exchange: records go straight into owner_blocks[o] (this sender's block in owner o's
        receive stripes, peer memory mapped into this process)."""
        ptrs = (C.c_void_p * n_owners)(*[_ptr(b) for b in owner_blocks])
        check(self._lib.lasv_graph_bin_reads_peers_device(self._h, _ptr(d_seq), _ptr(d_qual), n_bytes, quality_cutoff,
                                                          n_owners, bins_per_owner, stripe_capacity, spill_capacity, ptrs,
                                                          _ptr(d_counts), stream, 1 if append else 0))

    def insert_bins_device(self, d_records, d_counts, n_src, bins_per_src, stripe_capacity, spill_capacity,
                           stream=None):
        check(self._lib.lasv_graph_insert_bins_device(self._h, _ptr(d_records), _ptr(d_counts), n_src,
                                                      bins_per_src, stripe_capacity, spill_capacity, stream))

    def read_stats_device(self, d_seq, d_qual, d_offsets, n_reads, quality_cutoff, stream=None):
        check(self._lib.lasv_graph_read_stats_device(self._h, _ptr(d_seq), _ptr(d_qual), _ptr(d_offsets),
                                                     n_reads, quality_cutoff, stream))

    # -- queries
    def find(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1, self.N)
        n = len(keys)
        covg = np.zeros(n, dtype=np.uint32)
        edges = np.zeros(n, dtype=np.uint8)
        found = np.zeros(n, dtype=np.uint8)
        check(self._lib.lasv_graph_find(self._h, keys.ctypes.data, n, covg.ctypes.data,
                                        edges.ctypes.data, found.ctypes.data))
        return found.astype(bool), covg, edges

    def summary(self) -> dict:
        s = TableSummary()
        check(self._lib.lasv_graph_summary(self._h, C.byref(s)))
        return s.as_dict()

    # -- output
    def finalize(self):
        check(self._lib.lasv_graph_finalize(self._h))

    def export_table(self):
        """(Element[capacity], next_element[buckets]) in the reference's host layout."""
        table = np.zeros(self.capacity, dtype=element_dtype(self.N))
        nxt = np.zeros(1 << self.number_bits, dtype=np.int16)
        check(self._lib.lasv_graph_export_table(self._h, table.ctypes.data, nxt.ctypes.data))
        return table, nxt

    def import_table(self, table: np.ndarray):
        table = np.ascontiguousarray(table)
        assert table.nbytes == self.capacity * self.element_bytes
        check(self._lib.lasv_graph_import_table(self._h, table.ctypes.data))

    def records(self) -> np.ndarray:
        """`.ctx` records in table order."""
        n = self.unique_kmers
        out = np.zeros(max(n, 1), dtype=record_dtype(self.N))
        got = check(self._lib.lasv_graph_dump_ctx_records(self._h, out.ctypes.data, len(out)))
        return out[:got]

    def dump_binary(self, path, mean_read_len: int = 0, total_seq: int = 0):
        check(self._lib.lasv_graph_write_ctx(self._h, str(path).encode(), mean_read_len, total_seq))

    def append_binary_records(self, path) -> int:
        """The `.ctx` records of this table, appended to `path` (one shard of a sharded dump)."""
        return check(self._lib.lasv_graph_append_ctx_records(self._h, str(path).encode()))

    def output_supernodes(self, path, max_length: int = 10000, strict: bool = False) -> dict:
        """`--output_supernodes path` (graph_operations.c:1276-1293): the reference's FASTA of
        maximal unambiguous contigs for a traversal in table (= dumped .ctx) order."""
        st = _capi.SupernodeStats()
        fn = self._lib.lasv_graph_write_supernodes_strict if strict else self._lib.lasv_graph_write_supernodes
        check(fn(self._h, str(path).encode(), max_length, C.byref(st)))
        return st.as_dict()

    def health_check(self) -> dict:
        """`--health_check` (dB_graph_population.c:6274-6350): every edge leads to a node of the graph
        that holds the return arrow; problems are counted."""
        st = _capi.HealthStats()
        check(self._lib.lasv_graph_health_check(self._h, C.byref(st)))
        return st.as_dict()

    def clean(self, threshold: int, what: int = _capi.CLEAN_TIPS | _capi.CLEAN_SUPERNODES, max_length: int = 0) -> dict:
        """`--remove_low_coverage_supernodes threshold` (graph_operations.c:1069-1090) on the device table: tip
        clipping, then low-coverage supernode removal, exact for the table's slot order.  max_length > 0: the
        reference's walk limit (--max_var_len); LasvError(LASV_ERR_LIMIT), graph unchanged by that pass, where the
        reference would judge a piece of a longer supernode."""
        st = _capi.CleanStats()
        check(self._lib.lasv_graph_clean_limited(self._h, what, threshold, max_length, C.byref(st)))
        return st.as_dict()

    def output_branches(self, path, mark_visited: bool = False) -> dict:
        """`--mode build`'s ./branches.fa (graph_operations.c:1145-1152, print_branches.c:131-199) for a
        traversal in table order; mark_visited leaves the reference's `visited` marks on the device table."""
        st = _capi.BranchStats()
        check(self._lib.lasv_graph_write_branches(self._h, str(path).encode(), 1 if mark_visited else 0, C.byref(st)))
        return st.as_dict()

    def load_binary(self, path):
        mrl, ts = C.c_uint32(), C.c_uint64()
        check(self._lib.lasv_graph_load_ctx_file(self._h, str(path).encode(), C.byref(mrl), C.byref(ts)))
        return int(mrl.value), int(ts.value)

    def load_records(self, recs: np.ndarray):
        recs = np.ascontiguousarray(recs)
        check(self._lib.lasv_graph_load_ctx_records(self._h, recs.ctypes.data, len(recs)))

    # -- sequence files (file_reader.c)
    def load_se_seq_data(self, path, quality_cutoff: int) -> dict:
        st = LoadStats()
        check(self._lib.lasv_graph_load_se_file(self._h, str(path).encode(), quality_cutoff, C.byref(st)))
        return st.as_dict()

    def load_pe_seq_data(self, path1, path2, quality_cutoff: int) -> dict:
        st = LoadStats()
        check(self._lib.lasv_graph_load_pe_files(self._h, str(path1).encode(), str(path2).encode(),
                                                 quality_cutoff, C.byref(st)))
        return st.as_dict()

    def load_se_filelist(self, filelist, qual_thresh: int = 0, ascii_fq_offset: int = 33) -> dict:
        st, n = LoadStats(), C.c_uint(0)
        check(self._lib.lasv_graph_load_se_filelist(self._h, str(filelist).encode(), qual_thresh,
                                                    ascii_fq_offset, C.byref(n), C.byref(st)))
        return dict(st.as_dict(), files_loaded=int(n.value))

    def load_pe_filelists(self, list1, list2, qual_thresh: int = 0, ascii_fq_offset: int = 33) -> dict:
        st, n = LoadStats(), C.c_uint(0)
        check(self._lib.lasv_graph_load_pe_filelists(self._h, str(list1).encode(), str(list2).encode(),
                                                     qual_thresh, ascii_fq_offset, C.byref(n),
                                                     C.byref(st)))
        return dict(st.as_dict(), files_loaded=int(n.value))


def ctx_header(kmer_size: int, mean_read_len: int = 0, total_seq: int = 0) -> bytes:
    buf = np.zeros(128, dtype=np.uint8)
    n = _capi.lib().lasv_ctx_header(kmer_size, mean_read_len, total_seq, buf.ctypes.data)
    return buf[:n].tobytes()


def mean_contig_length(hist: np.ndarray) -> int:
    """calculate_mean_ulong (maths.c:68): integer mean of the contig-length histogram."""
    tot = int(hist.sum())
    if tot == 0:
        return 0
    return int((hist.astype(object) * np.arange(len(hist), dtype=object)).sum() // tot)


def seqfile_to_streams(path, cap_bytes: int = 1 << 26, cap_reads: int = 1 << 20):
    """Parse a sequence file with the loader's own reader (host only).  Returns
    (seq, qual or None, offsets)."""
    seq = np.zeros(cap_bytes, dtype=np.uint8)
    qual = np.zeros(cap_bytes, dtype=np.uint8)
    offs = np.zeros(cap_reads + 1, dtype=np.uint64)
    hq = C.c_int(0)
    n = check(_capi.lib().lasv_seqfile_to_streams(str(path).encode(), seq.ctypes.data, qual.ctypes.data,
                                                  cap_bytes, offs.ctypes.data, cap_reads, C.byref(hq)))
    nb = int(offs[n])
    return seq[:nb], (qual[:nb] if hq.value else None), offs[: n + 1]
