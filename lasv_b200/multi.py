"""One process, several GPUs: ctypes mirror of `lasv_multi_graph` (include/lasv_b200.h).

The hash-sharded build of SURVEY 8e driven by ONE host thread -- what the reference's single-threaded main()
reaches through the drop-in loaders with LASV_B200_DEVICES set (graph_operations.c:746-765).  `ShardedDBGraph`
(sharded.py) is the same scheme with one process per GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import LoadStats, TableSummary, check
from .graph import QUALITY_OFFSET_DEFAULT, _ptr, element_dtype, words_per_kmer


class MultiDBGraph:
    def __init__(self, kmer_size: int, number_bits: int, bucket_size: int, devices, max_rehash_tries: int = 15,
                 round_bytes: int = 0):
        self._lib = _capi.lib()
        self.kmer_size, self.number_bits, self.bucket_size = kmer_size, number_bits, bucket_size
        self.N = words_per_kmer(kmer_size)
        devs = (C.c_int * len(devices))(*devices)
        self._h = self._lib.lasv_multi_graph_new(kmer_size, number_bits, bucket_size, max_rehash_tries, devs,
                                                 len(devices), int(round_bytes))
        if not self._h:
            raise _capi.LasvError(-2, self._lib.lasv_last_error().decode(errors="replace"))
        self._h = C.c_void_p(self._h)

    def free(self):
        if self._h:
            self._lib.lasv_multi_graph_free(C.byref(self._h))
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.free()

    @property
    def n_shards(self) -> int:
        return int(self._lib.lasv_multi_graph_n_shards(self._h))

    def load_reads_host(self, seq, qual, offsets, quality_cutoff: int = QUALITY_OFFSET_DEFAULT):
        n_reads = len(offsets) - 1
        check(self._lib.lasv_multi_graph_load_reads_host(self._h, _ptr(seq), _ptr(qual), int(offsets[n_reads]),
                                                         _ptr(offsets), n_reads, quality_cutoff))

    def load_pe_filelists(self, list1, list2, qual_thresh: int, ascii_offset: int = 33) -> dict:
        st, n = LoadStats(), C.c_uint(0)
        check(self._lib.lasv_multi_graph_load_pe_filelists(self._h, str(list1).encode(), str(list2).encode(), qual_thresh,
                                                           ascii_offset, C.byref(n), C.byref(st)))
        return dict(st.as_dict(), files_loaded=int(n.value))

    def load_se_filelist(self, lst, qual_thresh: int, ascii_offset: int = 33) -> dict:
        st, n = LoadStats(), C.c_uint(0)
        check(self._lib.lasv_multi_graph_load_se_filelist(self._h, str(lst).encode(), qual_thresh, ascii_offset,
                                                          C.byref(n), C.byref(st)))
        return dict(st.as_dict(), files_loaded=int(n.value))

    def flush(self):
        check(self._lib.lasv_multi_graph_flush(self._h))

    def stats(self, hist_size: int = 0):
        st = LoadStats()
        hist = np.zeros(hist_size, dtype=np.uint64) if hist_size else None
        check(self._lib.lasv_multi_graph_stats(self._h, C.byref(st), _ptr(hist), hist_size))
        return (st.as_dict(), hist) if hist_size else st.as_dict()

    def rehash_histogram(self) -> np.ndarray:
        out = np.zeros(16, dtype=np.int64)
        check(self._lib.lasv_multi_graph_rehash_histogram(self._h, out.ctypes.data))
        return out

    def summary(self) -> dict:
        s = TableSummary()
        check(self._lib.lasv_multi_graph_summary(self._h, C.byref(s)))
        return s.as_dict()

    def export_table(self):
        """(Element[2^L * W], next_element[2^L], unique_kmers): ONE reference-layout table of all shards."""
        table = np.zeros((1 << self.number_bits) * self.bucket_size, dtype=element_dtype(self.N))
        nxt = np.zeros(1 << self.number_bits, dtype=np.int16)
        uniq = C.c_longlong(0)
        check(self._lib.lasv_multi_graph_export_table(self._h, table.ctypes.data, nxt.ctypes.data, C.byref(uniq)))
        return table, nxt, int(uniq.value)
