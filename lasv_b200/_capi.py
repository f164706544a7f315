"""ctypes binding of lasv_b200/liblasv_b200.so (include/lasv_b200.h).

The library is the product; this module only declares its entry points.  It
fails loudly if the library has not been built -- there is no Python or CPU
implementation behind any of these calls.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "liblasv_b200.so"
if os.environ.get("LASV_B200_LIBRARY"):   # experiments: a variant build of the same library (tools/variants.sh)
    LIB_PATH = Path(os.environ["LASV_B200_LIBRARY"]).resolve()

LASV_OK = 0
LASV_ERR_ARG = -1
LASV_ERR_CUDA = -2
LASV_ERR_TABLE_FULL = -3
LASV_ERR_IO = -4
LASV_ERR_STATE = -5
LASV_ERR_LIMIT = -7
LASV_ERR_CAPACITY = -6


class LoadStats(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("bad_reads", C.c_uint64), ("bases_read", C.c_uint64),
                ("bases_loaded", C.c_uint64), ("n_contigs", C.c_uint64),
                ("kmers_inserted", C.c_uint64), ("unique_kmers", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


class TableSummary(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("sum_covg", C.c_uint64), ("sum_edge_bits", C.c_uint64),
                ("fingerprint", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


class SupernodeStats(C.Structure):
    _fields_ = [("n_supernodes", C.c_uint64), ("n_nodes", C.c_uint64), ("n_interior", C.c_uint64),
                ("n_paths", C.c_uint64), ("n_cycles", C.c_uint64), ("n_not_printed", C.c_uint64),
                ("longest_edges", C.c_uint64), ("total_bases", C.c_uint64), ("over_max_length", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


class CleanStats(C.Structure):
    _fields_ = [("tips_clipped", C.c_uint64), ("tip_nodes_pruned", C.c_uint64), ("supernodes_pruned", C.c_uint64),
                ("nodes_pruned", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


class HealthStats(C.Structure):
    _fields_ = [("nodes_checked", C.c_uint64), ("edges_checked", C.c_uint64), ("missing_target", C.c_uint64),
                ("missing_return_arrow", C.c_uint64), ("zero_kmers", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


class BranchStats(C.Structure):
    _fields_ = [("n_branches", C.c_uint64), ("n_branching_nodes", C.c_uint64), ("n_nodes", C.c_uint64),
                ("n_refused", C.c_uint64), ("n_cut", C.c_uint64), ("bytes", C.c_uint64)]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f, _ in self._fields_}


CLEAN_TIPS, CLEAN_SUPERNODES = 1, 2


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genome_len", C.c_uint64), ("read_len", C.c_uint32),
                ("insert", C.c_uint32), ("err_q20", C.c_uint32), ("n_q20", C.c_uint32),
                ("lowq_q20", C.c_uint32), ("err_lowq_pct", C.c_uint32), ("tail_min", C.c_uint32),
                ("tail_max", C.c_uint32), ("tiling", C.c_uint32), ("tiling_step", C.c_uint32)]


class RefHashTable(C.Structure):
    """include/hash_table/open_hash/hash_table.h:40-50"""
    _fields_ = [("kmer_size", C.c_short), ("number_buckets", C.c_longlong), ("bucket_size", C.c_int),
                ("table", C.c_void_p), ("next_element", C.c_void_p), ("collisions", C.c_void_p),
                ("unique_kmers", C.c_longlong), ("max_rehash_tries", C.c_int)]


# every symbol include/lasv_b200.h declares (tests check the library exports them all)
EXPORTS = [
    "lasv_last_error", "lasv_device_count", "lasv_version", "lasv_graph_new", "lasv_graph_free",
    "lasv_graph_clear", "lasv_graph_unique_kmers", "lasv_graph_capacity", "lasv_graph_words_per_kmer",
    "lasv_graph_element_bytes", "lasv_graph_device_table", "lasv_graph_device_table_bytes", "lasv_graph_load_reads_device",
    "lasv_graph_load_reads_host", "lasv_graph_stats", "lasv_graph_rehash_histogram",
    "lasv_graph_route_reads_device", "lasv_graph_extract_records_device",
    "lasv_graph_insert_records_device", "lasv_graph_read_stats_device", "lasv_graph_find",
    "lasv_graph_summary", "lasv_graph_finalize", "lasv_graph_export_table", "lasv_graph_import_table",
    "lasv_graph_dump_ctx_records", "lasv_graph_write_ctx", "lasv_ctx_header", "lasv_graph_write_supernodes", "lasv_ref_hash_table_write_supernodes", "lasv_ref_hash_table_clean", "lasv_graph_insert_grouped_records_device", "lasv_graph_group_records_device", "lasv_graph_write_branches", "lasv_graph_append_ctx_records", "lasv_graph_clean", "lasv_graph_health_check", "lasv_ref_hash_table_write_branches",
    "lasv_graph_load_ctx_records", "lasv_graph_load_ctx_file", "lasv_graph_load_se_file",
    "lasv_graph_load_pe_files", "lasv_graph_load_se_filelist", "lasv_graph_load_pe_filelists",
    "load_se_filelist_into_graph_colour", "load_pe_filelists_into_graph_colour",
    "lasv_synth_reads_device", "lasv_synth_reads_host", "lasv_random_access_probe",
    "lasv_kernel_launches", "lasv_seqfile_to_streams", "lasv_random_access_probe_mode",
    "lasv_graph_record_buckets_device", "lasv_graph_bin_geometry", "lasv_graph_bin_reads_device",
    "lasv_graph_insert_bins_device", "lasv_graph_set_pipeline", "lasv_graph_flush",
    "lasv_graph_enable_kernel_timing", "lasv_graph_kernel_times", "lasv_graph_stats_async",
    "lasv_ref_hash_table_load_ctx_records", "lasv_seqfile_parse_check", "lasv_seqfile_inflate_check", "lasv_seqfile_parse_check_text",
    "lasv_graph_kernel_times_by_kind", "lasv_graph_set_insert_mode", "lasv_graph_insert_counts",
    "lasv_graph_set_accumulation", "lasv_graph_load_reads_host_async", "lasv_graph_set_staging",
    "lasv_graph_bin_reads_append_device",
    "lasv_multi_graph_new", "lasv_multi_graph_free", "lasv_multi_graph_n_shards", "lasv_multi_graph_load_reads_host",
    "lasv_multi_graph_load_se_filelist", "lasv_multi_graph_load_pe_filelists", "lasv_multi_graph_flush",
    "lasv_multi_graph_stats", "lasv_multi_graph_rehash_histogram", "lasv_multi_graph_summary",
    "lasv_multi_graph_export_table",
    "lasv_graph_write_supernodes_strict", "lasv_ref_hash_table_write_supernodes_strict", "lasv_graph_clean_limited",
    "lasv_ref_hash_table_clean_limited", "lasv_graph_bin_reads_peers_device",
]

INSERT_AUTO, INSERT_RANDOM, INSERT_STREAMED = 0, 1, 2
T_KINDS = ("partition", "insert", "sort", "stream", "drain")  # LASV_T_* of include/lasv_b200.h

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lasv_b200/csrc` (there is no fallback implementation)")
    L = C.CDLL(str(LIB_PATH))
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.lasv_last_error.restype = C.c_char_p
    L.lasv_version.restype = C.c_char_p
    L.lasv_device_count.restype = i32
    L.lasv_kernel_launches.restype = u64
    L.lasv_graph_new.restype = vp
    L.lasv_graph_new.argtypes = [i32, i32, i32, i32, i32]
    L.lasv_graph_free.argtypes = [C.POINTER(vp)]
    L.lasv_graph_clear.argtypes = [vp, vp]
    L.lasv_graph_set_pipeline.argtypes = [vp, i32]
    L.lasv_graph_flush.argtypes = [vp, vp]
    L.lasv_graph_enable_kernel_timing.argtypes = [vp, i32]
    L.lasv_graph_kernel_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(u64)]
    L.lasv_graph_kernel_times_by_kind.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(u64)]
    L.lasv_graph_set_insert_mode.argtypes = [vp, i32, C.c_double]
    L.lasv_graph_insert_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.lasv_graph_set_accumulation.argtypes = [vp, u64, C.POINTER(u64)]
    L.lasv_graph_unique_kmers.restype = C.c_longlong
    L.lasv_graph_unique_kmers.argtypes = [vp]
    L.lasv_graph_capacity.restype = C.c_longlong
    L.lasv_graph_capacity.argtypes = [vp]
    L.lasv_graph_words_per_kmer.argtypes = [vp]
    L.lasv_graph_element_bytes.restype = C.c_size_t
    L.lasv_graph_element_bytes.argtypes = [vp]
    L.lasv_graph_device_table.restype = vp
    L.lasv_graph_device_table.argtypes = [vp]
    L.lasv_graph_device_table_bytes.restype = C.c_uint64
    L.lasv_graph_device_table_bytes.argtypes = [vp]
    L.lasv_graph_load_reads_device.argtypes = [vp, vp, vp, u64, vp, u64, i32, vp]
    L.lasv_graph_load_reads_host.argtypes = [vp, vp, vp, u64, vp, u64, i32, C.POINTER(LoadStats)]
    L.lasv_graph_load_reads_host_async.argtypes = [vp, vp, vp, u64, vp, u64, i32, vp]
    L.lasv_graph_set_staging.argtypes = [vp, u64, C.POINTER(u64)]
    L.lasv_graph_stats.argtypes = [vp, C.POINTER(LoadStats), vp, u64, vp]
    L.lasv_graph_rehash_histogram.argtypes = [vp, vp]
    L.lasv_graph_stats_async.argtypes = [vp, vp, vp]
    L.lasv_graph_route_reads_device.argtypes = [vp, vp, vp, u64, i32, i32, vp, u64, vp, vp]
    L.lasv_graph_extract_records_device.argtypes = [vp, vp, vp, u64, i32, vp, u64, vp, vp]
    L.lasv_graph_insert_records_device.argtypes = [vp, vp, u64, vp]
    L.lasv_graph_read_stats_device.argtypes = [vp, vp, vp, vp, u64, i32, vp]
    u32p = C.POINTER(C.c_uint32)
    L.lasv_graph_bin_geometry.argtypes = [vp, i32, u64, u64, u32p, u32p, u32p, u32p, u32p]
    L.lasv_graph_bin_reads_device.argtypes = [vp, vp, vp, u64, i32, i32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.lasv_graph_bin_reads_append_device.argtypes = L.lasv_graph_bin_reads_device.argtypes
    L.lasv_graph_bin_reads_peers_device.argtypes = [vp, vp, vp, u64, i32, i32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                    C.POINTER(vp), vp, vp, i32]
    L.lasv_multi_graph_new.restype = vp
    L.lasv_multi_graph_new.argtypes = [i32, i32, i32, i32, C.POINTER(i32), i32, u64]
    L.lasv_multi_graph_free.argtypes = [C.POINTER(vp)]
    L.lasv_multi_graph_free.restype = None
    L.lasv_multi_graph_n_shards.argtypes = [vp]
    L.lasv_multi_graph_load_reads_host.argtypes = [vp, vp, vp, u64, vp, u64, i32]
    L.lasv_multi_graph_load_se_filelist.argtypes = [vp, C.c_char_p, i32, i32, C.POINTER(C.c_uint), C.POINTER(LoadStats)]
    L.lasv_multi_graph_load_pe_filelists.argtypes = [vp, C.c_char_p, C.c_char_p, i32, i32, C.POINTER(C.c_uint), C.POINTER(LoadStats)]
    L.lasv_multi_graph_flush.argtypes = [vp]
    L.lasv_multi_graph_stats.argtypes = [vp, C.POINTER(LoadStats), vp, u64]
    L.lasv_multi_graph_rehash_histogram.argtypes = [vp, vp]
    L.lasv_multi_graph_summary.argtypes = [vp, vp]
    L.lasv_multi_graph_export_table.argtypes = [vp, vp, vp, C.POINTER(C.c_longlong)]
    L.lasv_graph_insert_bins_device.argtypes = [vp, vp, vp, i32, C.c_uint32, C.c_uint32, C.c_uint32, vp]
    L.lasv_graph_find.argtypes = [vp, vp, u64, vp, vp, vp]
    L.lasv_graph_summary.argtypes = [vp, C.POINTER(TableSummary)]
    L.lasv_graph_finalize.argtypes = [vp]
    L.lasv_graph_export_table.argtypes = [vp, vp, vp]
    L.lasv_graph_import_table.argtypes = [vp, vp]
    L.lasv_graph_dump_ctx_records.restype = C.c_longlong
    L.lasv_graph_dump_ctx_records.argtypes = [vp, vp, u64]
    L.lasv_graph_append_ctx_records.restype = C.c_longlong
    L.lasv_graph_append_ctx_records.argtypes = [vp, C.c_char_p]
    L.lasv_graph_write_ctx.argtypes = [vp, C.c_char_p, C.c_uint32, u64]
    L.lasv_ctx_header.argtypes = [i32, C.c_uint32, u64, vp]
    L.lasv_graph_write_supernodes.argtypes = [vp, C.c_char_p, i32, C.POINTER(SupernodeStats)]
    L.lasv_graph_write_supernodes_strict.argtypes = [vp, C.c_char_p, i32, C.POINTER(SupernodeStats)]
    L.lasv_graph_load_ctx_records.argtypes = [vp, vp, u64]
    L.lasv_graph_load_ctx_file.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(u64)]
    L.lasv_graph_load_se_file.argtypes = [vp, C.c_char_p, i32, C.POINTER(LoadStats)]
    L.lasv_graph_load_pe_files.argtypes = [vp, C.c_char_p, C.c_char_p, i32, C.POINTER(LoadStats)]
    L.lasv_graph_load_se_filelist.argtypes = [vp, C.c_char_p, i32, i32, C.POINTER(C.c_uint),
                                              C.POINTER(LoadStats)]
    L.lasv_graph_load_pe_filelists.argtypes = [vp, C.c_char_p, C.c_char_p, i32, i32,
                                               C.POINTER(C.c_uint), C.POINTER(LoadStats)]
    ull_p = C.POINTER(C.c_ulonglong)
    L.load_se_filelist_into_graph_colour.restype = None
    L.load_se_filelist_into_graph_colour.argtypes = [
        C.c_char_p, i32, i32, C.c_byte, C.c_char, i32, C.POINTER(RefHashTable), C.c_char,
        C.POINTER(C.c_uint), ull_p, ull_p, ull_p, ull_p, vp, C.c_ulong]
    L.load_pe_filelists_into_graph_colour.restype = None
    L.load_pe_filelists_into_graph_colour.argtypes = [
        C.c_char_p, C.c_char_p, i32, i32, C.c_byte, C.c_char, i32, C.POINTER(RefHashTable), C.c_char,
        C.POINTER(C.c_uint), ull_p, ull_p, ull_p, ull_p, vp, C.c_ulong]
    L.lasv_synth_reads_device.argtypes = [C.POINTER(SynthParams), u64, u64, vp, vp, vp]
    L.lasv_synth_reads_host.argtypes = [C.POINTER(SynthParams), u64, u64, vp, vp]
    L.lasv_random_access_probe.argtypes = [vp, u64, u64, u64, vp]
    L.lasv_random_access_probe_mode.argtypes = [vp, u64, u64, u64, i32, vp]
    L.lasv_graph_record_buckets_device.argtypes = [vp, vp, u64, vp, vp]
    L.lasv_seqfile_to_streams.restype = C.c_longlong
    L.lasv_seqfile_to_streams.argtypes = [C.c_char_p, vp, vp, u64, vp, u64, C.POINTER(i32)]
    L.lasv_seqfile_parse_check.argtypes = [C.c_char_p, C.c_char_p, i32, u64, vp, C.POINTER(i32)]
    L.lasv_seqfile_inflate_check.argtypes = [C.c_char_p, i32, u64, vp]
    L.lasv_seqfile_parse_check_text.argtypes = [C.c_char_p, C.c_char_p, i32, u64, vp, C.POINTER(i32), vp]
    L.lasv_ref_hash_table_write_supernodes.argtypes = [C.POINTER(RefHashTable), C.c_char_p, i32, C.POINTER(SupernodeStats)]
    L.lasv_ref_hash_table_clean.argtypes = [C.POINTER(RefHashTable), i32, i32, C.POINTER(CleanStats)]
    L.lasv_graph_health_check.argtypes = [vp, C.POINTER(HealthStats)]
    L.lasv_graph_group_records_device.argtypes = [vp, vp, u64, C.c_uint32, vp, vp, vp]
    L.lasv_graph_insert_grouped_records_device.argtypes = [vp, vp, u64, vp, C.c_uint32, C.POINTER(C.c_uint64), vp]
    L.lasv_graph_clean.argtypes = [vp, i32, i32, C.POINTER(CleanStats)]
    L.lasv_graph_clean_limited.argtypes = [vp, i32, i32, i32, C.POINTER(CleanStats)]
    L.lasv_graph_write_branches.argtypes = [vp, C.c_char_p, i32, C.POINTER(BranchStats)]
    L.lasv_ref_hash_table_write_branches.argtypes = [C.POINTER(RefHashTable), C.c_char_p, C.POINTER(BranchStats)]
    L.lasv_ref_hash_table_load_ctx_records.restype = C.c_longlong
    L.lasv_ref_hash_table_load_ctx_records.argtypes = [C.POINTER(RefHashTable), C.c_char_p, C.c_long]
    _lib = L
    return L


class LasvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


class TableFullError(LasvError):
    """hash_table.c:309-317: a key met 16 full buckets (the reference die()s)."""


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().lasv_last_error().decode(errors="replace")
        raise (TableFullError if rc == LASV_ERR_TABLE_FULL else LasvError)(rc, msg)
    return rc
