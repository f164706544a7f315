"""lasv_b200 -- B200-native de Bruijn graph construction for laSV.

The product is the CUDA library lasv_b200/liblasv_b200.so (C ABI:
include/lasv_b200.h).  This package is its host-side mirror of the reference's
hash_table / file_reader interface; importing it never falls back to a CPU
implementation.
"""
from ._capi import LasvError, TableFullError, lib  # noqa: F401
from .graph import DBGraph, ctx_header, element_dtype, mean_contig_length, record_dtype, words_per_kmer  # noqa: F401
from . import synth  # noqa: F401
from .multi import MultiDBGraph  # noqa: F401

__all__ = ["DBGraph", "LasvError", "TableFullError", "ctx_header", "element_dtype", "record_dtype",
           "words_per_kmer", "mean_contig_length", "synth", "lib"]
