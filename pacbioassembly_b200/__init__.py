"""pacbioassembly_b200 -- B200-native (sm_100a) read-to-reference hot path of vmingchen/PacBioAssembly.

Spaced-seed extraction, seed-index probe and banded edit-distance alignment as hand-written CUDA kernels behind
the C ABI of include/pacbio_b200.h.  `api` is the ctypes host binding; `host/` holds the C++ classes that keep the
reference's dna_seq / seq_accessor / seq_aligner / locator interfaces.  There is no CPU fallback.
"""
from . import api  # noqa: F401
from .api import Context, PbError, parse_pattern  # noqa: F401

__all__ = ["api", "Context", "PbError", "parse_pattern"]
