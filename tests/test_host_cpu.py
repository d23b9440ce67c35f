"""CPU-side checks (-m "not gpu"): Philox known answers, and that the C-ABI library loads and exports every
symbol include/geneevolve_b200.h declares (no compute calls without a GPU; ge_create must fail loudly)."""
import ctypes
import os
import re

import pytest

from geneevolve_b200 import capi
from oracle.oracle import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_philox_known_answer():
    """Random123 kat_vectors for philox4x32-10."""
    assert philox(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox(0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "geneevolve_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ge_[a-zA-Z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from geneevolve_b200 import build
    build.build()
    lib = capi.load_library()
    syms = declared_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in include/geneevolve_b200.h but not exported"


def test_struct_sizes_match_header():
    # the ctypes mirrors must have the C layout (x86-64 SysV)
    assert ctypes.sizeof(capi.ge_config) == 64
    assert ctypes.sizeof(capi.ge_gen_params) == 40
    assert ctypes.sizeof(capi.ge_draws) == 13 * 8
    assert ctypes.sizeof(capi.ge_indiv_soa) == 12 * 8
    assert ctypes.sizeof(capi.ge_moments) == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GeneEvolveError) as e:
        capi.Engine(n_pop=1, n_chr=1, n_phen=1, capacity=8)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
