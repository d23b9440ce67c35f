"""ctypes loader of the CPU oracle (oracle/libge_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.  It reuses the
struct definitions and the numpy veneer of geneevolve_b200.capi (the oracle mirrors the C-ABI with a `go_`
prefix); the product package never imports anything from oracle/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from geneevolve_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libge_oracle.so")
GO_RNG_REF = 2
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def lib():
    global _lib
    if _lib is None:
        src = max(os.path.getmtime(os.path.join(HERE, f)) for f in ("ge_oracle.cpp", "ge_oracle.h"))
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < src:
            build()
        _lib = C.CDLL(LIB)
    return _lib


class OracleEngine(capi.Engine):
    def __init__(self, **kw):
        super().__init__(lib(), "go_", **kw)

    def haplotypes_from_segments(self, pop, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, self.n_loci[chr_]), np.uint8)
        self._call("download_haplotypes_from_segments", self.ctx, pop, chr_, capi._ptr(out, capi._u8p))
        return out

    def cv_alleles_bits(self, pop, phen, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, self.n_cv[(phen, chr_)]), np.uint8)
        self._call("download_cv_alleles_bits", self.ctx, pop, phen, chr_, capi._ptr(out, capi._u8p))
        return out

    def AD_raw(self, pop):
        n = self.population_size(pop)
        a, d = np.zeros((self.n_phen, n)), np.zeros((self.n_phen, n))
        self._call("download_AD_raw", self.ctx, pop, capi._ptr(a, capi._f64p), capi._ptr(d, capi._f64p))
        return a, d

    def e_raw(self, pop):
        n = self.population_size(pop)
        e = np.zeros((self.n_phen, n))
        self._call("download_e_raw", self.ctx, pop, capi._ptr(e, capi._f64p))
        return e


def philox(k0, k1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().go_philox4x32_10(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), out)
    return list(out)


def bench_propagate_bits(n_parents, n_offspring, n_loci, n_xo, seed=1):
    f = lib().go_bench_propagate_bits
    f.restype = C.c_double
    cs = C.c_uint64()
    t = f(C.c_uint64(n_parents), C.c_uint64(n_offspring), C.c_uint64(n_loci), C.c_uint64(n_xo), C.c_uint64(seed), C.byref(cs))
    return t, cs.value
