"""ctypes binding of the C-ABI in include/geneevolve_b200.h.

`Engine` is a thin numpy-in / numpy-out veneer over the entry points; every method maps 1:1 onto one C
function, which in turn replaces one private method of the reference's `class Simulation`
(src/Simulation.h:64-144, see the header for file:line of each).  The product always binds
libgeneevolve_b200.so (hand-written sm_100a CUDA); there is no CPU implementation behind this module and
loading fails loudly when the library has not been built (`python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

import numpy as np

GE_REP_BITS, GE_REP_SEGMENTS = 1, 2
GE_RNG_PHILOX, GE_RNG_REPLAY = 0, 1
GE_FLAG_SERIAL, GE_FLAG_SEG_WIDE_PARTS, GE_FLAG_SEG_VERBATIM, GE_FLAG_CV_FROM_SEGMENTS, GE_FLAG_NO_GRAPH = 1, 2, 4, 8, 16
GE_SEL = {"": 0, "logit": 1, "probit": 2, "stab": 3, "thr": 4}
GE_KERNEL_PROPAGATE_BITS, GE_KERNEL_RECOMBINE_SEGMENTS = 0, 1
GE_PHASES = {"mate": 2, "sample": 3, "cv_and_genetic_values": 4, "phenotype": 5}

_u64p, _u8p, _f64p, _i32p, _u32p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_uint32))


class ge_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_pop", C.c_int32), ("n_chr", C.c_int32), ("n_phen", C.c_int32),
                ("vt_type", C.c_int32), ("representation", C.c_int32), ("rng_mode", C.c_int32),
                ("flags", C.c_int32), ("seed", C.c_uint64), ("capacity", C.c_uint64),
                ("seg_capacity", C.c_uint64), ("rank", C.c_int32), ("world_size", C.c_int32)]


class ge_gen_params(C.Structure):
    _fields_ = [("pop_size", C.c_uint64), ("mat_cor", C.c_double), ("offspring_dist", C.c_int32),
                ("selection_func", C.c_int32), ("selection_par1", C.c_double), ("selection_par2", C.c_double)]


class ge_draws(C.Structure):
    _fields_ = [("n_offspring", C.c_uint64), ("father", _u64p), ("mother", _u64p), ("sex", _u8p),
                ("xo_off", _u64p), ("xo_bp", _u64p), ("start_hap", _u8p), ("mut_off", _u64p),
                ("mut_bp", _u64p), ("mut_gam", _u8p), ("e_raw", _f64p), ("common", _f64p), ("parental0", _f64p)]


class ge_mate_draws(C.Structure):
    _fields_ = [("thin_u", _f64p), ("mm_u", _f64p), ("trim_order", _u64p), ("n_trim_order", C.c_uint64), ("t1", _f64p), ("t2", _f64p),
                ("n_couples", C.c_uint64), ("family", _i32p), ("remainder_order", _u64p), ("n_remainder_order", C.c_uint64),
                ("rm_father_idx", _u64p), ("rm_mother_idx", _u64p), ("n_rm", C.c_uint64)]


class ge_indiv_soa(C.Structure):
    _fields_ = [("ids", _u64p), ("sex", _u8p)] + [(k, _f64p) for k in "ADGCEFP"] + \
               [("mv", _f64p), ("sv", _f64p), ("svf", _f64p)]


class ge_moments(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("var_A", "var_D", "var_G", "var_C", "var_E", "var_F", "var_P", "h2")]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)


class GeneEvolveError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else t()


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def gen_params(pop_size, mat_cor=0.0, offspring_dist="p", selection_func="logit", par1=0.0, par2=1.0):
    """One row of the generation-info file (src/Population.cpp:13-96)."""
    code = GE_SEL[selection_func] if isinstance(selection_func, str) else int(selection_func)
    od = ord(offspring_dist) if isinstance(offspring_dist, str) else int(offspring_dist)
    return ge_gen_params(int(pop_size), float(mat_cor), od, code, float(par1), float(par2))


class Draws:
    """Host-side `ge_draws` (numpy arrays kept alive next to the ctypes struct)."""
    FIELDS = [("father", np.uint64), ("mother", np.uint64), ("sex", np.uint8), ("xo_off", np.uint64),
              ("xo_bp", np.uint64), ("start_hap", np.uint8), ("mut_off", np.uint64), ("mut_bp", np.uint64),
              ("mut_gam", np.uint8), ("e_raw", np.float64), ("common", np.float64), ("parental0", np.float64)]
    _PT = {np.uint64: _u64p, np.uint8: _u8p, np.float64: _f64p}

    def __init__(self, n_offspring, **kw):
        self.n_offspring = int(n_offspring)
        self.arrays = {}
        for name, dt in self.FIELDS:
            self.arrays[name] = _arr(kw.get(name), dt)
        extra = set(kw) - {n for n, _ in self.FIELDS}
        if extra:
            raise TypeError(f"unknown draw fields {extra}")

    def struct(self):
        s = ge_draws()
        s.n_offspring = self.n_offspring
        for name, dt in self.FIELDS:
            setattr(s, name, _ptr(self.arrays[name], self._PT[dt]))
        return s


def default_library_path():
    # GE_LIBRARY: measurement aid — another build of the same CUDA library (A/B runs on one box)
    return os.environ.get("GE_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgeneevolve_b200.so")


def load_library(path=None):
    path = path or default_library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build the CUDA extension first (__graft_entry__.build()); "
                          "there is no CPU fallback")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


class Engine:
    """One context (= one GPU).  `prefix` selects the symbol family of `lib` ("ge_" for the product)."""

    def __init__(self, lib=None, prefix="ge_", *, n_pop=1, n_chr=1, n_phen=1, vt_type=1, device=0,
                 representation=GE_REP_BITS | GE_REP_SEGMENTS, rng_mode=GE_RNG_PHILOX, seed=1, capacity=0,
                 seg_capacity=0, rank=0, world_size=1, flags=0):
        self.lib = lib if lib is not None else load_library()
        self.prefix = prefix
        self.n_pop, self.n_chr, self.n_phen = n_pop, n_chr, n_phen
        self.n_loci = [0] * n_chr
        self.n_cv = {}
        self._fn("last_error").restype = C.c_char_p
        self.cfg = ge_config(device, n_pop, n_chr, n_phen, vt_type, representation, rng_mode, flags, seed,
                             capacity, seg_capacity, rank, world_size)
        self.ctx = C.c_void_p()
        self._call("create", C.byref(self.cfg), C.byref(self.ctx))

    # -- plumbing
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _call(self, name, *args):
        rc = self._fn(name)(*args)
        if rc != 0:
            raise GeneEvolveError(rc, self._fn("last_error")().decode())

    def close(self):
        if self.ctx:
            self._call("destroy", self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs
    def set_population(self, pop, avoid_inbreeding=False, random_mating=False, mm_percent=0.0):
        self._call("set_population", self.ctx, pop, int(avoid_inbreeding), int(random_mating), C.c_double(mm_percent))

    def set_genetic_map(self, pop, chr_, bp, recom_prob, bp_dist):
        bp, rp = _arr(bp, np.uint64), _arr(recom_prob, np.float64)
        self._call("set_genetic_map", self.ctx, pop, chr_, _ptr(bp, _u64p), _ptr(rp, _f64p), C.c_uint64(len(bp)), C.c_uint64(int(bp_dist)))

    def set_mutation_map(self, pop, chr_, bp, rate):
        bp, r = _arr(bp, np.uint64), _arr(rate, np.float64)
        self._call("set_mutation_map", self.ctx, pop, chr_, _ptr(bp, _u64p), _ptr(r, _f64p), C.c_uint64(len(bp)))

    def set_loci(self, chr_, pos):
        pos = _arr(pos, np.uint64)
        self.n_loci[chr_] = len(pos)
        self._call("set_loci", self.ctx, chr_, _ptr(pos, _u64p), C.c_uint64(len(pos)))

    def set_founder_panel(self, pop, chr_, alleles):
        al = _arr(alleles, np.uint8)
        assert al.ndim == 2 and al.shape[1] == self.n_loci[chr_]
        self._call("set_founder_panel", self.ctx, pop, chr_, _ptr(al, _u8p), C.c_uint64(al.shape[0]))

    def set_founder_panel_packed(self, pop, chr_, words):
        w = _arr(words, np.uint32)
        assert w.ndim == 2 and w.shape[1] == (self.n_loci[chr_] + 31) // 32
        self._call("set_founder_panel_packed", self.ctx, pop, chr_, _ptr(w, _u32p), C.c_uint64(w.shape[0]))

    def set_cv(self, pop, phen, chr_, bp, a, d, founder_cv):
        bp, a, d, v = _arr(bp, np.uint64), _arr(a, np.float64), _arr(d, np.float64), _arr(founder_cv, np.uint8)
        assert v.ndim == 2 and v.shape[1] == len(bp)
        self.n_cv[(phen, chr_)] = len(bp)
        self._call("set_cv", self.ctx, pop, phen, chr_, _ptr(bp, _u64p), _ptr(a, _f64p), _ptr(d, _f64p),
                   C.c_uint64(len(bp)), _ptr(v, _u8p), C.c_uint64(v.shape[0]))

    def set_pheno_scheme(self, pop, phen, va, vd, ve, vc=0.0, vf=0.0, omega=1.0, beta=0.0, lam=1.0):
        self._call("set_pheno_scheme", self.ctx, pop, phen, *[C.c_double(x) for x in (va, vd, ve, vc, vf, omega, beta, lam)])

    def set_chromosome_ids(self, global_ids):
        a = _arr(global_ids, np.int32)
        assert len(a) == self.n_chr
        self._call("set_chromosome_ids", self.ctx, _ptr(a, _i32p))

    def set_allreduce(self, fn):
        """fn(ptr:int, count:int, stream:int) -> None must sum `count` doubles at `ptr` over all ranks."""
        def thunk(user, buf, count, stream):
            try:
                fn(buf or 0, int(count), stream or 0)
                return 0
            except Exception as e:  # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        self._allreduce_cb = ALLREDUCE_FN(thunk)  # keep alive
        self._call("set_allreduce", self.ctx, self._allreduce_cb, None)

    def set_allreduce_nccl(self, comm, nccl_all_reduce):
        """comm: ncclComm_t of this rank (integer address); nccl_all_reduce: address of ncclAllReduce in the NCCL library that
        made it (dist.NcclComm provides both).  The library then issues the exchange itself and sharded generations replay as graphs."""
        self._call("set_allreduce_nccl", self.ctx, C.c_void_p(comm), C.c_void_p(nccl_all_reduce))

    def set_gamma(self, gamma):
        g = _arr(gamma, np.float64)
        self._call("set_gamma", self.ctx, _ptr(g, _f64p))

    # -- generations
    def _draws_array(self, draws):
        if draws is None:
            return None, None
        if isinstance(draws, Draws):
            draws = [draws]
        arr = (ge_draws * len(draws))(*[d.struct() for d in draws])
        return arr, draws

    def init_generation0(self, draws0=None):
        arr, keep = self._draws_array(draws0)
        self._call("init_generation0", self.ctx, arr)

    def mate(self, pop, gen, params):
        self._call("mate", self.ctx, pop, gen, C.byref(params))

    def mate_replay(self, pop, gen, params, *, thin_u, mm_u=None, trim_order=None, t1=None, t2=None, family=None, remainder_order=None,
                    rm_father_idx=None, rm_mother_idx=None):
        """The device mating kernels under the reference's own draws (ge_mate_replay); arrays as exported into tests/golden."""
        a = dict(thin_u=_arr(thin_u, np.float64), mm_u=_arr(mm_u, np.float64), trim_order=_arr(trim_order, np.uint64), t1=_arr(t1, np.float64),
                 t2=_arr(t2, np.float64), family=_arr(family, np.int32), remainder_order=_arr(remainder_order, np.uint64),
                 rm_father_idx=_arr(rm_father_idx, np.uint64), rm_mother_idx=_arr(rm_mother_idx, np.uint64))
        n = lambda k: 0 if a[k] is None else len(a[k])  # noqa: E731
        md = ge_mate_draws(_ptr(a["thin_u"], _f64p), _ptr(a["mm_u"], _f64p), _ptr(a["trim_order"], _u64p), n("trim_order"), _ptr(a["t1"], _f64p),
                           _ptr(a["t2"], _f64p), n("t1"), _ptr(a["family"], _i32p), _ptr(a["remainder_order"], _u64p), n("remainder_order"),
                           _ptr(a["rm_father_idx"], _u64p), _ptr(a["rm_mother_idx"], _u64p), n("rm_father_idx"))
        self._call("mate_replay", self.ctx, pop, gen, C.byref(params), C.byref(md))

    def set_couples(self, pop, pos_male, pos_female, inbreed, num_offspring):
        m, f = _arr(pos_male, np.uint64), _arr(pos_female, np.uint64)
        ib, no = _arr(inbreed, np.uint8), _arr(num_offspring, np.int32)
        self._call("set_couples", self.ctx, pop, _ptr(m, _u64p), _ptr(f, _u64p), _ptr(ib, _u8p), _ptr(no, _i32p), C.c_uint64(len(m)))

    def get_couples(self, pop):
        n = C.c_uint64()
        self._call("get_couples_count", self.ctx, pop, C.byref(n))
        n = n.value
        m, f = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        ib, no = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        self._call("get_couples", self.ctx, pop, _ptr(m, _u64p), _ptr(f, _u64p), _ptr(ib, _u8p), _ptr(no, _i32p))
        return dict(pos_male=m, pos_female=f, inbreed=ib, num_offspring=no)

    def reproduce(self, pop, gen, draws=None):
        s = draws.struct() if draws is not None else None
        self._call("reproduce", self.ctx, pop, gen, C.byref(s) if s is not None else None)

    def compute_AD(self, pop, gen):
        self._call("compute_AD", self.ctx, pop, gen)

    def scale_AD_compute_GEF(self, pop, gen, phen, e_raw=None):
        e = _arr(e_raw, np.float64)
        self._call("scale_AD_compute_GEF", self.ctx, pop, gen, phen, _ptr(e, _f64p))

    def environmental_effects_specific_to_each_population(self, phen):
        self._call("environmental_effects_specific_to_each_population", self.ctx, phen)

    def compute_mating_value_selection_value(self, pop, gen, params=None):
        self._call("compute_mating_value_selection_value", self.ctx, pop, gen, C.byref(params) if params is not None else None)

    def do_migration(self, gen, row):
        r = _arr(row, np.float64)
        self._call("do_migration", self.ctx, gen, _ptr(r, _f64p))

    def set_migration_sample(self, src_pop, positions):
        a = _arr(positions, np.uint64)
        self._call("set_migration_sample", self.ctx, src_pop, _ptr(a, _u64p), C.c_uint64(len(a)))

    def save_human_info_to_Pop_info_prev_gen(self, pop):
        self._call("save_human_info_to_Pop_info_prev_gen", self.ctx, pop)

    def step_generation(self, gen, params, migration_row=None, draws=None):
        parr = (ge_gen_params * len(params))(*params)
        mig = _arr(migration_row, np.float64)
        darr, keep = self._draws_array(draws)
        self._call("step_generation", self.ctx, gen, parr, _ptr(mig, _f64p), darr)

    # -- results
    def population_size(self, pop):
        n = C.c_uint64()
        self._call("get_population_size", self.ctx, pop, C.byref(n))
        return n.value

    @staticmethod
    def individual_bytes(n, n_phen):
        """Bytes ge_download_individuals moves for n individuals (the `.info` columns)."""
        return n * (7 * 8 + 1 + 7 * 8 * n_phen + 3 * 8)

    def individuals(self, pop, out=None):
        """`.info` columns of one population.  `out`: optional dict of preallocated (e.g. pinned) arrays
        sized for at least the current population; views of the right length are returned."""
        n, nf = self.population_size(pop), self.n_phen
        if out is None:
            out = {"ids": np.zeros((n, 7), np.uint64), "sex": np.zeros(n, np.uint8)}
            for k in "ADGCEFP":
                out[k] = np.zeros((nf, n), np.float64)
            for k in ("mv", "sv", "svf"):
                out[k] = np.zeros(n, np.float64)
        else:
            out = {"ids": out["ids"].reshape(-1)[:n * 7].reshape(n, 7), "sex": out["sex"][:n],
                   **{k: out[k].reshape(-1)[:nf * n].reshape(nf, n) for k in "ADGCEFP"},
                   **{k: out[k][:n] for k in ("mv", "sv", "svf")}}
        s = ge_indiv_soa()
        s.ids, s.sex = _ptr(out["ids"], _u64p), _ptr(out["sex"], _u8p)
        for k in list("ADGCEFP") + ["mv", "sv", "svf"]:
            setattr(s, k, _ptr(out[k], _f64p))
        self._call("download_individuals", self.ctx, pop, C.byref(s))
        return out

    def moments(self, pop, phen):
        m = ge_moments()
        self._call("get_moments", self.ctx, pop, phen, C.byref(m))
        return {k: getattr(m, k) for k, _ in ge_moments._fields_}

    def mv_sv_var(self, pop):
        a, b = C.c_double(), C.c_double()
        self._call("get_mv_sv_var", self.ctx, pop, C.byref(a), C.byref(b))
        return a.value, b.value

    def gen0_constants(self, pop, phen):
        v = [C.c_double() for _ in range(5)]
        self._call("get_gen0_constants", self.ctx, pop, phen, *[C.byref(x) for x in v])
        return dict(zip(("var_a0", "var_d0", "beta", "sv_mean0", "sv_var0"), [x.value for x in v]))

    def haplotypes(self, pop, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, self.n_loci[chr_]), np.uint8)
        self._call("download_haplotypes", self.ctx, pop, chr_, _ptr(out, _u8p))
        return out

    def haplotypes_from_segments(self, pop, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, self.n_loci[chr_]), np.uint8)
        self._call("download_haplotypes_from_segments", self.ctx, pop, chr_, _ptr(out, _u8p))
        return out

    def haplotypes_packed(self, pop, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, (self.n_loci[chr_] + 31) // 32), np.uint32)
        self._call("download_haplotypes_packed", self.ctx, pop, chr_, _ptr(out, _u32p))
        return out

    def segments(self, pop, chr_):
        n = self.population_size(pop)
        ns, nm = C.c_uint64(), C.c_uint64()
        self._call("get_segment_count", self.ctx, pop, chr_, C.byref(ns), C.byref(nm))
        off, seg = np.zeros(2 * n + 1, np.uint64), np.zeros((ns.value, 4), np.uint64)
        moff, mbp = np.zeros(2 * n + 1, np.uint64), np.zeros(nm.value, np.uint64)
        self._call("download_segments", self.ctx, pop, chr_, _ptr(off, _u64p), _ptr(seg, _u64p), _ptr(moff, _u64p), _ptr(mbp, _u64p))
        return dict(seg_off=off, seg=seg, mut_off=moff, mut_bp=mbp)

    def cv_alleles(self, pop, phen, chr_):
        n = self.population_size(pop)
        out = np.zeros((2 * n, self.n_cv[(phen, chr_)]), np.uint8)
        self._call("download_cv_alleles", self.ctx, pop, phen, chr_, _ptr(out, _u8p))
        return out

    def compact_segments(self, pop):
        """Merges adjacent same-founder parts (extension, see the header); returns (parts before, parts after)."""
        a, b = C.c_uint64(), C.c_uint64()
        self._call("compact_segments", self.ctx, pop, C.byref(a), C.byref(b))
        return a.value, b.value

    def rebase_founders(self, keep_history=True):
        """The current generation becomes the founder panel; every list restarts as one part (ge_rebase_founders)."""
        self._call("rebase_founders", self.ctx, int(bool(keep_history)))

    def segments_gen0(self, pop, chr_):
        """The lists of one chromosome against the generation-0 founders, composed through the re-basing history."""
        n = self.population_size(pop)
        ns = C.c_uint64()
        self._call("get_segment_count_gen0", self.ctx, pop, chr_, C.byref(ns))
        off, seg = np.zeros(2 * n + 1, np.uint64), np.zeros((ns.value, 4), np.uint64)
        self._call("download_segments_gen0", self.ctx, pop, chr_, _ptr(off, _u64p), _ptr(seg, _u64p))
        return dict(seg_off=off, seg=seg)

    def segment_format(self):
        """Bytes per part in device memory: 16 (the reference's part) or 8 (packed, end implied)."""
        b = C.c_int()
        self._call("get_segment_format", self.ctx, C.byref(b))
        return b.value

    def ibd_sharing(self, pop, chrom, ind_a, ind_b, min_bp=0):
        """(shared_bp, n_runs) per pair over the four haplotype combinations (ge_ibd_sharing)."""
        a, b = np.ascontiguousarray(ind_a, np.uint64), np.ascontiguousarray(ind_b, np.uint64)
        tot, runs = np.zeros(len(a), np.uint64), np.zeros(len(a), np.uint32)
        self._call("ibd_sharing", self.ctx, pop, chrom, _ptr(a, _u64p), _ptr(b, _u64p), C.c_uint64(len(a)), C.c_uint64(int(min_bp)), _ptr(tot, _u64p), _ptr(runs, _u32p))
        return tot, runs

    def recompute_cv_from_segments(self, pop):
        self._call("recompute_cv_from_segments", self.ctx, pop)

    def draws(self, pop):
        no, nx, nm = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._call("get_draw_counts", self.ctx, pop, C.byref(no), C.byref(nx), C.byref(nm))
        no, nx, nm = no.value, nx.value, nm.value
        C_ = self.n_chr
        d = dict(father=np.zeros(no, np.uint64), mother=np.zeros(no, np.uint64), sex=np.zeros(no, np.uint8),
                 xo_off=np.zeros(no * C_ * 2 + 1, np.uint64), xo_bp=np.zeros(nx, np.uint64),
                 start_hap=np.zeros(no * C_ * 2, np.uint8), mut_off=np.zeros(no * C_ + 1, np.uint64),
                 mut_bp=np.zeros(nm, np.uint64), mut_gam=np.zeros(nm, np.uint8))
        self._call("download_draws", self.ctx, pop, _ptr(d["father"], _u64p), _ptr(d["mother"], _u64p), _ptr(d["sex"], _u8p),
                   _ptr(d["xo_off"], _u64p), _ptr(d["xo_bp"], _u64p), _ptr(d["start_hap"], _u8p),
                   _ptr(d["mut_off"], _u64p), _ptr(d["mut_bp"], _u64p), _ptr(d["mut_gam"], _u8p))
        return d

    # -- measurement hooks (CUDA library only)
    def set_profiling(self, enabled=True):
        self._call("set_profiling", self.ctx, int(enabled))

    def kernel_time(self, kernel=GE_KERNEL_PROPAGATE_BITS):
        ms, n, b = C.c_double(), C.c_uint64(), C.c_uint64()
        self._call("get_kernel_time", self.ctx, kernel, C.byref(ms), C.byref(n), C.byref(b))
        return ms.value, n.value, b.value

    def reset_kernel_times(self):
        self._call("reset_kernel_times", self.ctx)

    def launch_count(self):
        n = C.c_uint64()
        self._call("get_launch_count", self.ctx, C.byref(n))
        return n.value

    def graph_replays(self):
        n = C.c_uint64()
        self._call("get_graph_replays", self.ctx, C.byref(n))
        return n.value

    def timer_start(self):
        self._call("timer_start", self.ctx)

    def timer_stop(self):
        ms = C.c_double()
        self._call("timer_stop", self.ctx, C.byref(ms))
        return ms.value

    def synchronize(self):
        self._call("synchronize", self.ctx)

    def device_memory_bytes(self):
        n = C.c_uint64()
        self._call("device_memory_bytes", self.ctx, C.byref(n))
        return n.value
