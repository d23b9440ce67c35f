// ge_oracle.cpp — CPU restatement of the GeneEvolve per-generation hot path.  TEST INFRASTRUCTURE ONLY:
// it is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, and is never
// imported, linked or executed by the product path (geneevolve_b200/, libgeneevolve_b200.so).
//
// Parity status: PINNED against tests/golden/*.npz (outputs of the real reference, see ge_oracle.h).
//
// Every function cites the reference lines it follows (paths relative to /root/reference).  Nothing is
// copied: the reference's AoS-of-nested-vectors model (src/Population.h) is restated on flat arrays that
// match the C-ABI in include/geneevolve_b200.h.  Three draw sources:
//   GO_RNG_REF     the reference's own engines (std::minstd_rand0, glibc rand(), libstdc++ distributions),
//                  seeded through the ras_glob_seed() topology of SURVEY.md §3.3 — used to pin the oracle;
//   GE_RNG_PHILOX  the counter-based streams the CUDA library uses (spec in DESIGN.md §RNG) — the oracle
//                  restates them so that GPU results can be checked bit for bit;
//   GE_RNG_REPLAY  draws supplied by the caller.
// Both chromosome representations are carried side by side (founder segments as the reference has them, and
// bit-packed haplotypes), so every test can also assert materialise(segments) == bits.

#include "ge_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

std::string g_err;
int fail(int code, const std::string &m) { g_err = m; return code; }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) and the stream layout shared with the CUDA library
// ------------------------------------------------------------------------------------------------
inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum Purpose : uint32_t {
    P_THIN = 1, P_RM_PAIR = 2, P_TRIM = 3, P_TEMPLATE = 4, P_POISSON = 5, P_REMAINDER = 6, P_XO = 7, P_MUT = 8,
    P_SEX = 9, P_ENOISE = 10, P_F0 = 11, P_COMMON = 12, P_MIGRATE = 13
};

inline double u01(uint32_t a, uint32_t b) {  // [0,1), 53 bits
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

struct Stream {
    uint32_t k0, k1;
    void draw(uint32_t purpose, int pop, int gen, uint64_t entity, uint32_t sub, uint32_t block, uint32_t w[4]) const {
        uint32_t c3 = (purpose << 24) | ((uint32_t)pop << 20) | ((uint32_t)gen & 0xFFFFFu);
        philox4x32_10(k0, k1, block, (uint32_t)entity, sub, c3, w);
    }
    void normal2(uint32_t purpose, int pop, int gen, uint64_t entity, uint32_t sub, double &z0, double &z1) const {
        uint32_t w[4];
        draw(purpose, pop, gen, entity, sub, 0, w);
        double u1 = 1.0 - u01(w[0], w[1]);  // (0,1]
        double u2 = u01(w[2], w[3]);
        double r = std::sqrt(-2.0 * std::log(u1));
        double th = 6.283185307179586476925286766559 * u2;
        z0 = r * std::cos(th);
        z1 = r * std::sin(th);
    }
};

// ------------------------------------------------------------------------------------------------
// data model (flat restatement of src/Population.h)
// ------------------------------------------------------------------------------------------------
struct Part {  // class part, src/Population.h:20-51 (gen0_indv is a function of root/hap and not stored)
    uint64_t st, en, hap, root;
    std::vector<uint64_t> mut;
    bool check_interval(uint64_t t) const { return st <= t && t < en; }  // :46-50
};
typedef std::vector<Part> PartList;

struct CvChr {
    std::vector<uint64_t> bp;
    std::vector<double> a, d;
    std::vector<uint8_t> val;  // [nhap][ncv]
    uint64_t nhap = 0;
};
struct Scheme { double va = 0, vd = 0, ve = 0, vc = 0, vf = 0, omega = 0, beta = 0, lambda = 0; };

struct Draws {
    std::vector<uint64_t> father, mother, xo_off, xo_bp, mut_off, mut_bp;
    std::vector<uint8_t> sex, start_hap, mut_gam;
    std::vector<double> e_raw, common;  // [nphen][n]
    void clear() { *this = Draws(); }
};

struct Pop {
    bool avoid_inbreeding = false, RM = false, has_mut = false;
    double MM = 0;
    std::vector<std::vector<uint64_t>> rmap_bp, mut_bp;
    std::vector<std::vector<double>> recom_prob, mut_rate;
    std::vector<uint64_t> bp_dist;
    std::vector<std::vector<CvChr>> cv;  // [phen][chr]
    std::vector<Scheme> scheme;
    std::vector<std::vector<uint8_t>> panel;  // [chr] bytes [nhap][nloci]
    uint64_t n_founder_haps = 0;
    // state
    uint64_t n = 0;
    std::vector<uint64_t> ids;  // n*7
    std::vector<uint8_t> sex;
    std::vector<double> A, D, G, C, E, F, P, mv, sv, svf, A_raw, D_raw;
    std::vector<PartList> segs;                 // slot (i*nchr + c)*2 + h
    std::vector<std::vector<uint32_t>> bits;    // [chr] packed [n*2][wc]
    std::vector<std::vector<uint64_t>> hmut;    // slot -> mutated positions of this haplotype (bit path)
    std::vector<std::vector<std::vector<uint8_t>>> cvbit, cvroot;  // [phen][chr] bytes [n*2][ncv]
    std::vector<double> prev_mv, prev_sv, prev_P, prev_F;  // Pop_phen_info, src/Simulation.h:22-29
    uint64_t prev_n = 0;
    std::vector<uint64_t> c_male, c_female;
    std::vector<uint8_t> c_inbreed;
    std::vector<int32_t> c_noff;
    std::vector<double> var_a0, var_d0, var_g0;
    double sv_mean0 = 0, sv_var0 = 0;
    Draws last;
};

double mean_(const std::vector<double> &x) {  // CommFunc::mean, src/CommFunc.cpp:38-45
    double s = 0;
    for (double v : x) s += v;
    return s / (double)x.size();
}
double var_(const std::vector<double> &x) {  // CommFunc::var, src/CommFunc.cpp:57-68
    size_t n = x.size();
    if (n <= 1) return 0.0;
    double mu = 0, s2 = 0;
    for (double v : x) mu += v;
    mu /= (double)n;
    for (double v : x) s2 += (v - mu) * (v - mu);
    return s2 / (double)(n - 1);
}
double var_slice(const std::vector<double> &x, uint64_t off, uint64_t n) {
    std::vector<double> t(x.begin() + off, x.begin() + off + n);
    return var_(t);
}

std::vector<uint64_t> ras_rank(const std::vector<double> &x) {  // CommFunc::ras_rank, src/CommFunc.cpp:152-161
    // O(n^2) in the reference; the same zero-based ranks are the stable argsort positions
    size_t n = x.size();
    std::vector<uint64_t> idx(n), r(n);
    for (size_t i = 0; i < n; i++) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return x[a] < x[b]; });
    for (size_t k = 0; k < n; k++) r[idx[k]] = k;
    return r;
}

int myrandom(int i) { return std::rand() % i; }  // src/Simulation.cpp:11-14

struct IndMV { uint64_t ind; double mating_value; };  // Ind_MatingValue, src/Population.h:156-161
bool ValueCmp(IndMV const &a, IndMV const &b) { return a.mating_value < b.mating_value; }  // :28-31

}  // namespace

struct go_ctx {
    ge_config cfg;
    int rng_mode;
    std::vector<Pop> pop;
    std::vector<std::vector<uint64_t>> loci;  // [chr] positions
    std::vector<double> gamma;
    std::default_random_engine glob_generator;  // src/Simulation.h:137
    Stream stream;
    std::vector<uint32_t> chr_ids;
    ge_allreduce_fn allreduce = nullptr; void *allreduce_user = nullptr;
    std::vector<std::vector<uint64_t>> mig_sample;  // replay: supplied migrant positions per source population
    std::default_random_engine mig_engine;  // the `static` engine of ras_SampleWithoutReplacement
    bool mig_engine_seeded = false;

    unsigned ras_glob_seed() {  // src/Simulation.cpp:17-21
        std::uniform_int_distribution<unsigned> distribution(1, 1000000);
        return distribution(glob_generator);
    }
    int nchr() const { return cfg.n_chr; }
    int nphen() const { return cfg.n_phen; }
    uint64_t wc(int c) const { return (loci[c].size() + 31) / 32; }

    // ---------------- generation 0 ----------------
    void resize_state(Pop &P, uint64_t n) {  // ras_allocate_memory_for_humans :2366-2392
        int nf = nphen();
        P.n = n;
        P.ids.assign(n * 7, 0);
        P.sex.assign(n, 0);
        for (auto *v : {&P.A, &P.D, &P.G, &P.C, &P.E, &P.F, &P.P, &P.A_raw, &P.D_raw}) v->assign((uint64_t)nf * n, 0.0);
        for (auto *v : {&P.mv, &P.sv, &P.svf}) v->assign(n, 0.0);
    }

    int initial_human_gen0(int p, const ge_draws *d0) {  // ras_initial_human_gen0 :3000-3072
        Pop &P = pop[p];
        unsigned seed = 0;
        if (rng_mode == GO_RNG_REF) { seed = ras_glob_seed(); std::srand(seed); }
        uint64_t nhaps = P.cv[0][0].nhap;  // :3009
        uint64_t n = nhaps / 2;
        resize_state(P, n);
        int C = nchr(), nf = nphen();
        P.segs.assign(n * C * 2, PartList());
        P.hmut.assign(n * C * 2, {});
        P.bits.assign(C, {});
        for (int c = 0; c < C; c++) P.bits[c].assign(n * 2 * wc(c), 0u);
        for (uint64_t i = 0; i < n; i++) {
            for (int c = 0; c < C; c++) {
                uint64_t st = P.rmap_bp[c].front(), en = P.rmap_bp[c].back();  // :3029-3030
                for (int h = 0; h < 2; h++) {
                    Part q; q.st = st; q.en = en; q.hap = 2 * i + h; q.root = (uint64_t)p;
                    P.segs[(i * C + c) * 2 + h].push_back(q);
                }
            }
            if (rng_mode == GO_RNG_REF) P.sex[i] = (uint8_t)((std::rand() % 2) + 1);  // :3036
            else if (rng_mode == GE_RNG_PHILOX) { uint32_t w[4]; stream.draw(P_SEX, p, 0, i, 0, 0, w); P.sex[i] = (uint8_t)((w[0] & 1u) + 1); }
            else P.sex[i] = d0 ? d0[p].sex[i] : 1;
            for (int k = 0; k < 7; k++) P.ids[i * 7 + k] = i;  // :3037-3043
        }
        // bit-packed generation 0: founder bits of covered loci only (a locus outside [rmap.bp[0], rmap.bp[last])
        // is inside no part and therefore reads 0 in ras_convert_interval_to_hap_matrix :1186-1230)
        for (int c = 0; c < C; c++) {
            if (P.panel.empty() || P.panel[c].empty()) continue;
            uint64_t nl = loci[c].size(), w = wc(c);
            uint64_t st = P.rmap_bp[c].front(), en = P.rmap_bp[c].back();
            for (uint64_t r = 0; r < 2 * n; r++)
                for (uint64_t s = 0; s < nl; s++)
                    if (loci[c][s] >= st && loci[c][s] < en && P.panel[c][r * nl + s]) P.bits[c][r * w + s / 32] |= 1u << (s % 32);
        }
        // CV planes for the bit path
        P.cvbit.assign(nf, std::vector<std::vector<uint8_t>>(C));
        P.cvroot.assign(nf, std::vector<std::vector<uint8_t>>(C));
        for (int f = 0; f < nf; f++)
            for (int c = 0; c < C; c++) {
                CvChr &cv = P.cv[f][c];
                uint64_t ncv = cv.bp.size();
                P.cvbit[f][c].assign(2 * n * ncv, 0);
                P.cvroot[f][c].assign(2 * n * ncv, (uint8_t)p);
                uint64_t st = P.rmap_bp[c].front(), en = P.rmap_bp[c].back();
                for (uint64_t r = 0; r < 2 * n; r++)
                    for (uint64_t k = 0; k < ncv; k++)
                        if (cv.bp[k] >= st && cv.bp[k] < en) P.cvbit[f][c][r * ncv + k] = cv.val[r * ncv + k];
            }
        // sibling-common effect for generation 0 :3054-3066
        for (int f = 0; f < nf; f++) {
            if (P.scheme[f].vc > 0) {
                if (rng_mode == GO_RNG_REF) {
                    unsigned s2 = ras_glob_seed();
                    std::default_random_engine generator(s2);
                    std::normal_distribution<double> distribution(0.0, std::sqrt(P.scheme[f].vc));
                    for (uint64_t i = 0; i < n; i++) P.C[(uint64_t)f * n + i] = distribution(generator);
                } else if (rng_mode == GE_RNG_PHILOX) {
                    for (uint64_t i = 0; i < n; i++) {
                        double z0, z1; stream.normal2(P_COMMON, p, 0, i, (uint32_t)f, z0, z1);
                        P.C[(uint64_t)f * n + i] = z0 * std::sqrt(P.scheme[f].vc);
                    }
                } else if (d0 && d0[p].common) {
                    for (uint64_t i = 0; i < n; i++) P.C[(uint64_t)f * n + i] = d0[p].common[(uint64_t)f * n + i];
                }
            }
        }
        (void)seed;
        return GE_OK;
    }

    // ---------------- mating ----------------
    int random_mate_ref(int p, const ge_gen_params &gp) {  // random_mate :2090-2157
        Pop &P = pop[p];
        unsigned seed = ras_glob_seed();
        std::default_random_engine generator(seed);
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        std::vector<uint64_t> pos_male, pos_female;
        for (uint64_t i = 0; i < P.n; i++) {
            double r = distribution(generator);
            if (r < P.svf[i]) {
                if (P.sex[i] == 1) pos_male.push_back(i);
                else if (P.sex[i] == 2) pos_female.push_back(i);
            }
        }
        if (pos_male.empty() || pos_female.empty()) return fail(GE_ERR_NO_MATES, "No one can marry");
        std::default_random_engine g_uint_f(seed + 1), g_uint_m(seed + 2);
        std::uniform_int_distribution<unsigned long int> d_uint_f(0, pos_male.size() - 1), d_uint_m(0, pos_female.size() - 1);
        uint64_t nc = gp.pop_size;
        P.c_male.assign(nc, 0); P.c_female.assign(nc, 0); P.c_inbreed.assign(nc, 0); P.c_noff.assign(nc, 1);
        for (uint64_t i = 0; i < nc; i++) {
            unsigned long int i_f = d_uint_f(g_uint_f), i_m = d_uint_m(g_uint_m);
            P.c_male[i] = pos_male[i_f]; P.c_female[i] = pos_female[i_m];
        }
        return GE_OK;
    }

    // shared tail of assort_mate (:2296-2355): pairing by template rank, inbreeding check
    void pair_by_rank(Pop &P, const std::vector<IndMV> &males, const std::vector<IndMV> &females,
                      const std::vector<double> &t1, const std::vector<double> &t2, uint64_t &n_inbreed) {
        uint64_t n2 = t1.size();
        std::vector<uint64_t> r1 = ras_rank(t1), r2 = ras_rank(t2);
        P.c_male.assign(n2, 0); P.c_female.assign(n2, 0); P.c_inbreed.assign(n2, 0); P.c_noff.assign(n2, 0);
        n_inbreed = 0;
        for (uint64_t i = 0; i < n2; i++) {
            uint64_t pm = males[r1[i]].ind, pf = females[r2[i]].ind;
            P.c_male[i] = pm; P.c_female[i] = pf;
            if (P.avoid_inbreeding) {  // :2304-2320; ids: 0 ID,1 F,2 M,3 FF,4 FM,5 MF,6 MM
                const uint64_t *a = &P.ids[pm * 7], *b = &P.ids[pf * 7];
                bool sib = a[1] == b[1];
                bool cousin = (a[3] == b[3] || a[3] == b[5] || a[5] == b[3] || a[5] == b[5] ||
                               a[4] == b[4] || a[4] == b[6] || a[6] == b[4] || a[6] == b[6]);
                P.c_inbreed[i] = sib || cousin;
                if (P.c_inbreed[i]) n_inbreed++;
            }
        }
    }

    int assort_mate_ref(int p, const ge_gen_params &gp) {  // assort_mate :2167-2360
        Pop &P = pop[p];
        unsigned seed = ras_glob_seed();
        std::srand(seed);
        std::default_random_engine generator(ras_glob_seed());
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        std::vector<IndMV> males, females;
        for (uint64_t i = 0; i < P.n; i++) {
            double r = distribution(generator);
            if (r < P.svf[i]) {
                IndMV e{i, P.mv[i]};
                if (P.sex[i] == 1) { males.push_back(e); double r2 = distribution(generator); if (r2 < P.MM) males.push_back(e); }
                else if (P.sex[i] == 2) { females.push_back(e); double r2 = distribution(generator); if (r2 < P.MM) females.push_back(e); }
            }
        }
        uint64_t nm = males.size(), nfm = females.size();
        if (std::min(nm, nfm) == 0) return fail(GE_ERR_NO_MATES, "couples=0");
        if (nm > nfm) { std::random_shuffle(males.begin(), males.end()); males.erase(males.begin(), males.begin() + (nm - nfm)); }
        else if (nm < nfm) { std::random_shuffle(females.begin(), females.end()); females.erase(females.begin(), females.begin() + (nfm - nm)); }
        std::sort(males.begin(), males.end(), ValueCmp);
        std::sort(females.begin(), females.end(), ValueCmp);
        uint64_t n2 = std::min(males.size(), females.size());
        // RasRandomNumber::ras_mvnorm (src/RasRandomNumber.cpp:15-53) with the 2x2 upper Cholesky factor
        // U = [[1, rho], [0, sqrt(1 - rho*rho)]] (Eigen llt of [[1,rho],[rho,1]]), product as RasMatrix::ras_prod_mat
        std::vector<double> t1(n2), t2(n2);
        {
            std::default_random_engine gen(ras_glob_seed());
            std::normal_distribution<double> nd(0.0, 1.0);
            double rho = gp.mat_cor, u11 = std::sqrt(1.0 - rho * rho);
            for (uint64_t i = 0; i < n2; i++) {
                double z0 = nd(gen), z1 = nd(gen);
                double s0 = 0; s0 = s0 + z0 * 1.0; s0 = s0 + z1 * 0.0;
                double s1 = 0; s1 = s1 + z0 * rho; s1 = s1 + z1 * u11;
                t1[i] = 0.0 + s0; t2[i] = 0.0 + s1;
            }
        }
        uint64_t n_inbreed = 0;
        pair_by_rank(P, males, females, t1, t2, n_inbreed);
        std::vector<uint64_t> can_marry;
        if (!P.avoid_inbreeding) for (uint64_t i = 0; i < n2; i++) can_marry.push_back(i);  // :2314-2318 (left empty otherwise)
        if (gp.offspring_dist == 'p' || gp.offspring_dist == 'P') {  // :2329-2337, ras_rpois src/RasRandomNumber.cpp:57-67
            double lam = (double)gp.pop_size / (double)(n2 - n_inbreed);
            std::default_random_engine gen(ras_glob_seed());
            std::poisson_distribution<int> pd(lam);
            for (uint64_t i = 0; i < n2; i++) P.c_noff[i] = pd(gen);
        } else {  // :2338-2355
            int nfix = (int)std::floor((double)gp.pop_size / (double)(n2 - n_inbreed));
            for (uint64_t i = 0; i < n2; i++) P.c_noff[i] = nfix;
            uint64_t remain = gp.pop_size - (uint64_t)nfix * (n2 - n_inbreed);
            std::random_shuffle(can_marry.begin(), can_marry.end(), myrandom);
            if (remain > can_marry.size()) return fail(GE_ERR_UNSUPPORTED, "reference indexes past pos_couple_can_marry here (avoid_inbreeding with 'f')");
            for (uint64_t i = 0; i < remain; i++) P.c_noff[can_marry[i]]++;
        }
        return GE_OK;
    }

    static uint64_t key64(const uint32_t w[4]) { return ((uint64_t)w[0] << 32) | w[1]; }

    int poisson_philox(double lam, int p, int gen, uint64_t entity) const {
        // exact Poisson(lam) as a sum of independent Poisson(<=32) chunks, each by sequential-search inversion
        int total = 0; uint32_t blk = 0; double rem = lam;
        while (rem > 0) {
            double l = rem > 32.0 ? 32.0 : rem;
            rem -= l;
            uint32_t w[4]; stream.draw(P_POISSON, p, gen, entity, 0, blk++, w);
            double u = u01(w[0], w[1]);
            double pk = std::exp(-l), F = pk; int k = 0;
            while (u >= F && k < 400) { k++; pk *= l / (double)k; F += pk; }
            total += k;
        }
        return total;
    }

    int thin_philox(int p, int gen, bool with_mm, std::vector<IndMV> &males, std::vector<IndMV> &females) {
        Pop &P = pop[p];
        for (uint64_t i = 0; i < P.n; i++) {
            uint32_t w[4]; stream.draw(P_THIN, p, gen, i, 0, 0, w);
            double r = u01(w[0], w[1]), r2 = u01(w[2], w[3]);
            if (r < P.svf[i]) {
                IndMV e{i, P.mv[i]};
                std::vector<IndMV> *dst = P.sex[i] == 1 ? &males : (P.sex[i] == 2 ? &females : nullptr);
                if (!dst) continue;
                dst->push_back(e);
                if (with_mm && r2 < P.MM) dst->push_back(e);
            }
        }
        return GE_OK;
    }

    int random_mate_philox(int p, int gen, const ge_gen_params &gp) {
        Pop &P = pop[p];
        std::vector<IndMV> males, females;
        thin_philox(p, gen, false, males, females);
        if (males.empty() || females.empty()) return fail(GE_ERR_NO_MATES, "No one can marry");
        uint64_t nc = gp.pop_size;
        P.c_male.assign(nc, 0); P.c_female.assign(nc, 0); P.c_inbreed.assign(nc, 0); P.c_noff.assign(nc, 1);
        for (uint64_t k = 0; k < nc; k++) {
            uint32_t w[4]; stream.draw(P_RM_PAIR, p, gen, k, 0, 0, w);
            P.c_male[k] = males[(uint64_t)(((uint64_t)w[0] * (uint64_t)males.size()) >> 32)].ind;
            P.c_female[k] = females[(uint64_t)(((uint64_t)w[1] * (uint64_t)females.size()) >> 32)].ind;
        }
        return GE_OK;
    }

    // remove the n_remove entries with the smallest (key, position); survivors keep their order
    void trim_philox(std::vector<IndMV> &v, uint64_t n_remove, int p, int gen, uint32_t sub) {
        uint64_t n = v.size();
        std::vector<uint64_t> key(n), idx(n);
        for (uint64_t k = 0; k < n; k++) { uint32_t w[4]; stream.draw(P_TRIM, p, gen, k, sub, 0, w); key[k] = key64(w); idx[k] = k; }
        std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return key[a] < key[b]; });
        std::vector<uint8_t> drop(n, 0);
        for (uint64_t k = 0; k < n_remove; k++) drop[idx[k]] = 1;
        std::vector<IndMV> out;
        for (uint64_t k = 0; k < n; k++) if (!drop[k]) out.push_back(v[k]);
        v.swap(out);
    }

    int assort_mate_philox(int p, int gen, const ge_gen_params &gp) {
        Pop &P = pop[p];
        std::vector<IndMV> males, females;
        thin_philox(p, gen, true, males, females);
        uint64_t nm = males.size(), nfm = females.size();
        if (std::min(nm, nfm) == 0) return fail(GE_ERR_NO_MATES, "couples=0");
        if (nm > nfm) trim_philox(males, nm - nfm, p, gen, 0);
        else if (nm < nfm) trim_philox(females, nfm - nm, p, gen, 1);
        std::stable_sort(males.begin(), males.end(), ValueCmp);
        std::stable_sort(females.begin(), females.end(), ValueCmp);
        uint64_t n2 = males.size();
        std::vector<double> t1(n2), t2(n2);
        double rho = gp.mat_cor, u11 = std::sqrt(1.0 - rho * rho);
        for (uint64_t i = 0; i < n2; i++) {
            double z0, z1; stream.normal2(P_TEMPLATE, p, gen, i, 0, z0, z1);
            t1[i] = z0; t2[i] = z0 * rho + z1 * u11;
        }
        uint64_t n_inbreed = 0;
        pair_by_rank(P, males, females, t1, t2, n_inbreed);
        if (n2 == n_inbreed) return fail(GE_ERR_NO_MATES, "every couple is inbred");
        if (gp.offspring_dist == 'p' || gp.offspring_dist == 'P') {
            double lam = (double)gp.pop_size / (double)(n2 - n_inbreed);
            for (uint64_t i = 0; i < n2; i++) P.c_noff[i] = poisson_philox(lam, p, gen, i);
        } else {
            int nfix = (int)std::floor((double)gp.pop_size / (double)(n2 - n_inbreed));
            for (uint64_t i = 0; i < n2; i++) P.c_noff[i] = nfix;
            uint64_t remain = gp.pop_size - (uint64_t)nfix * (n2 - n_inbreed);
            // intent of :2338-2355 — the remainder goes to distinct random couples that can marry
            std::vector<uint64_t> cand, key;
            for (uint64_t i = 0; i < n2; i++) if (!P.c_inbreed[i]) { uint32_t w[4]; stream.draw(P_REMAINDER, p, gen, i, 0, 0, w); cand.push_back(i); key.push_back(key64(w)); }
            std::vector<uint64_t> o(cand.size());
            for (uint64_t k = 0; k < o.size(); k++) o[k] = k;
            std::stable_sort(o.begin(), o.end(), [&](uint64_t a, uint64_t b) { return key[a] < key[b]; });
            for (uint64_t k = 0; k < remain && k < o.size(); k++) P.c_noff[cand[o[k]]]++;
        }
        return GE_OK;
    }

    int mate(int p, int gen, const ge_gen_params &gp) {
        Pop &P = pop[p];
        if (rng_mode == GO_RNG_REF) return P.RM ? random_mate_ref(p, gp) : assort_mate_ref(p, gp);
        if (rng_mode == GE_RNG_PHILOX) return P.RM ? random_mate_philox(p, gen, gp) : assort_mate_philox(p, gen, gp);
        return fail(GE_ERR_INVALID, "replay mode: supply couples with go_set_couples or offspring draws");
    }

    // ---------------- gamete formation ----------------
    // ras_sim_loc_rec :2973-2995 (returns the list WITH its sentinels, like the reference)
    std::vector<uint64_t> ras_sim_loc_rec_ref(Pop &P, int c, unsigned seed) {
        std::srand(seed);
        std::default_random_engine generator(seed + 1);
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        std::vector<uint64_t> locs;
        locs.push_back(P.rmap_bp[c][0]);
        for (uint64_t j = 0; j < P.recom_prob[c].size(); j++) {
            double r = distribution(generator);
            if (r < P.recom_prob[c][j]) locs.push_back(P.rmap_bp[c][j] + (uint64_t)(std::rand() % P.bp_dist[c]));
        }
        locs.push_back(P.rmap_bp[c].back());
        return locs;
    }

    // Philox skip-sampler: the same law as one Bernoulli(p_j) per map row (exact up to fp64 rounding of the
    // survival table T[k] = prod_{i<k} (1-p_i)), drawing one uniform per crossover instead of one per row.
    static std::vector<double> survival_table(const std::vector<double> &p, uint64_t first) {
        std::vector<double> T(p.size() + 1, 1.0);
        for (uint64_t k = 0; k < p.size(); k++) {
            double q = k < first ? 0.0 : p[k];
            if (q < 0) q = 0; if (q > 1) q = 1;
            T[k + 1] = T[k] * (1.0 - q);
        }
        return T;
    }
    static int64_t next_success(const std::vector<double> &T, uint64_t j, double v) {
        // smallest k >= j with T[k+1] < v, or -1
        uint64_t R = T.size() - 1;
        if (j >= R || !(T[R] < v)) return -1;
        uint64_t lo = j, hi = R - 1;
        while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (T[mid + 1] < v) hi = mid; else lo = mid + 1; }
        return (int64_t)lo;
    }
    void sample_xo_philox(Pop &P, int p, int gen, uint64_t i, int c, int g, std::vector<uint64_t> &xo, int &start) {
        std::vector<double> T = survival_table(P.recom_prob[c], 0);
        uint64_t R = P.recom_prob[c].size(), j = 0; uint32_t blk = 0;
        xo.clear();
        for (;;) {
            uint32_t w[4]; stream.draw(P_XO, p, gen, i, chr_ids[c] * 2u + (uint32_t)g, blk, w);
            if (blk == 0) start = (int)(w[3] & 1u);
            blk++;
            if (j >= R) break;
            double v = (1.0 - u01(w[0], w[1])) * T[j];
            int64_t k = next_success(T, j, v);
            if (k < 0) break;
            xo.push_back(P.rmap_bp[c][k] + (((uint64_t)w[2] * P.bp_dist[c]) >> 32));
            j = (uint64_t)k + 1;
        }
    }
    void sample_mut_philox(Pop &P, int p, int gen, uint64_t i, int c, std::vector<uint64_t> &mbp, std::vector<uint8_t> &mg) {
        std::vector<double> T = survival_table(P.mut_rate[c], 1);  // rows start at 1 (:2506)
        uint64_t R = P.mut_rate[c].size(), j = 1; uint32_t blk = 0;
        for (;;) {
            if (j >= R) break;
            uint32_t w[4]; stream.draw(P_MUT, p, gen, i, chr_ids[c], blk++, w);
            double v = (1.0 - u01(w[0], w[1])) * T[j];
            int64_t k = next_success(T, j, v);
            if (k < 0) break;
            uint64_t st = P.mut_bp[c][k - 1], en = P.mut_bp[c][k];  // uniform_int[st,en] inclusive (:2517-2519)
            mbp.push_back(st + (((uint64_t)w[2] * (en - st + 1)) >> 32));
            mg.push_back((uint8_t)(w[3] & 1u));
            j = (uint64_t)k + 1;
        }
    }

    static void modify_part_for_mutation_pos(Part &q) {  // :2961-2970
        std::vector<uint64_t> ret;
        for (uint64_t m : q.mut) if (q.check_interval(m)) ret.push_back(m);
        q.mut = ret;
    }
    // recombine :2903-2958; hap[0], hap[1] are the two haplotypes of the parent's chromosome
    static PartList recombine(const PartList *hap, int starting_haplotype, const std::vector<uint64_t> &locs) {
        PartList ret;
        int hi = starting_haplotype;
        if (locs.size() < 3) return hap[hi];
        for (uint64_t i1 = 1; i1 < locs.size(); i1++) {
            const PartList &H = hap[hi];
            uint64_t L = locs[i1 - 1], R = locs[i1], i2 = 0;
            while (H.size() > i2 && H[i2].en <= L) i2++;
            if (H.size() > i2 && H[i2].st < L && L < H[i2].en && R < H[i2].en) {
                Part q = H[i2]; q.st = L; q.en = R; modify_part_for_mutation_pos(q); ret.push_back(q); i2++;
            }
            if (H.size() > i2 && H[i2].st < L && L < H[i2].en && R >= H[i2].en) {
                Part q = H[i2]; q.st = L; modify_part_for_mutation_pos(q); ret.push_back(q); i2++;
            }
            while (H.size() > i2 && H[i2].en <= R && L <= H[i2].st) {
                Part q = H[i2]; modify_part_for_mutation_pos(q); ret.push_back(q); i2++;
            }
            if (H.size() > i2 && H[i2].st < R && R < H[i2].en) {
                Part q = H[i2]; q.en = R; modify_part_for_mutation_pos(q); ret.push_back(q);
            }
            hi = (hi + 1) % 2;
        }
        return ret;
    }

    // bit-packed gamete: locus s of the offspring haplotype takes parental haplotype
    //   start ^ (#{crossovers b_k <= pos[s]} & 1)
    // which is what recombine() + ras_convert_interval_to_hap_matrix give locus by locus (piece k covers
    // [b_k, b_{k+1}) and the haplotype flips after every piece, :2955).
    void propagate_bits(const Pop &P, const std::vector<std::vector<uint32_t>> &pbits, uint64_t parent, int c,
                        int start, const uint64_t *xo, uint64_t nxo, uint32_t *dst) const {
        uint64_t nl = loci[c].size(), w = wc(c);
        const uint32_t *h0 = &pbits[c][(parent * 2 + 0) * w], *h1 = &pbits[c][(parent * 2 + 1) * w];
        std::fill(dst, dst + w, 0u);
        uint64_t k = 0; int cur = start;
        for (uint64_t s = 0; s < nl; s++) {
            while (k < nxo && xo[k] <= loci[c][s]) { cur ^= 1; k++; }
            const uint32_t *src = cur ? h1 : h0;
            dst[s / 32] |= ((src[s / 32] >> (s % 32)) & 1u) << (s % 32);
        }
        (void)P;
    }
    static int parity_at(int start, const uint64_t *xo, uint64_t nxo, uint64_t pos) {
        int cur = start;
        for (uint64_t k = 0; k < nxo; k++) if (xo[k] <= pos) cur ^= 1;
        return cur;
    }

    int reproduce(int p, int gen, const ge_draws *dr) {  // reproduce :2394-2493
        Pop &P = pop[p];
        int C = nchr(), nf = nphen();
        unsigned seed = 0;
        if (rng_mode == GO_RNG_REF) { seed = ras_glob_seed(); std::srand(seed); }
        // offspring -> parents
        std::vector<uint64_t> father, mother, couple;
        uint64_t n_couples = P.c_male.size();
        std::vector<double> val_common;  // [nf][n_couples]
        if (dr) {
            father.assign(dr->father, dr->father + dr->n_offspring);
            mother.assign(dr->mother, dr->mother + dr->n_offspring);
        } else {
            for (uint64_t it = 0; it < n_couples; it++)
                if (!P.c_inbreed[it]) for (int s = 0; s < P.c_noff[it]; s++) { father.push_back(P.c_male[it]); mother.push_back(P.c_female[it]); couple.push_back(it); }
            val_common.assign((uint64_t)nf * n_couples, 0.0);
            if (rng_mode == GO_RNG_REF) {  // :2417-2429
                std::default_random_engine generator(seed + 1);
                for (int f = 0; f < nf; f++)
                    if (P.scheme[f].vc > 0) {
                        std::normal_distribution<double> distribution(0.0, std::sqrt(P.scheme[f].vc));
                        for (uint64_t it = 0; it < n_couples; it++) val_common[(uint64_t)f * n_couples + it] = distribution(generator);
                    }
            } else {
                for (int f = 0; f < nf; f++)
                    if (P.scheme[f].vc > 0)
                        for (uint64_t it = 0; it < n_couples; it++) {
                            double z0, z1; stream.normal2(P_COMMON, p, gen, it, (uint32_t)f, z0, z1);
                            val_common[(uint64_t)f * n_couples + it] = z0 * std::sqrt(P.scheme[f].vc);
                        }
            }
        }
        uint64_t n_off = father.size();
        // parent generation snapshot
        std::vector<PartList> psegs; psegs.swap(P.segs);
        std::vector<std::vector<uint32_t>> pbits; pbits.swap(P.bits);
        std::vector<std::vector<uint64_t>> phmut; phmut.swap(P.hmut);
        std::vector<std::vector<std::vector<uint8_t>>> pcvbit, pcvroot; pcvbit.swap(P.cvbit); pcvroot.swap(P.cvroot);
        std::vector<uint64_t> pids = P.ids;
        uint64_t n_par = P.n;
        resize_state(P, n_off);
        P.segs.assign(n_off * C * 2, PartList());
        P.hmut.assign(n_off * C * 2, {});
        P.bits.assign(C, {});
        for (int c = 0; c < C; c++) P.bits[c].assign(n_off * 2 * wc(c), 0u);
        P.cvbit.assign(nf, std::vector<std::vector<uint8_t>>(C));
        P.cvroot.assign(nf, std::vector<std::vector<uint8_t>>(C));
        for (int f = 0; f < nf; f++) for (int c = 0; c < C; c++) {
            P.cvbit[f][c].assign(2 * n_off * P.cv[f][c].bp.size(), 0);
            P.cvroot[f][c].assign(2 * n_off * P.cv[f][c].bp.size(), 0);
        }
        Draws &L = P.last; L.clear();
        L.father = father; L.mother = mother; L.xo_off.push_back(0); L.mut_off.push_back(0);
        L.common.assign((uint64_t)nf * n_off, 0.0);

        for (uint64_t i = 0; i < n_off; i++) {
            uint64_t par[2] = {father[i], mother[i]};
            if (par[0] >= n_par || par[1] >= n_par) return fail(GE_ERR_INVALID, "parent index out of range");
            for (int c = 0; c < C; c++) {
                PartList gam[2];
                std::vector<uint64_t> xo[2]; int start[2] = {0, 0};
                for (int g = 0; g < 2; g++) {
                    std::vector<uint64_t> locs;
                    if (rng_mode == GO_RNG_REF) {  // :2447-2456
                        unsigned seed_loc = std::rand();
                        locs = ras_sim_loc_rec_ref(P, c, seed_loc);
                        start[g] = std::rand() % 2;
                        xo[g].assign(locs.begin() + 1, locs.end() - 1);
                    } else {
                        if (dr) {
                            uint64_t slot = (i * C + c) * 2 + g;
                            xo[g].assign(dr->xo_bp + dr->xo_off[slot], dr->xo_bp + dr->xo_off[slot + 1]);
                            start[g] = dr->start_hap[slot];
                        } else sample_xo_philox(P, p, gen, i, c, g, xo[g], start[g]);
                        locs.push_back(P.rmap_bp[c][0]);
                        locs.insert(locs.end(), xo[g].begin(), xo[g].end());
                        locs.push_back(P.rmap_bp[c].back());
                    }
                    gam[g] = recombine(&psegs[(par[g] * C + c) * 2], start[g], locs);
                    // bit path
                    propagate_bits(P, pbits, par[g], c, start[g], xo[g].data(), xo[g].size(), &P.bits[c][(i * 2 + g) * wc(c)]);
                    for (int f = 0; f < nf; f++) {
                        CvChr &cv = P.cv[f][c]; uint64_t ncv = cv.bp.size();
                        for (uint64_t k = 0; k < ncv; k++) {
                            int h = parity_at(start[g], xo[g].data(), xo[g].size(), cv.bp[k]);
                            P.cvbit[f][c][(i * 2 + g) * ncv + k] = pcvbit[f][c][(par[g] * 2 + h) * ncv + k];
                            P.cvroot[f][c][(i * 2 + g) * ncv + k] = pcvroot[f][c][(par[g] * 2 + h) * ncv + k];
                        }
                    }
                    std::vector<uint64_t> &hm = P.hmut[(i * C + c) * 2 + g];
                    for (int h = 0; h < 2; h++)
                        for (uint64_t m : phmut[(par[g] * C + c) * 2 + h])
                            if (parity_at(start[g], xo[g].data(), xo[g].size(), m) == h) hm.push_back(m);
                    L.xo_bp.insert(L.xo_bp.end(), xo[g].begin(), xo[g].end());
                    L.xo_off.push_back(L.xo_bp.size());
                    L.start_hap.push_back((uint8_t)start[g]);
                }
                // mutation: ras_add_mutation :2497-2552
                if (P.has_mut) {
                    std::vector<uint64_t> mbp; std::vector<uint8_t> mg;
                    if (rng_mode == GO_RNG_REF) {
                        unsigned ms = ras_glob_seed();
                        std::srand(ms);
                        std::default_random_engine generator(ms + 1), generator_u(ms + 2);
                        std::uniform_real_distribution<double> distribution(0.0, 1.0);
                        for (unsigned k = 1; k < P.mut_rate[c].size(); k++) {
                            double r = distribution(generator_u);
                            if (r < P.mut_rate[c][k]) {
                                std::uniform_int_distribution<unsigned long int> dpos(P.mut_bp[c][k - 1], P.mut_bp[c][k]);
                                mbp.push_back(dpos(generator));
                                mg.push_back((uint8_t)(std::rand() % 2));
                            }
                        }
                    } else if (dr) {
                        if (dr->mut_off) for (uint64_t k = dr->mut_off[i * C + c]; k < dr->mut_off[i * C + c + 1]; k++) { mbp.push_back(dr->mut_bp[k]); mg.push_back(dr->mut_gam[k]); }
                    } else sample_mut_philox(P, p, gen, i, c, mbp, mg);
                    for (uint64_t k = 0; k < mbp.size(); k++) {
                        int g = mg[k]; uint64_t bpm = mbp[k];
                        bool landed = false;
                        for (Part &q : gam[g]) if (q.check_interval(bpm)) { q.mut.push_back(bpm); landed = true; }
                        if (landed) { L.mut_bp.push_back(bpm); L.mut_gam.push_back((uint8_t)g); }
                        // bit path: inside the covered range a hit toggles the allele once per lineage (the
                        // reference looks positions up with std::find, :1218-1222 / :2770-2774)
                        if (bpm >= P.rmap_bp[c].front() && bpm < P.rmap_bp[c].back()) {
                            std::vector<uint64_t> &hm = P.hmut[(i * C + c) * 2 + g];
                            bool seen = std::find(hm.begin(), hm.end(), bpm) != hm.end();
                            hm.push_back(bpm);
                            if (!seen) {
                                uint64_t w = wc(c);
                                for (uint64_t s = 0; s < loci[c].size(); s++)
                                    if (loci[c][s] == bpm) P.bits[c][(i * 2 + g) * w + s / 32] ^= 1u << (s % 32);
                                for (int f = 0; f < nf; f++) {
                                    CvChr &cv = P.cv[f][c]; uint64_t ncv = cv.bp.size();
                                    for (uint64_t kk = 0; kk < ncv; kk++) if (cv.bp[kk] == bpm) P.cvbit[f][c][(i * 2 + g) * ncv + kk] ^= 1;
                                }
                            }
                        }
                    }
                }
                L.mut_off.push_back(L.mut_bp.size());
                P.segs[(i * C + c) * 2 + 0] = gam[0];
                P.segs[(i * C + c) * 2 + 1] = gam[1];
            }
            // :2471-2484
            if (rng_mode == GO_RNG_REF) P.sex[i] = (uint8_t)((std::rand() % 2) + 1);
            else if (dr) P.sex[i] = dr->sex[i];
            else { uint32_t w[4]; stream.draw(P_SEX, p, gen, i, 0, 0, w); P.sex[i] = (uint8_t)((w[0] & 1u) + 1); }
            uint64_t *id = &P.ids[i * 7];
            const uint64_t *fa = &pids[par[0] * 7], *mo = &pids[par[1] * 7];
            id[0] = i; id[1] = fa[0]; id[2] = mo[0]; id[3] = fa[1]; id[4] = fa[2]; id[5] = mo[1]; id[6] = mo[2];
            for (int f = 0; f < nf; f++) {
                double cval = dr ? (dr->common ? dr->common[(uint64_t)f * n_off + i] : 0.0) : val_common[(uint64_t)f * n_couples + couple[i]];
                P.C[(uint64_t)f * n_off + i] = cval;
                L.common[(uint64_t)f * n_off + i] = cval;
            }
        }
        L.sex = P.sex;
        return GE_OK;
    }

    // ---------------- genetic values ----------------
    // ras_find_cv :2752-2815 for one haplotype
    void find_cv(const Pop &P, uint64_t i, int c, int f, int h, std::vector<uint8_t> &cvv, std::vector<double> &ga, std::vector<double> &gd) const {
        uint64_t ncv = P.cv[f][c].bp.size();
        cvv.assign(ncv, 0); ga.assign(ncv, 0.0); gd.assign(ncv, 0.0);
        for (const Part &q : P.segs[(i * nchr() + c) * 2 + h]) {
            const CvChr &rc = pop[q.root].cv[f][c];
            for (uint64_t k = 0; k < ncv; k++) {
                uint64_t bp = rc.bp[k];
                if (q.check_interval(bp)) {
                    uint8_t v = rc.val[q.hap * rc.bp.size() + k];
                    if (std::find(q.mut.begin(), q.mut.end(), bp) != q.mut.end()) v = !v;
                    cvv[k] = v; ga[k] = rc.a[k]; gd[k] = rc.d[k];
                }
            }
        }
    }

    int compute_AD(int p, int gen) {  // ras_compute_AD :2624-2749
        Pop &P = pop[p];
        int C = nchr(), nf = nphen();
        uint64_t n = P.n;
        std::vector<double> Achr((uint64_t)nf * C * n), Dchr((uint64_t)nf * C * n);
        for (int f = 0; f < nf; f++)
            for (int c = 0; c < C; c++) {
                uint64_t ncv = P.cv[f][c].bp.size();
                std::vector<uint8_t> cv0(n * ncv), cv1(n * ncv);
                std::vector<double> a0(n * ncv), a1(n * ncv), d0(n * ncv), d1(n * ncv);
                std::vector<uint8_t> tv; std::vector<double> ta, td;
                for (uint64_t i = 0; i < n; i++) {
                    find_cv(P, i, c, f, 0, tv, ta, td);
                    std::copy(tv.begin(), tv.end(), cv0.begin() + i * ncv); std::copy(ta.begin(), ta.end(), a0.begin() + i * ncv); std::copy(td.begin(), td.end(), d0.begin() + i * ncv);
                    find_cv(P, i, c, f, 1, tv, ta, td);
                    std::copy(tv.begin(), tv.end(), cv1.begin() + i * ncv); std::copy(ta.begin(), ta.end(), a1.begin() + i * ncv); std::copy(td.begin(), td.end(), d1.begin() + i * ncv);
                }
                std::vector<double> frq(ncv);
                for (uint64_t k = 0; k < ncv; k++) {
                    double fsum = 0;
                    for (uint64_t i = 0; i < n; i++) fsum += cv0[i * ncv + k] + cv1[i * ncv + k];
                    frq[k] = fsum / (2 * n);
                }
                for (uint64_t i = 0; i < n; i++) {
                    double A_chr = 0, D_chr = 0;
                    for (uint64_t k = 0; k < ncv; k++) {
                        double a = (a0[i * ncv + k] + a1[i * ncv + k]) / 2;
                        double d = (d0[i * ncv + k] + d1[i * ncv + k]) / 2;
                        if (P.scheme[f].vd == 0) d = 0;
                        unsigned t = cv0[i * ncv + k] + cv1[i * ncv + k];
                        double pp = frq[k], q = 1 - pp;
                        double alpha = a + d * (q - pp);
                        A_chr += ((double)t - 2 * pp) * alpha;
                        double c_t[3] = {-2 * pp * pp, 2 * pp * q, -2 * q * q};
                        D_chr += c_t[t] * d;
                    }
                    Achr[((uint64_t)f * C + c) * n + i] = A_chr;
                    Dchr[((uint64_t)f * C + c) * n + i] = D_chr;
                    if (std::isnan(A_chr) || std::isnan(D_chr)) return fail(GE_ERR_NAN, "A or D is nan");
                }
            }
        for (uint64_t i = 0; i < n; i++)
            for (int f = 0; f < nf; f++) {
                double bv = 0, add = 0, dom = 0;
                for (int c = 0; c < C; c++) {
                    double a = Achr[((uint64_t)f * C + c) * n + i], d = Dchr[((uint64_t)f * C + c) * n + i];
                    add += a; dom += d; bv += a + d;
                }
                P.A[(uint64_t)f * n + i] = add; P.D[(uint64_t)f * n + i] = dom; P.G[(uint64_t)f * n + i] = bv;
                P.A_raw[(uint64_t)f * n + i] = add; P.D_raw[(uint64_t)f * n + i] = dom;
            }
        if (allreduce) {  // chromosome-sharded: sum the partial genetic values over the ranks
            std::vector<double> buf;
            buf.insert(buf.end(), P.A.begin(), P.A.end()); buf.insert(buf.end(), P.D.begin(), P.D.end()); buf.insert(buf.end(), P.G.begin(), P.G.end());
            if (allreduce(allreduce_user, buf.data(), buf.size(), nullptr) != 0) return fail(GE_ERR_INVALID, "allreduce hook failed");
            uint64_t m = P.A.size();
            std::copy(buf.begin(), buf.begin() + m, P.A.begin()); std::copy(buf.begin() + m, buf.begin() + 2 * m, P.D.begin()); std::copy(buf.begin() + 2 * m, buf.end(), P.G.begin());
            P.A_raw = P.A; P.D_raw = P.D;
        }
        (void)gen;
        return GE_OK;
    }

    int scale_AD_compute_GEF(int p, int gen, int f, const double *e_in, const double *f0_in = nullptr) {  // ras_scale_AD_compute_GEF :3075-3206
        Pop &P = pop[p];
        uint64_t n = P.n;
        Scheme &S = P.scheme[f];
        std::vector<double> e(n), par_eff(n, 0.0);
        unsigned seed = 0;
        if (rng_mode == GO_RNG_REF) seed = ras_glob_seed();
        std::default_random_engine generator_e(seed), generator_f(seed + 1);
        std::normal_distribution<double> distribution_e(0.0, 1), distribution_f(0.0, std::sqrt(S.vf));
        for (uint64_t i = 0; i < n; i++) {
            if (rng_mode == GO_RNG_REF) e[i] = distribution_e(generator_e);
            else if (e_in) e[i] = e_in[i];
            else { double z0, z1; stream.normal2(P_ENOISE, p, gen, i, (uint32_t)f, z0, z1); e[i] = z0; }
            if (gen == 0) {
                if (S.vf > 0) {
                    if (rng_mode == GO_RNG_REF) par_eff[i] = distribution_f(generator_f);
                    else if (rng_mode == GE_RNG_PHILOX) { double z0, z1; stream.normal2(P_F0, p, 0, i, (uint32_t)f, z0, z1); par_eff[i] = z0 * std::sqrt(S.vf); }
                    else par_eff[i] = f0_in ? f0_in[i] : 0.0;
                }
            } else {
                uint64_t ind_f = P.ids[i * 7 + 1], ind_m = P.ids[i * 7 + 2];
                double ff = 0, fm = 0;
                const std::vector<double> &src = cfg.vt_type == 1 ? P.prev_P : P.prev_F;
                // The reference reads _Pop_info_prev_gen by parent ID unconditionally (:3118-3133) although the
                // arrays are stored by position (:3211-3236); after migration an ID can exceed the array (UB in
                // the reference).  The value only matters when vf > 0, so only then is it read (and checked).
                if ((cfg.vt_type == 1 || cfg.vt_type == 2) && S.vf > 0) {
                    if (ind_f >= P.prev_n || ind_m >= P.prev_n) return fail(GE_ERR_INVALID, "parent ID outside previous generation (reference reads out of bounds here)");
                    ff = src[(uint64_t)f * P.prev_n + ind_f]; fm = src[(uint64_t)f * P.prev_n + ind_m];
                }
                par_eff[i] = S.beta * (ff + fm);
            }
        }
        std::copy(e.begin(), e.end(), P.last.e_raw.begin() + (uint64_t)f * n);
        double s_a = 1;
        if (S.va > 0) s_a = std::sqrt(P.var_a0[f] / S.va);
        else if (S.va == -1) s_a = 1;
        double s_d = 0;
        if (S.vd > 0) s_d = std::sqrt(P.var_d0[f] / S.vd);
        else if (S.vd == -1) s_d = 1;
        double s_ev = 0;
        if (S.ve > 0) s_ev = std::sqrt(var_(e) / S.ve);
        for (uint64_t i = 0; i < n; i++) {
            uint64_t o = (uint64_t)f * n + i;
            P.E[o] = s_ev > 0 ? e[i] / s_ev : 0;
            P.A[o] = P.A[o] / s_a;
            P.D[o] = s_d > 0 ? P.D[o] / s_d : 0;
            P.G[o] = P.A[o] + P.D[o];
            P.F[o] = S.vf > 0 ? par_eff[i] : 0;
            P.P[o] = P.A[o] + P.D[o] + P.C[o] + P.E[o] + P.F[o];
        }
        return GE_OK;
    }

    double combined_variance(int f, double a) {  // ras_combined_variance :3254-3282
        std::vector<double> x, y;
        int np = cfg.n_pop;
        for (int p = 0; p < np; p++) {
            double bi = a * (2 * p / (np - 1) - 1);  // integer arithmetic, as in the reference
            for (uint64_t j = 0; j < pop[p].n; j++) { double v = pop[p].P[(uint64_t)f * pop[p].n + j]; x.push_back(v); y.push_back(v + bi); }
        }
        return var_(y) - (1 + gamma[f]) * var_(x);
    }
    double newton(int f, double x0, double precision, int depth) {  // NewtonRaphson :44-63, Fprime :35-39
        double fx0 = combined_variance(f, x0);
        const double dx = 0.001;
        double fp = (combined_variance(f, x0 + dx) - combined_variance(f, x0 - dx)) / (2 * dx);
        double x1 = x0 - fx0 / fp;
        double fx1 = combined_variance(f, x1);
        if (std::abs(fx1) < precision || depth > 200) return x1;
        return newton(f, x1, precision, depth + 1);
    }
    int env_effects(int f) {  // sim_environmental_effects_specific_to_each_population :3345-3381
        if (gamma.empty() || gamma[f] == 0) return GE_OK;
        if (cfg.n_pop < 2) return fail(GE_ERR_UNSUPPORTED, "--gamma with one population divides by zero in the reference (:3269)");
        double ah = newton(f, 10, 1e-4, 0);
        int np = cfg.n_pop;
        for (int p = 0; p < np; p++) {  // ras_add_environmental_effects_specific_to_each_population :3285-3297
            double gi = ah * (2 * p / (np - 1) - 1);
            for (uint64_t j = 0; j < pop[p].n; j++) pop[p].P[(uint64_t)f * pop[p].n + j] += gi;
        }
        return GE_OK;
    }

    static double selection_func(int gen, const ge_gen_params &gp, double z) {  // ras_selection_func :3386-3428
        if (gen == 0) return 1;
        switch (gp.selection_func) {
            case GE_SEL_DEFAULT: { double y = std::exp(0.0 + 1.0 * z); return y / (1 + y); }
            case GE_SEL_LOGIT: { double y = std::exp(gp.selection_par1 + gp.selection_par2 * z); return y / (1 + y); }
            case GE_SEL_PROBIT: return .5 * (1 + std::erf((z - gp.selection_par1) / (std::sqrt(2) * gp.selection_par2)));  // CommFunc::NormalCDF
            case GE_SEL_STAB: {  // CommFunc::NormalPDF with pi = 3.1415926 (src/CommFunc.cpp:4,266-270)
                const double pi = 3.1415926;
                return 1 / (std::sqrt(2.0 * pi) * gp.selection_par2) * std::exp(-0.5 * std::pow((z - gp.selection_par1) / gp.selection_par2, 2));
            }
            case GE_SEL_THR: return z <= gp.selection_par2 ? gp.selection_par1 : 1.0;
        }
        return 1;
    }
    int compute_mv_sv(int p, int gen, const ge_gen_params *gp) {  // ras_compute_mating_value_selection_value :3300-3342
        Pop &P = pop[p];
        uint64_t n = P.n; int nf = nphen();
        std::vector<double> x_sv(n);
        for (uint64_t i = 0; i < n; i++) {
            double mv = 0, sv = 0;
            for (int f = 0; f < nf; f++) { mv += P.scheme[f].omega * P.P[(uint64_t)f * n + i]; sv += P.scheme[f].lambda * P.P[(uint64_t)f * n + i]; }
            P.mv[i] = mv; x_sv[i] = sv;
        }
        if (gen == 0) { P.sv_var0 = var_(x_sv); P.sv_mean0 = mean_(x_sv); }
        ge_gen_params dummy{}; if (!gp) gp = &dummy;
        for (uint64_t i = 0; i < n; i++) {
            double z = x_sv[i] - P.sv_mean0;
            if (P.sv_var0 > 0) z = (x_sv[i] - P.sv_mean0) / std::sqrt(P.sv_var0);
            P.sv[i] = z;
            P.svf[i] = selection_func(gen, *gp, z);
        }
        return GE_OK;
    }

    void save_prev(int p) {  // ras_save_human_info_to_Pop_info_prev_gen :3211-3236
        Pop &P = pop[p];
        P.prev_n = P.n; P.prev_mv = P.mv; P.prev_sv = P.sv; P.prev_P = P.P; P.prev_F = P.F;
    }

    // ---------------- migration ----------------
    struct Indiv {  // one Human moved as a whole
        uint64_t ids[7]; uint8_t sex; std::vector<double> v;  // A D G C E F P A_raw D_raw per phen, then mv sv svf
        std::vector<PartList> segs; std::vector<std::vector<uint64_t>> hmut;
        std::vector<std::vector<uint32_t>> bits;  // [chr][2*wc]
        std::vector<std::vector<uint8_t>> cvbit, cvroot;  // [f*C+c][2*ncv]
    };
    Indiv take(const Pop &P, uint64_t i) const {
        Indiv h; int C = nchr(), nf = nphen();
        std::copy(&P.ids[i * 7], &P.ids[i * 7] + 7, h.ids); h.sex = P.sex[i];
        for (const std::vector<double> *a : {&P.A, &P.D, &P.G, &P.C, &P.E, &P.F, &P.P, &P.A_raw, &P.D_raw}) for (int f = 0; f < nf; f++) h.v.push_back((*a)[(uint64_t)f * P.n + i]);
        h.v.push_back(P.mv[i]); h.v.push_back(P.sv[i]); h.v.push_back(P.svf[i]);
        for (int c = 0; c < C; c++) {
            for (int hh = 0; hh < 2; hh++) { h.segs.push_back(P.segs[(i * C + c) * 2 + hh]); h.hmut.push_back(P.hmut[(i * C + c) * 2 + hh]); }
            uint64_t w = wc(c);
            h.bits.emplace_back(P.bits[c].begin() + i * 2 * w, P.bits[c].begin() + (i + 1) * 2 * w);
        }
        for (int f = 0; f < nf; f++) for (int c = 0; c < C; c++) {
            uint64_t ncv = P.cv[f][c].bp.size();
            h.cvbit.emplace_back(P.cvbit[f][c].begin() + i * 2 * ncv, P.cvbit[f][c].begin() + (i + 1) * 2 * ncv);
            h.cvroot.emplace_back(P.cvroot[f][c].begin() + i * 2 * ncv, P.cvroot[f][c].begin() + (i + 1) * 2 * ncv);
        }
        return h;
    }
    void rebuild(Pop &P, const std::vector<Indiv> &hs) {
        int C = nchr(), nf = nphen(); uint64_t n = hs.size();
        resize_state(P, n);
        P.segs.assign(n * C * 2, PartList()); P.hmut.assign(n * C * 2, {});
        for (int c = 0; c < C; c++) P.bits[c].assign(n * 2 * wc(c), 0u);
        for (int f = 0; f < nf; f++) for (int c = 0; c < C; c++) { P.cvbit[f][c].assign(2 * n * P.cv[f][c].bp.size(), 0); P.cvroot[f][c].assign(2 * n * P.cv[f][c].bp.size(), 0); }
        for (uint64_t i = 0; i < n; i++) {
            const Indiv &h = hs[i];
            std::copy(h.ids, h.ids + 7, &P.ids[i * 7]); P.sex[i] = h.sex;
            size_t k = 0;
            for (std::vector<double> *a : {&P.A, &P.D, &P.G, &P.C, &P.E, &P.F, &P.P, &P.A_raw, &P.D_raw}) for (int f = 0; f < nf; f++) (*a)[(uint64_t)f * n + i] = h.v[k++];
            P.mv[i] = h.v[k++]; P.sv[i] = h.v[k++]; P.svf[i] = h.v[k++];
            for (int c = 0; c < C; c++) {
                for (int hh = 0; hh < 2; hh++) { P.segs[(i * C + c) * 2 + hh] = h.segs[c * 2 + hh]; P.hmut[(i * C + c) * 2 + hh] = h.hmut[c * 2 + hh]; }
                std::copy(h.bits[c].begin(), h.bits[c].end(), P.bits[c].begin() + i * 2 * wc(c));
            }
            for (int f = 0; f < nf; f++) for (int c = 0; c < C; c++) {
                uint64_t ncv = P.cv[f][c].bp.size();
                std::copy(h.cvbit[f * C + c].begin(), h.cvbit[f * C + c].end(), P.cvbit[f][c].begin() + i * 2 * ncv);
                std::copy(h.cvroot[f * C + c].begin(), h.cvroot[f * C + c].end(), P.cvroot[f][c].begin() + i * 2 * ncv);
            }
        }
    }

    int do_migration(int gen, const double *row) {  // ras_do_migration :877-989
        int np = cfg.n_pop;
        std::vector<std::vector<uint64_t>> num_move(np, std::vector<uint64_t>(np, 0));
        for (int i = 0; i < np; i++) {
            double s = 0;
            for (int j = 0; j < np; j++) s += row[i * np + j];
            if (s < 0.99999 || s > 1.00001) return fail(GE_ERR_MIGRATION, "The sum of columns in transition matrix must be 1");
        }
        for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) if (i != j) num_move[i][j] = (uint64_t)std::round(row[i * np + j] * (double)pop[i].n);
        std::vector<std::vector<Indiv>> people(np);
        std::vector<std::vector<std::vector<Indiv>>> camp(np, std::vector<std::vector<Indiv>>(np));
        std::vector<std::vector<uint64_t>> samples(np);
        for (int i = 0; i < np; i++) {
            uint64_t s = 0; for (uint64_t v : num_move[i]) s += v;
            std::vector<uint64_t> sample(s);
            if (rng_mode == GO_RNG_REF) {
                // RasRandomNumber::ras_SampleWithoutReplacement (src/RasRandomNumber.cpp:90-120): Knuth algorithm S;
                // the engine is `static`, so only the first call's seed is honoured
                unsigned sd = ras_glob_seed();
                if (!mig_engine_seeded) { mig_engine.seed(sd); mig_engine_seeded = true; }
                static std::uniform_real_distribution<double> Dist(0, 1);
                int n = (int)s, N = (int)pop[i].n, t = 0, m = 0;
                while (m < n) {
                    double u = Dist(mig_engine);
                    if ((N - t) * u >= n - m) t++;
                    else { sample[m] = t; t++; m++; }
                }
            } else if (rng_mode == GE_RNG_REPLAY) {
                if ((int)mig_sample.size() <= i || mig_sample[i].size() != s) return fail(GE_ERR_INVALID, "replay mode: go_set_migration_sample must supply the migrants");
                sample = mig_sample[i];
            } else {
                // the s smallest (key, index) pairs: a uniform sample without replacement
                uint64_t N = pop[i].n;
                std::vector<uint64_t> key(N), o(N);
                for (uint64_t k = 0; k < N; k++) { uint32_t w[4]; stream.draw(P_MIGRATE, i, gen, k, 0, 0, w); key[k] = key64(w); o[k] = k; }
                std::stable_sort(o.begin(), o.end(), [&](uint64_t a, uint64_t b) { return key[a] < key[b]; });
                for (uint64_t k = 0; k < s; k++) sample[k] = o[k];
            }
            std::sort(sample.begin(), sample.end(), std::greater<uint64_t>());
            samples[i] = sample;
            // camps: consecutive slices of the sample go to consecutive destinations.  The reference does not
            // reset k per destination (:924-936); identical for <= 1 non-zero destination per source, which is
            // the only case in which the reference itself survives.
            uint64_t k = 0;
            for (int j = 0; j < np; j++) {
                if (i == j) continue;
                for (uint64_t it = 0; it < num_move[i][j]; it++) camp[i][j].push_back(take(pop[i], sample[k++]));
            }
        }
        for (int i = 0; i < np; i++) {
            std::vector<uint8_t> gone(pop[i].n, 0);
            for (uint64_t idx : samples[i]) gone[idx] = 1;
            for (uint64_t k = 0; k < pop[i].n; k++) if (!gone[k]) people[i].push_back(take(pop[i], k));
        }
        for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) if (i != j) for (auto &h : camp[i][j]) people[j].push_back(h);
        for (int i = 0; i < np; i++) rebuild(pop[i], people[i]);
        return GE_OK;
    }

    // ---------------- orchestration ----------------
    int init_generation0(const ge_draws *d0) {  // ras_init_generation0 :529-679
        int nf = nphen();
        for (int p = 0; p < cfg.n_pop; p++) {
            Pop &P = pop[p];
            int rc = initial_human_gen0(p, d0); if (rc) return rc;
            rc = compute_AD(p, 0); if (rc) return rc;
            P.prev_n = P.n; P.prev_mv.assign(P.n, 0); P.prev_sv.assign(P.n, 0);  // ras_fill_Pop_info_prev_gen_for_gen0_prev
            P.prev_P.assign((uint64_t)nf * P.n, 0); P.prev_F.assign((uint64_t)nf * P.n, 0);
            P.var_a0.assign(nf, 0); P.var_d0.assign(nf, 0); P.var_g0.assign(nf, 0);
            P.last.e_raw.assign((uint64_t)nf * P.n, 0.0);
            for (int f = 0; f < nf; f++) {
                P.var_g0[f] = var_slice(P.G, (uint64_t)f * P.n, P.n);
                P.var_a0[f] = var_slice(P.A, (uint64_t)f * P.n, P.n);
                P.var_d0[f] = var_slice(P.D, (uint64_t)f * P.n, P.n);
                bool rp = rng_mode == GE_RNG_REPLAY && d0;
                rc = scale_AD_compute_GEF(p, 0, f, (rp && d0[p].e_raw) ? d0[p].e_raw + (uint64_t)f * P.n : nullptr,
                                          (rp && d0[p].parental0) ? d0[p].parental0 + (uint64_t)f * P.n : nullptr);
                if (rc) return rc;
            }
        }
        for (int f = 0; f < nf; f++) env_effects(f);
        for (int p = 0; p < cfg.n_pop; p++) compute_mv_sv(p, 0, nullptr);
        for (int p = 0; p < cfg.n_pop; p++) save_prev(p);
        for (int p = 0; p < cfg.n_pop; p++) {  // :645-654 beta adjustment
            Pop &P = pop[p];
            for (int f = 0; f < nf; f++) {
                double vP = var_slice(P.P, (uint64_t)f * P.n, P.n), vF = var_slice(P.F, (uint64_t)f * P.n, P.n);
                if (cfg.vt_type == 1) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vP));
                else if (cfg.vt_type == 2) { if (vF > 0) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vF)); }
            }
        }
        return GE_OK;
    }

    int step_generation(int gen, const ge_gen_params *gp, const double *mig, const ge_draws *dr) {  // sim_next_generation :1890-2082
        int nf = nphen();
        for (int p = 0; p < cfg.n_pop; p++) {
            int rc;
            if (!(rng_mode == GE_RNG_REPLAY && dr)) { rc = mate(p, gen, gp[p]); if (rc) return rc; }
            rc = reproduce(p, gen, dr ? &dr[p] : nullptr); if (rc) return rc;
            rc = compute_AD(p, gen); if (rc) return rc;
            pop[p].last.e_raw.assign((uint64_t)nf * pop[p].n, 0.0);
            for (int f = 0; f < nf; f++) {
                rc = scale_AD_compute_GEF(p, gen, f, (dr && dr[p].e_raw) ? dr[p].e_raw + (uint64_t)f * pop[p].n : nullptr);
                if (rc) return rc;
            }
        }
        for (int f = 0; f < nf; f++) env_effects(f);
        for (int p = 0; p < cfg.n_pop; p++) compute_mv_sv(p, gen, &gp[p]);
        if (cfg.n_pop > 1 && mig) { int rc = do_migration(gen, mig); if (rc) return rc; }
        for (int p = 0; p < cfg.n_pop; p++) save_prev(p);
        return GE_OK;
    }
};

// ------------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------------
#define CHECK_POP(ctx, p) if (!(ctx) || (p) < 0 || (p) >= (ctx)->cfg.n_pop) return fail(GE_ERR_INVALID, "bad population index")

extern "C" {

const char *go_last_error(void) { return g_err.c_str(); }

int go_create(const ge_config *cfg, go_ctx **out) {
    if (!cfg || !out || cfg->n_pop < 1 || cfg->n_chr < 1 || cfg->n_phen < 1) return fail(GE_ERR_INVALID, "bad config");
    go_ctx *c = new go_ctx();
    c->cfg = *cfg; c->rng_mode = cfg->rng_mode;
    c->pop.resize(cfg->n_pop);
    c->loci.resize(cfg->n_chr);
    for (int k = 0; k < cfg->n_chr; k++) c->chr_ids.push_back((uint32_t)k);
    c->stream.k0 = (uint32_t)cfg->seed; c->stream.k1 = (uint32_t)(cfg->seed >> 32);
    c->glob_generator.seed((unsigned)cfg->seed);  // Simulation::run :75-76
    for (Pop &P : c->pop) {
        P.rmap_bp.resize(cfg->n_chr); P.recom_prob.resize(cfg->n_chr); P.bp_dist.resize(cfg->n_chr);
        P.mut_bp.resize(cfg->n_chr); P.mut_rate.resize(cfg->n_chr);
        P.cv.assign(cfg->n_phen, std::vector<CvChr>(cfg->n_chr));
        P.scheme.resize(cfg->n_phen);
        P.panel.resize(cfg->n_chr);
    }
    *out = c;
    return GE_OK;
}
int go_destroy(go_ctx *ctx) { delete ctx; return GE_OK; }

int go_set_population(go_ctx *ctx, int pop, int avoid_inbreeding, int random_mating, double mm) {
    CHECK_POP(ctx, pop);
    ctx->pop[pop].avoid_inbreeding = avoid_inbreeding; ctx->pop[pop].RM = random_mating; ctx->pop[pop].MM = mm;
    return GE_OK;
}
int go_set_genetic_map(go_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rp, uint64_t n, uint64_t bp_dist) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    P.rmap_bp[chr].assign(bp, bp + n); P.recom_prob[chr].assign(rp, rp + n); P.bp_dist[chr] = bp_dist;
    return GE_OK;
}
int go_set_mutation_map(go_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rate, uint64_t n) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    P.mut_bp[chr].assign(bp, bp + n); P.mut_rate[chr].assign(rate, rate + n); P.has_mut = true;
    return GE_OK;
}
int go_set_loci(go_ctx *ctx, int chr, const uint64_t *pos, uint64_t n) { ctx->loci[chr].assign(pos, pos + n); return GE_OK; }
int go_set_founder_panel(go_ctx *ctx, int pop, int chr, const uint8_t *al, uint64_t nh) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    P.panel[chr].assign(al, al + nh * ctx->loci[chr].size()); P.n_founder_haps = nh;
    return GE_OK;
}
int go_set_founder_panel_packed(go_ctx *ctx, int pop, int chr, const uint32_t *words, uint64_t nh) {  // same panel, bit-packed
    CHECK_POP(ctx, pop);
    uint64_t nl = ctx->loci[chr].size(), nw = (nl + 31) / 32;
    std::vector<uint8_t> al(nh * nl);
    for (uint64_t h = 0; h < nh; h++)
        for (uint64_t s = 0; s < nl; s++) al[h * nl + s] = (words[h * nw + (s >> 5)] >> (s & 31)) & 1u;
    return go_set_founder_panel(ctx, pop, chr, al.data(), nh);
}
int go_set_cv(go_ctx *ctx, int pop, int phen, int chr, const uint64_t *bp, const double *a, const double *d, uint64_t ncv, const uint8_t *val, uint64_t nh) {
    CHECK_POP(ctx, pop);
    CvChr &cv = ctx->pop[pop].cv[phen][chr];
    cv.bp.assign(bp, bp + ncv); cv.a.assign(a, a + ncv); cv.d.assign(d, d + ncv); cv.val.assign(val, val + nh * ncv); cv.nhap = nh;
    return GE_OK;
}
int go_set_pheno_scheme(go_ctx *ctx, int pop, int phen, double va, double vd, double ve, double vc, double vf, double omega, double beta, double lambda) {
    CHECK_POP(ctx, pop);
    Scheme &S = ctx->pop[pop].scheme[phen];
    S.va = va; S.vd = vd; S.ve = ve; S.vc = vc; S.vf = vf; S.omega = omega; S.beta = beta; S.lambda = lambda;
    return GE_OK;
}
int go_set_chromosome_ids(go_ctx *ctx, const int32_t *ids) { for (int k = 0; k < ctx->cfg.n_chr; k++) ctx->chr_ids[k] = (uint32_t)ids[k]; return GE_OK; }
int go_set_allreduce(go_ctx *ctx, ge_allreduce_fn fn, void *user) { ctx->allreduce = fn; ctx->allreduce_user = user; return GE_OK; }
int go_set_gamma(go_ctx *ctx, const double *g) { ctx->gamma.assign(g, g + ctx->cfg.n_phen); return GE_OK; }
int go_init_generation0(go_ctx *ctx, const ge_draws *d0) { return ctx->init_generation0(d0); }
int go_mate(go_ctx *ctx, int pop, int gen, const ge_gen_params *gp) { CHECK_POP(ctx, pop); return ctx->mate(pop, gen, *gp); }
int go_set_couples(go_ctx *ctx, int pop, const uint64_t *m, const uint64_t *f, const uint8_t *inb, const int32_t *no, uint64_t n) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    P.c_male.assign(m, m + n); P.c_female.assign(f, f + n); P.c_inbreed.assign(inb, inb + n); P.c_noff.assign(no, no + n);
    return GE_OK;
}
int go_get_couples_count(go_ctx *ctx, int pop, uint64_t *n) { CHECK_POP(ctx, pop); *n = ctx->pop[pop].c_male.size(); return GE_OK; }
int go_get_couples(go_ctx *ctx, int pop, uint64_t *m, uint64_t *f, uint8_t *inb, int32_t *no) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    std::copy(P.c_male.begin(), P.c_male.end(), m); std::copy(P.c_female.begin(), P.c_female.end(), f);
    std::copy(P.c_inbreed.begin(), P.c_inbreed.end(), inb); std::copy(P.c_noff.begin(), P.c_noff.end(), no);
    return GE_OK;
}
int go_reproduce(go_ctx *ctx, int pop, int gen, const ge_draws *dr) { CHECK_POP(ctx, pop); return ctx->reproduce(pop, gen, dr); }
int go_compute_AD(go_ctx *ctx, int pop, int gen) { CHECK_POP(ctx, pop); return ctx->compute_AD(pop, gen); }
int go_scale_AD_compute_GEF(go_ctx *ctx, int pop, int gen, int phen, const double *e) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    if (P.last.e_raw.size() != (uint64_t)ctx->cfg.n_phen * P.n) P.last.e_raw.assign((uint64_t)ctx->cfg.n_phen * P.n, 0.0);
    return ctx->scale_AD_compute_GEF(pop, gen, phen, e);
}
int go_environmental_effects_specific_to_each_population(go_ctx *ctx, int phen) { return ctx->env_effects(phen); }
int go_compute_mating_value_selection_value(go_ctx *ctx, int pop, int gen, const ge_gen_params *gp) { CHECK_POP(ctx, pop); return ctx->compute_mv_sv(pop, gen, gp); }
int go_do_migration(go_ctx *ctx, int gen, const double *row) { return ctx->do_migration(gen, row); }
int go_set_migration_sample(go_ctx *ctx, int src, const uint64_t *pos, uint64_t n) {
    CHECK_POP(ctx, src);
    ctx->mig_sample.resize(ctx->cfg.n_pop);
    ctx->mig_sample[src].assign(pos, pos + n);
    return GE_OK;
}
int go_save_human_info_to_Pop_info_prev_gen(go_ctx *ctx, int pop) { CHECK_POP(ctx, pop); ctx->save_prev(pop); return GE_OK; }
int go_step_generation(go_ctx *ctx, int gen, const ge_gen_params *gp, const double *mig, const ge_draws *dr) { return ctx->step_generation(gen, gp, mig, dr); }

int go_get_population_size(go_ctx *ctx, int pop, uint64_t *n) { CHECK_POP(ctx, pop); *n = ctx->pop[pop].n; return GE_OK; }
int go_download_individuals(go_ctx *ctx, int pop, ge_indiv_soa *o) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    auto cp = [](const std::vector<double> &s, double *d) { if (d) std::copy(s.begin(), s.end(), d); };
    if (o->ids) std::copy(P.ids.begin(), P.ids.end(), o->ids);
    if (o->sex) std::copy(P.sex.begin(), P.sex.end(), o->sex);
    cp(P.A, o->A); cp(P.D, o->D); cp(P.G, o->G); cp(P.C, o->C); cp(P.E, o->E); cp(P.F, o->F); cp(P.P, o->P);
    cp(P.mv, o->mv); cp(P.sv, o->sv); cp(P.svf, o->svf);
    return GE_OK;
}
int go_get_moments(go_ctx *ctx, int pop, int f, ge_moments *m) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; uint64_t o = (uint64_t)f * P.n, n = P.n;
    m->var_A = var_slice(P.A, o, n); m->var_D = var_slice(P.D, o, n); m->var_G = var_slice(P.G, o, n); m->var_C = var_slice(P.C, o, n);
    m->var_E = var_slice(P.E, o, n); m->var_F = var_slice(P.F, o, n); m->var_P = var_slice(P.P, o, n); m->h2 = m->var_A / m->var_P;
    return GE_OK;
}
int go_get_mv_sv_var(go_ctx *ctx, int pop, double *vm, double *vs) { CHECK_POP(ctx, pop); *vm = var_(ctx->pop[pop].mv); *vs = var_(ctx->pop[pop].sv); return GE_OK; }
int go_get_gen0_constants(go_ctx *ctx, int pop, int f, double *va0, double *vd0, double *beta, double *m0, double *v0) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    *va0 = P.var_a0[f]; *vd0 = P.var_d0[f]; *beta = P.scheme[f].beta; *m0 = P.sv_mean0; *v0 = P.sv_var0;
    return GE_OK;
}
int go_download_haplotypes(go_ctx *ctx, int pop, int c, uint8_t *al) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; uint64_t nl = ctx->loci[c].size(), w = ctx->wc(c);
    for (uint64_t r = 0; r < 2 * P.n; r++) for (uint64_t s = 0; s < nl; s++) al[r * nl + s] = (P.bits[c][r * w + s / 32] >> (s % 32)) & 1u;
    return GE_OK;
}
int go_download_haplotypes_from_segments(go_ctx *ctx, int pop, int c, uint8_t *al) {  // ras_convert_interval_to_hap_matrix :1186-1230
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; uint64_t nl = ctx->loci[c].size(); int C = ctx->nchr();
    std::fill(al, al + 2 * P.n * nl, 0);
    for (uint64_t i = 0; i < P.n; i++)
        for (int h = 0; h < 2; h++)
            for (const Part &q : P.segs[(i * C + c) * 2 + h]) {
                const std::vector<uint8_t> &panel = ctx->pop[q.root].panel[c];
                for (uint64_t s = 0; s < nl; s++) {
                    uint64_t pos = ctx->loci[c][s];
                    if (q.check_interval(pos)) {
                        if (q.hap >= ctx->pop[q.root].n_founder_haps) return fail(GE_ERR_INVALID, "p.hap_index is not in range");
                        uint8_t v = panel[q.hap * nl + s];
                        if (std::find(q.mut.begin(), q.mut.end(), pos) != q.mut.end()) v = !v;
                        al[(2 * i + h) * nl + s] = v;
                    }
                }
            }
    return GE_OK;
}
int go_get_segment_count(go_ctx *ctx, int pop, int c, uint64_t *ns, uint64_t *nm) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; int C = ctx->nchr(); uint64_t a = 0, b = 0;
    for (uint64_t i = 0; i < P.n; i++) for (int h = 0; h < 2; h++) for (const Part &q : P.segs[(i * C + c) * 2 + h]) { a++; b += q.mut.size(); }
    *ns = a; *nm = b;
    return GE_OK;
}
int go_download_segments(go_ctx *ctx, int pop, int c, uint64_t *off, uint64_t *seg, uint64_t *moff, uint64_t *mbp) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; int C = ctx->nchr(); uint64_t a = 0, b = 0;
    for (uint64_t i = 0; i < P.n; i++) for (int h = 0; h < 2; h++) {
        off[i * 2 + h] = a; if (moff) moff[i * 2 + h] = b;
        for (const Part &q : P.segs[(i * C + c) * 2 + h]) {
            seg[a * 4] = q.st; seg[a * 4 + 1] = q.en; seg[a * 4 + 2] = q.hap; seg[a * 4 + 3] = q.root; a++;
            if (mbp) for (uint64_t m : q.mut) mbp[b++]= m; else b += q.mut.size();
        }
    }
    off[2 * P.n] = a; if (moff) moff[2 * P.n] = b;
    return GE_OK;
}
int go_download_cv_alleles(go_ctx *ctx, int pop, int f, int c, uint8_t *out) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop]; uint64_t ncv = P.cv[f][c].bp.size();
    std::vector<uint8_t> v; std::vector<double> a, d;
    for (uint64_t i = 0; i < P.n; i++) for (int h = 0; h < 2; h++) { ctx->find_cv(P, i, c, f, h, v, a, d); std::copy(v.begin(), v.end(), out + (i * 2 + h) * ncv); }
    return GE_OK;
}
int go_download_cv_alleles_bits(go_ctx *ctx, int pop, int f, int c, uint8_t *out) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    std::copy(P.cvbit[f][c].begin(), P.cvbit[f][c].end(), out);
    return GE_OK;
}
int go_download_AD_raw(go_ctx *ctx, int pop, double *A, double *D) {
    CHECK_POP(ctx, pop);
    Pop &P = ctx->pop[pop];
    std::copy(P.A_raw.begin(), P.A_raw.end(), A); std::copy(P.D_raw.begin(), P.D_raw.end(), D);
    return GE_OK;
}
int go_get_draw_counts(go_ctx *ctx, int pop, uint64_t *no, uint64_t *nx, uint64_t *nm) {
    CHECK_POP(ctx, pop);
    Draws &L = ctx->pop[pop].last;
    *no = L.father.size(); *nx = L.xo_bp.size(); *nm = L.mut_bp.size();
    return GE_OK;
}
int go_download_draws(go_ctx *ctx, int pop, uint64_t *fa, uint64_t *mo, uint8_t *sex, uint64_t *xo_off, uint64_t *xo_bp, uint8_t *start, uint64_t *mut_off, uint64_t *mut_bp, uint8_t *mut_gam) {
    CHECK_POP(ctx, pop);
    Draws &L = ctx->pop[pop].last;
    if (fa) std::copy(L.father.begin(), L.father.end(), fa);
    if (mo) std::copy(L.mother.begin(), L.mother.end(), mo);
    if (sex) std::copy(L.sex.begin(), L.sex.end(), sex);
    if (xo_off) std::copy(L.xo_off.begin(), L.xo_off.end(), xo_off);
    if (xo_bp) std::copy(L.xo_bp.begin(), L.xo_bp.end(), xo_bp);
    if (start) std::copy(L.start_hap.begin(), L.start_hap.end(), start);
    if (mut_off) std::copy(L.mut_off.begin(), L.mut_off.end(), mut_off);
    if (mut_bp) std::copy(L.mut_bp.begin(), L.mut_bp.end(), mut_bp);
    if (mut_gam) std::copy(L.mut_gam.begin(), L.mut_gam.end(), mut_gam);
    return GE_OK;
}
int go_download_e_raw(go_ctx *ctx, int pop, double *e) { CHECK_POP(ctx, pop); std::copy(ctx->pop[pop].last.e_raw.begin(), ctx->pop[pop].last.e_raw.end(), e); return GE_OK; }

void go_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) { philox4x32_10(k0, k1, c0, c1, c2, c3, out); }

// Word-level bit-packed propagation on one core (the CPU "port" baseline for the HBM-bound kernel).
double go_bench_propagate_bits(uint64_t n_parents, uint64_t n_offspring, uint64_t n_loci, uint64_t n_xo, uint64_t seed, uint64_t *checksum) {
    uint64_t W = (n_loci + 31) / 32;
    std::vector<uint32_t> par(n_parents * 2 * W), off(n_offspring * 2 * W);
    std::mt19937_64 g(seed);
    for (auto &w : par) w = (uint32_t)g();
    std::vector<uint64_t> fa(n_offspring * 2), xo(n_offspring * 2 * n_xo);
    std::vector<uint8_t> st(n_offspring * 2);
    for (uint64_t i = 0; i < n_offspring * 2; i++) {
        fa[i] = g() % n_parents; st[i] = g() & 1;
        for (uint64_t k = 0; k < n_xo; k++) xo[i * n_xo + k] = g() % n_loci;
        std::sort(&xo[i * n_xo], &xo[i * n_xo] + n_xo);
    }
    auto t0 = std::chrono::steady_clock::now();
    for (uint64_t r = 0; r < n_offspring * 2; r++) {
        const uint32_t *h[2] = {&par[(fa[r] * 2) * W], &par[(fa[r] * 2 + 1) * W]};
        uint32_t *dst = &off[r * W];
        int cur = st[r]; uint64_t w0 = 0;
        for (uint64_t k = 0; k <= n_xo; k++) {
            uint64_t cut = k < n_xo ? xo[r * n_xo + k] : n_loci;  // loci [prev, cut) from cur
            uint64_t wend = cut / 32;
            if (wend > w0) { std::memcpy(dst + w0, h[cur] + w0, (wend - w0) * 4); w0 = wend; }
            if (k < n_xo && w0 < W) {
                uint32_t lowmask = (cut % 32) ? ((1u << (cut % 32)) - 1u) : 0u;
                // boundary word: low bits from cur, high bits decided by later pieces (start from the next hap)
                uint32_t word = (h[cur][w0] & lowmask) | (h[cur ^ 1][w0] & ~lowmask);
                // further cuts inside the same word
                int c2 = cur ^ 1; uint64_t kk = k + 1;
                while (kk < n_xo && xo[r * n_xo + kk] / 32 == w0) {
                    uint64_t cc = xo[r * n_xo + kk];
                    uint32_t m2 = (cc % 32) ? ((1u << (cc % 32)) - 1u) : 0u;
                    word = (word & m2) | (h[c2 ^ 1][w0] & ~m2);
                    c2 ^= 1; kk++;
                }
                dst[w0] = word; w0++;
                cur = c2; k = kk - 1;
            } else cur ^= 1;
        }
    }
    auto t1 = std::chrono::steady_clock::now();
    uint64_t cs = 0;
    for (uint64_t i = 0; i < off.size(); i += 97) cs = cs * 1315423911u + off[i];
    if (checksum) *checksum = cs;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
