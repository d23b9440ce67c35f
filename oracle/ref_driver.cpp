// ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from the product).
//
// Links the reference's own, unmodified objects (compiled from /root/reference/src by oracle/Makefile)
// and drives its *private* per-generation methods in the order of Simulation::sim_next_generation
// (src/Simulation.cpp:1890-2082), exporting
//   * the parsed inputs (genetic map, recombination probabilities, CV panels, generation table),
//   * per generation: the couples chosen by random_mate/assort_mate, and per offspring the crossover
//     lists, starting haplotypes, mutation hits, sex and N(0,1) environment draws,
//   * per generation: the resulting per-individual state (segments, A/D/G/C/E/F/P, MV/SV/SV_f, pedigree),
//   * optionally the materialised haplotype matrix (ras_convert_interval_to_hap_matrix, :1186-1230).
// These are "the reference's own exported parent pairs, breakpoints and mutation draws" that the
// fixed-draw parity tests replay through the CUDA path and through oracle/ge_oracle.cpp.
//
// The only restated reference logic is the ~40-line body of Simulation::reproduce (:2433-2488), needed
// because the crossover lists are locals of that function.  It is self-checked every generation: the
// untouched reproduce() runs first, then glob_generator is rewound and the loop is re-issued calling the
// reference's own ras_sim_loc_rec / recombine / ras_add_mutation; offspring must match exactly and the
// generator must end in the same state, otherwise the driver aborts.
//
// usage: ge_ref_export --export out.gex [--export_hap] [--write_info] <reference CLI flags...>

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#define private public
#include "Simulation.h"
#undef private
#include "CommFunc.h"
#include "gex.h"

static GexWriter W;

static std::string key(int gen, int pop, const std::string &name) {
    return "g" + std::to_string(gen) + ".p" + std::to_string(pop) + "." + name;
}

static void die(const std::string &msg) {
    std::cerr << "ref_driver: " << msg << std::endl;
    std::exit(2);
}

static int sel_code(const std::string &s) {
    if (s == "") return 0;
    if (s == "logit") return 1;
    if (s == "probit") return 2;
    if (s == "stab") return 3;
    if (s == "thr") return 4;
    return -1;
}

static void export_inputs(Simulation &sim, bool with_panel) {
    W.scalar_i("in.n_pop", sim._n_pop);
    W.scalar_i("in.tot_gen", sim._tot_gen);
    W.scalar_i("in.vt_type", sim._vt_type);
    W.scalar_i("in.seed", sim.par._seed);
    W.put("in.gamma", sim._gamma);
    {
        std::vector<double> m;
        for (auto &row : sim.migration_mat_gen) for (double x : row) m.push_back(x);
        if (!m.empty()) W.put2("in.migration", m, sim.migration_mat_gen.size(), (uint64_t)sim._n_pop * sim._n_pop);
    }
    for (int p = 0; p < sim._n_pop; p++) {
        Population &P = sim.population[p];
        std::string pre = "in.p" + std::to_string(p) + ".";
        int nchr = P._nchr, nphen = (int)P._pheno_scheme.size();
        W.scalar_i(pre + "nchr", nchr);
        W.scalar_i(pre + "nphen", nphen);
        W.scalar_i(pre + "avoid_inbreeding", P._avoid_inbreeding);
        W.scalar_i(pre + "RM", P._RM);
        W.scalar(pre + "MM_percent", P._MM_percent);
        W.scalar_i(pre + "has_mutation_map", P._mutation_map.size() > 0);
        W.scalar_i(pre + "n_founders", (int64_t)P._indv_id.size());
        std::vector<int64_t> chr_ids(P._all_active_chrs.begin(), P._all_active_chrs.end());
        W.put(pre + "chr_ids", chr_ids);
        W.put(pre + "pop_size", std::vector<uint64_t>(P._pop_size.begin(), P._pop_size.end()));
        W.put(pre + "mat_cor", P._mat_cor);
        std::vector<uint8_t> od; std::vector<int32_t> sf;
        for (auto &s : P._offspring_dist) od.push_back(s.empty() ? 0 : (uint8_t)s[0]);
        for (auto &s : P._selection_func) sf.push_back(sel_code(s));
        W.put(pre + "offspring_dist", od);
        W.put(pre + "selection_func", sf);
        W.put(pre + "selection_par1", P._selection_func_par1);
        W.put(pre + "selection_par2", P._selection_func_par2);
        std::vector<double> scheme;
        for (int f = 0; f < nphen; f++) {
            Phenotype_scheme &S = P._pheno_scheme[f];
            double v[8] = {S._va, S._vd, S._ve, S._vc, S._vf, S._omega, S._beta, S._lambda};
            scheme.insert(scheme.end(), v, v + 8);
        }
        W.put2(pre + "scheme", scheme, nphen, 8);  // va vd ve vc vf omega beta(lambda as given) lambda
        for (int c = 0; c < nchr; c++) {
            std::string cp = pre + "c" + std::to_string(c) + ".";
            W.put(cp + "rmap_bp", std::vector<uint64_t>(P._rmap[c].bp.begin(), P._rmap[c].bp.end()));
            W.put(cp + "rmap_cM", P._rmap[c].cM);
            W.put(cp + "recom_prob", P._recom_prob[c]);
            W.scalar_i(cp + "bp_dist", (int64_t)P._rmap[c].bp_dist_in_rmap);
            if (P._mutation_map.size() > 0) {
                W.put(cp + "mut_bp", std::vector<uint64_t>(P._mutation_map[c].bp.begin(), P._mutation_map[c].bp.end()));
                W.put(cp + "mut_rate", P._mutation_map[c].mutation_rate);
            }
            for (int f = 0; f < nphen; f++) {
                std::string fp = cp + "f" + std::to_string(f) + ".";
                CV_INFO &ci = P._pheno_scheme[f]._cv_info[c];
                W.put(fp + "cv_bp", std::vector<uint64_t>(ci.bp.begin(), ci.bp.end()));
                W.put(fp + "cv_a", ci.genetic_value_a);
                W.put(fp + "cv_d", ci.genetic_value_d);
                auto &val = P._pheno_scheme[f]._cvs[c].val;
                uint64_t nh = val.size(), ncv = nh ? val[0].size() : 0;
                std::vector<uint8_t> flat(nh * ncv);
                for (uint64_t h = 0; h < nh; h++) for (uint64_t k = 0; k < ncv; k++) flat[h * ncv + k] = val[h][k];
                W.put2(fp + "cv_val", flat, nh, ncv);
            }
        }
    }
    if (with_panel) {
        int nchr = sim.population[0]._nchr;
        for (int c = 0; c < nchr; c++) {
            std::vector<Legend> legs(sim._n_pop);
            std::vector<Hap_SNP> haps(sim._n_pop);
            if (!sim.ras_read_hap_legend_sample_chr(legs, haps, c)) die("cannot read founder panel");
            for (int p = 0; p < sim._n_pop; p++) {
                std::string cp = "in.p" + std::to_string(p) + ".c" + std::to_string(c) + ".";
                W.put(cp + "panel_pos", std::vector<uint64_t>(legs[p].pos.begin(), legs[p].pos.end()));
                uint64_t nh = haps[p].hap.size(), ns = nh ? haps[p].hap[0].size() : 0;
                std::vector<uint8_t> flat(nh * ns);
                for (uint64_t h = 0; h < nh; h++) for (uint64_t s = 0; s < ns; s++) flat[h * ns + s] = haps[p].hap[h][s];
                W.put2(cp + "panel", flat, nh, ns);
            }
        }
    }
}

static void export_state(Simulation &sim, int gen, int p) {
    Population &P = sim.population[p];
    uint64_t n = P.h.size();
    int nchr = P._nchr, nphen = (int)P._pheno_scheme.size();
    W.scalar_i(key(gen, p, "n"), (int64_t)n);
    std::vector<uint64_t> ids(n * 7);
    std::vector<uint8_t> sex(n);
    std::vector<double> mv(n), sv(n), svf(n);
    for (uint64_t i = 0; i < n; i++) {
        Human &h = P.h[i];
        uint64_t v[7] = {h.ID, h.ID_Father, h.ID_Mother, h.ID_Fathers_Father, h.ID_Fathers_Mother,
                         h.ID_Mothers_Father, h.ID_Mothers_Mother};
        std::copy(v, v + 7, ids.begin() + i * 7);
        sex[i] = (uint8_t)h.sex;
        mv[i] = h.mating_value; sv[i] = h.selection_value; svf[i] = h.selection_value_func;
    }
    W.put2(key(gen, p, "ids"), ids, n, 7);
    W.put(key(gen, p, "sex"), sex);
    W.put(key(gen, p, "mv"), mv);
    W.put(key(gen, p, "sv"), sv);
    W.put(key(gen, p, "svf"), svf);
    const char *names[7] = {"A", "D", "G", "C", "E", "F", "P"};
    for (int k = 0; k < 7; k++) {
        std::vector<double> a((uint64_t)nphen * n);
        for (int f = 0; f < nphen; f++)
            for (uint64_t i = 0; i < n; i++) {
                Human &h = P.h[i];
                double x = 0;
                switch (k) {
                    case 0: x = h.additive[f]; break;
                    case 1: x = h.dominance[f]; break;
                    case 2: x = h.bv[f]; break;
                    case 3: x = h.common_sibling[f]; break;
                    case 4: x = h.e_noise[f]; break;
                    case 5: x = h.parental_effect[f]; break;
                    case 6: x = h.phen[f]; break;
                }
                a[(uint64_t)f * n + i] = x;
            }
        W.put2(key(gen, p, names[k]), a, nphen, n);
    }
    std::vector<double> sc;
    for (int f = 0; f < nphen; f++) {
        sc.push_back(P._var_a_gen0[f]); sc.push_back(P._var_d_gen0[f]); sc.push_back(P._pheno_scheme[f]._beta);
    }
    W.put2(key(gen, p, "var_a0_var_d0_beta"), sc, nphen, 3);
    W.scalar(key(gen, p, "gen0_SV_mean"), sim._gen0_SV_mean[p]);
    W.scalar(key(gen, p, "gen0_SV_var"), sim._gen0_SV_var[p]);
    for (int c = 0; c < nchr; c++) {
        std::vector<uint64_t> off(1, 0), seg, moff(1, 0), mbp;
        for (uint64_t i = 0; i < n; i++)
            for (int hh = 0; hh < 2; hh++) {
                auto &parts = P.h[i].chr[c].Hap[hh];
                for (auto &pt : parts) {
                    seg.push_back(pt.st); seg.push_back(pt.en); seg.push_back(pt.hap_index);
                    seg.push_back((uint64_t)pt.root_population);
                    for (auto m : pt.mutation_pos) mbp.push_back(m);
                    moff.push_back(mbp.size());
                }
                off.push_back(seg.size() / 4);
            }
        std::string cp = "c" + std::to_string(c) + ".";
        W.put(key(gen, p, cp + "seg_off"), off);
        W.put2(key(gen, p, cp + "seg"), seg, seg.size() / 4, 4);
        W.put(key(gen, p, cp + "segmut_off"), moff);
        W.put(key(gen, p, cp + "segmut_bp"), mbp);
    }
}

static void export_hap(Simulation &sim, int gen) {
    int nchr = sim.population[0]._nchr;
    for (int c = 0; c < nchr; c++) {
        std::vector<Legend> legs(sim._n_pop);
        std::vector<Hap_SNP> haps(sim._n_pop);
        if (!sim.ras_read_hap_legend_sample_chr(legs, haps, c)) die("cannot read founder panel");
        for (int p = 0; p < sim._n_pop; p++) {
            Hap_SNP out;
            if (!sim.ras_convert_interval_to_hap_matrix(p, haps, legs, c, out)) die("hap conversion failed");
            uint64_t nh = out.hap.size(), ns = nh ? out.hap[0].size() : 0;
            std::vector<uint8_t> flat(nh * ns);
            for (uint64_t h = 0; h < nh; h++) for (uint64_t s = 0; s < ns; s++) flat[h * ns + s] = out.hap[h][s];
            W.put2(key(gen, p, "c" + std::to_string(c) + ".hap"), flat, nh, ns);
        }
    }
}

static bool parts_equal(const std::vector<part> &a, const std::vector<part> &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) {
        if (a[i].st != b[i].st || a[i].en != b[i].en || a[i].hap_index != b[i].hap_index ||
            a[i].root_population != b[i].root_population || a[i].mutation_pos != b[i].mutation_pos ||
            a[i].gen0_indv != b[i].gen0_indv)
            return false;
    }
    return true;
}

// The draws random_mate (:2090-2157) / assort_mate (:2167-2360) consumed, captured by re-deriving their engines: the functions seed
// everything from successive ras_glob_seed() calls, so with glob_generator rewound to its state before the call the same seeds come
// out again, and the reference's OWN helpers (RasRandomNumber::ras_mvnorm, ras_rpois, std::random_shuffle over rand()) give the same
// values.  Nothing of the pairing logic is restated here; the exported couples stay those of the untouched function.  Self-checks:
// glob_generator must end in the state the real call left, and the Poisson family sizes must equal the couples' num_offspring.
extern int myrandom(int i);   // src/Simulation.cpp:11-14
static void export_mating_draws(Simulation &sim, int p, int gen, const std::default_random_engine &g_before) {
    Population &P = sim.population[p];
    const std::default_random_engine g_after = sim.glob_generator;
    sim.glob_generator = g_before;
    const uint64_t n_h = P.h.size();
    const double nan = std::nan("");
    std::vector<double> thin_u(n_h, nan), mm_u(n_h, nan);
    if (P._RM) {
        unsigned seed = sim.ras_glob_seed();
        std::default_random_engine generator(seed);
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        uint64_t n_m = 0, n_f = 0;
        for (uint64_t i = 0; i < n_h; i++) {
            double r = distribution(generator);
            thin_u[i] = r;
            if (r < P.h[i].selection_value_func) { if (P.h[i].sex == 1) n_m++; else if (P.h[i].sex == 2) n_f++; }
        }
        std::default_random_engine g_uint_f(seed + 1), g_uint_m(seed + 2);
        std::uniform_int_distribution<unsigned long int> d_uint_f(0, n_m - 1), d_uint_m(0, n_f - 1);
        const uint64_t nc = P._pop_size[gen - 1];
        std::vector<uint64_t> i_f(nc), i_m(nc);
        for (uint64_t i = 0; i < nc; i++) { i_f[i] = d_uint_f(g_uint_f); i_m[i] = d_uint_m(g_uint_m); }
        W.put(key(gen, p, "mate.rm_father_idx"), i_f);
        W.put(key(gen, p, "mate.rm_mother_idx"), i_m);
    } else {
        unsigned seed = sim.ras_glob_seed();
        std::srand(seed);
        std::default_random_engine generator(sim.ras_glob_seed());
        std::uniform_real_distribution<double> distribution(0.0, 1.0);
        uint64_t n_m = 0, n_f = 0;
        for (uint64_t i = 0; i < n_h; i++) {
            double r = distribution(generator);
            thin_u[i] = r;
            if (r < P.h[i].selection_value_func && (P.h[i].sex == 1 || P.h[i].sex == 2)) {
                double r2 = distribution(generator);
                mm_u[i] = r2;
                (P.h[i].sex == 1 ? n_m : n_f) += (r2 < P._MM_percent) ? 2 : 1;
            }
        }
        std::vector<uint64_t> trim_order;
        if (n_m != n_f) {   // the longer list after std::random_shuffle (:2235 / :2242): positions in the unshuffled list
            trim_order.resize(std::max(n_m, n_f));
            for (uint64_t k = 0; k < trim_order.size(); k++) trim_order[k] = k;
            std::random_shuffle(trim_order.begin(), trim_order.end());
        }
        W.put(key(gen, p, "mate.trim_order"), trim_order);
        const uint64_t n2 = std::min(n_m, n_f);
        std::vector<double> mu(2, 0);
        std::vector<std::vector<double> > corr(2, std::vector<double>(2, 0));
        corr[0][0] = 1; corr[0][1] = P._mat_cor[gen - 1]; corr[1][0] = P._mat_cor[gen - 1]; corr[1][1] = 1;
        std::vector<std::vector<double> > tpl = RasRandomNumber::ras_mvnorm(n2, mu, corr, sim.ras_glob_seed());
        std::vector<double> t1(n2), t2(n2);
        for (uint64_t i = 0; i < n2; i++) { t1[i] = tpl[i][0]; t2[i] = tpl[i][1]; }
        W.put(key(gen, p, "mate.t1"), t1);
        W.put(key(gen, p, "mate.t2"), t2);
        if (P._couples_info.size() != n2) die("mating draws: couple count differs from the reference's");
        uint64_t n_inbreed = 0;
        for (auto &c : P._couples_info) n_inbreed += c.inbreed ? 1 : 0;
        const std::string &od = P._offspring_dist[gen - 1];
        std::vector<int32_t> family;
        std::vector<uint64_t> remainder_order;
        if (od == "p" || od == "P") {
            double lam = (double)P._pop_size[gen - 1] / (n2 - n_inbreed);
            std::vector<int> no = RasRandomNumber::ras_rpois(n2, lam, sim.ras_glob_seed());
            family.assign(no.begin(), no.end());
            for (uint64_t i = 0; i < n2; i++) if (family[i] != P._couples_info[i].num_offspring) die("mating draws: Poisson family sizes differ from the reference's couples");
        } else if (!P._avoid_inbreeding) {   // pos_couple_can_marry is only filled without --avoid_inbreeding (:2322-2326)
            remainder_order.resize(n2);
            for (uint64_t k = 0; k < n2; k++) remainder_order[k] = k;
            std::random_shuffle(remainder_order.begin(), remainder_order.end(), myrandom);
        }
        W.put(key(gen, p, "mate.family"), family);
        W.put(key(gen, p, "mate.remainder_order"), remainder_order);
    }
    W.put(key(gen, p, "mate.thin_u"), thin_u);
    W.put(key(gen, p, "mate.mm_u"), mm_u);
    if (!(sim.glob_generator == g_after)) die("mating draws: glob_generator ended in a different state than the reference's call left");
    sim.glob_generator = g_after;
}

// Runs the untouched reproduce(), then replays its loop body (:2433-2488) to capture the local draws.
static std::vector<Human> reproduce_with_export(Simulation &sim, int p, int gen) {
    Population &P = sim.population[p];
    std::default_random_engine g0 = sim.glob_generator;
    std::vector<Human> h_ref = sim.reproduce(p, gen);
    std::default_random_engine g1 = sim.glob_generator;

    sim.glob_generator = g0;
    unsigned seed = sim.ras_glob_seed();
    std::srand(seed);
    uint64_t n_couples = P._couples_info.size();
    int nchr = (int)P.h[0].chr.size();
    bool has_mut = P._mutation_map.size() > 0;

    std::vector<uint64_t> c_male, c_female, c_noff; std::vector<uint8_t> c_inbreed;
    for (auto &c : P._couples_info) {
        c_male.push_back(c.pos_male); c_female.push_back(c.pos_female);
        c_noff.push_back((uint64_t)c.num_offspring); c_inbreed.push_back(c.inbreed);
    }
    W.put(key(gen, p, "couple_male"), c_male);
    W.put(key(gen, p, "couple_female"), c_female);
    W.put(key(gen, p, "couple_noff"), c_noff);
    W.put(key(gen, p, "couple_inbreed"), c_inbreed);

    std::vector<uint64_t> off_father, off_mother, off_couple, xo_off(1, 0), xo_bp, mut_off(1, 0), mut_bp;
    std::vector<uint8_t> start_hap, mut_gam, sex;
    uint64_t i_people = 0;
    for (uint64_t it = 0; it < n_couples; it++) {
        if (P._couples_info[it].inbreed) continue;
        uint64_t pm = P._couples_info[it].pos_male, pf = P._couples_info[it].pos_female;
        Human h_pat = P.h[pm];
        Human h_mat = P.h[pf];
        for (int ns = 0; ns < P._couples_info[it].num_offspring; ns++) {
            if (i_people >= h_ref.size()) die("replay produced more offspring than reproduce()");
            for (int c = 0; c < nchr; c++) {
                unsigned seed_loc = std::rand();
                std::vector<unsigned long int> loc_pat = sim.ras_sim_loc_rec(P._recom_prob[c], P._rmap[c], seed_loc);
                int s_pat = std::rand() % 2;
                std::vector<part> hp = sim.recombine(h_pat.chr[c], s_pat, loc_pat);
                seed_loc = std::rand();
                std::vector<unsigned long int> loc_mat = sim.ras_sim_loc_rec(P._recom_prob[c], P._rmap[c], seed_loc);
                int s_mat = std::rand() % 2;
                std::vector<part> hm = sim.recombine(h_mat.chr[c], s_mat, loc_mat);
                if (has_mut) {
                    std::vector<size_t> np, nm;
                    for (auto &q : hp) np.push_back(q.mutation_pos.size());
                    for (auto &q : hm) nm.push_back(q.mutation_pos.size());
                    sim.ras_add_mutation(p, c, hp, hm);
                    for (size_t j = 0; j < hp.size(); j++)
                        for (size_t k = np[j]; k < hp[j].mutation_pos.size(); k++) { mut_bp.push_back(hp[j].mutation_pos[k]); mut_gam.push_back(0); }
                    for (size_t j = 0; j < hm.size(); j++)
                        for (size_t k = nm[j]; k < hm[j].mutation_pos.size(); k++) { mut_bp.push_back(hm[j].mutation_pos[k]); mut_gam.push_back(1); }
                }
                mut_off.push_back(mut_bp.size());
                if (!parts_equal(hp, h_ref[i_people].chr[c].Hap[0]) || !parts_equal(hm, h_ref[i_people].chr[c].Hap[1]))
                    die("replay of reproduce() diverged from the reference (segments)");
                // crossovers without the two sentinels the reference adds (:2983, :2993)
                for (size_t k = 1; k + 1 < loc_pat.size(); k++) xo_bp.push_back(loc_pat[k]);
                xo_off.push_back(xo_bp.size());
                start_hap.push_back((uint8_t)s_pat);
                for (size_t k = 1; k + 1 < loc_mat.size(); k++) xo_bp.push_back(loc_mat[k]);
                xo_off.push_back(xo_bp.size());
                start_hap.push_back((uint8_t)s_mat);
            }
            int sx = (std::rand() % 2) + 1;
            if (sx != h_ref[i_people].sex) die("replay of reproduce() diverged from the reference (sex)");
            sex.push_back((uint8_t)sx);
            off_father.push_back(pm); off_mother.push_back(pf); off_couple.push_back(it);
            i_people++;
        }
    }
    if (i_people != h_ref.size()) die("replay offspring count mismatch");
    if (!(sim.glob_generator == g1)) die("replay left glob_generator in a different state");
    W.put(key(gen, p, "off_father"), off_father);
    W.put(key(gen, p, "off_mother"), off_mother);
    W.put(key(gen, p, "off_couple"), off_couple);
    W.put(key(gen, p, "off_sex"), sex);
    W.put(key(gen, p, "xo_off"), xo_off);
    W.put(key(gen, p, "xo_bp"), xo_bp);
    W.put(key(gen, p, "start_hap"), start_hap);
    W.put(key(gen, p, "mut_off"), mut_off);
    W.put(key(gen, p, "mut_bp"), mut_bp);
    W.put(key(gen, p, "mut_gam"), mut_gam);
    return h_ref;
}

// Peeks the N(0,1) draws that ras_scale_AD_compute_GEF (:3078-3102) is about to make, without
// disturbing glob_generator; self-checked afterwards against the scaled e_noise the reference stored.
static std::vector<double> peek_e_draws(Simulation &sim, uint64_t n) {
    std::default_random_engine g = sim.glob_generator;
    std::uniform_int_distribution<unsigned> distribution(1, 1000000);
    unsigned seed = distribution(g);
    std::default_random_engine generator_e(seed);
    std::normal_distribution<double> distribution_e(0.0, 1);
    std::vector<double> e(n);
    for (uint64_t i = 0; i < n; i++) e[i] = distribution_e(generator_e);
    return e;
}

static void phenotypes_with_export(Simulation &sim, int gen, int p) {
    Population &P = sim.population[p];
    int nphen = (int)P._pheno_scheme.size();
    uint64_t n = P.h.size();
    std::vector<double> eraw((uint64_t)nphen * n), araw((uint64_t)nphen * n), draw((uint64_t)nphen * n);
    for (int f = 0; f < nphen; f++)
        for (uint64_t i = 0; i < n; i++) {
            araw[(uint64_t)f * n + i] = P.h[i].additive[f];
            draw[(uint64_t)f * n + i] = P.h[i].dominance[f];
        }
    W.put2(key(gen, p, "A_raw"), araw, nphen, n);   // ras_compute_AD output before scaling
    W.put2(key(gen, p, "D_raw"), draw, nphen, n);
    for (int f = 0; f < nphen; f++) {
        if (gen == 0) {
            // :555-565 — generation-0 scaling constants
            P._var_bv_gen0[f] = CommFunc::var(P.get_bv(f));
            P._var_a_gen0[f] = CommFunc::var(P.get_additive(f));
            P._var_d_gen0[f] = CommFunc::var(P.get_dominance(f));
        }
        std::vector<double> e = peek_e_draws(sim, n);
        if (!sim.ras_scale_AD_compute_GEF(gen, p, f, P._var_a_gen0[f], P._var_d_gen0[f])) die("ras_scale_AD_compute_GEF failed");
        if (P._pheno_scheme[f]._ve > 0) {
            double s_ev = std::sqrt(CommFunc::var(e) / P._pheno_scheme[f]._ve);
            for (uint64_t i = 0; i < n; i++)
                if (e[i] / s_ev != P.h[i].e_noise[f]) die("peeked environment draws do not match the reference");
        }
        std::copy(e.begin(), e.end(), eraw.begin() + (uint64_t)f * n);
    }
    W.put2(key(gen, p, "e_raw"), eraw, nphen, n);
}

int main(int argc, char **argv) {
    std::string export_path;
    bool export_hap_flag = false, write_info = false, quiet = true;
    std::vector<std::string> vec_arg;
    for (int i = 0; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--export" && i + 1 < argc) { export_path = argv[++i]; continue; }
        if (a == "--export_hap") { export_hap_flag = true; continue; }
        if (a == "--write_info") { write_info = true; continue; }
        if (a == "--verbose") { quiet = false; continue; }
        vec_arg.push_back(a);
    }
    if (export_path.empty()) die("--export <file> is required");
    vec_arg.push_back("nothing"); vec_arg.push_back("nothing");  // as src/Main.cpp:55-56

    std::streambuf *old_buf = std::cout.rdbuf();
    std::ofstream devnull("/dev/null");
    if (quiet) std::cout.rdbuf(devnull.rdbuf());

    Parameters par;
    if (!par.read(vec_arg) || !par.check()) die("bad reference parameters");
    Simulation sim;
    sim.par = par;
    sim.glob_generator.seed(par._seed);               // :75-76
    if (!sim.ras_init_parameters()) die("ras_init_parameters failed");
    if (!W.open(export_path)) die("cannot open export file");
    export_inputs(sim, export_hap_flag);

    int nphen = (int)par._va[0].size();
    // ---- generation 0, order of ras_init_generation0 (:529-679)
    for (int p = 0; p < sim._n_pop; p++) {
        sim.ras_initial_human_gen0(p);
        if (!sim.ras_compute_AD(p, 0)) die("ras_compute_AD failed");
        sim.ras_fill_Pop_info_prev_gen_for_gen0_prev(p);
        phenotypes_with_export(sim, 0, p);
    }
    for (int f = 0; f < nphen; f++) sim.sim_environmental_effects_specific_to_each_population(f);
    for (int p = 0; p < sim._n_pop; p++) sim.ras_compute_mating_value_selection_value(0, p);
    for (int p = 0; p < sim._n_pop; p++) sim.ras_save_human_info_to_Pop_info_prev_gen(p);
    for (int p = 0; p < sim._n_pop; p++) {
        Population &P = sim.population[p];
        if (write_info) P.ras_save_human_info(0);
        for (int f = 0; f < nphen; f++) {                // :645-654 — beta adjustment
            double var_P = CommFunc::var(P.get_phen(f));
            double var_F = CommFunc::var(P.get_parental_effect(f));
            if (sim._vt_type == 1) P._pheno_scheme[f]._beta = std::sqrt(P._pheno_scheme[f]._vf / (2 * var_P));
            else if (sim._vt_type == 2) { if (var_F > 0) P._pheno_scheme[f]._beta = std::sqrt(P._pheno_scheme[f]._vf / (2 * var_F)); }
        }
        export_state(sim, 0, p);
    }
    if (export_hap_flag) export_hap(sim, 0);

    // ---- generations 1.., order of sim_next_generation (:1890-2082)
    for (int gen = 1; gen <= sim._tot_gen; gen++) {
        for (int p = 0; p < sim._n_pop; p++) {
            Population &P = sim.population[p];
            const std::default_random_engine g_before_mating = sim.glob_generator;
            bool ok = P._RM ? sim.random_mate(p, gen - 1) : sim.assort_mate(p, gen - 1);
            if (!ok) die("mating failed");
            export_mating_draws(sim, p, gen, g_before_mating);
            P.h = reproduce_with_export(sim, p, gen);
            if (!sim.ras_compute_AD(p, gen)) die("ras_compute_AD failed");
            phenotypes_with_export(sim, gen, p);
        }
        for (int f = 0; f < nphen; f++) sim.sim_environmental_effects_specific_to_each_population(f);
        for (int p = 0; p < sim._n_pop; p++) sim.ras_compute_mating_value_selection_value(gen, p);
        // state before migration (what the offspring kernels must reproduce)
        for (int p = 0; p < sim._n_pop; p++) {
            uint64_t n = sim.population[p].h.size();
            std::vector<uint64_t> idv(n);
            for (uint64_t i = 0; i < n; i++) idv[i] = sim.population[p].h[i].ID;
            W.put(key(gen, p, "premig_ids"), idv);
        }
        if (sim._n_pop > 1) if (!sim.ras_do_migration(gen - 1)) die("migration failed");
        for (int p = 0; p < sim._n_pop; p++) sim.ras_save_human_info_to_Pop_info_prev_gen(p);
        for (int p = 0; p < sim._n_pop; p++) {
            if (write_info) sim.population[p].ras_save_human_info(gen);
            export_state(sim, gen, p);
        }
        if (export_hap_flag) export_hap(sim, gen);
    }
    W.close();
    std::cout.rdbuf(old_buf);
    std::cout << "ref_driver: exported " << sim._tot_gen << " generations to " << export_path << std::endl;
    return 0;
}
