// ge_host.cpp — readers of the reference's input formats, the generation loop over the C-ABI and the writers of
// its per-generation outputs.  See ge_host.hpp for the map onto the reference's host code.
#include "ge_host.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <mutex>
#include <sstream>
#include <thread>

namespace gehost {

// ------------------------------------------------------------------------------------------------
// command line (Parameters::read / check, src/parameters.cpp:15-382)
// ------------------------------------------------------------------------------------------------
const char *Options::usage() {
    return "geneevolve_b200_cli — GeneEvolve's reproduction loop on B200 (flags as in the reference)\n"
           "  per population (separate populations with --next_population):\n"
           "    --file_gen_info F --file_hap_name F --file_recom_map F [--file_mutation_map F] [--RM] [--MM x]\n"
           "    per phenotype: --file_cv_info F --file_cvs F [--va x --vd x --vc x --ve x --vf x --omega x --beta x --lambda x]\n"
           "  global: [--gamma x]... [--file_migration F] [--vt_type 1|2] [--avoid_inbreeding] [--seed n] [--prefix P]\n"
           "          [--out_hap] [--out_plink] [--out_plink01] [--out_interval] [--file_output_generations F]\n"
           "          [--device n] [--gpus N] [--quiet]\n"
           "          [--compact_segments]  (extension: merge adjacent same-founder segments after every generation)\n"
           "  rejected: --file_ref_vcf (libStatGen VCF reader), --out_vcf (the reference refuses it for hap-format founders too)\n";
}

bool Options::parse(const std::vector<std::string> &a) {
    size_t npop = 1;
    for (const auto &s : a) if (s == "--next_population") npop++;
    pop.assign(npop, PopOptions());
    size_t ip = 0;
    auto need = [&](size_t &i) -> const std::string & {
        static const std::string empty;
        if (i + 1 >= a.size()) { error = "Error: missing value after [" + a[i] + "]."; return empty; }
        return a[++i];
    };
    auto num = [&](size_t &i) { const std::string &v = need(i); return error.empty() ? std::strtod(v.c_str(), nullptr) : 0.0; };
    for (size_t i = 0; i < a.size() && error.empty(); i++) {
        const std::string &f = a[i];
        PopOptions &P = pop[ip];
        if (f == "--next_population") ip++;
        else if (f == "--file_gen_info") P.file_gen_info = need(i);
        else if (f == "--file_hap_name") P.file_hap_name = need(i);
        else if (f == "--file_recom_map") P.file_recom_map = need(i);
        else if (f == "--file_mutation_map") P.file_mutation_map = need(i);
        else if (f == "--MM") P.MM = num(i);
        else if (f == "--RM") P.RM = true;
        else if (f == "--vt_type") vt_type = (int)num(i);
        else if (f == "--file_cv_info") P.file_cv_info.push_back(need(i));
        else if (f == "--file_cvs") P.file_cvs.push_back(need(i));
        else if (f == "--va") P.va.push_back(num(i));
        else if (f == "--vd") P.vd.push_back(num(i));
        else if (f == "--vc") P.vc.push_back(num(i));
        else if (f == "--ve") P.ve.push_back(num(i));
        else if (f == "--vf") P.vf.push_back(num(i));
        else if (f == "--omega") P.omega.push_back(num(i));
        else if (f == "--beta") P.beta.push_back(num(i));
        else if (f == "--lambda") P.lambda.push_back(num(i));
        else if (f == "--gamma") gamma.push_back(num(i));
        else if (f == "--file_migration") file_migration = need(i);
        else if (f == "--avoid_inbreeding") avoid_inbreeding = true;
        else if (f == "--seed") seed = (uint64_t)num(i);
        else if (f == "--prefix") prefix = need(i);
        else if (f == "--out_hap") out_hap = true;
        else if (f == "--out_interval") out_interval = true;
        else if (f == "--out_plink") out_plink = true;
        else if (f == "--out_plink01") out_plink01 = true;
        else if (f == "--file_output_generations") file_output_generations = need(i);
        else if (f == "--device") device = (int)num(i);
        else if (f == "--gpus") gpus = (int)num(i);
        else if (f == "--compact_segments") compact_segments = true;
        else if (f == "--quiet") quiet = true;
        else if (f == "--debug") {}
        else if (f == "--help" || f == "-h" || f == "?") help = true;
        else if (f == "--out_vcf")   // with hap-format founders the reference refuses too (src/Simulation.cpp:1071-1075)
            error = "Error: current version can't convert to VCF output format! [--out_vcf]";
        else if (f == "--file_ref_vcf")
            error = "Error: [" + f + "] needs libStatGen's VCF reader, which is outside the GPU reproduction path; use the hap formats.";
        else error = "Error: unknown option [" + f + "].";
    }
    if (!error.empty() || help) return error.empty();
    size_t nphen = pop[0].file_cv_info.size();
    for (size_t p = 0; p < npop; p++) {  // the reference's defaults and checks (src/parameters.cpp:153-356)
        PopOptions &P = pop[p];
        std::string where = " in population " + std::to_string(p + 1) + ".";
        if (P.file_gen_info.empty()) return (error = "Error: missing parameter [--file_gen_info]" + where, false);
        if (P.file_hap_name.empty()) return (error = "Error: missing the reference file. Check the parameter [--file_hap_name]" + where, false);
        if (P.file_recom_map.empty()) return (error = "Error: missing parameter [--file_recom_map]" + where, false);
        size_t k = P.file_cv_info.size();
        if (k == 0) return (error = "Error: missing parameter [--file_cv_info]" + where, false);
        if (P.file_cvs.size() != k) return (error = "Error: each phenotype needs one [--file_cvs]. Error" + where, false);
        if (k != nphen) return (error = "Error: The number of phenotypes should be the same for each population.", false);
        struct D { std::vector<double> *v; double def; const char *name; };
        for (D d : {D{&P.va, -1, "--va"}, D{&P.vd, -1, "--vd"}, D{&P.vc, 0, "--vc"}, D{&P.ve, 1, "--ve"}, D{&P.vf, 0, "--vf"},
                    D{&P.omega, 1, "--omega"}, D{&P.beta, 1, "--beta"}, D{&P.lambda, 1, "--lambda"}}) {
            if (d.v->empty()) d.v->assign(k, d.def);
            if (d.v->size() != k) return (error = std::string("Error: each phenotype needs one [") + d.name + "]. Error" + where, false);
        }
        for (size_t f = 0; f < k; f++) {
            if (!(P.va[f] > 0 || P.va[f] == -1)) return (error = "Error: The parameter [--va] should be positive. Error" + where, false);
            if (!(P.vd[f] >= 0 || P.vd[f] == -1)) return (error = "Error: The parameter [--vd] should not be negetive. Error" + where, false);
            if (P.vc[f] < 0 || P.ve[f] < 0 || P.vf[f] < 0) return (error = "Error: [--vc], [--ve], [--vf] should not be negetive. Error" + where, false);
        }
        if (P.MM < 0 || P.MM > 1) return (error = "Error: The parameter [--MM] should be between 0 and 1. Error" + where, false);
    }
    if (gamma.empty()) gamma.assign(nphen, 0.0);
    if (gamma.size() != nphen) return (error = "Error: the number of [--gamma] must be equal to the number of phenotypes (" + std::to_string(nphen) + ").", false);
    if (npop > 1 && file_migration.empty())
        return (error = "Error: When you have more than one populations, you must specify the [--file_migration] option.", false);
    if (gpus < 1) return (error = "Error: [--gpus] must be at least 1.", false);
    if (seed == 0) seed = 1;  // the reference seeds from the clock here (src/parameters.cpp:206-209); a fixed default keeps runs reproducible
    return true;
}

// ------------------------------------------------------------------------------------------------
// readers
// ------------------------------------------------------------------------------------------------
static std::vector<std::string> split_ws(const std::string &line) {
    std::vector<std::string> t;
    std::istringstream is(line);
    std::string s;
    while (is >> s) t.push_back(s);
    return t;
}
static int chr_index(const std::vector<ChrFiles> &chrs, int chr) {
    for (size_t j = 0; j < chrs.size(); j++) if (chrs[j].chr == chr) return (int)j;
    return -1;
}
static bool open_in(const std::string &path, std::ifstream &f, std::string &err) {
    f.open(path.c_str());
    if (!f) { err = "Error: can not open the file [" + path + "] to read."; return false; }
    return true;
}

bool read_generation_info(const std::string &path, std::vector<GenRow> &out, std::string &err) {  // src/Population.cpp:13-96
    std::ifstream f;
    if (!open_in(path, f, err)) return false;
    std::string line;
    std::getline(f, line);  // header
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.empty()) continue;
        if (t.size() != 6) { err = "Error: file [" + path + "] must have 6 columns: pop_size, mat_cor, offspring_dist, selection_func, selection_func_par1 and selection_func_par2."; return false; }
        GenRow r;
        r.pop_size = (uint64_t)std::strtod(t[0].c_str(), nullptr);  // "3e+05" is allowed
        r.mat_cor = std::strtod(t[1].c_str(), nullptr);
        if (r.mat_cor > 1 || r.mat_cor < -1) { std::cout << " Warning in file [" << path << "]: mate_corr should be in range [-1,1]. We set it to 0." << std::endl; r.mat_cor = 0; }
        r.offspring_dist = (t[2] == "f") ? 'f' : 'p';
        if (t[2] != "p" && t[2] != "f") std::cout << " Warning in file [" << path << "]: offspring_dist should be [p] or [f]. We set it to [p]." << std::endl;
        r.par1 = std::strtod(t[4].c_str(), nullptr); r.par2 = std::strtod(t[5].c_str(), nullptr);
        if (t[3] == "logit") r.selection_func = GE_SEL_LOGIT;
        else if (t[3] == "probit") r.selection_func = GE_SEL_PROBIT;
        else if (t[3] == "stab") r.selection_func = GE_SEL_STAB;
        else if (t[3] == "thr") r.selection_func = GE_SEL_THR;
        else {
            std::cout << " Warning in file [" << path << "]: selection_func should be [logit,probit,stab,thr]. We set it to [logit 0 1]." << std::endl;
            r.selection_func = GE_SEL_LOGIT; r.par1 = 0; r.par2 = 1;
        }
        out.push_back(r);
    }
    if (out.empty()) { err = "Error: file [" + path + "] holds no generation."; return false; }
    return true;
}

bool read_hap_address(const std::string &path, std::vector<ChrFiles> &out, std::string &err) {  // src/Population.cpp:103-142
    std::ifstream f;
    if (!open_in(path, f, err)) return false;
    std::string line;
    std::getline(f, line);  // header
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.empty()) continue;
        if (t.size() < 4) { err = "Error: file [" + path + "] needs 4 columns: chr hap legend sample."; return false; }
        out.push_back({std::atoi(t[0].c_str()), t[1], t[2], t[3]});
    }
    if (out.empty()) { err = "Error: file [" + path + "] lists no chromosome."; return false; }
    return true;
}

bool read_recombination_map(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<GeneticMap> &out, std::string &err) {
    std::ifstream f;  // src/Population.cpp:349-414 and ras_compute_recom_prob :471-507
    if (!open_in(path, f, err)) return false;
    out.assign(chrs.size(), GeneticMap());
    std::string line;
    std::getline(f, line);  // header
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.size() < 3) continue;
        int j = chr_index(chrs, std::atoi(t[0].c_str()));
        if (j < 0) continue;  // only active chromosomes
        out[j].bp.push_back((uint64_t)std::strtod(t[1].c_str(), nullptr));
        out[j].cM.push_back(std::strtod(t[2].c_str(), nullptr));
    }
    for (size_t j = 0; j < out.size(); j++) {
        GeneticMap &m = out[j];
        if (m.bp.size() < 2) { err = "Error: the recombination map [" + path + "] has fewer than two rows for chromosome " + std::to_string(chrs[j].chr) + "."; return false; }
        m.bp_dist = m.bp[1] - m.bp[0];
        m.recom_prob.assign(m.cM.size(), 0.0);
        for (size_t k = 1; k < m.cM.size(); k++) m.recom_prob[k] = (m.cM[k] - m.cM[k - 1]) * .01;
    }
    return true;
}

bool read_mutation_map(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<MutationMap> &out, std::string &err) {
    std::ifstream f;  // src/Population.cpp:420-468
    if (!open_in(path, f, err)) return false;
    out.assign(chrs.size(), MutationMap());
    std::string line;
    std::getline(f, line);
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.size() < 3) continue;
        int j = chr_index(chrs, std::atoi(t[0].c_str()));
        if (j < 0) continue;
        double r = std::strtod(t[2].c_str(), nullptr);
        if (r < 0 || r > 1) r = 0;
        out[j].bp.push_back((uint64_t)std::strtod(t[1].c_str(), nullptr));
        out[j].rate.push_back(r);
    }
    return true;
}

bool read_cv_info(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<CvBlock> &out, std::string &err) {
    std::ifstream f;  // src/Population.cpp:197-260
    if (!open_in(path, f, err)) return false;
    out.assign(chrs.size(), CvBlock());
    std::string line;
    std::getline(f, line);
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.empty()) continue;
        if (t.size() != 4) { err = "Error: file [" + path + "] should have 4 columns."; return false; }
        int chr = std::atoi(t[0].c_str()), j = chr_index(chrs, chr);
        if (j < 0) { err = "Error:  In file [" + path + "]. Chromosome [" + std::to_string(chr) + "] is not defined in the --file_hap_name [file]."; return false; }
        out[j].bp.push_back((uint64_t)std::strtod(t[1].c_str(), nullptr));
        out[j].a.push_back(std::strtod(t[2].c_str(), nullptr));
        out[j].d.push_back(std::strtod(t[3].c_str(), nullptr));
    }
    return true;
}

bool count_hap_columns(const std::string &path, uint64_t &n_hap, std::string &err) {
    std::ifstream f;
    if (!open_in(path, f, err)) return false;
    std::string line;
    std::getline(f, line);
    n_hap = 0;
    for (size_t i = 0; i < line.size(); i += 2) if (line[i] == '0' || line[i] == '1') n_hap++; else break;
    return true;
}

// one text row per SNP, alleles at even byte offsets (src/format_hap.cpp:62-121); rows -> bit s of hap-major words
bool read_hap_packed(const std::string &path, uint64_t n_hap, uint64_t n_snp, std::vector<uint32_t> &words, std::string &err) {
    std::ifstream f;
    if (!open_in(path, f, err)) return false;
    uint64_t nw = (n_snp + 31) / 32;
    words.assign(n_hap * nw, 0u);
    std::string line;
    uint64_t s = 0;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        if (s >= n_snp) { err = "Error: in file [" + path + "]: more rows than SNPs in the legend."; return false; }
        if (line.size() < 2 * n_hap - 1) { err = "Error: in file [" + path + "], line number:" + std::to_string(s) + " is too short."; return false; }
        const uint32_t bit = 1u << (s & 31);
        const uint64_t w = s >> 5;
        for (uint64_t h = 0; h < n_hap; h++) {
            char ch = line[2 * h];
            if (ch == '1') words[h * nw + w] |= bit;
            else if (ch != '0') { err = std::string("Error: undefined character [") + ch + "] in file [" + path + "], line number:" + std::to_string(s); return false; }
        }
        s++;
    }
    if (s != n_snp) { err = "Error: in file [" + path + "]: " + std::to_string(s) + " rows, expected " + std::to_string(n_snp) + "."; return false; }
    return true;
}

bool read_cvs(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<CvBlock> &io, std::string &err) {
    std::ifstream f;  // src/Population.cpp:280-343; no header
    if (!open_in(path, f, err)) return false;
    std::string line;
    std::vector<std::string> name(chrs.size());
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.size() < 2) continue;
        int j = chr_index(chrs, std::atoi(t[0].c_str()));
        if (j >= 0) name[j] = t[1];
    }
    for (size_t j = 0; j < chrs.size(); j++) {
        CvBlock &b = io[j];
        if (name[j].empty()) { if (!b.bp.empty()) { err = "Error: no CV haplotype file for chromosome " + std::to_string(chrs[j].chr) + " in [" + path + "]."; return false; } continue; }
        uint64_t nh = 0;
        if (!count_hap_columns(name[j], nh, err)) return false;
        std::ifstream g;
        if (!open_in(name[j], g, err)) return false;
        uint64_t ncv = b.bp.size(), k = 0;
        b.n_hap = nh;
        b.val.assign(nh * ncv, 0);
        while (std::getline(g, line)) {
            if (line.empty()) continue;
            if (k >= ncv) { err = "Error reading file [" + name[j] + "]: more rows than CVs in the cv_info file."; return false; }
            if (line.size() < 2 * nh - 1) { err = "Error reading file [" + name[j] + "]"; return false; }
            for (uint64_t h = 0; h < nh; h++) {
                char ch = line[2 * h];
                if (ch != '0' && ch != '1') { err = std::string("Error: undefined character [") + ch + "] in file [" + name[j] + "]"; return false; }
                b.val[h * ncv + k] = ch == '1';
            }
            k++;
        }
        if (k != ncv) { err = "Error reading file [" + name[j] + "]: " + std::to_string(k) + " rows for " + std::to_string(ncv) + " CVs."; return false; }
    }
    return true;
}

bool read_legend(const std::string &path, std::vector<std::string> &id, std::vector<uint64_t> &pos, std::vector<std::string> *al0,
                 std::vector<std::string> *al1, std::string &err) {
    std::ifstream f;  // src/format_hap.cpp:125-156, header then id pos allele0 allele1
    if (!open_in(path, f, err)) return false;
    std::string a, b, c, d;
    f >> a >> b >> c >> d;
    uint64_t p;
    while (f >> a >> p >> c >> d) {
        id.push_back(a); pos.push_back(p);
        if (al0) al0->push_back(c);
        if (al1) al1->push_back(d);
    }
    return true;
}
bool read_indv(const std::string &path, std::vector<std::string> &out, std::string &err) {
    std::ifstream f;  // src/format_hap.cpp:160-183, no header
    if (!open_in(path, f, err)) return false;
    std::string s;
    while (f >> s) out.push_back(s);
    return true;
}
bool read_migration(const std::string &path, int n_pop, size_t n_gen, std::vector<std::vector<double>> &out, std::string &err) {
    std::ifstream f;  // :839-874
    if (!open_in(path, f, err)) return false;
    std::string line;
    while (std::getline(f, line)) {
        auto t = split_ws(line);
        if (t.empty()) continue;
        if ((int)t.size() < n_pop * n_pop) { err = "Error: The file [" + path + "] must have n^2 columns, where n is the number of populations."; return false; }
        std::vector<double> row(n_pop * n_pop);
        for (int k = 0; k < n_pop * n_pop; k++) row[k] = std::strtod(t[k].c_str(), nullptr);
        out.push_back(row);
    }
    if (out.size() != n_gen) { err = "Error: The file [" + path + "] must have " + std::to_string(n_gen) + " lines, equal to the number of generations."; return false; }
    return true;
}
bool read_output_generations(const std::string &path, std::vector<int> &out, std::string &err) {
    std::ifstream f;  // :3481-3512
    if (!open_in(path, f, err)) return false;
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        char *end = nullptr;
        double d = std::strtod(line.c_str(), &end);
        if (end == line.c_str()) { err = "Error: Invalid or blank input number in [file_output_generations]!"; return false; }
        out.push_back((int)d);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// the simulation driver
// ------------------------------------------------------------------------------------------------
HostSimulation::HostSimulation(const Options &o, int rank_, int world_, Collective *coll_) : opt(o), rank(rank_), world(world_), coll(coll_) {
    if (rank != 0) opt.quiet = true;   // one narrator
}
int HostSimulation::allreduce_hook(void *user, double *buf, uint64_t count, void *stream) {
    HostSimulation *self = static_cast<HostSimulation *>(user);
    return self->coll->allreduce_sum(self->rank, buf, count, stream);
}
std::vector<std::vector<int>> HostSimulation::assign_chromosomes(const std::vector<double> &weight, int world) {
    std::vector<int> order(weight.size());
    for (size_t c = 0; c < order.size(); c++) order[c] = (int)c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<std::vector<int>> out(world);
    std::vector<double> load(world, 0.0);
    for (int c : order) {
        int r = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        out[r].push_back(c);
        load[r] += weight[c];
    }
    for (auto &v : out) std::sort(v.begin(), v.end());
    return out;
}
HostSimulation::~HostSimulation() { info_writer.reset(); if (ctx) ge_destroy(ctx); }
bool HostSimulation::gfail(const char *what) { err = std::string(what) + ": " + ge_last_error(); return false; }

bool HostSimulation::load_inputs() {
    n_pop = (int)opt.pop.size();
    n_phen = (int)opt.pop[0].file_cv_info.size();
    need_panel = opt.out_hap || opt.out_plink || opt.out_plink01;
    in.assign(n_pop, PopInputs());
    for (int p = 0; p < n_pop; p++) {
        const PopOptions &O = opt.pop[p];
        PopInputs &I = in[p];
        if (!read_generation_info(O.file_gen_info, I.gens, err)) return false;
        if (!read_hap_address(O.file_hap_name, I.chrs, err)) return false;
        if (!read_recombination_map(O.file_recom_map, I.chrs, I.rmap, err)) return false;
        if (!O.file_mutation_map.empty() && !read_mutation_map(O.file_mutation_map, I.chrs, I.mutmap, err)) return false;
        I.cv.resize(n_phen);
        for (int f = 0; f < n_phen; f++) {
            if (!read_cv_info(O.file_cv_info[f], I.chrs, I.cv[f], err)) return false;
            if (!read_cvs(O.file_cvs[f], I.chrs, I.cv[f], err)) return false;
        }
        if (!read_indv(I.chrs[0].sample, I.indv_id, err)) return false;
        if (p == 0) { n_chr = (int)I.chrs.size(); tot_gen = (int)I.gens.size(); }
        if ((int)I.chrs.size() != n_chr) return fail("Error: every population must list the same chromosomes.");
        if ((int)I.gens.size() != tot_gen) return fail("Error: every population needs the same number of generations.");
        if (need_panel) {
            I.legend_pos.resize(n_chr); I.legend_id.resize(n_chr); I.legend_al0.resize(n_chr); I.legend_al1.resize(n_chr);
            for (int c = 0; c < n_chr; c++)
                if (!read_legend(I.chrs[c].legend, I.legend_id[c], I.legend_pos[c], &I.legend_al0[c], &I.legend_al1[c], err)) return false;
        }
    }
    if (n_pop > 1 && !read_migration(opt.file_migration, n_pop, (size_t)tot_gen, migration, err)) return false;
    if (!opt.file_output_generations.empty() && !read_output_generations(opt.file_output_generations, output_generations, err)) return false;
    return true;
}

bool HostSimulation::upload() {
    uint64_t cap = 0;
    for (int p = 0; p < n_pop; p++) {
        cap = std::max<uint64_t>(cap, in[p].cv[0][0].n_hap / 2);
        for (const GenRow &r : in[p].gens) cap = std::max(cap, r.pop_size);
    }
    cap = (uint64_t)((double)cap * (n_pop > 1 ? 1.6 : 1.0) + 8.0 * std::sqrt((double)cap)) + 256;  // Poisson family sizes and migration
    // multi-GPU: every rank holds all individuals but only its chromosomes
    if (world > n_chr) return fail("Error: [--gpus] exceeds the number of chromosomes.");
    std::vector<double> weight(n_chr);
    for (int c = 0; c < n_chr; c++) weight[c] = (double)(in[0].rmap[c].bp.back() - in[0].rmap[c].bp.front());
    mine = assign_chromosomes(weight, world)[rank];
    const int n_loc = (int)mine.size();
    ge_config cfg = {};
    cfg.device = opt.device + rank; cfg.n_pop = n_pop; cfg.n_chr = n_loc; cfg.n_phen = n_phen; cfg.vt_type = opt.vt_type;
    cfg.representation = (need_panel ? GE_REP_BITS : 0) | ((opt.out_interval || !need_panel) ? GE_REP_SEGMENTS : 0);
    cfg.rng_mode = GE_RNG_PHILOX; cfg.seed = opt.seed; cfg.capacity = cap; cfg.rank = rank; cfg.world_size = world;
    if (ge_create(&cfg, &ctx) != GE_OK) return gfail("ge_create");
    if (world > 1) {
        std::vector<int32_t> ids(mine.begin(), mine.end());
        if (ge_set_chromosome_ids(ctx, ids.data()) != GE_OK) return gfail("ge_set_chromosome_ids");
        if (ge_set_allreduce(ctx, &HostSimulation::allreduce_hook, this) != GE_OK) return gfail("ge_set_allreduce");
    }
    if (ge_set_gamma(ctx, opt.gamma.data()) != GE_OK) return gfail("ge_set_gamma");
    for (int k = 0; k < n_loc && need_panel; k++)
        if (ge_set_loci(ctx, k, in[0].legend_pos[mine[k]].data(), in[0].legend_pos[mine[k]].size()) != GE_OK) return gfail("ge_set_loci");
    for (int p = 0; p < n_pop; p++) {
        const PopOptions &O = opt.pop[p];
        PopInputs &I = in[p];
        if (ge_set_population(ctx, p, opt.avoid_inbreeding, O.RM, O.MM) != GE_OK) return gfail("ge_set_population");
        for (int k = 0; k < n_loc; k++) {
            const int c = mine[k];
            const GeneticMap &m = I.rmap[c];
            if (ge_set_genetic_map(ctx, p, k, m.bp.data(), m.recom_prob.data(), m.bp.size(), m.bp_dist) != GE_OK) return gfail("ge_set_genetic_map");
            if (!I.mutmap.empty() && !I.mutmap[c].bp.empty() &&
                ge_set_mutation_map(ctx, p, k, I.mutmap[c].bp.data(), I.mutmap[c].rate.data(), I.mutmap[c].bp.size()) != GE_OK) return gfail("ge_set_mutation_map");
            for (int f = 0; f < n_phen; f++) {
                const CvBlock &b = I.cv[f][c];
                if (ge_set_cv(ctx, p, f, k, b.bp.data(), b.a.data(), b.d.data(), b.bp.size(), b.val.data(), b.n_hap) != GE_OK) return gfail("ge_set_cv");
            }
            if (need_panel) {
                uint64_t nh = 0;
                if (!count_hap_columns(I.chrs[c].hap, nh, err)) return false;
                std::vector<uint32_t> words;
                if (!opt.quiet) std::cout << "    reading founder panel [" << I.chrs[c].hap << "]" << std::endl;
                if (!read_hap_packed(I.chrs[c].hap, nh, I.legend_pos[c].size(), words, err)) return false;
                if (ge_set_founder_panel_packed(ctx, p, k, words.data(), nh) != GE_OK) return gfail("ge_set_founder_panel_packed");
            }
        }
        for (int f = 0; f < n_phen; f++)
            if (ge_set_pheno_scheme(ctx, p, f, O.va[f], O.vd[f], O.ve[f], O.vc[f], O.vf[f], O.omega[f], O.beta[f], O.lambda[f]) != GE_OK) return gfail("ge_set_pheno_scheme");
    }
    return true;
}

// The reference writes `<prefix>.info.popK.genG.txt` for every individual after every generation, single-threaded
// inside the generation loop.  Here the loop only downloads the columns (one pinned-speed copy); formatting and
// writing run on a background thread that splits the rows over the host's cores, so the text dump overlaps the next
// generations on the GPU.  Numbers are printed with "%g", which is what `ostream << double` produces.
class InfoWriter {
public:
    InfoWriter() : worker([this] { loop(); }) {}
    ~InfoWriter() { finish(); }
    void submit(std::function<void()> job) {
        std::unique_lock<std::mutex> l(m);
        cv_room.wait(l, [this] { return q.size() < 3; });   // bounded: at most three generations of columns in flight
        q.push_back(std::move(job));
        cv_job.notify_one();
    }
    void finish() {
        {
            std::unique_lock<std::mutex> l(m);
            if (done) return;
            done = true;
            cv_job.notify_one();
        }
        worker.join();
    }
    std::string error;   // first failure, read after finish()

private:
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> l(m);
                cv_job.wait(l, [this] { return done || !q.empty(); });
                if (q.empty()) return;
                job = std::move(q.front());
                q.pop_front();
                cv_room.notify_one();
            }
            job();
        }
    }
    std::mutex m;
    std::condition_variable cv_job, cv_room;
    std::deque<std::function<void()>> q;
    bool done = false;
    std::thread worker;
};

struct InfoColumns {  // one generation of one population, as downloaded
    uint64_t n = 0;
    int n_phen = 0;
    std::vector<uint64_t> ids;
    std::vector<uint8_t> sex;
    std::vector<double> col[7], mv, sv, svf;
};

static void format_info_rows(const InfoColumns &c, uint64_t i0, uint64_t i1, std::string &out) {
    char buf[64];
    out.reserve((size_t)(i1 - i0) * (60 + 80 * c.n_phen));
    auto put_u = [&](uint64_t v) { out.append(buf, (size_t)std::snprintf(buf, sizeof buf, "%llu ", (unsigned long long)v)); };
    auto put_d = [&](double v, char end) { int k = std::snprintf(buf, sizeof buf, "%g", v); buf[k] = end; out.append(buf, (size_t)k + 1); };
    for (uint64_t i = i0; i < i1; i++) {
        for (int k = 0; k < 7; k++) put_u(c.ids[i * 7 + k] + 1);  // IDs start from 1 in the files
        put_u(c.sex[i]);
        for (int j = 0; j < c.n_phen; j++)
            for (int k = 0; k < 7; k++) put_d(c.col[k][(uint64_t)j * c.n + i], ' ');
        put_d(c.mv[i], ' '); put_d(c.sv[i], ' '); put_d(c.svf[i], '\n');
    }
}

bool HostSimulation::write_info(int pop, int gen) {  // Population::ras_save_human_info, src/Population.cpp:510-568
    auto cols = std::make_shared<InfoColumns>();
    InfoColumns &c = *cols;
    if (ge_get_population_size(ctx, pop, &c.n) != GE_OK) return gfail("ge_get_population_size");
    c.n_phen = n_phen;
    c.ids.resize(c.n * 7); c.sex.resize(c.n); c.mv.resize(c.n); c.sv.resize(c.n); c.svf.resize(c.n);
    for (auto &v : c.col) v.resize(c.n * n_phen);
    ge_indiv_soa s = {c.ids.data(), c.sex.data(), c.col[0].data(), c.col[1].data(), c.col[2].data(), c.col[3].data(), c.col[4].data(), c.col[5].data(),
                      c.col[6].data(), c.mv.data(), c.sv.data(), c.svf.data()};
    if (ge_download_individuals(ctx, pop, &s) != GE_OK) return gfail("ge_download_individuals");
    std::string path = opt.prefix + ".info.pop" + std::to_string(pop + 1) + ".gen" + std::to_string(gen) + ".txt";
    if (!info_writer) info_writer.reset(new InfoWriter());
    InfoWriter *w = info_writer.get();
    w->submit([cols, path, w] {
        std::ofstream o(path.c_str(), std::ios::binary);
        if (!o) { if (w->error.empty()) w->error = "Error: can not open the file [" + path + "] to write."; return; }
        o << "ID ID_Father ID_Mother ID_Fathers_Father ID_Fathers_Mother ID_Mothers_Father ID_Mothers_Mother sex ";
        for (int j = 0; j < cols->n_phen; j++)
            for (const char *k : {"_A", "_D", "_G", "_C", "_E", "_F", "_P"}) o << "ph" << j + 1 << k << ' ';
        o << "MV SV SV_f\n";
        unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), (unsigned)(cols->n / 4096 + 1)));
        std::vector<std::string> part(nt);
        std::vector<std::future<void>> fut;
        for (unsigned t = 0; t < nt; t++)
            fut.push_back(std::async(std::launch::async, [&, t] { format_info_rows(*cols, cols->n * t / nt, cols->n * (t + 1) / nt, part[t]); }));
        for (unsigned t = 0; t < nt; t++) { fut[t].get(); o.write(part[t].data(), (std::streamsize)part[t].size()); }
    });
    return true;
}

bool HostSimulation::after_generation(int gen) {
    for (int p = 0; p < n_pop && rank == 0; p++) {   // every rank holds identical per-individual columns: rank 0 reports
        if (!write_info(p, gen)) return false;
        SummaryRow r;
        r.m.resize(n_phen);
        for (int f = 0; f < n_phen; f++) if (ge_get_moments(ctx, p, f, &r.m[f]) != GE_OK) return gfail("ge_get_moments");
        if (ge_get_mv_sv_var(ctx, p, &r.var_mv, &r.var_sv) != GE_OK) return gfail("ge_get_mv_sv_var");
        summary[p].push_back(r);
        if (!opt.quiet) {
            uint64_t n = 0;
            ge_get_population_size(ctx, p, &n);
            std::cout << "  generation " << gen << ", population " << p + 1 << ": n=" << n << ", var_A=" << r.m[0].var_A << ", var_P=" << r.m[0].var_P
                      << ", h2=" << r.m[0].h2 << std::endl;
        }
    }
    bool out = gen == tot_gen && gen > 0;  // the last generation is always written (:144), others on request (:2059-2063)
    for (int g : output_generations) out |= g == gen;
    if (out && (need_panel || opt.out_interval)) return write_genotypes(gen);
    return true;
}

bool HostSimulation::write_summary() {  // Simulation::ras_save_res :782-834
    for (int p = 0; p < n_pop; p++) {
        std::string path = opt.prefix + ".pop" + std::to_string(p + 1) + ".summary";
        std::ofstream o(path.c_str());
        if (!o) return fail("Error: can not open the file [" + path + "] to write.");
        const char *sep = " ";
        o << "gen" << sep;
        for (int f = 0; f < n_phen; f++)
            for (const char *k : {"_var_A", "_var_D", "_var_G", "_var_C", "_var_E", "_var_F", "_var_P", "_h2", "_var_G_std"}) o << "ph" << f + 1 << k << sep;
        o << "var_mating_value" << sep << "var_selection_value" << std::endl;
        for (size_t g = 0; g < summary[p].size(); g++) {
            const SummaryRow &r = summary[p][g];
            o << g << sep;
            for (int f = 0; f < n_phen; f++) {
                const ge_moments &m = r.m[f];
                o << m.var_A << sep << m.var_D << sep << m.var_G << sep << m.var_C << sep << m.var_E << sep << m.var_F << sep << m.var_P << sep << m.h2 << sep
                  << m.var_G / summary[p][0].m[f].var_G << sep;
            }
            o << r.var_mv << sep << r.var_sv << std::endl;
        }
    }
    return true;
}

bool HostSimulation::write_genotypes(int gen) {
    for (int p = 0; p < n_pop; p++) {
        uint64_t n = 0;
        if (ge_get_population_size(ctx, p, &n) != GE_OK) return gfail("ge_get_population_size");
        std::vector<uint64_t> ids(n * 7);
        std::vector<uint8_t> sex(n);
        ge_indiv_soa s = {};
        s.ids = ids.data(); s.sex = sex.data();
        if (ge_download_individuals(ctx, p, &s) != GE_OK) return gfail("ge_download_individuals");
        for (int k = 0; k < (int)mine.size(); k++) {   // each rank writes the files of its own chromosomes
            const int c = mine[k];
            std::string base = opt.prefix + ".pop" + std::to_string(p + 1) + ".gen" + std::to_string(gen) + ".chr" + std::to_string(in[0].chrs[c].chr);
            std::vector<uint8_t> m;   // alleles of this chromosome, [2n haplotypes][ns loci], materialised on the device
            const uint64_t ns = need_panel ? in[0].legend_pos[c].size() : 0;
            if (need_panel) {
                m.resize(2 * n * ns);
                if (ge_download_haplotypes(ctx, p, k, m.data()) != GE_OK) return gfail("ge_download_haplotypes");
            }
            if (opt.out_hap) {  // ras_write_hap_legend_sample :1142-1182 -> format_hap::write_hap / write_indv (src/format_hap.cpp:6-53)
                std::ofstream o((base + ".hap").c_str());
                if (!o) return fail("Error: can not open the file [" + base + ".hap] to write.");
                std::string line(2 * 2 * n, ' ');
                for (uint64_t k = 0; k < ns; k++) {
                    for (uint64_t h = 0; h < 2 * n; h++) line[2 * h] = m[h * ns + k] ? '1' : '0';
                    o << line << '\n';
                }
                std::ofstream oi((base + ".indv").c_str());
                for (uint64_t i = 0; i < n; i++) oi << ids[i * 7] + 1 << '\n';
            }
            // ras_write_hap_to_plink_format :1254-1303 -> format_plink::write_ped_map / write_ped01_map (src/format_plink.cpp:5-135).
            // Both flags name the same two files in the reference, the 0/1 coding is written last and wins; same here.
            for (int pass = 0; pass < 2; pass++) {
                const bool hap01 = pass == 1;
                if (!(hap01 ? opt.out_plink01 : opt.out_plink)) continue;
                const PopInputs &I = in[p];
                std::ofstream o((base + ".ped").c_str());
                if (!o) return fail("Error: can not open the file [" + base + ".ped] to write.");
                std::string line;
                for (uint64_t i = 0; i < n; i++) {   // FID IID PID MID sex phen, IDs + 1 because 0 is PLINK's missing value (:1393-1404)
                    const uint64_t *q = &ids[i * 7];
                    line = std::to_string(q[1] + 1) + ' ' + std::to_string(q[0] + 1) + ' ' + std::to_string(q[1] + 1) + ' ' + std::to_string(q[2] + 1) + ' ' +
                           std::to_string((int)sex[i]) + " -9";
                    const uint8_t *h0 = &m[(2 * i) * ns], *h1 = &m[(2 * i + 1) * ns];
                    for (uint64_t j = 0; j < ns; j++) {
                        line += ' ';
                        if (hap01) line += h0[j] ? '1' : '0'; else line += h0[j] ? I.legend_al1[c][j] : I.legend_al0[c][j];
                        line += ' ';
                        if (hap01) line += h1[j] ? '1' : '0'; else line += h1[j] ? I.legend_al1[c][j] : I.legend_al0[c][j];
                    }
                    o << line << '\n';
                }
                std::ofstream om((base + ".map").c_str());
                if (!om) return fail("Error: can not open the file [" + base + ".map] to write.");
                for (uint64_t j = 0; j < ns; j++) om << in[0].chrs[c].chr << ' ' << I.legend_id[c][j] << " 0 " << I.legend_pos[c][j] << '\n';
            }
            if (opt.out_interval) {  // ras_write_hap_to_interval_format :1582-1639
                uint64_t nseg = 0, nmut = 0;
                if (ge_get_segment_count(ctx, p, k, &nseg, &nmut) != GE_OK) return gfail("ge_get_segment_count");
                std::vector<uint64_t> off(2 * n + 1), seg(4 * std::max<uint64_t>(nseg, 1));
                if (ge_download_segments(ctx, p, k, off.data(), seg.data(), nullptr, nullptr) != GE_OK) return gfail("ge_download_segments");
                std::ofstream o((base + ".int").c_str());
                if (!o) return fail("Error: can not open the file [" + base + ".int] to write.");
                o << "h_ID chr hap st en hap_index gen0_indv root_pop" << std::endl;
                for (uint64_t i = 0; i < n; i++)
                    for (int h = 0; h < 2; h++)
                        for (uint64_t e = off[2 * i + h]; e < off[2 * i + h + 1]; e++) {
                            const uint64_t *q = &seg[4 * e];  // st en hap_index root_population
                            const std::vector<std::string> &names = in[q[3]].indv_id;
                            std::string who = (q[2] / 2 < names.size() ? names[q[2] / 2] : std::string("NA")) + (q[2] % 2 ? ".2" : ".1");  // :3031-3033
                            o << ids[i * 7] + 1 << ' ' << in[0].chrs[c].chr << ' ' << h << ' ' << q[0] << ' ' << q[1] << ' ' << q[2] + 1 << ' ' << who << ' ' << q[3] + 1 << '\n';
                        }
            }
        }
    }
    return true;
}

bool HostSimulation::run() {
    // phase timers like the reference's "Time taken for ..." lines (src/Simulation.cpp:70-114), in milliseconds
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        auto t1 = std::chrono::steady_clock::now();
        if (!opt.quiet) std::cout << "  Time taken for " << what << ": " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms" << std::endl;
        t0 = t1;
    };
    if (!load_inputs()) return false;
    if (!opt.quiet) std::cout << "  populations: " << n_pop << ", chromosomes: " << n_chr << ", phenotypes: " << n_phen << ", generations: " << tot_gen << std::endl;
    lap("reading the input files");
    if (!upload()) return false;
    lap("creating the device context and uploading the inputs");
    summary.assign(n_pop, {});
    if (ge_init_generation0(ctx, nullptr) != GE_OK) return gfail("ge_init_generation0");
    if (!after_generation(0)) return false;
    lap("initializing generation 0");
    std::vector<ge_gen_params> gp(n_pop);
    for (int gen = 1; gen <= tot_gen; gen++) {  // ras_main_sim :684-702
        for (int p = 0; p < n_pop; p++) {
            const GenRow &r = in[p].gens[gen - 1];
            gp[p].pop_size = r.pop_size; gp[p].mat_cor = r.mat_cor; gp[p].offspring_dist = r.offspring_dist;
            gp[p].selection_func = r.selection_func; gp[p].selection_par1 = r.par1; gp[p].selection_par2 = r.par2;
        }
        if (ge_step_generation(ctx, gen, gp.data(), n_pop > 1 ? migration[gen - 1].data() : nullptr, nullptr) != GE_OK) return gfail("ge_step_generation");
#ifndef GE_HOST_NO_COMPACT
        if (opt.compact_segments && (opt.out_interval || !need_panel))
            for (int p = 0; p < n_pop; p++) if (ge_compact_segments(ctx, p, nullptr, nullptr) != GE_OK) return gfail("ge_compact_segments");
#endif
        if (!after_generation(gen)) return false;
    }
    lap("the main body of simulation");
    if (info_writer) {
        info_writer->finish();
        if (!info_writer->error.empty()) return fail(info_writer->error);
    }
    bool ok = rank != 0 || write_summary();
    lap("finishing the output files");
    return ok;
}

}  // namespace gehost
