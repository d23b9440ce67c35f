// ge_host.hpp — C++ host of libgeneevolve_b200.so: the reference's command line and file formats for the
// per-generation reproduction path, on top of the C-ABI (include/geneevolve_b200.h).
//
// The reference keeps parameters.cpp, the text loaders of Population.cpp and the format_* writers on the host and
// runs the generation loop itself; this host keeps the same flags and formats and hands the loop to the GPU:
//   Options::parse          <- Parameters::read / check            (src/parameters.cpp:15-382)
//   read_* functions        <- Population::ras_read_* loaders      (src/Population.cpp:13-468), format_hap readers
//   HostSimulation::run     <- Simulation::run / ras_main_sim      (src/Simulation.cpp:68-161, 684-702)
//   write_info / summary    <- Population::ras_save_human_info     (src/Population.cpp:510-568), ras_save_res (:782-834)
//   write_hap / write_int   <- ras_write_hap_legend_sample (:1142-1182), ras_write_hap_to_interval_format (:1582-1639)
//   write_ped / write_map   <- ras_write_hap_to_plink_format (:1254-1303), format_plink::write_ped_map / write_ped01_map
// Nothing is computed here: every number in the outputs comes out of the library.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "geneevolve_b200.h"

namespace gehost {

struct GenRow { uint64_t pop_size; double mat_cor; char offspring_dist; int selection_func; double par1, par2; };

struct ChrFiles { int chr; std::string hap, legend, sample; };

struct GeneticMap { std::vector<uint64_t> bp; std::vector<double> cM, recom_prob; uint64_t bp_dist = 0; };
struct MutationMap { std::vector<uint64_t> bp; std::vector<double> rate; };
struct CvBlock { std::vector<uint64_t> bp; std::vector<double> a, d; std::vector<uint8_t> val; uint64_t n_hap = 0; };  // val[h*ncv + k]

struct PopOptions {  // the per-population flag group (flags before / between --next_population)
    std::string file_gen_info, file_hap_name, file_recom_map, file_mutation_map;
    std::vector<std::string> file_cv_info, file_cvs;
    std::vector<double> va, vd, vc, ve, vf, omega, beta, lambda;
    double MM = 0;
    bool RM = false;
};

struct Options {
    std::vector<PopOptions> pop;
    std::vector<double> gamma;
    std::string file_migration, file_output_generations, prefix = "out";
    int vt_type = 1, device = 0, gpus = 1;   // --gpus N: chromosomes are spread over devices device .. device+N-1
    bool avoid_inbreeding = false, out_hap = false, out_interval = false, out_plink = false, out_plink01 = false, quiet = false, help = false;
    bool compact_segments = false;           // extension (ge_compact_segments): the .int output is then not the reference's
    uint64_t seed = 0;
    std::string error;
    bool parse(const std::vector<std::string> &args);  // false + error on a bad command line
    static const char *usage();
};

struct PopInputs {
    std::vector<GenRow> gens;
    std::vector<ChrFiles> chrs;
    std::vector<std::string> indv_id;                 // founder sample names (gen0_indv labels of the .int file)
    std::vector<GeneticMap> rmap;                     // [chr]
    std::vector<MutationMap> mutmap;                  // [chr] or empty
    std::vector<std::vector<CvBlock>> cv;             // [phen][chr]
    std::vector<std::vector<uint64_t>> legend_pos;    // [chr] (only when genotypes are needed)
    std::vector<std::vector<std::string>> legend_id, legend_al0, legend_al1;
};

// text readers; each returns false and fills err on failure
bool read_generation_info(const std::string &path, std::vector<GenRow> &out, std::string &err);
bool read_hap_address(const std::string &path, std::vector<ChrFiles> &out, std::string &err);
bool read_recombination_map(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<GeneticMap> &out, std::string &err);
bool read_mutation_map(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<MutationMap> &out, std::string &err);
bool read_cv_info(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<CvBlock> &out, std::string &err);
bool read_cvs(const std::string &path, const std::vector<ChrFiles> &chrs, std::vector<CvBlock> &io, std::string &err);
bool read_legend(const std::string &path, std::vector<std::string> &id, std::vector<uint64_t> &pos, std::vector<std::string> *al0,
                 std::vector<std::string> *al1, std::string &err);
bool read_indv(const std::string &path, std::vector<std::string> &out, std::string &err);
// IMPUTE2 .hap (rows = SNPs, columns = haplotypes) -> bit-packed hap-major words (ge_set_founder_panel_packed layout)
bool read_hap_packed(const std::string &path, uint64_t n_hap, uint64_t n_snp, std::vector<uint32_t> &words, std::string &err);
bool count_hap_columns(const std::string &path, uint64_t &n_hap, std::string &err);
bool read_migration(const std::string &path, int n_pop, size_t n_gen, std::vector<std::vector<double>> &out, std::string &err);
bool read_output_generations(const std::string &path, std::vector<int> &out, std::string &err);

class InfoWriter;  // background formatter/writer of the per-generation .info files (ge_host.cpp)

// Sum-allreduce between the per-GPU contexts of one run (the library's only exchange step, ge_set_allreduce).  The
// product binary links the NCCL implementation (ge_collective_nccl.cpp); the CPU tests of the host link a
// thread-barrier implementation over host buffers.
class Collective {
public:
    virtual ~Collective() {}
    virtual bool init(int world, int first_device, std::string &err) = 0;   // once, before the rank threads start
    virtual int allreduce_sum(int rank, double *buf, uint64_t count, void *stream) = 0;
    virtual void abort() = 0;                                               // a rank failed: release the others
};
Collective *make_collective();

class HostSimulation {
public:
    explicit HostSimulation(const Options &o, int rank = 0, int world = 1, Collective *coll = nullptr);
    ~HostSimulation();
    bool run();                       // Simulation::run
    const std::string &error() const { return err; }
    // longest-processing-time assignment of chromosomes (weights: map span in bp) to ranks
    static std::vector<std::vector<int>> assign_chromosomes(const std::vector<double> &weight, int world);

private:
    Options opt;
    std::vector<PopInputs> in;
    std::vector<std::vector<double>> migration;   // [gen][n_pop*n_pop]
    std::vector<int> output_generations;
    ge_ctx *ctx = nullptr;
    std::unique_ptr<InfoWriter> info_writer;
    int rank = 0, world = 1;
    Collective *coll = nullptr;
    std::vector<int> mine;            // global indices of the chromosomes this rank owns (all of them when world == 1)
    static int allreduce_hook(void *user, double *buf, uint64_t count, void *stream);
    int n_pop = 0, n_chr = 0, n_phen = 0, tot_gen = 0;
    bool need_panel = false;
    std::string err;
    struct SummaryRow { std::vector<ge_moments> m; double var_mv, var_sv; };
    std::vector<std::vector<SummaryRow>> summary;  // [pop][gen]

    bool load_inputs();               // ras_init_parameters (:164-525)
    bool upload();                    // flat arrays -> ge_set_*
    bool after_generation(int gen);   // ras_save_human_info + the variance report (:2014-2055) + genotype output (:2059-2063)
    bool write_info(int pop, int gen);
    bool write_summary();
    bool write_genotypes(int gen);
    bool fail(const std::string &m) { err = m; return false; }
    bool gfail(const char *what);
};

}  // namespace gehost
