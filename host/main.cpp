// geneevolve_b200_cli — the reference's command line (src/Main.cpp:26-103) over libgeneevolve_b200.so.
// With --gpus N one host thread drives one context per GPU (chromosome shards); the threads meet in the
// sum-allreduce of the partial genetic values (NCCL) once per population and generation.
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ge_host.hpp"

int main(int argc, char **argv) {
    std::vector<std::string> args(argv + 1, argv + argc);
    gehost::Options opt;
    if (!opt.parse(args)) {
        std::cout << opt.error << std::endl;
        return -1;
    }
    if (opt.help || args.empty()) {
        std::cout << gehost::Options::usage();
        return 0;
    }
    auto t0 = std::chrono::steady_clock::now();
    if (opt.gpus == 1) {
        gehost::HostSimulation sim(opt);
        if (!sim.run()) {  // like the reference: message on stdout, exit code -1 (src/Main.cpp:84-88)
            std::cout << sim.error() << std::endl;
            return -1;
        }
    } else {
        std::unique_ptr<gehost::Collective> coll(gehost::make_collective());
        std::string err;
        if (!coll || !coll->init(opt.gpus, opt.device, err)) {
            std::cout << "Error: cannot set up the collective for " << opt.gpus << " GPUs: " << err << std::endl;
            return -1;
        }
        std::vector<std::string> errors(opt.gpus);
        std::atomic<int> failed(0);
        std::vector<std::thread> th;
        for (int r = 0; r < opt.gpus; r++)
            th.emplace_back([&, r] {
                gehost::HostSimulation sim(opt, r, opt.gpus, coll.get());
                if (!sim.run()) {
                    errors[r] = sim.error();
                    failed++;
                    coll->abort();   // the other ranks may be waiting in the allreduce
                }
            });
        for (auto &t : th) t.join();
        if (failed) {
            for (int r = 0; r < opt.gpus; r++) if (!errors[r].empty()) std::cout << "rank " << r << ": " << errors[r] << std::endl;
            return -1;
        }
    }
    if (!opt.quiet)
        std::cout << "  Time taken for simulation: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " seconds." << std::endl;
    return 0;
}
