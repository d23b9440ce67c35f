// geneevolve_b200_cli — the reference's command line (src/Main.cpp:26-103) over libgeneevolve_b200.so.
#include <chrono>
#include <iostream>
#include <string>
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
    gehost::HostSimulation sim(opt);
    if (!sim.run()) {  // like the reference: message on stdout, exit code -1 (src/Main.cpp:84-88)
        std::cout << sim.error() << std::endl;
        return -1;
    }
    if (!opt.quiet)
        std::cout << "  Time taken for simulation: " << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " seconds." << std::endl;
    return 0;
}
