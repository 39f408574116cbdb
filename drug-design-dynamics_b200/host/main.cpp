// main.cpp -- the three entry points of metropolisMethod/main.go (:15-19) behind one binary:
//   ddd_metropolis run  [dir]               RunMultipleLigands()  (default dir Data/mol2_files)
//   ddd_metropolis rmsd [dir]               TestMethodRMSD()
//   ddd_metropolis <protein.mol2> <ligand.mol2>...   RShinyAppMain(os.Args): writes ../output/simulations.csv
#include "metropolis.hpp"
#include <iostream>

int main(int argc, char **argv) {
    using namespace ddd_host;
    try {
        std::vector<std::string> args(argv, argv + argc);
        if (argc >= 2 && args[1] == "run") return RunMultipleLigands(argc > 2 ? args[2] : "Data/mol2_files");
        if (argc >= 2 && args[1] == "rmsd") return TestMethodRMSD(argc > 2 ? args[2] : "Data/mol2_files");
        return RShinyAppMain(args);
    } catch (const GoPanic &e) {
        std::cerr << "panic: " << e.what() << std::endl;          // the reference's Check() panics (main.go:179-183)
        return 2;
    }
}
