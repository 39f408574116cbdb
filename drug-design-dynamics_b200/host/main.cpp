// main.cpp -- the three entry points of metropolisMethod/main.go (:15-19) behind one binary:
//   ddd_metropolis run  [dir]               RunMultipleLigands()  (default dir Data/mol2_files)
//   ddd_metropolis rmsd [dir]               TestMethodRMSD()
//   ddd_metropolis screen <protein.mol2> <ligand dir> <output dir> [iterations] [numProcs]   screen-scale driver: parallel mol2
//                                           ingest, one resident screen, argmin + UpdateMol2Coordinates + simulations.csv
//   ddd_metropolis parse <file.mol2>...     host only: ParseMol2Batch -> "<atoms> <sum x> <sum y> <sum z> <sum q>" per file (%.17g)
//   ddd_metropolis update <orig.mol2> <out.mol2> <dx>   host only: ParseMol2, x += dx, UpdateMol2Coordinates
//   ddd_metropolis <protein.mol2> <ligand.mol2>...   RShinyAppMain(os.Args): writes ../output/simulations.csv
#include "metropolis.hpp"
#include <cstdio>
#include <iostream>

int main(int argc, char **argv) {
    using namespace ddd_host;
    try {
        std::vector<std::string> args(argv, argv + argc);
        if (argc >= 2 && args[1] == "run") return RunMultipleLigands(argc > 2 ? args[2] : "Data/mol2_files");
        if (argc >= 2 && args[1] == "rmsd") return TestMethodRMSD(argc > 2 ? args[2] : "Data/mol2_files");
        if (argc >= 3 && args[1] == "parse") {
            std::vector<std::string> files(args.begin() + 2, args.end());
            std::vector<int64_t> off;
            std::vector<Atom> flat;
            ParseMol2Batch(files, off, flat, 0);
            for (size_t i = 0; i < files.size(); ++i) {
                double sx = 0, sy = 0, sz = 0, sq = 0;
                for (int64_t a = off[i]; a < off[i + 1]; ++a) { sx += flat[(size_t)a].Position.X; sy += flat[(size_t)a].Position.Y; sz += flat[(size_t)a].Position.Z; sq += flat[(size_t)a].Charge; }
                std::printf("%lld %.17g %.17g %.17g %.17g\n", (long long)(off[i + 1] - off[i]), sx, sy, sz, sq);
            }
            return 0;
        }
        if (argc == 5 && args[1] == "update") {
            Molecule m = ParseMol2(args[2]);
            for (auto &a : m.atoms) a.Position.X += std::stod(args[4]);
            UpdateMol2Coordinates(args[2], args[3], m);
            return 0;
        }
        if (argc >= 5 && args[1] == "screen")
            return RunScreen(args[2], args[3], args[4], argc > 5 ? std::stoi(args[5]) : 3000, argc > 6 ? std::stoi(args[6]) : 8);
        return RShinyAppMain(args);
    } catch (const GoPanic &e) {
        std::cerr << "panic: " << e.what() << std::endl;          // the reference's Check() panics (main.go:179-183)
        return 2;
    }
}
