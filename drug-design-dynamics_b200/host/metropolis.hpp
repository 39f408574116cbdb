// metropolis.hpp -- C++ host-side mirror of the reference's Go API for the Monte-Carlo path
// (metropolisMethod/: datatypes.go, metropolis.go, validateResults.go, parsing.go, the entry points of main.go).
// Same names, argument order and error behaviour as the Go functions; the heavy ones call libddd_mc.so
// (include/ddd_mc.h).  The reference's toolchain (Go) is absent from the build image, so this is the
// compiled-language host side; go/metropolis_cuda.go is its Go twin.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace ddd_host {

// datatypes.go:6-12
constexpr double K = 9e9;
constexpr double THRESHOLD = 5;
constexpr double MINDISTANCE = 0.5;
constexpr double MAXANGLE = 0.79;
constexpr double TEMPERATURE = 310.15;

struct Position3d {                       // datatypes.go:23-25, :33-60
    double X = 0, Y = 0, Z = 0;
    void Normalize();                     // no-op when the magnitude is 0
    double Magnitude() const;
    double Dot(const Position3d &o) const;
    Position3d Add(const Position3d &o) const;
    Position3d Scale(double s) const;
    bool operator==(const Position3d &o) const { return X == o.X && Y == o.Y && Z == o.Z; }
};

struct Atom {                             // datatypes.go:18-21 -- 4 doubles: the C ABI's xyzq
    Position3d Position;
    double Charge = 0;
};
static_assert(sizeof(Atom) == 32, "Atom must be 4 doubles (X, Y, Z, Charge)");

struct Molecule {                         // datatypes.go:14-16
    std::vector<Atom> atoms;
};

struct MultipleLigandSimulationOutput {   // datatypes.go:27-30
    std::vector<Molecule> Ligand;
    std::vector<double> Energy;
};

struct GoPanic : std::runtime_error {     // Check() -> panic (main.go:179-183), integer divide by zero, index out of range
    using std::runtime_error::runtime_error;
};

void Seed(uint64_t seed);                 // pins the Philox seed (Go: rand.Seed); default is random like Go >= 1.20

// ---- metropolis.go (device-backed)
std::pair<std::vector<Molecule>, std::vector<double>> SimulateMultipleLigands(const Molecule &protein, std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs);          // :11
std::pair<std::vector<Molecule>, std::vector<double>> SimulateMultipleLigandsParallel(const Molecule &protein, const std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs); // :27
MultipleLigandSimulationOutput SimulateLigandMinimizationOneProc(const Molecule &protein, const std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs);                  // :56 (channel -> return value)
Molecule SimulateEnergyMinimization(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature);                                                                            // :72
Molecule SimulateEnergyMinimizationParallel(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature, int numProcs);                                                      // :94
Molecule SimulateEnergyMinimizationOneProc(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature);                                                                     // :116 (channel -> return value)
double CalculateEnergy(const Molecule &protein, const Molecule &ligand);                                                                                                                                         // :247
Molecule &ShiftLigandCloserByThreshold(Molecule &ligand, const Molecule &protein, double threshold);        // :267, in place
double FindClosestAtomDistance(const Molecule &ligand, const Molecule &protein, Position3d *ligandAtom, Position3d *proteinAtom);   // :290

// ---- metropolis.go (host scalar helpers, as in the reference)
bool AcceptMove(double currentEnergy, double newEnergy, double temperature);                                // :141
Molecule &JitterLigand(Molecule &ligand, double minDistance);                                               // :154, in place
Molecule &JitterAndRotateLigand(Molecule &ligand, double minDistance, double maxAngle);                     // :172, in place
Molecule &RotateLigand(Molecule &ligand, double maxAngle);                                                  // :189, in place
Position3d RotateAtom(const Position3d &pos, const Position3d &axis, double theta);                         // :213
bool IsCollisionFree(const Molecule &ligand, double minDistance);                                           // :229
double Distance(const Position3d &a, const Position3d &b);                                                  // :309
Molecule CopyLigand(const Molecule &ligand);                                                                // :319

// ---- validateResults.go
double CalculateRMSD(const Molecule &simulated, const Molecule &reference);                                 // :50
std::vector<double> CalculateRMSDBatch(const std::vector<Molecule> &simulated, const std::vector<Molecule> &reference);
Molecule &RandomizeLigandPose(Molecule &ligand);                                                            // :70, in place
double CompareRMSD(const Molecule &protein, Molecule ligand, int iterations, bool rotate, double temperature, int numProcs);        // :40
double MultipleProteinRMSD(const std::string &dir, int iterations, bool rotate, int numProteins, int numProcs);                     // :14 (returns the average; plot omitted)
double average(const std::vector<double> &arr);                                                             // :84

// ---- parsing.go / plotting.go helpers on the path
Molecule ParseMol2(const std::string &filename);                                                            // parsing.go:50
// screen-scale ingest: all files parsed on host threads straight into the C ABI's CSR staging buffers (threads <= 0: all cores)
void ParseMol2Batch(const std::vector<std::string> &files, std::vector<int64_t> &offsets, std::vector<Atom> &flat, int threads);
std::vector<std::string> findFilesWithSubstring(const std::string &rootDir, const std::string &searchString);   // parsing.go:142
std::string ExtractFileLabel(const std::string &filePath);                                                  // plotting.go:95
void UpdateMol2Coordinates(const std::string &originalFile, const std::string &updatedFile, const Molecule &newMolecule);       // plotting.go:126
size_t SaveMinimumEnergyLigand(const std::vector<double> &energyList, const std::vector<std::string> &ligandFiles, const std::string &fileName,
                               const std::vector<Molecule> &minLigands);                                  // plotting.go:109 (returns minIndex)

// ---- main.go entry points
int RunMultipleLigands(const std::string &dir);                                                             // main.go:125
int TestMethodRMSD(const std::string &dir);                                                                 // main.go:116
int RShinyAppMain(const std::vector<std::string> &args);                                                    // main.go:21 (argv -> ../output/simulations.csv)
int RunScreen(const std::string &proteinFile, const std::string &ligandDir, const std::string &outputDir, int iterations, int numProcs);   // screen-scale driver (no reference row)

}  // namespace ddd_host
