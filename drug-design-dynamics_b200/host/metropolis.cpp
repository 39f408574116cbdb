// metropolis.cpp -- see metropolis.hpp.  Heavy functions go through the C ABI; there is no CPU path for them.
#include "metropolis.hpp"
#include "../../include/ddd_mc.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <thread>

namespace ddd_host {
namespace {

std::once_flag g_once;
std::atomic<uint64_t> g_calls{0};
uint64_t g_seed = std::random_device{}() * 0x9E3779B97F4A7C15ull + std::random_device{}();
std::mt19937_64 g_host_rng{std::random_device{}()};          // host helpers only (Go: the global math/rand source)

double rand_float64() { return (double)(g_host_rng() >> 11) * (1.0 / 9007199254740992.0); }

void check(int rc) {
    if (rc != DDD_OK) throw GoPanic(std::string("ddd_mc: ") + ddd_last_error());
}

void init() {
    std::call_once(g_once, [] { check(ddd_init(nullptr, 0)); });
}

uint64_t next_seed() { return g_seed + 0x9E3779B97F4A7C15ull * g_calls.fetch_add(1); }

ddd_params params() {
    ddd_params p;
    ddd_params_default(&p);
    return p;
}

const double *ptr(const Molecule &m) { return m.atoms.empty() ? nullptr : &m.atoms[0].Position.X; }
double *ptr(Molecule &m) { return m.atoms.empty() ? nullptr : &m.atoms[0].Position.X; }

struct Receptor {
    ddd_receptor *h = nullptr;
    explicit Receptor(const Molecule &protein) {
        init();
        check(ddd_receptor_create(ptr(protein), (int64_t)protein.atoms.size(), &h));
    }
    ~Receptor() { ddd_receptor_destroy(h); }
};

void pack(const std::vector<Molecule> &ligands, std::vector<int64_t> &offsets, std::vector<Atom> &flat) {
    offsets.assign(ligands.size() + 1, 0);
    for (size_t i = 0; i < ligands.size(); ++i) offsets[i + 1] = offsets[i] + (int64_t)ligands[i].atoms.size();
    flat.clear();
    flat.reserve((size_t)offsets.back());
    for (auto &l : ligands) flat.insert(flat.end(), l.atoms.begin(), l.atoms.end());
}

const double *ptr(const std::vector<Atom> &a) { return a.empty() ? nullptr : &a[0].Position.X; }
double *ptr(std::vector<Atom> &a) { return a.empty() ? nullptr : &a[0].Position.X; }

Molecule one_chain(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature) {
    Receptor rec(protein);
    int64_t offsets[2] = {0, (int64_t)ligand.atoms.size()};
    Molecule out;
    out.atoms.resize(ligand.atoms.size());
    ddd_chain_stats st;
    ddd_params p = params();
    check(ddd_run_chains(rec.h, offsets, ptr(ligand), 1, 1, iterations, rotate ? 1 : 0, temperature, &p, next_seed(), 0,
                         nullptr, nullptr, ptr(out), &st, nullptr));
    return out;
}

}  // namespace

void Seed(uint64_t seed) { g_seed = seed; g_calls = 0; g_host_rng.seed(seed); }

// ---- datatypes.go
double Position3d::Magnitude() const { return std::sqrt(X * X + Y * Y + Z * Z); }
void Position3d::Normalize() { const double mag = Magnitude(); if (mag > 0) { X /= mag; Y /= mag; Z /= mag; } }
double Position3d::Dot(const Position3d &o) const { return X * o.X + Y * o.Y + Z * o.Z; }
Position3d Position3d::Add(const Position3d &o) const { return {X + o.X, Y + o.Y, Z + o.Z}; }
Position3d Position3d::Scale(double s) const { return {X * s, Y * s, Z * s}; }

// ---- metropolis.go, device-backed
std::pair<std::vector<Molecule>, std::vector<double>> SimulateMultipleLigands(const Molecule &protein, std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs) {
    for (auto &l : ligands) ShiftLigandCloserByThreshold(l, protein, THRESHOLD);          // :14-16, in place
    return SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs);
}

std::pair<std::vector<Molecule>, std::vector<double>> SimulateMultipleLigandsParallel(const Molecule &protein, const std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs) {
    if (numProcs == 0) throw GoPanic("runtime error: integer divide by zero");                // :34
    Receptor rec(protein);
    std::vector<int64_t> offsets;
    std::vector<Atom> flat;
    pack(ligands, offsets, flat);
    std::vector<Atom> out(flat.size());
    std::vector<double> energy(ligands.size());
    ddd_params p = params();
    check(ddd_minimize_batch(rec.h, offsets.data(), ptr(flat), (int64_t)ligands.size(), iterations, rotate ? 1 : 0, temperature,
                             numProcs, &p, next_seed(), 0, ptr(out), energy.data(), nullptr, nullptr));
    std::vector<Molecule> minLigands(ligands.size());
    for (size_t i = 0; i < ligands.size(); ++i) minLigands[i].atoms.assign(out.begin() + offsets[i], out.begin() + offsets[i + 1]);
    return {minLigands, energy};
}

MultipleLigandSimulationOutput SimulateLigandMinimizationOneProc(const Molecule &protein, const std::vector<Molecule> &ligands, int iterations, bool rotate, double temperature, int numProcs) {
    auto r = SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs);
    return {r.first, r.second};
}

Molecule SimulateEnergyMinimizationParallel(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature, int numProcs) {
    return SimulateMultipleLigandsParallel(protein, {ligand}, iterations, rotate, temperature, numProcs).first[0];
}

Molecule SimulateEnergyMinimization(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature) {
    return one_chain(protein, ligand, iterations, rotate, temperature);      // dead code in the reference (:72); kept for API parity
}

Molecule SimulateEnergyMinimizationOneProc(const Molecule &protein, const Molecule &ligand, int iterations, bool rotate, double temperature) {
    return one_chain(protein, ligand, iterations, rotate, temperature);
}

double CalculateEnergy(const Molecule &protein, const Molecule &ligand) {
    Receptor rec(protein);
    int64_t offsets[2] = {0, (int64_t)ligand.atoms.size()};
    ddd_params p = params();
    p.precision = DDD_PRECISION_F64;                                          // float64 with the reference's per-term arithmetic
    double e = 0;
    check(ddd_energy_batch(rec.h, offsets, ptr(ligand), 1, &p, &e, nullptr));
    return e;
}

Molecule &ShiftLigandCloserByThreshold(Molecule &ligand, const Molecule &protein, double threshold) {
    Receptor rec(protein);
    int64_t offsets[2] = {0, (int64_t)ligand.atoms.size()};
    check(ddd_shift_closer_batch(rec.h, offsets, ptr(ligand), 1, threshold));
    return ligand;
}

double FindClosestAtomDistance(const Molecule &ligand, const Molecule &protein, Position3d *ligandAtom, Position3d *proteinAtom) {
    Receptor rec(protein);
    int64_t offsets[2] = {0, (int64_t)ligand.atoms.size()};
    double d = 0;
    int64_t li = -1, pj = -1;
    check(ddd_closest_atom_batch(rec.h, offsets, ptr(ligand), 1, &d, &li, &pj));
    if (ligandAtom) *ligandAtom = li >= 0 ? ligand.atoms[(size_t)li].Position : Position3d{};
    if (proteinAtom) *proteinAtom = pj >= 0 ? protein.atoms[(size_t)pj].Position : Position3d{};
    return d;
}

// ---- metropolis.go, host scalar helpers
bool AcceptMove(double currentEnergy, double newEnergy, double temperature) {
    if (newEnergy < currentEnergy) return true;
    const double deltaE = newEnergy - currentEnergy;
    const double probability = std::exp(-deltaE / temperature);
    return rand_float64() < probability;
}

static void jitter(Molecule &m, double width) {
    for (auto &a : m.atoms) {
        a.Position.X += (rand_float64() - 0.5) * width;
        a.Position.Y += (rand_float64() - 0.5) * width;
        a.Position.Z += (rand_float64() - 0.5) * width;
    }
}

Molecule &JitterLigand(Molecule &ligand, double minDistance) {
    for (;;) { jitter(ligand, 0.1); if (IsCollisionFree(ligand, minDistance)) return ligand; }
}

Molecule &JitterAndRotateLigand(Molecule &ligand, double minDistance, double maxAngle) {
    for (;;) { RotateLigand(ligand, maxAngle); jitter(ligand, 0.1); if (IsCollisionFree(ligand, minDistance)) return ligand; }
}

Molecule &RotateLigand(Molecule &ligand, double maxAngle) {
    Position3d axis{rand_float64() * 2 - 1, rand_float64() * 2 - 1, rand_float64() * 2 - 1};
    axis.Normalize();
    const double theta = (rand_float64() * 2 - 1) * maxAngle;
    for (auto &a : ligand.atoms) a.Position = RotateAtom(a.Position, axis, theta);
    return ligand;
}

Position3d RotateAtom(const Position3d &pos, const Position3d &axis, double theta) {
    const double cosTheta = std::cos(theta), sinTheta = std::sin(theta), dot = pos.Dot(axis);
    return {pos.X * cosTheta + sinTheta * (axis.Y * pos.Z - axis.Z * pos.Y) + axis.X * dot * (1 - cosTheta),
            pos.Y * cosTheta + sinTheta * (axis.Z * pos.X - axis.X * pos.Z) + axis.Y * dot * (1 - cosTheta),
            pos.Z * cosTheta + sinTheta * (axis.X * pos.Y - axis.Y * pos.X) + axis.Z * dot * (1 - cosTheta)};
}

bool IsCollisionFree(const Molecule &ligand, double minDistance) {
    for (size_t i = 0; i < ligand.atoms.size(); ++i)
        for (size_t j = i + 1; j < ligand.atoms.size(); ++j)
            if (Distance(ligand.atoms[i].Position, ligand.atoms[j].Position) < minDistance) return false;
    return true;
}

double Distance(const Position3d &a, const Position3d &b) {
    const double dx = a.X - b.X, dy = a.Y - b.Y, dz = a.Z - b.Z;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}

Molecule CopyLigand(const Molecule &ligand) { return Molecule{ligand.atoms}; }

// ---- validateResults.go
std::vector<double> CalculateRMSDBatch(const std::vector<Molecule> &simulated, const std::vector<Molecule> &reference) {
    init();
    std::vector<int64_t> offsets;
    std::vector<Atom> sim, ref;
    pack(simulated, offsets, sim);
    for (size_t i = 0; i < simulated.size(); ++i) {
        if (reference[i].atoms.size() < simulated[i].atoms.size()) throw GoPanic("runtime error: index out of range");
        ref.insert(ref.end(), reference[i].atoms.begin(), reference[i].atoms.begin() + (long)simulated[i].atoms.size());
    }
    std::vector<double> out(simulated.size());
    check(ddd_rmsd_batch(offsets.data(), ptr(sim), ptr(ref), 4, (int64_t)simulated.size(), out.data()));
    return out;
}

double CalculateRMSD(const Molecule &simulated, const Molecule &reference) { return CalculateRMSDBatch({simulated}, {reference})[0]; }

Molecule &RandomizeLigandPose(Molecule &ligand) {
    jitter(ligand, 10.0);                                                    // :73-75
    return RotateLigand(ligand, 3.14159265358979323846);                     // :78
}

double CompareRMSD(const Molecule &protein, Molecule ligand, int iterations, bool rotate, double temperature, int numProcs) {
    const Molecule reference = CopyLigand(ligand);
    RandomizeLigandPose(ligand);
    const Molecule simulated = SimulateEnergyMinimizationParallel(protein, ligand, iterations, rotate, temperature, numProcs);
    return CalculateRMSD(simulated, reference);
}

double average(const std::vector<double> &arr) {
    if (arr.empty()) return 0;
    double sum = 0;
    for (double v : arr) sum += v;
    return sum / (double)arr.size();
}

double MultipleProteinRMSD(const std::string &dir, int iterations, bool rotate, int numProteins, int numProcs) {
    auto proteinFiles = findFilesWithSubstring(dir, "protein");
    if ((int)proteinFiles.size() < numProteins) throw GoPanic("runtime error: slice bounds out of range");
    proteinFiles.resize((size_t)numProteins);
    std::vector<double> rmsd;
    for (auto &pf : proteinFiles) {
        const Molecule protein = ParseMol2(pf);
        const std::string label = ExtractFileLabel(pf);
        const Molecule ligand = ParseMol2(dir + "/" + label + "_ligand.mol2");
        rmsd.push_back(CompareRMSD(protein, ligand, iterations, rotate, TEMPERATURE, numProcs));
        std::cout << label << " RMSD " << rmsd.back() << "\n";
    }
    const double avg = average(rmsd);
    std::cout << "The average RMSD value was: " << avg << std::endl;
    return avg;
}

// ---- parsing.go
Molecule ParseMol2(const std::string &filename) {
    std::ifstream f(filename);
    if (!f) throw GoPanic("open " + filename + ": no such file or directory");
    Molecule m;
    std::string line;
    bool inAtoms = false;
    auto num = [](const std::string &s) { try { size_t n = 0; double v = std::stod(s, &n); return n == s.size() ? v : 0.0; } catch (...) { return 0.0; } };
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.rfind("@<TRIPOS>ATOM", 0) == 0) { inAtoms = true; continue; }
        if (inAtoms && line.rfind("@<TRIPOS>", 0) == 0) break;
        if (!inAtoms) continue;
        std::istringstream ss(line);
        std::vector<std::string> fields;
        for (std::string t; ss >> t;) fields.push_back(t);
        if (fields.size() < 9) continue;
        m.atoms.push_back({{num(fields[2]), num(fields[3]), num(fields[4])}, num(fields.back())});
    }
    return m;
}

std::vector<std::string> findFilesWithSubstring(const std::string &rootDir, const std::string &searchString) {
    std::vector<std::string> out;
    for (auto &e : std::filesystem::recursive_directory_iterator(rootDir))
        if (e.is_regular_file() && e.path().filename().string().find(searchString) != std::string::npos) out.push_back(e.path().string());
    std::sort(out.begin(), out.end());                                       // filepath.Walk visits in lexical order
    return out;
}

std::string ExtractFileLabel(const std::string &filePath) {
    std::string base = std::filesystem::path(filePath).filename().string();
    std::string label = base.substr(0, base.find('_'));
    if (label.size() >= 5 && label.compare(label.size() - 5, 5, ".mol2") == 0) label.resize(label.size() - 5);
    return label;
}

// ---- main.go
int RunMultipleLigands(const std::string &dir) {
    auto ligandFiles = findFilesWithSubstring(dir, "ligand");
    const int iterations = 3000;
    const bool rotate = true;
    const int numProcs = (int)std::max(1u, std::thread::hardware_concurrency());      // runtime.NumCPU()
    if (ligandFiles.size() < 5) throw GoPanic("runtime error: slice bounds out of range [:5]");
    ligandFiles.resize(5);
    std::vector<Molecule> ligands;
    for (auto &f : ligandFiles) ligands.push_back(ParseMol2(f));
    const std::string proteinFile = "223l_protein.mol2";
    const Molecule protein = ParseMol2(dir + "/" + proteinFile);
    std::cout << "Starting simulation" << std::endl;
    const auto t0 = std::chrono::steady_clock::now();
    auto res = SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, TEMPERATURE, numProcs);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Time taken for simulation:  " << sec << "s" << std::endl;
    // plotting.go (gonum PNG) is out of scope; the label -> energy table it plots is printed instead
    for (size_t i = 0; i < ligandFiles.size(); ++i) std::printf("%s %.6f\n", ExtractFileLabel(ligandFiles[i]).c_str(), res.second[i]);
    return 0;
}

int TestMethodRMSD(const std::string &dir) {
    MultipleProteinRMSD(dir, 1000, false, 2, (int)std::max(1u, std::thread::hardware_concurrency()));
    return 0;
}

int RShinyAppMain(const std::vector<std::string> &args) {
    if (args.size() < 3) { std::cout << "Please provide right inputs." << std::endl; return 0; }
    const std::string proteinFilePath = args[1];
    const Molecule protein = ParseMol2(proteinFilePath);
    std::vector<std::string> results;
    double minEnergy = 1e20;
    std::string minEnergyLigand;
    const int numProcs = (int)std::max(1u, std::thread::hardware_concurrency());
    for (size_t i = 2; i < args.size(); ++i) {
        const Molecule ligand = ParseMol2(args[i]);
        const Molecule newLigand = SimulateEnergyMinimizationParallel(protein, ligand, 3000, true, TEMPERATURE, numProcs);
        const double newEnergy = CalculateEnergy(protein, newLigand);
        if (newEnergy < minEnergy) { minEnergy = newEnergy; minEnergyLigand = args[i]; }
        char buf[64];
        std::snprintf(buf, sizeof buf, "%.6f", newEnergy);                    // strconv.FormatFloat(e, 'f', 6, 64)
        results.push_back(buf);
        std::cout << "Finish i ligand here!: " << (i - 2) << std::endl;
    }
    namespace fs = std::filesystem;
    const fs::path outputDir = "../output";
    std::error_code ec;
    if (!minEnergyLigand.empty()) {
        fs::copy_file(minEnergyLigand, outputDir / fs::path(minEnergyLigand).filename(), fs::copy_options::overwrite_existing, ec);
        if (ec) { std::cerr << "Failed to copy the ligand file: " << ec.message() << std::endl; return 1; }
        std::cout << "Ligand with minimum energy copied to: " << (outputDir / fs::path(minEnergyLigand).filename()).string() << std::endl;
    }
    fs::copy_file(proteinFilePath, outputDir / fs::path(proteinFilePath).filename(), fs::copy_options::overwrite_existing, ec);
    if (ec) { std::cerr << "Failed to copy the protein file: " << ec.message() << std::endl; return 1; }
    std::ofstream csv(outputDir / "simulations.csv");
    if (!csv) { std::cerr << "Failed to create output file" << std::endl; return 1; }
    csv << "BindingEnergy\n";                                                 // main.go:101, read by Rshiny/app.R:186-188
    for (auto &r : results) csv << r << "\n";
    std::cout << "Binding Energy data written to simulations.csv" << std::endl;
    return 0;
}

}  // namespace ddd_host
