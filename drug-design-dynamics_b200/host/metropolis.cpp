// metropolis.cpp -- see metropolis.hpp.  Heavy functions go through the C ABI; there is no CPU path for them.
#include "metropolis.hpp"
#include "../../include/ddd_mc.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
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

// The Go functions receive the protein on every call (metropolis.go:247, :267, :290, :94 as RShinyAppMain calls it per ligand,
// main.go:42-48): the library's content-keyed cache keeps it resident instead of rebuilding it per call.
struct Receptor {
    ddd_receptor *h = nullptr;
    explicit Receptor(const Molecule &protein) {
        init();
        check(ddd_receptor_acquire(ptr(protein), (int64_t)protein.atoms.size(), &h));
    }
    ~Receptor() { ddd_receptor_release(h); }
    Receptor(const Receptor &) = delete;
    Receptor &operator=(const Receptor &) = delete;
};

// a chain that needed more than max_retries collision retries in one step stopped there; the reference would still be
// retrying (metropolis.go:155, :173) -- say so instead of returning a truncated chain silently
void warn_if_truncated(const char *where) {
    int64_t n = 0;
    if (ddd_last_chain_status(&n) & DDD_STATUS_RETRY_LIMIT)
        std::cerr << "ddd_mc warning: " << n << " chain(s) of " << where << " stopped at the collision-retry bound; the reference would "
                     "still be retrying there (metropolis.go:173)" << std::endl;
}

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
    warn_if_truncated("SimulateEnergyMinimizationOneProc");
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
    // :14-16 shift every ligand (in place, as the reference's slice aliasing does), then :17-20 minimise: one upload, the
    // shift, the chains, the replica pick and the gather all run on the resident screen
    if (numProcs == 0) throw GoPanic("runtime error: integer divide by zero");                // :97
    Receptor rec(protein);
    std::vector<int64_t> offsets;
    std::vector<Atom> flat;
    pack(ligands, offsets, flat);
    std::vector<Atom> out(flat.size());
    std::vector<double> energy(ligands.size());
    ddd_params p = params();
    const uint64_t seed = next_seed();
    ddd_screen *scr = nullptr;
    check(ddd_screen_create(rec.h, offsets.data(), ptr(flat), (int64_t)ligands.size(), numProcs, 0, &scr));
    struct Guard { ddd_screen *s; ~Guard() { ddd_screen_destroy(s); } } guard{scr};
    check(ddd_screen_shift_closer(scr, THRESHOLD));
    check(ddd_screen_download_start(scr, ptr(flat)));
    for (size_t i = 0; i < ligands.size(); ++i) std::copy(flat.begin() + offsets[i], flat.begin() + offsets[i + 1], ligands[i].atoms.begin());
    check(ddd_screen_run(scr, iterations / numProcs, rotate ? 1 : 0, temperature, &p, seed));
    check(ddd_screen_pick(scr, temperature, seed, 0, ptr(out), energy.data(), nullptr, nullptr));
    std::vector<Molecule> minLigands(ligands.size());
    for (size_t i = 0; i < ligands.size(); ++i) minLigands[i].atoms.assign(out.begin() + offsets[i], out.begin() + offsets[i + 1]);
    return {minLigands, energy};
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
    warn_if_truncated("SimulateMultipleLigandsParallel");
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
    // :41 reference = CopyLigand(ligand); :42 RandomizeLigandPose; :43 SimulateEnergyMinimizationParallel; :44 CalculateRMSD --
    // on the resident screen: the uploaded pose stays behind as the reference, nothing but the RMSD comes back
    if (numProcs == 0) throw GoPanic("runtime error: integer divide by zero");
    Receptor rec(protein);
    const int64_t offsets[2] = {0, (int64_t)ligand.atoms.size()};
    ddd_params p = params();
    const uint64_t seed = next_seed();
    ddd_screen *scr = nullptr;
    check(ddd_screen_create(rec.h, offsets, ptr(ligand), 1, numProcs, 0, &scr));
    struct Guard { ddd_screen *s; ~Guard() { ddd_screen_destroy(s); } } guard{scr};
    check(ddd_screen_randomize(scr, seed, 0));
    check(ddd_screen_run(scr, iterations / numProcs, rotate ? 1 : 0, temperature, &p, seed));
    double rmsd = 0;
    check(ddd_screen_pick(scr, temperature, seed, 0, nullptr, nullptr, nullptr, &rmsd));
    return rmsd;
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
namespace {
// strconv.ParseFloat(s, 64) with the error discarded (parsing.go:85-88): a field that is not entirely a number yields 0
double parse_field(const char *b, const char *e) {
    if (b == e) return 0.0;
    char buf[64];
    const size_t n = (size_t)(e - b);
    if (n >= sizeof buf) return 0.0;
    std::memcpy(buf, b, n); buf[n] = '\0';
    char *end = nullptr;
    const double v = std::strtod(buf, &end);
    return (end == buf + n) ? v : 0.0;
}
bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// parsing.go:50-106 over a file already in memory: lines between "@<TRIPOS>ATOM" and the next "@<TRIPOS>", strings.Fields
// split, >= 9 fields, x/y/z = fields 2..4, charge = the LAST field
void parse_mol2_buffer(const std::string &text, std::vector<Atom> &atoms) {
    bool inAtoms = false;
    const char *p = text.data(), *end = p + text.size();
    while (p < end) {
        const char *eol = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        const size_t len = (size_t)(le - p);
        if (len >= 13 && std::memcmp(p, "@<TRIPOS>ATOM", 13) == 0) inAtoms = true;
        else if (inAtoms && len >= 9 && std::memcmp(p, "@<TRIPOS>", 9) == 0) break;
        else if (inAtoms) {
            const char *fb[4] = {nullptr, nullptr, nullptr, nullptr}, *fe[4] = {nullptr, nullptr, nullptr, nullptr};   // fields 2, 3, 4, last
            int nf = 0;
            const char *q = p;
            while (q < le) {
                while (q < le && is_space(*q)) ++q;
                if (q >= le) break;
                const char *b = q;
                while (q < le && !is_space(*q)) ++q;
                if (nf >= 2 && nf <= 4) { fb[nf - 2] = b; fe[nf - 2] = q; }
                fb[3] = b; fe[3] = q;
                ++nf;
            }
            if (nf >= 9) atoms.push_back({{parse_field(fb[0], fe[0]), parse_field(fb[1], fe[1]), parse_field(fb[2], fe[2])}, parse_field(fb[3], fe[3])});
        }
        p = eol ? eol + 1 : end;
    }
}

bool read_file(const std::string &filename, std::string &text) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0);
    text.resize((size_t)std::max<std::streamoff>(n, 0));
    if (n > 0) f.read(&text[0], n);
    return true;
}
}  // namespace

Molecule ParseMol2(const std::string &filename) {
    std::string text;
    if (!read_file(filename, text)) throw GoPanic("open " + filename + ": no such file or directory");
    Molecule m;
    parse_mol2_buffer(text, m.atoms);
    return m;
}

// Screen-scale ingest: every file parsed on its own host thread straight into the CSR staging buffers the C ABI takes
// (offsets[n+1] + flat xyzq), same order as `files`.  A 100 000-ligand sweep otherwise spends longer in ParseMol2 than on the GPU.
void ParseMol2Batch(const std::vector<std::string> &files, std::vector<int64_t> &offsets, std::vector<Atom> &flat, int threads) {
    const size_t n = files.size();
    std::vector<std::vector<Atom>> parsed(n);
    std::vector<std::string> errors(n);
    const int nt = (int)std::max<size_t>(1, std::min<size_t>(n, threads > 0 ? (size_t)threads : std::max(1u, std::thread::hardware_concurrency())));
    std::atomic<size_t> next{0};
    auto work = [&]() {
        std::string text;
        for (size_t i; (i = next.fetch_add(1)) < n;) {
            if (!read_file(files[i], text)) { errors[i] = "open " + files[i] + ": no such file or directory"; continue; }
            parse_mol2_buffer(text, parsed[i]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    for (auto &e : errors) if (!e.empty()) throw GoPanic(e);
    offsets.assign(n + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[i + 1] = offsets[i] + (int64_t)parsed[i].size();
    flat.resize((size_t)offsets[n]);
    for (size_t i = 0; i < n; ++i) std::copy(parsed[i].begin(), parsed[i].end(), flat.begin() + offsets[i]);
}

// ---- plotting.go:126-189: rewrite the ATOM section of `originalFile` with the molecule's coordinates ("%.4f"), fields joined
// by single spaces as strings.Join does; everything else copied line by line
void UpdateMol2Coordinates(const std::string &originalFile, const std::string &updatedFile, const Molecule &newMolecule) {
    std::string text;
    if (!read_file(originalFile, text)) throw GoPanic("failed to open file: open " + originalFile + ": no such file or directory");
    std::ofstream out(updatedFile, std::ios::binary);
    if (!out) throw GoPanic("failed to create output file: open " + updatedFile);
    size_t atomIndex = 0;
    bool inAtomSection = false;
    std::string acc;
    acc.reserve(text.size() + 64);
    const char *p = text.data(), *end = p + text.size();
    while (p < end) {                                                         // bufio.Scanner: lines without their terminator
        const char *eol = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        const char *le = eol ? eol : end;
        std::string line(p, le);
        if (!line.empty() && line.back() == '\r') line.pop_back();            // ScanLines drops a trailing \r
        p = eol ? eol + 1 : end;
        if (line.rfind("@<TRIPOS>ATOM", 0) == 0) { inAtomSection = true; acc += line; acc += '\n'; continue; }
        else if (line.rfind("@<TRIPOS>", 0) == 0 && inAtomSection) inAtomSection = false;
        if (inAtomSection && atomIndex < newMolecule.atoms.size()) {
            std::vector<std::string> parts;
            for (size_t i = 0; i < line.size();) {
                while (i < line.size() && is_space(line[i])) ++i;
                if (i >= line.size()) break;
                size_t j = i;
                while (j < line.size() && !is_space(line[j])) ++j;
                parts.emplace_back(line, i, j - i);
                i = j;
            }
            if (parts.size() >= 9) {
                const Atom &a = newMolecule.atoms[atomIndex];
                char buf[64];
                std::snprintf(buf, sizeof buf, "%.4f", a.Position.X); parts[2] = buf;
                std::snprintf(buf, sizeof buf, "%.4f", a.Position.Y); parts[3] = buf;
                std::snprintf(buf, sizeof buf, "%.4f", a.Position.Z); parts[4] = buf;
                for (size_t k = 0; k < parts.size(); ++k) { if (k) acc += ' '; acc += parts[k]; }
                acc += '\n';
                ++atomIndex;
            } else { acc += line; acc += '\n'; }
        } else { acc += line; acc += '\n'; }
    }
    if (atomIndex != newMolecule.atoms.size()) throw GoPanic("mismatch between number of atoms in molecule and file");   // before Flush: nothing written
    out << acc;
}

// plotting.go:109-121 -- note minEnergy starts at 0.0: only a negative energy can displace index 0
size_t SaveMinimumEnergyLigand(const std::vector<double> &energyList, const std::vector<std::string> &ligandFiles, const std::string &fileName,
                               const std::vector<Molecule> &minLigands) {
    size_t minIndex = 0;
    double minEnergy = 0.0;
    for (size_t i = 0; i < energyList.size(); ++i)
        if (energyList[i] < minEnergy) { minIndex = i; minEnergy = energyList[i]; }
    if (minIndex >= ligandFiles.size() || minIndex >= minLigands.size()) throw GoPanic("runtime error: index out of range");
    UpdateMol2Coordinates(ligandFiles[minIndex], fileName, minLigands[minIndex]);
    std::cout << "Wrote min ligand file to " << fileName;
    return minIndex;
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
    {
        std::vector<int64_t> off;
        std::vector<Atom> flat;
        ParseMol2Batch(ligandFiles, off, flat, 0);
        for (size_t i = 0; i < ligandFiles.size(); ++i) ligands.push_back(Molecule{{flat.begin() + off[i], flat.begin() + off[i + 1]}});
    }
    const std::string proteinFile = "223l_protein.mol2";
    const Molecule protein = ParseMol2(dir + "/" + proteinFile);
    std::cout << "Starting simulation" << std::endl;
    const auto t0 = std::chrono::steady_clock::now();
    auto res = SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, TEMPERATURE, numProcs);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Time taken for simulation:  " << sec << "s" << std::endl;
    // plotting.go (gonum PNG) is out of scope; the label -> energy table it plots is printed instead
    for (size_t i = 0; i < ligandFiles.size(); ++i) std::printf("%s %.6f\n", ExtractFileLabel(ligandFiles[i]).c_str(), res.second[i]);
    const std::string proteinPDB = ExtractFileLabel(proteinFile);
    const std::string outputDir = "Output/" + proteinPDB + "/";                // main.go:152-154
    std::error_code ec;
    std::filesystem::create_directories(outputDir, ec);
    if (ec) throw GoPanic("mkdir " + outputDir + ": " + ec.message());
    SaveMinimumEnergyLigand(res.second, ligandFiles, outputDir + "minLigand_" + proteinFile, res.first);      // main.go:157
    std::cout << std::endl;
    return 0;
}

// Screen-scale entry: every *ligand* mol2 file under `ligandDir` against one protein -- parallel ingest into the
// CSR staging buffers, one resident screen (chains, replica pick, SaveMinimumEnergyLigand's argmin on the device), the
// winner's mol2 rewritten by UpdateMol2Coordinates, the energies as the Shiny CSV (header BindingEnergy, %.6f).
int RunScreen(const std::string &proteinFile, const std::string &ligandDir, const std::string &outputDir, int iterations, int numProcs) {
    const auto t0 = std::chrono::steady_clock::now();
    auto ligandFiles = findFilesWithSubstring(ligandDir, "ligand");
    if (ligandFiles.empty()) throw GoPanic("no ligand files under " + ligandDir);
    if (numProcs <= 0) throw GoPanic("runtime error: integer divide by zero");
    const Molecule protein = ParseMol2(proteinFile);
    std::vector<int64_t> offsets;
    std::vector<Atom> flat;
    ParseMol2Batch(ligandFiles, offsets, flat, 0);
    const double t_parse = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    Receptor rec(protein);
    ddd_params p = params();
    const uint64_t seed = next_seed();
    ddd_screen *scr = nullptr;
    check(ddd_screen_create(rec.h, offsets.data(), ptr(flat), (int64_t)ligandFiles.size(), numProcs, 0, &scr));
    struct Guard { ddd_screen *s; ~Guard() { ddd_screen_destroy(s); } } guard{scr};
    check(ddd_screen_run(scr, iterations / numProcs, 1, TEMPERATURE, &p, seed));
    std::vector<Atom> out(flat.size());
    std::vector<double> energy(ligandFiles.size());
    check(ddd_screen_pick(scr, TEMPERATURE, seed, 0, ptr(out), energy.data(), nullptr, nullptr));
    int64_t minIndex = 0;
    double minEnergy = 0;
    check(ddd_screen_argmin(scr, DDD_ARGMIN_SAVE_MIN_LIGAND, &minIndex, &minEnergy));       // plotting.go:109-121 on the device
    const double t_gpu = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - t_parse;
    std::error_code ec;
    std::filesystem::create_directories(outputDir, ec);
    Molecule best{{out.begin() + offsets[(size_t)minIndex], out.begin() + offsets[(size_t)minIndex + 1]}};
    const std::string minFile = outputDir + "/minLigand_" + std::filesystem::path(proteinFile).filename().string();
    UpdateMol2Coordinates(ligandFiles[(size_t)minIndex], minFile, best);
    std::ofstream csv(outputDir + "/simulations.csv");
    csv << "BindingEnergy\n";
    char buf[64];
    for (double e : energy) { std::snprintf(buf, sizeof buf, "%.6f", e); csv << buf << "\n"; }
    std::printf("screen: %zu ligands (%lld atoms) parsed in %.3f s, minimised in %.3f s; min ligand %s (%.6f) -> %s\n", ligandFiles.size(),
                (long long)offsets.back(), t_parse, t_gpu, ExtractFileLabel(ligandFiles[(size_t)minIndex]).c_str(), minEnergy, minFile.c_str());
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
