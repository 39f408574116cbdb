// metropolis_test.cpp -- metropolisMethod/metropolis_test.go restated against the C++ drop-in API, test for test
// (same mocks, same assertions).  Needs a GPU: run by tests/test_gpu_host_cpp.py on the B200 box.
#include "metropolis.hpp"
#include <cmath>
#include <cstdio>
#include <functional>

using namespace ddd_host;

static int failures = 0;
#define EXPECT(cond, msg) do { if (!(cond)) { std::printf("    FAIL %s:%d %s\n", __FILE__, __LINE__, msg); ++failures; } } while (0)

static Molecule createMockProtein(double c1, double c2) { return {{{{0, 0, 0}, c1}, {{1, 1, 1}, c2}}}; }                 // :153
static Molecule createMockLigandWithCustomCharge(double c1, double c2) { return {{{{4, 4, 4}, c1}, {{3, 3, 3}, c2}}}; }  // :187
static Molecule createMockLigand() { return createMockLigandWithCustomCharge(1, -1); }                                   // :176
static std::vector<Molecule> createMockLigands(int n) { return std::vector<Molecule>((size_t)n, createMockLigand()); }   // :165
static bool almostEqual(double a, double b, double eps) { return std::fabs(a - b) <= eps; }                              // :200

static void TestSimulateMultipleLigands() {
    auto protein = createMockProtein(1.0, -1.0);
    auto ligands = createMockLigands(3);
    auto r = SimulateMultipleLigands(protein, ligands, 1000, true, 300.0, 4);
    EXPECT(r.first.size() == ligands.size() && r.second.size() == ligands.size(), "Expected 3 ligands and energies");
}
static void TestSimulateMultipleLigandsParallel() {
    auto protein = createMockProtein(1.0, -1.0);
    auto ligands = createMockLigands(5);
    auto r = SimulateMultipleLigandsParallel(protein, ligands, 2000, true, 300.0, 2);
    EXPECT(r.first.size() == 5 && r.second.size() == 5, "Expected 5 ligands and energies");
    for (size_t i = 0; i < 5; ++i) EXPECT(almostEqual(CalculateEnergy(protein, r.first[i]), r.second[i], 1e-6 * std::fabs(r.second[i])), "energy[i] == CalculateEnergy(ligand[i])");
}
static void TestCalculateEnergyPositive() {
    const double e = CalculateEnergy(createMockProtein(1.0, 1.0), createMockLigandWithCustomCharge(1.0, 1.0));
    EXPECT(e > 0, "Expected positive energy");
    EXPECT(almostEqual(e, 7361215932.1677294, 1e-5), "derived known answer (SURVEY.md section 4)");
}
static void TestCalculateEnergyNegative() {
    const double e = CalculateEnergy(createMockProtein(1.0, 1.0), createMockLigandWithCustomCharge(-1.0, -1.0));
    EXPECT(e < 0, "Expected negative energy");
}
static void TestAcceptMove() { EXPECT(AcceptMove(10.0, 5.0, 300.0), "Expected move to be accepted"); }
static void TestJitterLigand() {
    auto ligand = createMockLigand();
    EXPECT(IsCollisionFree(JitterLigand(ligand, 0.5), 0.5), "Expected collision-free ligand after jittering");
}
static void TestRotateLigand() {
    auto ligand = createMockLigand();
    const auto old = createMockLigand();
    RotateLigand(ligand, M_PI / 4);
    EXPECT(!(ligand.atoms[0].Position == old.atoms[0].Position), "Expected ligand to be rotated");
}
static void TestShiftLigandCloserByThreshold() {
    auto protein = createMockProtein(1.0, -1.0);
    auto ligand = createMockLigand();
    auto &shifted = ShiftLigandCloserByThreshold(ligand, protein, 5.0);
    const double d1 = Distance(shifted.atoms[0].Position, protein.atoms[0].Position);
    const double d2 = Distance(shifted.atoms[1].Position, protein.atoms[1].Position);
    EXPECT(!(d1 > 5.0 && d2 > 5.0), "Expected shifted ligand to be within threshold distance");
}
static void TestCopyLigand() {
    auto ligand = createMockLigandWithCustomCharge(2.3, -8.7);
    auto copied = CopyLigand(ligand);
    EXPECT(&ligand.atoms[0] != &copied.atoms[0], "Expected ligand to be deep copied");
    EXPECT(copied.atoms[0].Charge == 2.3 && copied.atoms[1].Charge == -8.7, "Ligand atom fields don't match");
}
static void TestDistance() { EXPECT(Distance({1, 2, 3}, {4, 6, 8}) == std::sqrt(50.0), "Expected distance sqrt(50) exactly"); }
static void TestRotateAtom() {
    const auto r = RotateAtom({1, 0, 0}, {0, 0, 1}, M_PI / 2);
    EXPECT(almostEqual(r.X, 0, 1e-6) && almostEqual(r.Y, 1, 1e-6) && almostEqual(r.Z, 0, 1e-6), "Expected rotated position (0,1,0)");
}
static void TestFindClosestAtomDistance() {
    const double d = FindClosestAtomDistance(createMockLigand(), createMockProtein(2.0, 4.0), nullptr, nullptr);
    EXPECT(d == std::sqrt(12.0), "Expected distance sqrt(12) exactly");
}
static void TestGoPanics() {        // error behaviour of the boundary (not in metropolis_test.go)
    bool threw = false;
    try { SimulateEnergyMinimizationParallel(createMockProtein(1, 1), createMockLigand(), 10, true, 300.0, 0); } catch (const GoPanic &) { threw = true; }
    EXPECT(threw, "numProcs == 0 must panic (integer divide by zero)");
    threw = false;
    try { ParseMol2("/nonexistent.mol2"); } catch (const GoPanic &) { threw = true; }
    EXPECT(threw, "missing file must panic like Check(err)");
    EXPECT(CalculateRMSD(createMockLigand(), createMockLigand()) == 0.0, "RMSD of identical poses is 0");
}

int main() {
    Seed(1);
    const std::pair<const char *, std::function<void()>> tests[] = {
        {"TestSimulateMultipleLigands", TestSimulateMultipleLigands}, {"TestSimulateMultipleLigandsParallel", TestSimulateMultipleLigandsParallel},
        {"TestCalculateEnergyPositive", TestCalculateEnergyPositive}, {"TestCalculateEnergyNegative", TestCalculateEnergyNegative},
        {"TestAcceptMove", TestAcceptMove}, {"TestJitterLigand", TestJitterLigand}, {"TestRotateLigand", TestRotateLigand},
        {"TestShiftLigandCloserByThreshold", TestShiftLigandCloserByThreshold}, {"TestCopyLigand", TestCopyLigand},
        {"TestDistance", TestDistance}, {"TestRotateAtom", TestRotateAtom}, {"TestFindClosestAtomDistance", TestFindClosestAtomDistance},
        {"TestGoPanics", TestGoPanics}};
    for (auto &t : tests) {
        const int before = failures;
        try { t.second(); } catch (const std::exception &e) { std::printf("    EXCEPTION %s\n", e.what()); ++failures; }
        std::printf("%s %s\n", failures == before ? "--- PASS:" : "--- FAIL:", t.first);
    }
    std::printf(failures ? "FAIL\n" : "PASS\n");
    return failures ? 1 : 0;
}
