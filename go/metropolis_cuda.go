// metropolis_cuda.go -- drop-in replacement for metropolisMethod/metropolis.go (and for CalculateRMSD /
// RandomizeLigandPose's heavy callers in validateResults.go) that runs the Monte-Carlo hot path on B200 GPUs
// through the C ABI of libddd_mc.so (include/ddd_mc.h).
//
// HOW TO USE (INTEGRATION.md has the full recipe):
//   1. copy this file and ddd_mc.h next to the reference's main.go / datatypes.go / parsing.go / plotting.go /
//      validateResults.go, and DELETE metropolis.go (every function it defined is re-defined here with the same
//      signature) and the body of CalculateRMSD in validateResults.go (see validate_cuda.go);
//   2. CGO_LDFLAGS="-L<repo>/drug-design-dynamics_b200/csrc -lddd_mc" go build
//   main.go, the tests in metropolis_test.go and the R Shiny driver (argv -> ../output/simulations.csv) run unchanged.
//
// STATUS: written against go1.23 / cgo rules but NOT COMPILED in the build container (no Go toolchain there,
// SURVEY.md F1).  The C side it binds is exercised by tests/ through the identical ctypes binding
// (drug-design-dynamics_b200/capi.py), argument for argument.
//
// What changes versus the reference, and why it is still a drop-in:
//   - Molecule.atoms []Atom is Go's {X,Y,Z float64; Charge float64} = 4 x float64 per atom, exactly the `xyzq`
//     layout of the C ABI, so &mol.atoms[0] is passed without repacking (cgo pins it for the duration of the call).
//   - SimulateMultipleLigandsParallel makes ONE batched call for all ligands x replica chains instead of
//     numProcs x numProcs goroutines (metropolis.go:27-51, :94-111); results come back in input order.
//   - every replica chain starts from a private copy of the ligand, which removes the reference's data race on the
//     shared backing array (SURVEY.md F5); the replica pick (:102-109) runs in replica-index order.
//   - the scalar helpers (AcceptMove, RotateAtom, Distance, JitterLigand, ...) stay pure Go, as in the reference.
package main

/*
#cgo CFLAGS: -I${SRCDIR}
#cgo LDFLAGS: -lddd_mc
#include <stdlib.h>
#include "ddd_mc.h"
*/
import "C"

import (
	"math"
	"math/rand"
	"sync"
	"sync/atomic"
	"unsafe"
)

var (
	dddOnce sync.Once
	dddCall uint64 // successive top-level calls use fresh Philox seeds, like Go's advancing global source
	dddSeed = rand.Uint64()
)

func dddInit() {
	dddOnce.Do(func() {
		if rc := C.ddd_init(nil, 0); rc != 0 { // every visible GPU; chains shard over them inside the library
			panic("ddd_mc: " + C.GoString(C.ddd_last_error())) // the reference's Check() panics too (main.go:179-183)
		}
	})
}

func dddCheck(rc C.int) {
	if rc != 0 {
		panic("ddd_mc: " + C.GoString(C.ddd_last_error()))
	}
}

func dddNextSeed() uint64 {
	n := atomic.AddUint64(&dddCall, 1) - 1
	return dddSeed + 0x9E3779B97F4A7C15*n
}

func dddParams() C.ddd_params {
	var p C.ddd_params
	C.ddd_params_default(&p) // K, THRESHOLD, MINDISTANCE, MAXANGLE, 0.1 jitter, 1e-6 clamp (datatypes.go:6-12)
	return p
}

// atomsPtr returns a pointer to the first atom (4 float64 each) or nil for an empty molecule.
func atomsPtr(m Molecule) *C.double {
	if len(m.atoms) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&m.atoms[0]))
}

type dddReceptor struct{ h *C.ddd_receptor }

func newReceptor(protein Molecule) dddReceptor {
	dddInit()
	var r dddReceptor
	dddCheck(C.ddd_receptor_create(atomsPtr(protein), C.int64_t(len(protein.atoms)), &r.h))
	return r
}

func (r dddReceptor) close() { C.ddd_receptor_destroy(r.h) }

// packLigands flattens []Molecule into the CSR batch of the C ABI (one contiguous []Atom + offsets).
func packLigands(ligands []Molecule) ([]C.int64_t, []Atom) {
	offsets := make([]C.int64_t, len(ligands)+1)
	total := 0
	for i, l := range ligands {
		total += len(l.atoms)
		offsets[i+1] = C.int64_t(total)
	}
	flat := make([]Atom, 0, total)
	for _, l := range ligands {
		flat = append(flat, l.atoms...)
	}
	return offsets, flat
}

func flatPtr(a []Atom) *C.double {
	if len(a) == 0 {
		return nil
	}
	return (*C.double)(unsafe.Pointer(&a[0]))
}

// SimulateMultipleLigands: metropolis.go:11-22 (serial variant: shift every ligand first, in place).
func SimulateMultipleLigands(protein Molecule, ligands []Molecule, iterations int, rotate bool, temperature float64, numProcs int) ([]Molecule, []float64) {
	for i, ligand := range ligands {
		ligands[i] = ShiftLigandCloserByThreshold(ligand, protein, THRESHOLD)
	}
	return SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs)
}

// SimulateMultipleLigandsParallel: metropolis.go:27-51 + :56-67 + :94-111 in one ddd_minimize_batch call.
func SimulateMultipleLigandsParallel(protein Molecule, ligands []Molecule, iterations int, rotate bool, temperature float64, numProcs int) ([]Molecule, []float64) {
	_ = len(ligands) / numProcs // the reference panics here with "integer divide by zero" when numProcs == 0 (:34)
	rec := newReceptor(protein)
	defer rec.close()
	offsets, flat := packLigands(ligands)
	out := make([]Atom, len(flat))
	energy := make([]float64, len(ligands))
	p := dddParams()
	rot := C.int32_t(0)
	if rotate {
		rot = 1
	}
	if len(ligands) > 0 {
		var ePtr *C.double = (*C.double)(unsafe.Pointer(&energy[0]))
		dddCheck(C.ddd_minimize_batch(rec.h, &offsets[0], flatPtr(flat), C.int64_t(len(ligands)), C.int64_t(iterations), rot,
			C.double(temperature), C.int32_t(numProcs), &p, C.uint64_t(dddNextSeed()), 0, flatPtr(out), ePtr, nil, nil))
	}
	minLigands := make([]Molecule, len(ligands))
	for i := range ligands {
		minLigands[i] = Molecule{atoms: out[offsets[i]:offsets[i+1]:offsets[i+1]]}
	}
	return minLigands, energy
}

// SimulateLigandMinimizationOneProc: metropolis.go:56-67.
func SimulateLigandMinimizationOneProc(protein Molecule, ligands []Molecule, iterations int, rotate bool, temperature float64, numProcs int, ligandChannel chan MultipleLigandSimulationOutput) {
	l, e := SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs)
	ligandChannel <- MultipleLigandSimulationOutput{Ligand: l, Energy: e}
}

// SimulateEnergyMinimizationParallel: metropolis.go:94-111.
func SimulateEnergyMinimizationParallel(protein, ligand Molecule, iterations int, rotate bool, temperature float64, numProcs int) Molecule {
	l, _ := SimulateMultipleLigandsParallel(protein, []Molecule{ligand}, iterations, rotate, temperature, numProcs)
	return l[0]
}

func runOneChain(protein, ligand Molecule, iterations int, rotate bool, temperature float64) Molecule {
	rec := newReceptor(protein)
	defer rec.close()
	offsets := []C.int64_t{0, C.int64_t(len(ligand.atoms))}
	out := make([]Atom, len(ligand.atoms))
	var st C.ddd_chain_stats
	p := dddParams()
	rot := C.int32_t(0)
	if rotate {
		rot = 1
	}
	dddCheck(C.ddd_run_chains(rec.h, &offsets[0], atomsPtr(ligand), 1, 1, C.int64_t(iterations), rot, C.double(temperature),
		&p, C.uint64_t(dddNextSeed()), 0, nil, nil, flatPtr(out), &st, nil))
	return Molecule{atoms: out}
}

// SimulateEnergyMinimization: metropolis.go:72-89 (dead code in the reference; kept for API parity as one chain).
func SimulateEnergyMinimization(protein, ligand Molecule, iterations int, rotate bool, temperature float64) Molecule {
	return runOneChain(protein, ligand, iterations, rotate, temperature)
}

// SimulateEnergyMinimizationOneProc: metropolis.go:116-136.
func SimulateEnergyMinimizationOneProc(protein, ligand Molecule, iterations int, rotate bool, temperature float64, c chan Molecule) {
	c <- runOneChain(protein, ligand, iterations, rotate, temperature)
}

// CalculateEnergy: metropolis.go:247-262, evaluated on device in float64 with the reference's per-term arithmetic.
func CalculateEnergy(protein, ligand Molecule) float64 {
	rec := newReceptor(protein)
	defer rec.close()
	offsets := []C.int64_t{0, C.int64_t(len(ligand.atoms))}
	p := dddParams()
	p.precision = C.DDD_PRECISION_F64
	var e C.double
	dddCheck(C.ddd_energy_batch(rec.h, &offsets[0], atomsPtr(ligand), 1, &p, &e, nil))
	return float64(e)
}

// FindClosestAtomDistance: metropolis.go:290-304.
func FindClosestAtomDistance(ligand, protein Molecule) (float64, Position3d, Position3d) {
	rec := newReceptor(protein)
	defer rec.close()
	offsets := []C.int64_t{0, C.int64_t(len(ligand.atoms))}
	var d C.double
	var li, pj C.int64_t
	dddCheck(C.ddd_closest_atom_batch(rec.h, &offsets[0], atomsPtr(ligand), 1, &d, &li, &pj))
	var lp, pp Position3d
	if li >= 0 && pj >= 0 {
		lp, pp = ligand.atoms[li].Position, protein.atoms[pj].Position
	}
	return float64(d), lp, pp
}

// ShiftLigandCloserByThreshold: metropolis.go:267-285; mutates the caller's atoms in place like the reference.
func ShiftLigandCloserByThreshold(ligand, protein Molecule, threshold float64) Molecule {
	rec := newReceptor(protein)
	defer rec.close()
	offsets := []C.int64_t{0, C.int64_t(len(ligand.atoms))}
	dddCheck(C.ddd_shift_closer_batch(rec.h, &offsets[0], atomsPtr(ligand), 1, C.double(threshold)))
	return ligand
}

// ---- scalar helpers: unchanged host code (a kernel launch for these would be absurd, SURVEY.md 8b) ----

// AcceptMove: metropolis.go:141-149.
func AcceptMove(currentEnergy, newEnergy, temperature float64) bool {
	if newEnergy < currentEnergy {
		return true
	}
	deltaE := newEnergy - currentEnergy
	probability := math.Exp(-deltaE / temperature)
	return rand.Float64() < probability
}

// JitterLigand: metropolis.go:154-167.
func JitterLigand(ligand Molecule, minDistance float64) Molecule {
	for {
		newLigand := ligand
		for i := range newLigand.atoms {
			newLigand.atoms[i].Position.X += (rand.Float64() - 0.5) * 0.1
			newLigand.atoms[i].Position.Y += (rand.Float64() - 0.5) * 0.1
			newLigand.atoms[i].Position.Z += (rand.Float64() - 0.5) * 0.1
		}
		if IsCollisionFree(newLigand, minDistance) {
			return newLigand
		}
	}
}

// JitterAndRotateLigand: metropolis.go:172-184.
func JitterAndRotateLigand(ligand Molecule, minDistance float64, maxAngle float64) Molecule {
	for {
		newLigand := RotateLigand(ligand, maxAngle)
		for i := range newLigand.atoms {
			newLigand.atoms[i].Position.X += (rand.Float64() - 0.5) * 0.1
			newLigand.atoms[i].Position.Y += (rand.Float64() - 0.5) * 0.1
			newLigand.atoms[i].Position.Z += (rand.Float64() - 0.5) * 0.1
		}
		if IsCollisionFree(newLigand, minDistance) {
			return newLigand
		}
	}
}

// RotateLigand: metropolis.go:189-208.
func RotateLigand(ligand Molecule, maxAngle float64) Molecule {
	newLigand := ligand
	axis := Position3d{X: rand.Float64()*2 - 1, Y: rand.Float64()*2 - 1, Z: rand.Float64()*2 - 1}
	axis.Normalize()
	theta := (rand.Float64()*2 - 1) * maxAngle
	for i := range newLigand.atoms {
		newLigand.atoms[i].Position = RotateAtom(newLigand.atoms[i].Position, axis, theta)
	}
	return newLigand
}

// RotateAtom: metropolis.go:213-224.
func RotateAtom(pos, axis Position3d, theta float64) Position3d {
	cosTheta := math.Cos(theta)
	sinTheta := math.Sin(theta)
	dot := pos.Dot(axis)
	x := pos.X*cosTheta + sinTheta*(axis.Y*pos.Z-axis.Z*pos.Y) + axis.X*dot*(1-cosTheta)
	y := pos.Y*cosTheta + sinTheta*(axis.Z*pos.X-axis.X*pos.Z) + axis.Y*dot*(1-cosTheta)
	z := pos.Z*cosTheta + sinTheta*(axis.X*pos.Y-axis.Y*pos.X) + axis.Z*dot*(1-cosTheta)
	return Position3d{X: x, Y: y, Z: z}
}

// IsCollisionFree: metropolis.go:229-242.
func IsCollisionFree(ligand Molecule, minDistance float64) bool {
	for i := 0; i < len(ligand.atoms); i++ {
		for j := i + 1; j < len(ligand.atoms); j++ {
			if Distance(ligand.atoms[i].Position, ligand.atoms[j].Position) < minDistance {
				return false
			}
		}
	}
	return true
}

// Distance: metropolis.go:309-314.
func Distance(a, b Position3d) float64 {
	dx := a.X - b.X
	dy := a.Y - b.Y
	dz := a.Z - b.Z
	return math.Sqrt(dx*dx + dy*dy + dz*dz)
}

// CopyLigand: metropolis.go:319-334.
func CopyLigand(ligand Molecule) Molecule {
	newAtoms := make([]Atom, len(ligand.atoms))
	copy(newAtoms, ligand.atoms)
	return Molecule{atoms: newAtoms}
}
