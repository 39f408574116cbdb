// metropolis_cuda.go -- drop-in replacement for metropolisMethod/metropolis.go (and for CalculateRMSD /
// RandomizeLigandPose's heavy callers in validateResults.go) that runs the Monte-Carlo hot path on B200 GPUs
// through the C ABI of libddd_mc.so (include/ddd_mc.h).
//
// HOW TO USE (INTEGRATION.md has the full recipe):
//   1. python go/strip_heavy.py <reference>/metropolisMethod/metropolis.go > metropolis_helpers.go
//      -- this keeps the reference's OWN scalar helpers (AcceptMove, JitterLigand, JitterAndRotateLigand, RotateLigand,
//      RotateAtom, IsCollisionFree, Distance, CopyLigand) byte for byte and removes only the functions re-defined below, so
//      this overlay carries no line of the reference;
//   2. put metropolis_helpers.go, this file, validate_cuda.go and ddd_mc.h next to the reference's main.go / datatypes.go /
//      parsing.go / plotting.go / validateResults.go, delete metropolis.go, and delete the bodies of CalculateRMSD and
//      CompareRMSD in validateResults.go (validate_cuda.go re-defines them);
//   3. CGO_LDFLAGS="-L<repo>/drug-design-dynamics_b200/csrc -lddd_mc" go build
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
//   - the scalar helpers (AcceptMove, RotateAtom, Distance, JitterLigand, ...) stay the reference's own Go code
//     (metropolis_helpers.go, produced by go/strip_heavy.py); a kernel launch for them would be absurd (SURVEY.md 8b).
//   - the protein arrives on every call: the library's content-keyed receptor cache (ddd_receptor_acquire) keeps it resident,
//     so CalculateEnergy / FindClosestAtomDistance / ShiftLigandCloserByThreshold and RShinyAppMain's per-ligand loop
//     (main.go:42-48) do not rebuild it.
package main

/*
#cgo CFLAGS: -I${SRCDIR}
#cgo LDFLAGS: -lddd_mc
#include <stdlib.h>
#include "ddd_mc.h"
*/
import "C"

import (
	"fmt"
	"math/rand"
	"os"
	"runtime"
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

// newReceptor borrows the resident copy of `protein` from the library's content-keyed cache (built on first use).
func newReceptor(protein Molecule) dddReceptor {
	dddInit()
	var r dddReceptor
	dddCheck(C.ddd_receptor_acquire(atomsPtr(protein), C.int64_t(len(protein.atoms)), &r.h))
	return r
}

func (r dddReceptor) close() { C.ddd_receptor_release(r.h) }

// dddWarnIfTruncated: a chain that needed more than max_retries collision retries in ONE step stopped there; the reference
// would still be retrying (metropolis.go:155, :173).  Say so instead of returning a truncated chain silently.
// (The status is thread-local in the library: the callers hold runtime.LockOSThread across the batch call and this one.)
func dddWarnIfTruncated(where string) {
	var n C.int64_t
	if bits := C.ddd_last_chain_status(&n); bits&C.DDD_STATUS_RETRY_LIMIT != 0 {
		fmt.Fprintf(os.Stderr, "ddd_mc warning: %d chain(s) of %s stopped at the collision-retry bound; the reference would still be retrying there (metropolis.go:173)\n", int64(n), where)
	}
}

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

// SimulateMultipleLigands: metropolis.go:11-22 (serial variant: shift every ligand first, in place, then minimise).
// One upload: the shift (:14-16), the chains, the replica pick (:102-109) and the gather run on the resident screen.
func SimulateMultipleLigands(protein Molecule, ligands []Molecule, iterations int, rotate bool, temperature float64, numProcs int) ([]Molecule, []float64) {
	width := iterations / numProcs // panics with "integer divide by zero" when numProcs == 0, like :97
	rec := newReceptor(protein)
	defer rec.close()
	offsets, flat := packLigands(ligands)
	out := make([]Atom, len(flat))
	energy := make([]float64, len(ligands))
	if len(ligands) == 0 {
		return []Molecule{}, energy
	}
	p := dddParams()
	seed := C.uint64_t(dddNextSeed())
	var scr *C.ddd_screen
	dddCheck(C.ddd_screen_create(rec.h, &offsets[0], flatPtr(flat), C.int64_t(len(ligands)), C.int32_t(numProcs), 0, &scr))
	defer C.ddd_screen_destroy(scr)
	dddCheck(C.ddd_screen_shift_closer(scr, C.double(THRESHOLD)))
	dddCheck(C.ddd_screen_download_start(scr, flatPtr(flat)))
	for i := range ligands { // the reference shifts the caller's atoms in place (slice alias, SURVEY F4)
		copy(ligands[i].atoms, flat[offsets[i]:offsets[i+1]])
	}
	dddCheck(C.ddd_screen_run(scr, C.int64_t(width), boolToC(rotate), C.double(temperature), &p, seed))
	dddCheck(C.ddd_screen_pick(scr, C.double(temperature), seed, 0, flatPtr(out), (*C.double)(unsafe.Pointer(&energy[0])), nil, nil))
	minLigands := make([]Molecule, len(ligands))
	for i := range ligands {
		minLigands[i] = Molecule{atoms: out[offsets[i]:offsets[i+1]:offsets[i+1]]}
	}
	return minLigands, energy
}

func boolToC(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
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
		runtime.LockOSThread() // the chain status below is thread-local in the library: stay on this OS thread between the two calls
		defer runtime.UnlockOSThread()
		var ePtr *C.double = (*C.double)(unsafe.Pointer(&energy[0]))
		dddCheck(C.ddd_minimize_batch(rec.h, &offsets[0], flatPtr(flat), C.int64_t(len(ligands)), C.int64_t(iterations), rot,
			C.double(temperature), C.int32_t(numProcs), &p, C.uint64_t(dddNextSeed()), 0, flatPtr(out), ePtr, nil, nil))
		dddWarnIfTruncated("SimulateMultipleLigandsParallel")
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
	runtime.LockOSThread() // the chain status below is thread-local in the library
	defer runtime.UnlockOSThread()
	dddCheck(C.ddd_run_chains(rec.h, &offsets[0], atomsPtr(ligand), 1, 1, C.int64_t(iterations), rot, C.double(temperature),
		&p, C.uint64_t(dddNextSeed()), 0, nil, nil, flatPtr(out), &st, nil))
	dddWarnIfTruncated("SimulateEnergyMinimizationOneProc")
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
