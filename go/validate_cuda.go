// validate_cuda.go -- replaces CalculateRMSD in metropolisMethod/validateResults.go (:50-65) with the batched
// device kernel, and adds the batched form used by large validation sweeps (TestMethodRMSD shape: one
// (simulated, reference) pair per protein; config 5 runs 100 000 pairs in one call).
// Delete the reference's CalculateRMSD and CompareRMSD when adding this file; RandomizeLigandPose, MultipleProteinRMSD and
// average stay as they are.  CompareRMSD (:40-45) becomes one resident pipeline: upload the ligand once, RandomizeLigandPose,
// the replica chains, the replica pick and CalculateRMSD against the uploaded pose all run on the device, and only the RMSD
// comes back.  NOT COMPILED in the build container (no Go toolchain).
package main

/*
#include "ddd_mc.h"
*/
import "C"

import "unsafe"

// CalculateRMSD: validateResults.go:50-65.  Same panic as the reference when `reference` is shorter.
func CalculateRMSD(simulated, reference Molecule) float64 {
	if n := len(simulated.atoms); n > 0 {
		_ = reference.atoms[n-1] // index out of range panic when `reference` is shorter, as in the reference's loop (:56); a slice expression would only check the capacity
	}
	out := CalculateRMSDBatch([]Molecule{simulated}, []Molecule{reference})
	return out[0]
}

// CalculateRMSDBatch evaluates CalculateRMSD for every pair in one kernel launch (ddd_rmsd_batch).
func CalculateRMSDBatch(simulated, reference []Molecule) []float64 {
	dddInit()
	n := len(simulated)
	out := make([]float64, n)
	if n == 0 {
		return out
	}
	offsets := make([]C.int64_t, n+1)
	total := 0
	for i := range simulated {
		total += len(simulated[i].atoms)
		offsets[i+1] = C.int64_t(total)
	}
	sim := make([]Atom, 0, total)
	ref := make([]Atom, 0, total)
	for i := range simulated {
		sim = append(sim, simulated[i].atoms...)
		ref = append(ref, reference[i].atoms[:len(simulated[i].atoms)]...)
	}
	dddCheck(C.ddd_rmsd_batch(&offsets[0], flatPtr(sim), flatPtr(ref), 4, C.int64_t(n), (*C.double)(unsafe.Pointer(&out[0]))))
	return out
}

// CompareRMSD: validateResults.go:40-45 on the resident screen.  reference = the pose as uploaded (:41); the randomisation (:42)
// draws from the library's Philox stream instead of math/rand (the reference's stream is unpinned anyway, SURVEY F10).
func CompareRMSD(protein Molecule, ligand Molecule, iterations int, rotate bool, temperature float64, numProcs int) float64 {
	width := iterations / numProcs // integer divide by zero panic when numProcs == 0, as :97
	rec := newReceptor(protein)
	defer rec.close()
	offsets := []C.int64_t{0, C.int64_t(len(ligand.atoms))}
	p := dddParams()
	seed := C.uint64_t(dddNextSeed())
	var scr *C.ddd_screen
	dddCheck(C.ddd_screen_create(rec.h, &offsets[0], atomsPtr(ligand), 1, C.int32_t(numProcs), 0, &scr))
	defer C.ddd_screen_destroy(scr)
	dddCheck(C.ddd_screen_randomize(scr, seed, 0))
	dddCheck(C.ddd_screen_run(scr, C.int64_t(width), boolToC(rotate), C.double(temperature), &p, seed))
	var rmsd C.double
	dddCheck(C.ddd_screen_pick(scr, C.double(temperature), seed, 0, nil, nil, nil, &rmsd))
	return float64(rmsd)
}

// MinimumEnergyIndexResident is SaveMinimumEnergyLigand's choice (plotting.go:109-121) made on the device for a screen whose
// replica pick has run; SaveMinimumEnergyLigand itself stays the reference's code (it only scans a []float64).
func MinimumEnergyIndexResident(scr *C.ddd_screen) (int, float64) {
	var idx C.int64_t
	var e C.double
	dddCheck(C.ddd_screen_argmin(scr, C.DDD_ARGMIN_SAVE_MIN_LIGAND, &idx, &e))
	return int(idx), float64(e)
}
