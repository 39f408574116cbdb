// validate_cuda.go -- replaces CalculateRMSD in metropolisMethod/validateResults.go (:50-65) with the batched
// device kernel, and adds the batched form used by large validation sweeps (TestMethodRMSD shape: one
// (simulated, reference) pair per protein; config 5 runs 100 000 pairs in one call).
// Delete the reference's CalculateRMSD when adding this file; CompareRMSD, RandomizeLigandPose,
// MultipleProteinRMSD and average stay as they are (they call SimulateEnergyMinimizationParallel and
// CalculateRMSD, both now device-backed).  NOT COMPILED in the build container (no Go toolchain).
package main

/*
#include "ddd_mc.h"
*/
import "C"

import "unsafe"

// CalculateRMSD: validateResults.go:50-65.  Same panic as the reference when `reference` is shorter.
func CalculateRMSD(simulated, reference Molecule) float64 {
	_ = reference.atoms[:len(simulated.atoms)] // index out of range panic, as in the reference's loop
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
