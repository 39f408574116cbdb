// record_stream_test.go -- the recorded-random-stream parity harness of BASELINE.json's north_star
// ("identical move/accept sequences when both sides are driven from the same recorded random stream").
//
// Drop this file into an UNMODIFIED copy of metropolisMethod/ (with metropolis.go present) and run
//     go test -run TestRecordStream -v
// It (1) seeds math/rand, pre-draws N uniforms and writes them to stream.f64 (little-endian float64);
// (2) re-seeds with the same seed and runs the untouched reference SimulateEnergyMinimizationOneProc for one
// chain, so the chain consumes exactly the recorded uniforms in metropolis.go's own order; (3) writes the start
// pose, the final pose and CalculateEnergy of both to chain.json.
// tests/test_gpu_chains.py::test_recorded_stream_replay shows the consumer side: ddd_run_chains(recorded=...)
// replays such a stream on the device (DDD_PRECISION_F64) and must reproduce the final pose to 1e-9 A.
// To compare against a Go run, point tools/replay_go_record.py at the two files.
// NOT COMPILED in the build container (no Go toolchain, SURVEY.md F1).
package main

import (
	"encoding/binary"
	"encoding/json"
	"math"
	"math/rand"
	"os"
	"testing"
)

func TestRecordStream(t *testing.T) {
	const seed, iterations, rotate = 20260101, 400, true
	protein, err := ParseMol2("Data/mol2_files/223l_protein.mol2")
	Check(err)
	ligand, err := ParseMol2("Data/mol2_files/223l_ligand.mol2")
	Check(err)

	// (1) record: worst case (4 + 3 nL) draws per attempt, generous retry allowance, + 1 accept draw per step
	n := iterations * ((4+3*len(ligand.atoms))*8 + 1)
	rand.Seed(seed)
	f, err := os.Create("stream.f64")
	Check(err)
	buf := make([]byte, 8)
	for i := 0; i < n; i++ {
		binary.LittleEndian.PutUint64(buf, math.Float64bits(rand.Float64()))
		_, err = f.Write(buf)
		Check(err)
	}
	f.Close()

	// (2) replay inside the reference itself
	rand.Seed(seed)
	start := CopyLigand(ligand)
	c := make(chan Molecule, 1)
	SimulateEnergyMinimizationOneProc(protein, CopyLigand(ligand), iterations, rotate, TEMPERATURE, c)
	final := <-c

	// (3) dump
	flat := func(m Molecule) [][4]float64 {
		out := make([][4]float64, len(m.atoms))
		for i, a := range m.atoms {
			out[i] = [4]float64{a.Position.X, a.Position.Y, a.Position.Z, a.Charge}
		}
		return out
	}
	js, _ := json.Marshal(map[string]interface{}{
		"seed": seed, "iterations": iterations, "rotate": rotate, "temperature": TEMPERATURE,
		"start": flat(start), "final": flat(final),
		"start_energy": CalculateEnergy(protein, start), "final_energy": CalculateEnergy(protein, final),
	})
	Check(os.WriteFile("chain.json", js, 0644))
}
