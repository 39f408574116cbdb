"""BASELINE config 1 end to end: the shipped sample structures (rebuilt as mol2 files from tests/golden/sample_mol2.npz, the
arrays the reference's parser yields) through the compiled drop-in binary, beside the CPU restatement of the same calls.

    python tools/config1_e2e.py > profiles/r02_config1_e2e.json

  run    = RunMultipleLigands (main.go:125-158): 223l protein x first 5 ligands, 3000 iterations, rotate, numProcs = cores
  shiny  = RShinyAppMain (main.go:21-114): <protein> <ligand>... -> ../output/simulations.csv, one ligand at a time
Config 1 has 5 x numProcs chains: it cannot fill a B200 -- these numbers are latencies, not throughput."""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ddd_loader
from _oracle import Oracle


def write_mol2(path, xyzq):
    lines = ["@<TRIPOS>MOLECULE", path.stem, f" {len(xyzq)} 0 0 0 0", "SMALL", "GASTEIGER", "", "@<TRIPOS>ATOM"]
    for i, (x, y, z, q) in enumerate(xyzq):
        lines.append(f"{i + 1:7d} A{i + 1:<6d} {x:10.4f} {y:10.4f} {z:10.4f} C.3     1  UNL1      {q:8.4f}")
    lines += ["@<TRIPOS>BOND"]
    path.write_text("\n".join(lines) + "\n")


def main():
    ddd = ddd_loader.load()
    ddd.build.build_host()
    exe = ROOT / "drug-design-dynamics_b200" / "host" / "ddd_metropolis"
    z = np.load(ROOT / "tests" / "golden" / "sample_mol2.npz")
    cores = os.cpu_count() or 1
    out = {"host_cores": cores, "numProcs": cores, "iterations": 3000}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        data = tmp / "work" / "Data" / "mol2_files"
        data.mkdir(parents=True)
        (tmp / "output").mkdir()
        for k in z.files:
            if not k.endswith("_types"):
                write_mol2(data / f"{k}.mol2", z[k])
        ligs = sorted(str(p) for p in data.glob("*_ligand.mol2"))[:5]
        prot = str(data / "223l_protein.mol2")
        for name, cmd in (("run", [str(exe), "run", str(data)]), ("shiny", [str(exe), prot] + ligs)):
            times = []
            for rep in range(3):                         # first call pays CUDA context creation; report all three
                t0 = time.perf_counter()
                res = subprocess.run(cmd, cwd=tmp / "work", capture_output=True, text=True, timeout=600)
                times.append(time.perf_counter() - t0)
                assert res.returncode == 0, res.stderr
            sim = [l for l in res.stdout.split("\n") if "Time taken" in l]
            out[name] = {"process_wall_s": times, "stdout_simulation_line": sim[0] if sim else None}
        csv = (tmp / "output" / "simulations.csv").read_text().split()
        out["shiny"]["csv"] = csv
        assert csv[0] == "BindingEnergy" and len(csv) == 6
        # in-process (library already initialised): the same two entry points through the ctypes mirror
        met = ddd.metropolis
        met.Seed(1)
        protein = met.ParseMol2(prot); ligands = [met.ParseMol2(f) for f in ligs]
        met.SimulateMultipleLigandsParallel(protein, ligands, 3000, True, met.TEMPERATURE, cores)       # warm
        t0 = time.perf_counter(); met.SimulateMultipleLigandsParallel(protein, ligands, 3000, True, met.TEMPERATURE, cores)
        out["run"]["in_process_call_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        for l in ligands:
            nl = met.SimulateEnergyMinimizationParallel(protein, l, 3000, True, met.TEMPERATURE, cores); met.CalculateEnergy(protein, nl)
        out["shiny"]["in_process_loop_s"] = time.perf_counter() - t0
        # the CPU restatement of the same call (C float64, NOT Go), all cores, both square variants
        off, xyzq = ddd.capi.pack_ligands([l.xyzq for l in ligands])
        for variant in ("gopow", ""):
            orc = Oracle(variant)
            t0 = time.perf_counter()
            rc, _, en, st = orc.simulate_multiple_ligands(protein.xyzq, off, xyzq, 3000, True, met.TEMPERATURE, cores, 1, n_threads=cores)
            out.setdefault("cpu_restatement", {})[variant or "xx"] = {"seconds": time.perf_counter() - t0, "steps": int(st["steps"].sum())}
    out["note"] = ("process_wall_s includes process start, CUDA context creation (~0.3-1 s) and mol2 parsing; in_process = the library call alone. "
                   "The CPU rows run one thread per LIGAND (5 busy threads at most), replicas sequential inside, as orc_simulate_multiple_ligands does.")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
