"""-m gpu: the compiled (C++) host mirror of the reference's Go API, end to end on the B200:
  * host/ddd_metropolis_test = metropolis_test.go restated test for test against the drop-in API;
  * host/ddd_metropolis      = main.go's entry points, incl. the R Shiny contract (argv -> ../output/simulations.csv,
    header `BindingEnergy`, one %.6f per ligand in argv order; main.go:21-114, Rshiny/app.R:180-188)."""
import csv
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def write_mol2(path, xyzq, name="MOL"):
    """mol2 in the shape Open Babel gave the reference's inputs (9 fields per ATOM line, charge last)."""
    with open(path, "w") as f:
        f.write(f"@<TRIPOS>MOLECULE\n{name}\n {len(xyzq)} 0 0 0 0\nSMALL\nGASTEIGER\n\n@<TRIPOS>ATOM\n")
        for i, (x, y, z, q) in enumerate(xyzq):
            f.write(f"{i + 1:7d} C{i + 1:<6d} {x:9.4f} {y:9.4f} {z:9.4f} C.3   1  LIG1   {q:9.4f}\n")
        f.write("@<TRIPOS>BOND\n")


@pytest.fixture(scope="module")
def host(ddd):
    return ddd.build.build_host()


def test_go_unit_tests_restated_in_cpp(host, gpu):
    res = subprocess.run([str(host[1])], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.strip().endswith("PASS") and res.stdout.count("--- PASS:") == 13


def test_rshiny_argv_to_csv_contract(host, gpu, sample, tmp_path):
    work = tmp_path / "metropolis"; work.mkdir(); (tmp_path / "output").mkdir()
    write_mol2(work / "223l_protein.mol2", sample["223l_protein"])
    names = ["223l", "227l", "7loc"]
    for n in names:
        write_mol2(work / f"{n}_ligand.mol2", sample[n + "_ligand"])
    res = subprocess.run([str(host[0]), "223l_protein.mol2"] + [f"{n}_ligand.mol2" for n in names], cwd=work,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("Finish i ligand here!:") == 3 and "Binding Energy data written to simulations.csv" in res.stdout
    rows = list(csv.reader(open(tmp_path / "output" / "simulations.csv")))
    assert rows[0] == ["BindingEnergy"] and len(rows) == 4
    energies = [float(r[0]) for r in rows[1:]]
    assert all(len(r[0].split(".")[1]) == 6 for r in rows[1:])               # %.6f
    # greedy descent from the file pose: the final energy can only be <= CalculateEnergy(start)
    off, lig = gpu.pack_ligands([sample[n + "_ligand"] for n in names])
    with gpu.Receptor(sample["223l_protein"]) as rec:
        e0 = gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64))
    assert all(e <= s + 1e-3 * abs(s) for e, s in zip(energies, e0))          # mol2 files carry 4 decimals
    assert (tmp_path / "output" / "223l_protein.mol2").exists()
    assert any((tmp_path / "output" / f"{n}_ligand.mol2").exists() for n in names)
    # too few arguments: the reference prints a hint and returns (main.go:22-25)
    res = subprocess.run([str(host[0]), "only_protein.mol2"], cwd=work, capture_output=True, text=True)
    assert res.returncode == 0 and "Please provide right inputs." in res.stdout
    # a missing file panics like Check(err) (main.go:179-183)
    res = subprocess.run([str(host[0]), "nope.mol2", "223l_ligand.mol2"], cwd=work, capture_output=True, text=True)
    assert res.returncode != 0 and "panic:" in res.stderr


def test_run_multiple_ligands_and_rmsd_entry_points(host, gpu, sample, tmp_path):
    d = tmp_path / "Data" / "mol2_files"; d.mkdir(parents=True)
    for k, v in sample.items():
        if k != "1a4h_protein":
            write_mol2(d / f"{k}.mol2", v)
    res = subprocess.run([str(host[0]), "run", str(d)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "Starting simulation" in res.stdout and "Time taken for simulation:" in res.stdout
    labels = [ln.split()[0] for ln in res.stdout.splitlines() if len(ln.split()) == 2 and ln.split()[0][0].isdigit()]
    assert labels == ["1a0i", "223l", "227l", "421p", "6tol"]               # filepath.Walk order, [:5] (main.go:127-132)
    res = subprocess.run([str(host[0]), "rmsd", str(d)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "The average RMSD value was:" in res.stdout
