"""Regenerates the committed golden fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

1. sample_mol2.npz  -- the reference's shipped sample structures (metropolisMethod/Data/mol2_files:
   2 proteins + 8 ligands, and Rshiny/MCdata/1a4h_protein) parsed with the oracle's ParseMol2
   restatement (parsing.go:50-106) into float64 xyzq arrays.  Derived data, not reference source.
2. oracle_golden.json -- float64 outputs of the C oracle on those inputs and on seeded synthetic
   inputs: CalculateEnergy (incl. SURVEY App. B), chain end states / accept traces for fixed Philox
   seeds, replica picks, RMSD, closest-atom, shift, randomise.  The GPU tests compare the CUDA
   path against these on the GPU box, where /root/reference and (for the chains) minutes of CPU
   time are not available.
The reference itself (Go) cannot run here, so these goldens pin the ORACLE, not the Go binary;
tests/test_oracle_reference_kats.py pins the oracle against the reference's own test KATs.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ddd_loader  # noqa: E402
from _oracle import Oracle  # noqa: E402

REF = Path("/root/reference")
MOL2 = REF / "metropolisMethod" / "Data" / "mol2_files"


def main():
    orc = Oracle()
    synth = ddd_loader.load().synth
    arrays = {}
    for f in sorted(MOL2.glob("*.mol2")):
        arrays[f.stem] = orc.parse_mol2(f)
    arrays["1a4h_protein"] = orc.parse_mol2(REF / "Rshiny" / "MCdata" / "1a4h_protein.mol2")
    # SYBYL atom types of the same atoms (LJ extension; the reference's parser reads past this column)
    M = ddd_loader.load().metropolis
    types = {f.stem + "_types": np.array(M.ParseMol2Types(f)) for f in sorted(MOL2.glob("*.mol2"))}
    for k, v in types.items():
        assert len(v) == len(arrays[k[:-6]])
    np.savez_compressed(HERE / "sample_mol2.npz", **arrays, **types)

    g = {"note": "float64 outputs of oracle/ddd_oracle.c (C restatement, NOT Go); see make_golden.py"}
    ligs = ["1a0i", "223l", "227l", "421p", "6tol", "721p", "7loc", "821p"]
    g["sample_energy"] = {}
    for prot in ("223l_protein", "227l_protein"):
        for l in ligs:
            e, a = orc.energy(arrays[prot], arrays[l + "_ligand"], with_abs=True)
            g["sample_energy"][f"{prot}|{l}"] = {"energy": e, "abs_sum": a, "n_lig": int(len(arrays[l + "_ligand"]))}

    # config 1 (RunMultipleLigands, main.go:125-158) at numProcs = 8 with pinned seeds
    first5 = ["1a0i", "223l", "227l", "421p", "6tol"]            # filepath.Walk order, [:5]
    off = np.zeros(6, dtype=np.int64); off[1:] = np.cumsum([len(arrays[l + "_ligand"]) for l in first5])
    xyzq = np.concatenate([arrays[l + "_ligand"] for l in first5])
    g["config1"] = {}
    for seed in (1, 2):
        rc, out, en, st = orc.simulate_multiple_ligands(arrays["223l_protein"], off, xyzq, 3000, True, 310.15, 8, seed, n_threads=8)
        assert rc == 0
        g["config1"][str(seed)] = {"energy": en.tolist(), "accepted": st["accepted"].tolist(), "retries": st["retries"].tolist(),
                                   "draws": st["draws"].tolist(), "final_energy": st["final_energy"].tolist(),
                                   "pose_checksum": [float(out[off[i]:off[i + 1], :3].sum()) for i in range(5)]}

    # single chains on the shipped data: accept traces + end states (rotate on/off)
    g["chains"] = []
    for lig, rotate, steps, seed, cid in (("223l", False, 400, 5, 0), ("223l", True, 400, 5, 1), ("7loc", False, 400, 9, 3),
                                          ("6tol", True, 300, 2, 0), ("227l", False, 600, 4, 2)):
        pose, st, tr, et = orc.chain(arrays["223l_protein"], arrays[lig + "_ligand"], steps, rotate, 310.15,
                                     orc.philox_stream(seed, cid), want_trace=True)
        g["chains"].append({"ligand": lig, "rotate": rotate, "steps": steps, "seed": seed, "chain_id": cid,
                            "accept_trace": "".join(map(str, tr.tolist())), "accepted": int(st["accepted"]),
                            "retries": int(st["retries"]), "draws": int(st["draws"]), "final_energy": float(st["final_energy"]),
                            "start_energy": float(st["start_energy"]), "pose": pose[:, :3].tolist()})

    # synthetic small receptor: energies / closest / shift / randomise / rmsd / minimise
    prot = synth.make_receptor(600, seed=11)
    offs, lig = synth.make_ligands([12, 7, 20, 1, 33, 2, 16, 40], prot)
    e, a = orc.energy_batch(prot, offs, lig, with_abs=True)
    s = {"energy": e.tolist(), "abs_sum": a.tolist()}
    cl = [orc.find_closest(lig[offs[i]:offs[i + 1]], prot) for i in range(8)]
    s["closest"] = [[c[0], c[1], c[2]] for c in cl]
    far = lig.copy(); far[:, :3] += 25.0
    s["shift_checksum"] = [float(orc.shift_closer(far[offs[i]:offs[i + 1]], prot, 5.0)[:, :3].sum()) for i in range(8)]
    rz = [orc.randomize_pose(lig[offs[i]:offs[i + 1]], orc.philox_stream(5, 0x2000000000000000 + i)) for i in range(8)]
    s["randomize_checksum"] = [float(r[:, :3].sum()) for r in rz]
    s["rmsd"] = [orc.rmsd(rz[i], lig[offs[i]:offs[i + 1]]) for i in range(8)]
    rc, out, en, st = orc.simulate_multiple_ligands(prot, offs, lig, 1000, True, 310.15, 4, 21)
    s["minimize_rotate"] = {"energy": en.tolist(), "accepted": st["accepted"].tolist()}
    rc, out, en, st = orc.simulate_multiple_ligands(prot, offs, lig, 1000, False, 310.15, 4, 21)
    s["minimize_jitter"] = {"energy": en.tolist(), "accepted": st["accepted"].tolist(), "pose_checksum": float(out[:, :3].sum())}
    g["synth600"] = s

    # cutoff extension (no reference row; pins the builder's own restatement): shipped data, 10 A shifted-Coulomb cutoff
    pc = orc.params(cutoff=10.0)
    c = {"cutoff": 10.0, "energy": {}, "chains": []}
    for l in ligs:
        e, a = orc.energy(arrays["223l_protein"], arrays[l + "_ligand"], pc, with_abs=True)
        c["energy"][l] = {"energy": e, "abs_sum": a}
    for lig, rotate, steps, seed, cid in (("223l", True, 300, 5, 1), ("421p", False, 300, 7, 0)):
        pose, st, tr, et = orc.chain(arrays["223l_protein"], arrays[lig + "_ligand"], steps, rotate, 310.15,
                                     orc.philox_stream(seed, cid), pc, want_trace=True)
        c["chains"].append({"ligand": lig, "rotate": rotate, "steps": steps, "seed": seed, "chain_id": cid,
                            "accept_trace": "".join(map(str, tr.tolist())), "accepted": int(st["accepted"]),
                            "retries": int(st["retries"]), "draws": int(st["draws"]), "final_energy": float(st["final_energy"]),
                            "pose_checksum": float(pose[:, :3].sum())})
    g["cutoff10"] = c

    (HERE / "oracle_golden.json").write_text(json.dumps(g, indent=1))
    print("wrote", HERE / "sample_mol2.npz", HERE / "oracle_golden.json")


if __name__ == "__main__":
    main()
