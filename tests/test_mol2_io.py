"""mol2 ingest / egress at screen scale: ParseMol2 (parsing.go:50-106), the batch parser that fills the C ABI's CSR
staging buffers, UpdateMol2Coordinates (plotting.go:126-189) and SaveMinimumEnergyLigand (plotting.go:109-121).  Host logic only:
the Python mirror, the compiled C++ mirror and the C oracle's parser must agree on synthetic files that exercise the parser's
corners (short lines, a junk number, CRLF, a section after ATOM, charge = LAST field)."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _write_mol2(path, atoms, crlf=False, junk_line=True):
    nl = "\r\n" if crlf else "\n"
    lines = ["@<TRIPOS>MOLECULE", "synthetic", f" {len(atoms)} 0 0 0 0", "SMALL", "GASTEIGER", "", "@<TRIPOS>ATOM"]
    for i, (x, y, z, q) in enumerate(atoms):
        lines.append(f"{i + 1:7d} C{i + 1:<6d} {x:10.4f} {y:10.4f} {z:10.4f} C.3     1  LIG1      {q:8.4f}")
        if junk_line and i == 1:
            lines.append("      too short line")                          # < 9 fields: skipped
    lines += ["@<TRIPOS>BOND", "     1     1     2    1", "     2     2     3   ar"]
    path.write_text(nl.join(lines) + nl)


@pytest.fixture()
def mol2_files(tmp_path):
    rng = np.random.default_rng(4)
    files, mols = [], []
    for k in range(23):
        n = int(rng.integers(1, 60))
        a = np.round(np.concatenate([rng.normal(0, 20, size=(n, 3)), rng.normal(0, 0.2, size=(n, 1))], axis=1), 4)
        f = tmp_path / f"{k:03d}x_ligand.mol2"
        _write_mol2(f, a, crlf=(k % 5 == 0))
        files.append(str(f)); mols.append(a)
    return files, mols


def test_parsers_agree(ddd, orc, mol2_files):
    files, mols = mol2_files
    met = ddd.metropolis
    off, xyzq = met.ParseMol2Batch(files)
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(m) for m in mols])]).tolist()
    for i, f in enumerate(files):
        one = met.ParseMol2(f).xyzq
        assert np.array_equal(one, mols[i]) and np.array_equal(xyzq[off[i]:off[i + 1]], one)
        assert np.array_equal(orc.parse_mol2(f), one)                      # the C restatement of parsing.go:50-106
    assert met.ParseMol2Batch([])[0].tolist() == [0]
    with pytest.raises(OSError):
        met.ParseMol2Batch(files + ["/nonexistent/x_ligand.mol2"])


def test_parser_corner_cases(ddd, orc, tmp_path):
    f = tmp_path / "odd_ligand.mol2"
    f.write_text("\n".join([
        "@<TRIPOS>ATOM",
        "1 N 1.5 abc 3.25 N.3 1 LIG 0.25",                                  # junk y -> 0 (strconv error discarded, parsing.go:85-88)
        "2 C 1.0 2.0 3.0 C.3 1 LIG 0.5 extra -0.75",                        # charge = LAST field
        "3 C 1.0 2.0",                                                      # short: skipped
        "@<TRIPOS>BOND",
        "4 C 9.0 9.0 9.0 C.3 1 LIG 0.1",                                    # after the section: ignored
    ]) + "\n")
    want = np.array([[1.5, 0.0, 3.25, 0.25], [1.0, 2.0, 3.0, -0.75]])
    assert np.array_equal(ddd.metropolis.ParseMol2(str(f)).xyzq, want)
    assert np.array_equal(orc.parse_mol2(f), want)


def test_update_mol2_coordinates_roundtrip(ddd, mol2_files, tmp_path):
    files, mols = mol2_files
    met = ddd.metropolis
    rng = np.random.default_rng(1)
    for i in (0, 5, 7):
        new = mols[i].copy(); new[:, :3] += rng.normal(0, 3, size=(len(new), 3))
        out = tmp_path / f"upd{i}.mol2"
        met.UpdateMol2Coordinates(files[i], str(out), met.Molecule(xyzq=new))
        back = met.ParseMol2(str(out)).xyzq
        assert np.allclose(back[:, :3], new[:, :3], rtol=0, atol=0.5e-4 + 1e-12) and np.array_equal(back[:, 3], new[:, 3])
        text = out.read_text()
        assert "\r" not in text and text.count("@<TRIPOS>") == 3           # \n endings (bufio.ScanLines drops \r), sections kept
        assert ("too short line" in text) == (len(new) >= 2)                # lines that are not atom records are copied
        atom_lines = [l for l in text.split("\n") if l.startswith("1 C1 ")]
        assert atom_lines and "  " not in atom_lines[0]                     # strings.Join(parts, " ")
    with pytest.raises(ValueError):                                         # plotting.go:183-185
        met.UpdateMol2Coordinates(files[0], str(tmp_path / "bad.mol2"), met.Molecule(xyzq=np.zeros((len(mols[0]) + 1, 4))))


def test_save_minimum_energy_ligand_rule(ddd, orc, mol2_files, tmp_path, capsys):
    files, mols = mol2_files
    met = ddd.metropolis
    ligs = [met.Molecule(xyzq=m) for m in mols[:4]]
    out = tmp_path / "min.mol2"
    assert met.SaveMinimumEnergyLigand([5.0, 3.0, 7.0, 1.0], files[:4], str(out), ligs) == 0      # nothing below 0.0: index 0 (plotting.go:111)
    assert np.allclose(met.ParseMol2(str(out)).xyzq, mols[0], atol=1e-4)
    assert met.SaveMinimumEnergyLigand([5.0, -3.0, -7.0, -7.0], files[:4], str(out), ligs) == 2    # strict <: the first minimum
    assert np.allclose(met.ParseMol2(str(out)).xyzq, mols[2], atol=1e-4)
    assert "Wrote min ligand file to" in capsys.readouterr().out
    for e in ([5.0, 3.0], [1.0, -2.0, -2.0], [], [float("nan"), -1.0]):
        assert orc.min_energy_index(e, 0)[0] == (met.SaveMinimumEnergyLigand(e, files[:4], str(out), ligs) if e else 0)


def test_cpp_mirror_parses_and_rewrites_like_the_python_mirror(ddd, mol2_files, tmp_path):
    files, mols = mol2_files
    ddd.build.build_host()
    exe = ROOT / "drug-design-dynamics_b200" / "host" / "ddd_metropolis"
    res = subprocess.run([str(exe), "parse"] + files, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    rows = [l.split() for l in res.stdout.strip().split("\n")]
    assert len(rows) == len(files)
    for r, m in zip(rows, mols):
        assert int(r[0]) == len(m)
        sums = [0.0, 0.0, 0.0, 0.0]
        for a in m:                                                         # same left-to-right accumulation as the binary
            for k in range(4):
                sums[k] += a[k]
        assert [float(x) for x in r[1:]] == sums
    out_c, out_p = tmp_path / "c.mol2", tmp_path / "p.mol2"
    res = subprocess.run([str(exe), "update", files[3], str(out_c), "1.25"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    new = mols[3].copy(); new[:, 0] += 1.25
    ddd.metropolis.UpdateMol2Coordinates(files[3], str(out_p), ddd.metropolis.Molecule(xyzq=new))
    assert out_c.read_bytes() == out_p.read_bytes()
    res = subprocess.run([str(exe), "parse", "/nonexistent/q_ligand.mol2"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "panic" in res.stderr                    # Check() -> panic (main.go:179-183)
