#!/usr/bin/env python
"""strip_heavy.py -- produce metropolis_helpers.go from the reference's own metropolis.go.

    python go/strip_heavy.py <reference>/metropolisMethod/metropolis.go > metropolis_helpers.go

The cgo overlay (metropolis_cuda.go) re-defines the heavy functions of metropolis.go.  Everything else in that file -- the
scalar helpers a kernel launch would be absurd for -- must stay the reference's own code.  Instead of shipping
a copy of those helpers, this script deletes ONLY the re-defined functions (with their doc comments) from the reference file and
prints the rest unchanged, so the overlay contains no line of the reference.  Unused imports are dropped so that `go build`
does not reject the result.
"""
import re
import sys

# functions re-defined by go/metropolis_cuda.go (file:line in the reference's metropolis.go)
HEAVY = [
    "SimulateMultipleLigands",             # :11
    "SimulateMultipleLigandsParallel",     # :27
    "SimulateLigandMinimizationOneProc",   # :56
    "SimulateEnergyMinimization",          # :72
    "SimulateEnergyMinimizationParallel",  # :94
    "SimulateEnergyMinimizationOneProc",   # :116
    "CalculateEnergy",                     # :247
    "ShiftLigandCloserByThreshold",        # :267
    "FindClosestAtomDistance",             # :290
]
# what must survive (checked by tests/test_go_overlay.py)
KEPT = ["AcceptMove", "JitterLigand", "JitterAndRotateLigand", "RotateLigand", "RotateAtom", "IsCollisionFree", "Distance", "CopyLigand"]


def strip(src: str) -> str:
    lines = src.split("\n")
    out, i = [], 0
    while i < len(lines):
        m = re.match(r"func\s+(\w+)\s*\(", lines[i])
        if m and m.group(1) in HEAVY:
            while out and out[-1].lstrip().startswith("//"):      # the function's doc comment goes with it
                out.pop()
            depth, started = 0, False
            while i < len(lines):                                  # skip to the closing brace of the function body
                depth += lines[i].count("{") - lines[i].count("}")
                started = started or "{" in lines[i]
                i += 1
                if started and depth == 0:
                    break
            while i < len(lines) and lines[i].strip() == "":       # and the blank line(s) after it
                i += 1
            continue
        out.append(lines[i])
        i += 1
    text = "\n".join(out)
    # drop imports the surviving helpers no longer use
    body = re.sub(r"import\s*\((.*?)\)", "", text, flags=re.S)
    def keep_import(mm):
        pkgs = [l.strip() for l in mm.group(1).split("\n") if l.strip()]
        used = [p for p in pkgs if re.search(r"\b" + re.escape(p.strip('"').split("/")[-1]) + r"\.", body)]
        return "import (\n" + "".join(f"\t{p}\n" for p in used) + ")"
    return re.sub(r"import\s*\((.*?)\)", keep_import, text, count=1, flags=re.S)


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    sys.stdout.write(strip(open(sys.argv[1]).read()))
