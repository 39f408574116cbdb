"""The Go overlay (go/) cannot be compiled here (no Go toolchain), but its shape can be checked: go/strip_heavy.py must remove from
the reference's metropolis.go exactly the functions go/metropolis_cuda.go re-defines and keep the reference's scalar helpers byte
for byte, the overlay must define every heavy function of SURVEY 8(b) and none of the helpers, and every C symbol it calls must be
declared in include/ddd_mc.h."""
import importlib.util
import re
from pathlib import Path

import pytest

from conftest import ROOT

REF = Path("/root/reference/metropolisMethod/metropolis.go")


def _strip_mod():
    spec = importlib.util.spec_from_file_location("strip_heavy", ROOT / "go" / "strip_heavy.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _go_funcs(text):
    return re.findall(r"^func\s+(\w+)\s*\(", text, flags=re.M)


def test_overlay_defines_the_heavy_functions_and_no_helper():
    m = _strip_mod()
    overlay = (ROOT / "go" / "metropolis_cuda.go").read_text()
    defined = _go_funcs(overlay)
    for f in m.HEAVY:
        assert f in defined, f
    for f in m.KEPT:
        assert f not in defined, f"{f} must stay the reference's own code (metropolis_helpers.go)"
    val = _go_funcs((ROOT / "go" / "validate_cuda.go").read_text())
    assert "CalculateRMSD" in val and "CompareRMSD" in val


def test_overlay_calls_only_declared_c_symbols():
    header = (ROOT / "include" / "ddd_mc.h").read_text()
    declared = set(re.findall(r"\b(ddd_\w+)\s*\(", header))
    for name in ("metropolis_cuda.go", "validate_cuda.go"):
        text = (ROOT / "go" / name).read_text()
        for sym in set(re.findall(r"C\.(ddd_\w+)\(", text)):
            assert sym in declared, f"{name} calls C.{sym}, which include/ddd_mc.h does not declare"


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")
def test_strip_heavy_keeps_the_references_helpers_verbatim():
    m = _strip_mod()
    src = REF.read_text()
    out = m.strip(src)
    assert sorted(_go_funcs(out)) == sorted(m.KEPT)
    assert sorted(set(_go_funcs(src)) - set(_go_funcs(out))) == sorted(m.HEAVY)
    # every kept function body is the reference's, byte for byte
    for f in m.KEPT:
        body = re.search(r"^func\s+" + f + r"\s*\(.*?^}", src, flags=re.M | re.S).group(0)
        assert body in out, f
    assert out.count("{") == out.count("}")
    imports = re.search(r"import\s*\((.*?)\)", out, flags=re.S).group(1)
    assert '"math"' in imports and '"math/rand"' in imports            # the helpers' only dependencies
    body = out.replace(imports, "")
    for pkg in re.findall(r'"([\w/]+)"', imports):
        assert re.search(r"\b" + pkg.split("/")[-1] + r"\.", body), f"unused import {pkg} would fail go build"
