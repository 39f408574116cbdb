"""Imports the hyphen-named package directory `drug-design-dynamics_b200/` as module `ddd_b200`."""
import importlib.util
import sys
from pathlib import Path

_NAME = "ddd_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    pkg_dir = Path(__file__).resolve().parent / "drug-design-dynamics_b200"
    spec = importlib.util.spec_from_file_location(_NAME, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
