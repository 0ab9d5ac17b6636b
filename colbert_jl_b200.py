"""Import alias: the package directory is named `colbert.jl_b200/` (after the reference,
ColBERT.jl); a dot cannot appear in a Python module name, so `import colbert_jl_b200`
loads that directory as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "colbert.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "colbert_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["colbert_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
