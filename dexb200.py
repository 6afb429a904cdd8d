"""Import shim: loads the package directory ``dynamicexpressions.jl_b200/`` under
the importable name ``dexb200`` (a directory name containing a dot cannot be
imported with a plain ``import`` statement).

    import dexb200
    y, ok = dexb200.eval_tree_array(tree, X, operators)
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "dynamicexpressions.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "dexb200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dexb200"] = _mod
_spec.loader.exec_module(_mod)
