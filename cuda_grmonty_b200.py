"""Import shim: the package directory is named `cuda-grmonty_b200/` (not a Python identifier), so
`import cuda_grmonty_b200` loads its __init__.py from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-grmonty_b200")
_spec = importlib.util.spec_from_file_location("cuda_grmonty_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cuda_grmonty_b200"] = _mod
_spec.loader.exec_module(_mod)
