"""Import alias: `import msacl_b200` loads the package that lives in the directory named after
the upstream project (hyphens are not valid in Python identifiers)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "multi-step-actor-critic-learning-with-lyapunov-certificates-for-exponentially-stabilizing-control_b200")
_spec = importlib.util.spec_from_file_location("msacl_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["msacl_b200"] = _mod
_spec.loader.exec_module(_mod)
