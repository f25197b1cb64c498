"""Loader shim: the package directory is named `pdynamo-mirror_b200` (not a valid Python identifier), so
`import pdynamo_mirror_b200` from the repository root resolves here and is replaced by the real package."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdynamo-mirror_b200")
_spec = importlib.util.spec_from_file_location("pdynamo_mirror_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pdynamo_mirror_b200"] = _mod
_spec.loader.exec_module(_mod)
