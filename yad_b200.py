"""Import shim: the package directory is named ``yolo-inspired-audio-activity-detection_b200`` (not a
valid Python identifier), so this module loads it under the importable name ``yad_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "yolo-inspired-audio-activity-detection_b200")
_spec = importlib.util.spec_from_file_location(
    "yad_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["yad_b200"] = _mod
_spec.loader.exec_module(_mod)
