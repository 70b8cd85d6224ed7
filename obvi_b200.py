"""Import helper: the package directory is `obvi-slam_b200/` (hyphenated, as the task names it), which the
`import` statement cannot spell.  `import obvi_b200 as ob` gives the package; `ob.synth` the generator."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

_pkg = importlib.import_module("obvi-slam_b200")
synth = importlib.import_module("obvi-slam_b200.synth")
schedule = importlib.import_module("obvi-slam_b200.schedule")
pg_state_io = importlib.import_module("obvi-slam_b200.pg_state_io")
ltm_extraction = importlib.import_module("obvi-slam_b200.ltm_extraction")
vslam_dataset_io = importlib.import_module("obvi-slam_b200.vslam_dataset_io")

globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
