"""`import breedgym` drop-in alias: resolves `gym.make("breedgym:BreedGym")` and
`from breedgym.vector import VecBreedGym` to the B200-native implementation."""
import importlib
import sys

import breedgym_b200 as _impl

for _name in ("breedgym", "wrappers", "vector", "vector.vec_env", "vector.vec_wrappers",
              "vector.breeding_programs_env", "utils", "utils.paths", "utils.index_functions"):
    sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"breedgym_b200.{_name}")

from breedgym_b200 import breedgym, utils, vector, wrappers  # noqa: E402,F401

__version__ = _impl.__version__
