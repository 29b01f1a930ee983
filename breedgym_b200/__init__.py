"""breedgym_b200 -- B200-native breeding-simulation engine behind BreedGym's Gymnasium surface.

Registers the six ids of the reference (breedgym/__init__.py:4-34).  Importing
the package needs the C-ABI shared library (built in-tree by
`python -m breedgym_b200.build`); there is no CPU fallback.
"""
from .gym_compat import register

__version__ = "0.1.0"

for _id, _entry in (
    ("BreedGym", "breedgym_b200.breedgym:BreedGym"),
    ("SimplifiedBreedGym", "breedgym_b200.wrappers:SimplifiedBreedGym"),
    ("KBestBreedGym", "breedgym_b200.wrappers:KBestBreedGym"),
    ("VecBreedGym", "breedgym_b200.vector.vec_env:VecBreedGym"),
    ("SelectionScores", "breedgym_b200.vector.vec_wrappers:SelectionScores"),
    ("PairScores", "breedgym_b200.vector.vec_wrappers:PairScores"),
):
    register(id=_id, entry_point=_entry)
