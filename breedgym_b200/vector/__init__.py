"""Vector environments (drop-in for `breedgym.vector`, reference breedgym/vector/__init__.py:1-13)."""
from .vec_env import VecBreedGym
from .vec_wrappers import PairScores, RavelIndex, SelectionScores

from .breeding_programs_env import WheatBreedGym  # isort: skip
from .distributed import DistributedBreedGym  # isort: skip
from .sharded import ShardedVecBreedGym  # isort: skip

__all__ = [
    "VecBreedGym",
    "SelectionScores",
    "PairScores",
    "RavelIndex",
    "WheatBreedGym",
    "DistributedBreedGym",
    "ShardedVecBreedGym",
]
