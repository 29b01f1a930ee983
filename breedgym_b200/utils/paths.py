"""Package paths (reference: breedgym/utils/paths.py)."""
import pathlib

__all__ = ["PROJECT_PATH", "DATA_PATH", "CODE_PATH"]

CODE_PATH = pathlib.Path(__file__).resolve().parents[1]
PROJECT_PATH = CODE_PATH.parent
DATA_PATH = CODE_PATH / "data"
FIGURE_PATH = PROJECT_PATH / "figures"
