"""pabi_b200 -- B200-native batched legal move generation and perft, a drop-in for
that path of kirillbobyrev/pabi."""
from .chess import Engine, Move, Position, default_engine, perft, perft_multi, pinned_empty  # noqa: F401
from .game import Games  # noqa: F401
from ._native import PabiCudaError, POSITION_DTYPE, Timing  # noqa: F401

__all__ = ["Engine", "Move", "Position", "Games", "perft", "perft_multi", "pinned_empty", "default_engine", "PabiCudaError", "POSITION_DTYPE", "Timing"]
