"""Batched mirror of pabi's `Game` environment (src/chess/game.rs:21-111, `Environment` trait of
src/environment.rs:56-74): n games advance in lock-step on the GPU.

The Syzygy adjudication branch of `Game::result` (game.rs:72-96) is not provided -- it belongs to
the third-party shakmaty-syzygy prober -- so results are those of the reference with no tablebase.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._native import POSITION_DTYPE, PabiCudaError, Timing
from .chess import Engine, Move, _as_array

NO_MOVE = 0xFFFF
RESULTS = (None, "win", "draw", "loss")  # GameResult, src/environment.rs:56-61, from the root player's perspective


class Games:
    """`Game::new` for every root (game.rs:30-47).  `history_cap` bounds the actions a game may receive."""

    def __init__(self, engine: Engine, roots, history_cap: int = 1024):
        self.engine = engine
        arr = _as_array(roots)
        self.n = len(arr)
        self._games = C.c_void_p()
        rc = _native.lib().pabi_cuda_games_create(engine._ctx, arr.ctypes.data, self.n, history_cap, C.byref(self._games))
        engine._check(rc)
        self.last_timing: Optional[Timing] = None

    def close(self) -> None:
        if getattr(self, "_games", None):
            _native.lib().pabi_cuda_games_destroy(self._games)
            self._games = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def actions(self) -> Tuple[np.ndarray, np.ndarray]:
        """`Game::actions` (game.rs:50-52) of every game as CSR (offsets[n+1], moves)."""
        offsets = np.empty(self.n + 1, dtype=np.uint32)
        moves = np.empty(max(256, 48 * self.n), dtype=np.uint16)
        L = _native.lib()
        rc = L.pabi_cuda_games_actions(self._games, offsets.ctypes.data, moves.ctypes.data, len(moves))
        if rc == -2:
            moves = np.empty(int(offsets[self.n]), dtype=np.uint16)
            rc = L.pabi_cuda_games_actions(self._games, offsets.ctypes.data, moves.ctypes.data, len(moves))
        self.engine._check(rc)
        return offsets, moves[: int(offsets[self.n])]

    def apply(self, actions: Sequence) -> None:
        """`Game::apply` (game.rs:54-59): one action per game; None / NO_MOVE leaves a game untouched."""
        raw = np.array([NO_MOVE if a is None else (a.raw if isinstance(a, Move) else int(a)) for a in actions], dtype=np.uint16)
        if len(raw) != self.n:
            raise ValueError("one action per game")
        self.engine._check(_native.lib().pabi_cuda_games_apply(self._games, raw.ctypes.data))

    def result(self) -> List[Optional[str]]:
        """`Game::result` (game.rs:61-110): None, "win", "draw" or "loss" per game."""
        out = np.zeros(self.n, dtype=np.uint8)
        self.engine._check(_native.lib().pabi_cuda_games_result(self._games, out.ctypes.data))
        return [RESULTS[r] for r in out]

    def positions(self) -> np.ndarray:
        out = np.empty(self.n, dtype=POSITION_DTYPE)
        self.engine._check(_native.lib().pabi_cuda_games_positions(self._games, out.ctypes.data))
        return out

    def play_random(self, seeds, max_plies: int) -> Tuple[List[Optional[str]], np.ndarray]:
        """Uniformly random self-play on the device until every game has a result or max_plies actions."""
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        if len(sd) != self.n:
            raise ValueError("one seed per game")
        results, plies, t = np.zeros(self.n, dtype=np.uint8), np.zeros(self.n, dtype=np.uint32), Timing()
        self.engine._check(_native.lib().pabi_cuda_games_play_random(self._games, sd.ctypes.data, max_plies, results.ctypes.data,
                                                                     plies.ctypes.data, C.byref(t)))
        self.last_timing = t
        return [RESULTS[r] for r in results], plies
