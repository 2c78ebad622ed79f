"""ctypes loader for the CPU oracle (oracle/liboracle.so).

Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"


class Position(C.Structure):
    """Image of pabi's `Position` (position.rs:45-63); 112 bytes."""

    _fields_ = [
        ("white", C.c_uint64 * 6),
        ("black", C.c_uint64 * 6),
        ("hash", C.c_uint64),
        ("fullmove_counter", C.c_uint16),
        ("castling", C.c_uint8),
        ("side_to_move", C.c_uint8),
        ("halfmove_clock", C.c_uint8),
        ("en_passant_square", C.c_uint8),
        ("_pad", C.c_uint8 * 2),
    ]


POSITION_DTYPE = np.dtype(
    [("white", "<u8", 6), ("black", "<u8", 6), ("hash", "<u8"), ("fullmove_counter", "<u2"),
     ("castling", "u1"), ("side_to_move", "u1"), ("halfmove_clock", "u1"),
     ("en_passant_square", "u1"), ("_pad", "u1", 2)], align=True)
assert POSITION_DTYPE.itemsize == 112 and C.sizeof(Position) == 112


class AttackInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("attacks", "checkers", "pins", "xrays", "safe_king_squares")]


def build(force: bool = False) -> Path:
    so = ORACLE_DIR / "liboracle.so"
    srcs = [ORACLE_DIR / f for f in ("tables.c", "position.c", "movegen.c", "drivers.c", "game.c",
                                     "pabi_oracle.h", "oracle_internal.h", "Makefile")]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(str(build()))
    P = C.POINTER(Position)
    L.oracle_table.restype = C.POINTER(C.c_uint64)
    L.oracle_table.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
    for name in ("oracle_rook_attacks", "oracle_bishop_attacks"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [C.c_int, C.c_uint64]
    L.oracle_ray.restype = C.c_uint64
    L.oracle_ray.argtypes = [C.c_int, C.c_int]
    L.oracle_position_starting.argtypes = [P]
    L.oracle_position_from_fen.argtypes = [C.c_char_p, P, C.c_char_p, C.c_size_t]
    L.oracle_position_parse.argtypes = [C.c_char_p, P, C.c_char_p, C.c_size_t]
    L.oracle_position_to_fen.argtypes = [P, C.c_char_p, C.c_size_t]
    L.oracle_attack_info.argtypes = [P, C.POINTER(AttackInfo)]
    L.oracle_generate_moves.restype = C.c_uint32
    L.oracle_generate_moves.argtypes = [P, C.POINTER(C.c_uint16)]
    L.oracle_make_move.argtypes = [P, C.c_uint16]
    L.oracle_in_check.argtypes = [P]
    L.oracle_perft.restype = C.c_uint64
    L.oracle_perft.argtypes = [P, C.c_uint8]
    L.oracle_perft_mt.restype = C.c_uint64
    L.oracle_perft_mt.argtypes = [P, C.c_uint8, C.c_int]
    L.oracle_compute_hash.restype = C.c_uint64
    L.oracle_compute_hash.argtypes = [P]
    L.oracle_move_from_uci.argtypes = [C.c_char_p, C.POINTER(C.c_uint16)]
    L.oracle_move_to_uci.argtypes = [C.c_uint16, C.c_char_p]
    L.oracle_generate_moves_batch.restype = C.c_uint64
    L.oracle_generate_moves_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
    L.oracle_random_playout.restype = C.c_size_t
    L.oracle_random_playout.argtypes = [P, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    L.oracle_game_new.restype = C.c_void_p
    L.oracle_game_new.argtypes = [P]
    L.oracle_game_free.argtypes = [C.c_void_p]
    L.oracle_game_actions.restype = C.c_uint32
    L.oracle_game_actions.argtypes = [C.c_void_p, C.POINTER(C.c_uint16)]
    L.oracle_game_apply.argtypes = [C.c_void_p, C.c_uint16]
    L.oracle_game_result.argtypes = [C.c_void_p]
    L.oracle_game_position.argtypes = [C.c_void_p, P]
    L.oracle_game_play_random.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint32)]
    _lib = L
    return L


class OracleError(ValueError):
    pass


def table(name: str) -> np.ndarray:
    n = C.c_size_t()
    ptr = lib().oracle_table(name.encode(), C.byref(n))
    if not ptr:
        raise KeyError(name)
    return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()


def starting() -> Position:
    p = Position()
    lib().oracle_position_starting(C.byref(p))
    return p


def from_fen(fen: str) -> Position:
    """Position::from_fen (position.rs:157-275)."""
    p, err = Position(), C.create_string_buffer(512)
    if lib().oracle_position_from_fen(fen.encode("utf-8"), C.byref(p), err, 512) != 0:
        raise OracleError(err.value.decode("utf-8", "replace"))
    return p


def parse(text: str) -> Position:
    """Position::try_from(&str) (position.rs:809-821); "startpos" is a test-suite alias."""
    if text == "startpos":
        return starting()
    p, err = Position(), C.create_string_buffer(512)
    if lib().oracle_position_parse(text.encode("utf-8"), C.byref(p), err, 512) != 0:
        raise OracleError(err.value.decode("utf-8", "replace"))
    return p


def to_fen(p: Position) -> str:
    buf = C.create_string_buffer(128)
    lib().oracle_position_to_fen(C.byref(p), buf, 128)
    return buf.value.decode()


def generate_moves(p: Position) -> list:
    out = (C.c_uint16 * 256)()
    n = lib().oracle_generate_moves(C.byref(p), out)
    return list(out[:n])


def move_to_uci(mv: int) -> str:
    buf = C.create_string_buffer(6)
    lib().oracle_move_to_uci(mv, buf)
    return buf.value.decode()


def move_from_uci(uci: str) -> int:
    mv = C.c_uint16()
    if lib().oracle_move_from_uci(uci.encode(), C.byref(mv)) != 0:
        raise OracleError(f"bad UCI move {uci}")
    return mv.value


def make_move(p: Position, mv: int) -> None:
    lib().oracle_make_move(C.byref(p), mv)


def copy(p: Position) -> Position:
    q = Position()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(Position))
    return q


def attack_info(p: Position) -> AttackInfo:
    ai = AttackInfo()
    lib().oracle_attack_info(C.byref(p), C.byref(ai))
    return ai


def in_check(p: Position) -> bool:
    return bool(lib().oracle_in_check(C.byref(p)))


def perft(p: Position, depth: int, threads: int = 1) -> int:
    if threads > 1:
        return lib().oracle_perft_mt(C.byref(p), depth, threads)
    return lib().oracle_perft(C.byref(p), depth)


def positions_array(ps) -> np.ndarray:
    """List of Position -> contiguous structured array (what the C ABI takes)."""
    arr = np.zeros(len(ps), dtype=POSITION_DTYPE)
    for i, p in enumerate(ps):
        C.memmove(arr.ctypes.data + i * 112, C.byref(p), 112)
    return arr


def generate_moves_batch(arr: np.ndarray):
    """CSR move lists (offsets[n+1] u32, moves u16) in the reference's order."""
    arr = np.ascontiguousarray(arr)
    n = len(arr)
    offsets = np.zeros(n + 1, dtype=np.uint32)
    moves = np.zeros(max(1, 256 * min(n, 4096)), dtype=np.uint16)
    total = lib().oracle_generate_moves_batch(arr.ctypes.data, n, offsets.ctypes.data, moves.ctypes.data, len(moves))
    if total > len(moves):
        moves = np.zeros(total, dtype=np.uint16)
        lib().oracle_generate_moves_batch(arr.ctypes.data, n, offsets.ctypes.data, moves.ctypes.data, len(moves))
    return offsets, moves[:total].copy()


def random_playout(start: Position, seed: int, max_plies: int, stride: int = 1) -> np.ndarray:
    cap = max_plies // stride + 2
    arr = np.zeros(cap, dtype=POSITION_DTYPE)
    n = lib().oracle_random_playout(C.byref(start), seed, max_plies, arr.ctypes.data, cap, stride)
    return arr[:n].copy()


RESULTS = {0: None, 1: "win", 2: "draw", 3: "loss"}


class Game:
    """pabi's `Game` (game.rs:21-111) without the Syzygy branch, on the CPU oracle."""

    def __init__(self, root: Position):
        self._g = lib().oracle_game_new(C.byref(root))

    def __del__(self):
        if getattr(self, "_g", None):
            lib().oracle_game_free(self._g)
            self._g = None

    def actions(self) -> list:
        out = (C.c_uint16 * 256)()
        n = lib().oracle_game_actions(self._g, out)
        return list(out[:n])

    def apply(self, move: int) -> None:
        lib().oracle_game_apply(self._g, move)

    def result(self):
        return RESULTS[lib().oracle_game_result(self._g)]

    def position(self) -> Position:
        p = Position()
        lib().oracle_game_position(self._g, C.byref(p))
        return p

    def play_random(self, seed: int, max_plies: int):
        plies = C.c_uint32()
        r = lib().oracle_game_play_random(self._g, seed, max_plies, C.byref(plies))
        return RESULTS[r], plies.value
