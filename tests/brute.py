"""ctypes loader for tests/native/libbrute.so: an independent mailbox legal-move generator (pseudo-legal moves by
stepping, legality by playing the move and walking out from the king).  It plays the part shakmaty plays for the
reference (tests/chess.rs:555-578): a second opinion that shares no code, tables or data structures with the oracle
or the product.  Test infrastructure only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

DIR = Path(__file__).resolve().parent / "native"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", str(DIR), "libbrute.so"], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(str(DIR / "libbrute.so"))
        L.brute_legal_moves.restype = C.c_uint32
        L.brute_legal_moves.argtypes = [C.c_void_p, C.POINTER(C.c_uint16)]
        L.brute_in_check.argtypes = [C.c_void_p]
        L.brute_compare_batch.restype = C.c_uint64
        L.brute_compare_batch.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def legal_moves(position) -> list:
    """Sorted u16 moves of one position (a ctypes struct or a 112-byte numpy record)."""
    raw = bytes(position) if not isinstance(position, np.void) else position.tobytes()
    buf = C.create_string_buffer(raw, 112)
    out = (C.c_uint16 * 256)()
    n = lib().brute_legal_moves(buf, out)
    return list(out[:n])


def in_check(position) -> bool:
    raw = bytes(position) if not isinstance(position, np.void) else position.tobytes()
    return bool(lib().brute_in_check(C.create_string_buffer(raw, 112)))


def compare_batch(positions: np.ndarray, offsets: np.ndarray, moves: np.ndarray, checks=None):
    """(positions that differ, index of the first, total brute-force moves): the candidate CSR lists are sorted per
    position and compared as sets with the brute-force generator, like tests/chess.rs:562-575."""
    positions = np.ascontiguousarray(positions)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    moves = np.ascontiguousarray(moves, dtype=np.uint16)
    flags = None if checks is None else np.ascontiguousarray(checks, dtype=np.uint8)
    first, total = C.c_uint64(), C.c_uint64()
    bad = lib().brute_compare_batch(positions.ctypes.data, len(positions), offsets.ctypes.data, moves.ctypes.data,
                                    None if flags is None else flags.ctypes.data, C.byref(first), C.byref(total))
    return bad, first.value, total.value
