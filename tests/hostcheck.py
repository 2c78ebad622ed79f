"""ctypes loader for tests/native/libhostcheck.so: the product's device
arithmetic (pabi_b200/csrc/chess.cuh) compiled for the host.  Test
infrastructure only."""
import ctypes as C
import subprocess
from pathlib import Path

import oracle

DIR = Path(__file__).resolve().parent / "native"
GEN_FN = C.CFUNCTYPE(C.c_uint32, C.POINTER(oracle.Position), C.POINTER(C.c_uint16))
MAKE_FN = C.CFUNCTYPE(None, C.POINTER(oracle.Position), C.c_uint16)
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        subprocess.run(["make", "-s", "-C", str(DIR)], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(str(DIR / "libhostcheck.so"))
        P = C.POINTER(oracle.Position)
        L.hostcheck_count_moves.restype = C.c_uint32
        L.hostcheck_count_moves.argtypes = [P]
        L.hostcheck_generate_moves.restype = C.c_uint32
        L.hostcheck_generate_moves.argtypes = [P, C.POINTER(C.c_uint16)]
        L.hostcheck_make_move.argtypes = [P, C.c_uint16]
        L.hostcheck_in_check.argtypes = [P]
        L.hostcheck_compute_hash.restype = C.c_uint64
        L.hostcheck_compute_hash.argtypes = [P]
        L.hostcheck_analysis.argtypes = [P, C.POINTER(C.c_uint64)]
        L.hostcheck_attack_info.argtypes = [P, C.POINTER(C.c_uint64)]
        L.hostcheck_parse.argtypes = [C.c_char_p, C.c_size_t, P]
        L.hostcheck_incremental.argtypes = [P, C.c_int, C.POINTER(C.c_uint64)]
        L.hostcheck_perft.restype = C.c_uint64
        L.hostcheck_perft.argtypes = [P, C.c_int]
        L.hostcheck_walk_compare.restype = C.c_uint64
        L.hostcheck_walk_compare.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p, P, C.POINTER(C.c_uint64)]
        L.hostcheck_table.restype = C.POINTER(C.c_uint64)
        L.hostcheck_table.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
        L.hostcheck_bishop.restype = C.c_uint64
        L.hostcheck_bishop.argtypes = [C.c_int, C.c_uint64]
        L.hostcheck_rook.restype = C.c_uint64
        L.hostcheck_rook.argtypes = [C.c_int, C.c_uint64]
        L.hostcheck_slider_entry.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                             C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def walk_compare(p: oracle.Position, depth: int):
    """Compare generate_moves (ordered), count_moves and make_move with the oracle at
    every node of the perft tree to `depth`.  Returns (mismatches, visited, first_bad_fen)."""
    O = oracle.lib()
    gen = C.cast(O.oracle_generate_moves, C.c_void_p)
    mk = C.cast(O.oracle_make_move, C.c_void_p)
    bad, visited = oracle.Position(), C.c_uint64()
    n = lib().hostcheck_walk_compare(C.byref(p), depth, gen, mk, C.byref(bad), C.byref(visited))
    return n, visited.value, (oracle.to_fen(bad) if n else None)
