"""ctypes binding of libpabi_b200.so (include/pabi_cuda.h).

The library is built in-tree by `make -C pabi_b200/csrc` (see
`__graft_entry__.build`).  There is no Python or CPU fallback: if the shared
library is missing, or no CUDA device is present when a device entry point is
called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libpabi_b200.so"
MAX_MOVES = 256


class PositionStruct(C.Structure):
    """`pabi_position_t`: image of pabi's `Position` (src/chess/position.rs:45-63)."""

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
assert POSITION_DTYPE.itemsize == 112 and C.sizeof(PositionStruct) == 112


class Timing(C.Structure):
    """`pabi_timing_t`."""

    _fields_ = [
        ("device_ms", C.c_double),
        ("host_ms", C.c_double),
        ("algorithmic_bytes", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("interior_nodes", C.c_uint64),
        ("leaf_kernel_ms", C.c_double),
        ("leaf_kernel_launches", C.c_uint64),
        ("leaf_parents", C.c_uint64),
        ("leaf_children", C.c_uint64),
        ("merged_records", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {"device_ms": self.device_ms, "host_ms": self.host_ms, "algorithmic_bytes": self.algorithmic_bytes,
                "kernel_launches": self.kernel_launches, "interior_nodes": self.interior_nodes,
                "leaf_kernel_ms": self.leaf_kernel_ms, "leaf_kernel_launches": self.leaf_kernel_launches,
                "leaf_parents": self.leaf_parents, "leaf_children": self.leaf_children,
                "merged_records": self.merged_records}


class PabiCudaError(RuntimeError):
    """A device entry point failed (status > 0: cudaError_t; < 0: argument error)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"pabi_cuda status {status}: {message}")
        self.status = status


EXPORTS = (
    "pabi_position_starting", "pabi_position_from_fen", "pabi_position_parse", "pabi_position_to_fen",
    "pabi_move_from_uci", "pabi_move_to_uci", "pabi_move_flip_perspective", "pabi_position_hash", "pabi_position_compute_hash",
    "pabi_position_num_pieces", "pabi_position_halfmove_clock_expired", "pabi_zobrist_keys", "pabi_cuda_perft_batch_depths", "pabi_cuda_bounds_violation",
    "pabi_cuda_create", "pabi_cuda_destroy", "pabi_cuda_last_error",
    "pabi_cuda_device", "pabi_cuda_host_alloc", "pabi_cuda_host_free", "pabi_cuda_perft", "pabi_cuda_perft_hashed", "pabi_cuda_perft_batch", "pabi_cuda_perft_shard", "pabi_cuda_perft_multi", "pabi_cuda_perft_divide",
    "pabi_cuda_generate_moves_batch", "pabi_cuda_count_moves_batch", "pabi_cuda_in_check_batch",
    "pabi_cuda_attack_info_batch", "pabi_cuda_make_moves_batch", "pabi_cuda_expand_batch", "pabi_cuda_pack_records",
    "pabi_cuda_generate_moves_device", "pabi_cuda_random_playouts",
    "pabi_cuda_parse_batch", "pabi_position_pack64", "pabi_position_unpack64",
    "pabi_cuda_games_create", "pabi_cuda_games_destroy", "pabi_cuda_games_actions", "pabi_cuda_games_apply",
    "pabi_cuda_games_result", "pabi_cuda_games_positions", "pabi_cuda_games_play_random",
)

_lib = None


def build(force: bool = False) -> Path:
    """Compile the CUDA library for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(PKG_DIR / "csrc")] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `make -C pabi_b200/csrc` (there is no fallback implementation)")
    L = C.CDLL(str(LIB_PATH))
    P = C.POINTER(PositionStruct)
    T = C.POINTER(Timing)
    vp, sz, u8, u32, u64p = C.c_void_p, C.c_size_t, C.c_uint8, C.c_uint32, C.POINTER(C.c_uint64)
    L.pabi_position_starting.argtypes = [P]
    L.pabi_position_starting.restype = None
    L.pabi_position_from_fen.argtypes = [C.c_char_p, P, C.c_char_p, sz]
    L.pabi_position_parse.argtypes = [C.c_char_p, P, C.c_char_p, sz]
    L.pabi_position_to_fen.argtypes = [P, C.c_char_p, sz]
    L.pabi_move_from_uci.argtypes = [C.c_char_p, C.POINTER(C.c_uint16)]
    L.pabi_move_to_uci.argtypes = [C.c_uint16, C.c_char_p]
    L.pabi_move_flip_perspective.argtypes = [C.c_uint16]
    L.pabi_move_flip_perspective.restype = C.c_uint16
    L.pabi_position_hash.argtypes = [P]
    L.pabi_position_hash.restype = C.c_uint64
    L.pabi_position_compute_hash.argtypes = [P]
    L.pabi_position_compute_hash.restype = C.c_uint64
    L.pabi_position_num_pieces.argtypes = [P]
    L.pabi_position_num_pieces.restype = C.c_uint32
    L.pabi_position_halfmove_clock_expired.argtypes = [P]
    L.pabi_zobrist_keys.argtypes = [u64p]
    L.pabi_zobrist_keys.restype = None
    L.pabi_cuda_perft_batch_depths.argtypes = [vp, vp, vp, sz, vp, T]
    L.pabi_cuda_create.argtypes = [C.c_int, sz, C.POINTER(vp)]
    L.pabi_cuda_destroy.argtypes = [vp]
    L.pabi_cuda_destroy.restype = None
    L.pabi_cuda_last_error.argtypes = [vp]
    L.pabi_cuda_last_error.restype = C.c_char_p
    L.pabi_cuda_device.argtypes = [vp]
    L.pabi_cuda_bounds_violation.argtypes = [vp]
    L.pabi_cuda_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.pabi_cuda_host_free.argtypes = [vp]
    L.pabi_cuda_host_free.restype = None
    L.pabi_cuda_perft.argtypes = [vp, P, u8, u64p, T]
    L.pabi_cuda_perft_hashed.argtypes = [vp, P, u8, u64p, T]
    L.pabi_cuda_perft_batch.argtypes = [vp, vp, sz, u8, vp, T]
    L.pabi_cuda_perft_shard.argtypes = [vp, P, u8, u32, u32, u32, u64p, T]
    L.pabi_cuda_perft_multi.argtypes = [C.POINTER(vp), sz, P, u8, u32, u64p, C.POINTER(C.c_double)]
    L.pabi_cuda_perft_divide.argtypes = [vp, P, u8, vp, vp, C.POINTER(u32), T]
    L.pabi_cuda_generate_moves_batch.argtypes = [vp, vp, sz, vp, vp, sz, T]
    L.pabi_cuda_count_moves_batch.argtypes = [vp, vp, sz, vp, T]
    L.pabi_cuda_in_check_batch.argtypes = [vp, vp, sz, vp, T]
    L.pabi_cuda_attack_info_batch.argtypes = [vp, vp, sz, vp, T]
    L.pabi_cuda_make_moves_batch.argtypes = [vp, vp, vp, sz, vp, T]
    L.pabi_cuda_expand_batch.argtypes = [vp, vp, sz, vp, vp, sz, T]
    L.pabi_cuda_random_playouts.argtypes = [vp, vp, vp, sz, u32, u32, u32, vp, vp, T]
    L.pabi_cuda_parse_batch.argtypes = [vp, C.c_char_p, vp, sz, vp, vp, T]
    L.pabi_position_pack64.argtypes = [P, C.POINTER(C.c_uint64)]
    L.pabi_position_pack64.restype = None
    L.pabi_position_unpack64.argtypes = [C.POINTER(C.c_uint64), P]
    L.pabi_position_unpack64.restype = None
    L.pabi_cuda_games_create.argtypes = [vp, vp, sz, u32, C.POINTER(vp)]
    L.pabi_cuda_games_destroy.argtypes = [vp]
    L.pabi_cuda_games_destroy.restype = None
    L.pabi_cuda_games_actions.argtypes = [vp, vp, vp, sz]
    L.pabi_cuda_games_apply.argtypes = [vp, vp]
    L.pabi_cuda_games_result.argtypes = [vp, vp]
    L.pabi_cuda_games_positions.argtypes = [vp, vp]
    L.pabi_cuda_games_play_random.argtypes = [vp, vp, u32, vp, vp, T]
    L.pabi_cuda_pack_records.argtypes = [vp, vp, sz, vp, sz]
    L.pabi_cuda_generate_moves_device.argtypes = [vp, vp, sz, sz, vp, vp, sz, u64p, T]
    _lib = L
    return L
