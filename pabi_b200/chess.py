"""Host-side mirror of pabi's `pabi::chess` API for the perft / move-generation path.

Names and semantics follow the Rust crate (paths relative to its root):

    Position::starting / from_fen / try_from / Display   src/chess/position.rs:78-91,157-275,809-860
    Position::generate_moves / make_move / in_check       src/chess/position.rs:317-433,735-746
    perft(&Position, depth)                               src/chess/position.rs:909-924
    Move::{new, from_uci, from, to, promotion}, Display   src/chess/core.rs:24-134

Everything that generates or makes moves runs on the GPU through the C ABI of
`include/pabi_cuda.h`; this module only marshals buffers.  The batched entry
points on :class:`Engine` are what a datagen / book / perft driver should use;
the single-position methods on :class:`Position` are one-element batches.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from ._native import POSITION_DTYPE, PabiCudaError, PositionStruct, Timing

NO_SQUARE = 64
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"


class Move:
    """16-bit packed move (src/chess/core.rs:24-66): bits 0-5 from, 6-11 to, 12-14 promotion."""

    __slots__ = ("raw",)
    PROMOTIONS = {None: 0, "n": 1, "b": 2, "r": 3, "q": 4}

    def __init__(self, from_square: int, to_square: int, promotion: Optional[str] = None):
        self.raw = (from_square & 63) | ((to_square & 63) << 6) | (self.PROMOTIONS[promotion] << 12)

    @classmethod
    def from_raw(cls, raw: int) -> "Move":
        m = cls.__new__(cls)
        m.raw = int(raw)
        return m

    @classmethod
    def from_uci(cls, uci: str) -> "Move":
        out = C.c_uint16()
        if _native.lib().pabi_move_from_uci(uci.encode(), C.byref(out)) != 0:
            raise ValueError(f"UCI move should be 4 or 5 characters of squares and promotion, got {uci!r}")
        return cls.from_raw(out.value)

    def from_(self) -> int:
        return self.raw & 63

    def to(self) -> int:
        return (self.raw >> 6) & 63

    def promotion(self) -> Optional[str]:
        return (None, "n", "b", "r", "q")[(self.raw >> 12) & 7]

    def flip_perspective(self) -> "Move":
        """`Move::flip_perspective` (src/chess/core.rs:91-97)."""
        return Move.from_raw(_native.lib().pabi_move_flip_perspective(self.raw))

    def __str__(self) -> str:
        buf = C.create_string_buffer(6)
        _native.lib().pabi_move_to_uci(self.raw, buf)
        return buf.value.decode()

    def __repr__(self) -> str:
        return f"Move({self})"

    def __eq__(self, other) -> bool:
        return isinstance(other, Move) and other.raw == self.raw

    def __hash__(self) -> int:
        return hash(self.raw)


def pinned_empty(shape, dtype) -> np.ndarray:
    """An uninitialised numpy array in page-locked host memory (`pabi_cuda_host_alloc`): positions, move lists and
    counts that live in such arrays cross the PCIe link without the driver's staging copy."""
    import weakref
    dt = np.dtype(dtype)
    count = int(np.prod(shape))
    ptr = C.c_void_p()
    rc = _native.lib().pabi_cuda_host_alloc(max(1, count * dt.itemsize), C.byref(ptr))
    if rc != 0:
        raise PabiCudaError(rc, "pabi_cuda_host_alloc failed")
    buf = (C.c_uint8 * max(1, count * dt.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dt, count=count).reshape(shape)
    weakref.finalize(buf, _native.lib().pabi_cuda_host_free, ptr.value)  # buf stays alive as long as a view of it does
    return arr


def _as_array(positions) -> np.ndarray:
    """Sequence of Position / structured array -> contiguous POSITION_DTYPE array."""
    if isinstance(positions, np.ndarray):
        if positions.dtype.itemsize != 112:
            raise TypeError("positions array must have 112-byte records (pabi_position_t)")
        return np.ascontiguousarray(positions)
    arr = np.zeros(len(positions), dtype=POSITION_DTYPE)
    for i, p in enumerate(positions):
        C.memmove(arr.ctypes.data + 112 * i, C.byref(p.raw), 112)
    return arr


class Engine:
    """One CUDA context of the library (`pabi_cuda_ctx`): a device, a stream, the
    staged tables and the breadth-first frontier buffers."""

    def __init__(self, device: int = -1, frontier_capacity: int = 0):
        self._ctx = C.c_void_p()
        L = _native.lib()
        rc = L.pabi_cuda_create(device, frontier_capacity, C.byref(self._ctx))
        if rc != 0:
            raise PabiCudaError(rc, "pabi_cuda_create failed (a CUDA device is required; there is no CPU fallback)")
        self.last_timing: Optional[Timing] = None

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            _native.lib().pabi_cuda_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device(self) -> int:
        return _native.lib().pabi_cuda_device(self._ctx)

    def bounds_violation(self) -> int:
        """Checking builds (-DPABI_BOUNDS): source line of the first out-of-range index since the last call, 0 = none;
        -1 in the shipped build."""
        return _native.lib().pabi_cuda_bounds_violation(self._ctx)

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise PabiCudaError(rc, _native.lib().pabi_cuda_last_error(self._ctx).decode())

    # -- perft ---------------------------------------------------------------
    def perft(self, position: "Position", depth: int) -> int:
        nodes, t = C.c_uint64(), Timing()
        self._check(_native.lib().pabi_cuda_perft(self._ctx, C.byref(position.raw), depth, C.byref(nodes), C.byref(t)))
        self.last_timing = t
        return nodes.value

    def perft_hashed(self, position: "Position", depth: int) -> int:
        """perft with transposition merging (`pabi_cuda_perft_hashed`): same result, shared subtrees walked once."""
        nodes, t = C.c_uint64(), Timing()
        self._check(_native.lib().pabi_cuda_perft_hashed(self._ctx, C.byref(position.raw), depth, C.byref(nodes), C.byref(t)))
        self.last_timing = t
        return nodes.value

    def perft_batch(self, positions, depth: int) -> np.ndarray:
        arr = _as_array(positions)
        out, t = np.zeros(len(arr), dtype=np.uint64), Timing()
        self._check(_native.lib().pabi_cuda_perft_batch(self._ctx, arr.ctypes.data, len(arr), depth, out.ctypes.data,
                                                        C.byref(t)))
        self.last_timing = t
        return out

    def perft_batch_depths(self, positions, depths: Sequence[int]) -> np.ndarray:
        """`perft(positions[i], depths[i])` for every i in one call (`pabi_cuda_perft_batch_depths`): the shape of the
        reference's perft suite, where every vector has its own depth."""
        arr = _as_array(positions)
        dp = np.ascontiguousarray(np.asarray(depths, dtype=np.uint8))
        if len(dp) != len(arr):
            raise ValueError("one depth per position")
        out, t = np.zeros(len(arr), dtype=np.uint64), Timing()
        self._check(_native.lib().pabi_cuda_perft_batch_depths(self._ctx, arr.ctypes.data, dp.ctypes.data, len(arr),
                                                               out.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out

    def perft_shard(self, position: "Position", depth: int, split_ply: int, shard_index: int, shard_count: int) -> int:
        nodes, t = C.c_uint64(), Timing()
        self._check(_native.lib().pabi_cuda_perft_shard(self._ctx, C.byref(position.raw), depth, split_ply, shard_index,
                                                        shard_count, C.byref(nodes), C.byref(t)))
        self.last_timing = t
        return nodes.value

    def divide(self, position: "Position", depth: int) -> List[Tuple[Move, int]]:
        moves = (C.c_uint16 * _native.MAX_MOVES)()
        nodes = (C.c_uint64 * _native.MAX_MOVES)()
        n, t = C.c_uint32(), Timing()
        self._check(_native.lib().pabi_cuda_perft_divide(self._ctx, C.byref(position.raw), depth, moves, nodes, C.byref(n),
                                                         C.byref(t)))
        self.last_timing = t
        return [(Move.from_raw(moves[i]), int(nodes[i])) for i in range(n.value)]

    # -- batched move generation ----------------------------------------------
    def generate_moves_batch(self, positions, offsets_out: Optional[np.ndarray] = None,
                             moves_out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """CSR move lists: (offsets[n+1] u32, moves u16) in the reference's order.  `offsets_out` / `moves_out`
        (e.g. from `pinned_empty`) are used when given and large enough."""
        arr = _as_array(positions)
        n = len(arr)
        offsets = offsets_out if offsets_out is not None and len(offsets_out) >= n + 1 else np.empty(n + 1, dtype=np.uint32)
        moves = moves_out if moves_out is not None else np.empty(max(256, 40 * n), dtype=np.uint16)  # ~28 moves per playout position
        t = Timing()
        L = _native.lib()
        rc = L.pabi_cuda_generate_moves_batch(self._ctx, arr.ctypes.data, n, offsets.ctypes.data, moves.ctypes.data,
                                              len(moves), C.byref(t))
        if rc == -2:
            moves = np.empty(int(offsets[n]), dtype=np.uint16)
            rc = L.pabi_cuda_generate_moves_batch(self._ctx, arr.ctypes.data, n, offsets.ctypes.data, moves.ctypes.data,
                                                  len(moves), C.byref(t))
        self._check(rc)
        self.last_timing = t
        return offsets[: n + 1], moves[: int(offsets[n])]

    def pack_records(self, positions, device_records: int, stride: int) -> None:
        """Positions -> 64-byte SoA records at a device address (`pabi_cuda_pack_records`): word w of record i at
        device_records[w * stride + i]."""
        arr = _as_array(positions)
        self._check(_native.lib().pabi_cuda_pack_records(self._ctx, arr.ctypes.data, len(arr), device_records, stride))

    def generate_moves_resident(self, device_records: int, stride: int, n: int, device_offsets: int, device_moves: int,
                                moves_cap: int) -> Tuple[int, Timing]:
        """Move lists of records that already live on the device, written to device arrays (`pabi_cuda_generate_moves_device`):
        returns (total moves, timing)."""
        total, t = C.c_uint64(), Timing()
        self._check(_native.lib().pabi_cuda_generate_moves_device(self._ctx, device_records, stride, n, device_offsets, device_moves,
                                                                  moves_cap, C.byref(total), C.byref(t)))
        self.last_timing = t
        return total.value, t

    def count_moves_batch(self, positions) -> np.ndarray:
        arr = _as_array(positions)
        out, t = np.zeros(len(arr), dtype=np.uint32), Timing()
        self._check(_native.lib().pabi_cuda_count_moves_batch(self._ctx, arr.ctypes.data, len(arr), out.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out

    def in_check_batch(self, positions) -> np.ndarray:
        arr = _as_array(positions)
        out, t = np.zeros(len(arr), dtype=np.uint8), Timing()
        self._check(_native.lib().pabi_cuda_in_check_batch(self._ctx, arr.ctypes.data, len(arr), out.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out.astype(bool)

    def attack_info_batch(self, positions) -> np.ndarray:
        """AttackInfo (src/chess/attacks.rs:89-97) per position: columns attacks, checkers, pins, xrays,
        safe_king_squares."""
        arr = _as_array(positions)
        out, t = np.zeros((len(arr), 5), dtype=np.uint64), Timing()
        self._check(_native.lib().pabi_cuda_attack_info_batch(self._ctx, arr.ctypes.data, len(arr), out.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out

    def make_moves_batch(self, positions, moves: Sequence[int]) -> np.ndarray:
        arr = _as_array(positions)
        mv = np.ascontiguousarray(np.asarray(moves, dtype=np.uint16))
        if len(mv) != len(arr):
            raise ValueError("one move per position")
        out, t = np.zeros(len(arr), dtype=POSITION_DTYPE), Timing()
        self._check(_native.lib().pabi_cuda_make_moves_batch(self._ctx, arr.ctypes.data, mv.ctypes.data, len(arr),
                                                             out.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out

    def random_playouts(self, starts, seeds, max_plies: int, stride: int = 1) -> List[np.ndarray]:
        """Seeded random playouts on the device (`pabi_cuda_random_playouts`): one array of recorded
        positions per game."""
        arr = _as_array(starts)
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        if len(sd) != len(arr):
            raise ValueError("one seed per game")
        cap = max_plies // stride + 1
        out = np.empty(len(arr) * cap, dtype=POSITION_DTYPE)
        counts, t = np.zeros(len(arr), dtype=np.uint32), Timing()
        self._check(_native.lib().pabi_cuda_random_playouts(self._ctx, arr.ctypes.data, sd.ctypes.data, len(arr), max_plies, stride,
                                                            cap, out.ctypes.data, counts.ctypes.data, C.byref(t)))
        self.last_timing = t
        return [out[g * cap: g * cap + int(counts[g])] for g in range(len(arr))]

    def parse_batch(self, lines: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
        """FEN / EPD lines parsed on the device (`pabi_cuda_parse_batch`, `Position::try_from` semantics):
        (positions, status) with status 0 = accepted, < 0 = the rejection class."""
        encoded = [l.encode("utf-8") for l in lines]
        offsets = np.zeros(len(encoded) + 1, dtype=np.uint64)
        np.cumsum([len(b) for b in encoded], out=offsets[1:])
        text = b"".join(encoded)
        out = np.empty(len(encoded), dtype=POSITION_DTYPE)
        status, t = np.zeros(len(encoded), dtype=np.int8), Timing()
        self._check(_native.lib().pabi_cuda_parse_batch(self._ctx, text, offsets.ctypes.data, len(encoded), out.ctypes.data,
                                                        status.ctypes.data, C.byref(t)))
        self.last_timing = t
        return out, status

    def expand_batch(self, positions) -> Tuple[np.ndarray, np.ndarray]:
        """Every child of every position in generate_moves order: (offsets[n+1], children)."""
        arr = _as_array(positions)
        n = len(arr)
        offsets = np.zeros(n + 1, dtype=np.uint32)
        L, t = _native.lib(), Timing()
        self._check(L.pabi_cuda_expand_batch(self._ctx, arr.ctypes.data, n, offsets.ctypes.data, None, 0, C.byref(t)))
        children = np.zeros(int(offsets[n]), dtype=POSITION_DTYPE)
        if len(children):
            self._check(L.pabi_cuda_expand_batch(self._ctx, arr.ctypes.data, n, offsets.ctypes.data, children.ctypes.data,
                                                 len(children), C.byref(t)))
        self.last_timing = t
        return offsets, children


def perft_multi(engines: Sequence[Engine], position: "Position", depth: int, split_ply: int = 4) -> Tuple[int, List[float]]:
    """perft over several contexts (normally one per GPU) from one process: `pabi_cuda_perft_multi`.
    Returns (nodes, per-context device milliseconds)."""
    n = len(engines)
    ctxs = (C.c_void_p * n)(*[e._ctx for e in engines])
    nodes, ms = C.c_uint64(), (C.c_double * n)()
    rc = _native.lib().pabi_cuda_perft_multi(ctxs, n, C.byref(position.raw), depth, split_ply, C.byref(nodes), ms)
    if rc != 0:
        raise PabiCudaError(rc, "; ".join(_native.lib().pabi_cuda_last_error(e._ctx).decode() for e in engines))
    return nodes.value, list(ms)


_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


class Position:
    """pabi's `Position` (src/chess/position.rs:45-63) as a 112-byte C struct."""

    __slots__ = ("raw",)

    def __init__(self, raw: Optional[PositionStruct] = None):
        self.raw = raw if raw is not None else PositionStruct()

    @classmethod
    def starting(cls) -> "Position":
        p = cls()
        _native.lib().pabi_position_starting(C.byref(p.raw))
        return p

    @classmethod
    def from_fen(cls, fen: str) -> "Position":
        """Strict FEN / operation-less EPD (position.rs:157-275); raises ValueError with the reference's message."""
        p, err = cls(), C.create_string_buffer(512)
        if _native.lib().pabi_position_from_fen(fen.encode("utf-8"), C.byref(p.raw), err, 512) != 0:
            raise ValueError(err.value.decode("utf-8", "replace"))
        return p

    @classmethod
    def try_from(cls, text: str) -> "Position":
        """`impl TryFrom<&str> for Position` (position.rs:809-821): trims, accepts a "fen "/"epd " prefix."""
        p, err = cls(), C.create_string_buffer(512)
        if _native.lib().pabi_position_parse(text.encode("utf-8"), C.byref(p.raw), err, 512) != 0:
            raise ValueError(err.value.decode("utf-8", "replace"))
        return p

    @classmethod
    def from_record(cls, record) -> "Position":
        return cls(PositionStruct.from_buffer_copy(np.asarray(record).tobytes()))

    def pack64(self) -> np.ndarray:
        """The 64-byte record of the position (`pabi_position_pack64`): the device / wire format."""
        out = np.zeros(8, dtype=np.uint64)
        _native.lib().pabi_position_pack64(C.byref(self.raw), out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out

    @classmethod
    def unpack64(cls, words) -> "Position":
        w = np.ascontiguousarray(np.asarray(words, dtype=np.uint64))
        p = cls()
        _native.lib().pabi_position_unpack64(w.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(p.raw))
        return p

    def clone(self) -> "Position":
        q = Position()
        C.memmove(C.byref(q.raw), C.byref(self.raw), 112)
        return q

    def __str__(self) -> str:
        buf = C.create_string_buffer(128)
        _native.lib().pabi_position_to_fen(C.byref(self.raw), buf, 128)
        return buf.value.decode()

    def __repr__(self) -> str:
        return f"Position({self})"

    def __eq__(self, other) -> bool:
        return isinstance(other, Position) and bytes(self.raw)[:96] == bytes(other.raw)[:96] and str(self) == str(other)

    def us(self) -> int:
        return self.raw.side_to_move

    def they(self) -> int:
        return self.raw.side_to_move ^ 1

    def num_pieces(self) -> int:
        """`Position::num_pieces` (position.rs:122-124)."""
        return _native.lib().pabi_position_num_pieces(C.byref(self.raw))

    def halfmove_clock_expired(self) -> bool:
        """`Position::halfmove_clock_expired` (position.rs:750-752)."""
        return bool(_native.lib().pabi_position_halfmove_clock_expired(C.byref(self.raw)))

    def hash(self) -> int:
        """`Position::hash` (position.rs:110): the cached Zobrist hash, set by the parsers and maintained by make_move."""
        return _native.lib().pabi_position_hash(C.byref(self.raw))

    def compute_hash(self) -> int:
        """`compute_hash` (position.rs:776-806) over the current fields."""
        return _native.lib().pabi_position_compute_hash(C.byref(self.raw))

    def generate_moves(self, engine: Optional[Engine] = None) -> List[Move]:
        _, moves = (engine or default_engine()).generate_moves_batch([self])
        return [Move.from_raw(m) for m in moves]

    def make_move(self, move: Move, engine: Optional[Engine] = None) -> None:
        out = (engine or default_engine()).make_moves_batch([self], [move.raw])
        C.memmove(C.byref(self.raw), out.ctypes.data, 112)  # the hash field comes back updated, as make_move does

    def attack_info(self, engine: Optional[Engine] = None) -> dict:
        """`Position::attack_info()` (position.rs:285-292)."""
        row = (engine or default_engine()).attack_info_batch([self])[0]
        return dict(zip(("attacks", "checkers", "pins", "xrays", "safe_king_squares"), (int(x) for x in row)))

    def in_check(self, engine: Optional[Engine] = None) -> bool:
        return bool((engine or default_engine()).in_check_batch([self])[0])


def zobrist_keys() -> np.ndarray:
    """The 781 Zobrist keys in the layout of `pabi_zobrist_keys` (include/pabi_cuda.h)."""
    out = np.zeros(781, dtype=np.uint64)
    _native.lib().pabi_zobrist_keys(out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def perft(position: Position, depth: int, engine: Optional[Engine] = None) -> int:
    """`perft(&Position, depth) -> u64` (src/chess/position.rs:909-924) on the GPU."""
    return (engine or default_engine()).perft(position, depth)
