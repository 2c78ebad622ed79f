//! FFI binding of `libpabi_b200.so` for the pabi crate (to live at `src/chess/cuda.rs`).
//!
//! Source only: this build image has no Rust toolchain, so the file is not compiled here.  It
//! mirrors `include/pabi_cuda.h` one-to-one; INTEGRATION.md explains the mapping onto
//! `pabi::chess` (position.rs / core.rs) and the `build.rs` hook that runs nvcc.

use std::ffi::{c_char, CStr, CString};

use crate::chess::core::{Move, MoveList};
use crate::chess::position::Position;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RawPosition {
    pub white: [u64; 6], // king, queens, rooks, bishops, knights, pawns (bitboard.rs:350-357)
    pub black: [u64; 6],
    pub hash: u64,
    pub fullmove_counter: u16,
    pub castling: u8,          // core.rs:604-612
    pub side_to_move: u8,      // environment.rs:13-16
    pub halfmove_clock: u8,
    pub en_passant_square: u8, // 0..=63, 64 = None
    pub _pad: [u8; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct Timing {
    pub device_ms: f64,
    pub host_ms: f64,
    pub algorithmic_bytes: u64,
    pub kernel_launches: u64,
    pub interior_nodes: u64,
    pub leaf_kernel_ms: f64,
    pub leaf_kernel_launches: u64,
    pub leaf_parents: u64,
    pub leaf_children: u64,
    pub merged_records: u64,
}

#[repr(C)]
pub struct Ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct Games {
    _private: [u8; 0],
}

extern "C" {
    pub fn pabi_position_starting(out: *mut RawPosition);
    pub fn pabi_position_from_fen(fen: *const c_char, out: *mut RawPosition, err: *mut c_char, cap: usize) -> i32;
    pub fn pabi_position_parse(text: *const c_char, out: *mut RawPosition, err: *mut c_char, cap: usize) -> i32;
    pub fn pabi_position_to_fen(p: *const RawPosition, out: *mut c_char, cap: usize) -> i32;
    pub fn pabi_move_from_uci(uci: *const c_char, out: *mut u16) -> i32;
    pub fn pabi_move_to_uci(mv: u16, out: *mut c_char) -> i32;
    pub fn pabi_move_flip_perspective(mv: u16) -> u16; // Move::flip_perspective, core.rs:91-97
    pub fn pabi_position_hash(p: *const RawPosition) -> u64; // Position::hash, position.rs:110
    pub fn pabi_position_compute_hash(p: *const RawPosition) -> u64; // compute_hash, position.rs:776-806
    pub fn pabi_position_num_pieces(p: *const RawPosition) -> u32; // position.rs:122-124
    pub fn pabi_position_halfmove_clock_expired(p: *const RawPosition) -> i32; // position.rs:750-752
    pub fn pabi_zobrist_keys(out: *mut u64); // 781 keys, layout in include/pabi_cuda.h

    pub fn pabi_cuda_create(device: i32, frontier_capacity: usize, out: *mut *mut Ctx) -> i32;
    pub fn pabi_cuda_destroy(ctx: *mut Ctx);
    pub fn pabi_cuda_last_error(ctx: *mut Ctx) -> *const c_char;
    pub fn pabi_cuda_device(ctx: *const Ctx) -> i32;
    pub fn pabi_cuda_bounds_violation(ctx: *mut Ctx) -> i32; // checking builds only; -1 in the shipped build
    pub fn pabi_cuda_host_alloc(bytes: usize, out: *mut *mut std::ffi::c_void) -> i32;
    pub fn pabi_cuda_host_free(ptr: *mut std::ffi::c_void);

    pub fn pabi_cuda_perft(ctx: *mut Ctx, root: *const RawPosition, depth: u8, nodes: *mut u64, t: *mut Timing) -> i32;
    pub fn pabi_cuda_perft_hashed(ctx: *mut Ctx, root: *const RawPosition, depth: u8, nodes: *mut u64, t: *mut Timing) -> i32;
    pub fn pabi_cuda_perft_batch(ctx: *mut Ctx, roots: *const RawPosition, n: usize, depth: u8, nodes_out: *mut u64,
                                 t: *mut Timing) -> i32;
    pub fn pabi_cuda_perft_batch_depths(ctx: *mut Ctx, roots: *const RawPosition, depths: *const u8, n: usize, nodes_out: *mut u64,
                                        t: *mut Timing) -> i32;
    pub fn pabi_cuda_perft_shard(ctx: *mut Ctx, root: *const RawPosition, depth: u8, split_ply: u32, shard_index: u32,
                                 shard_count: u32, nodes: *mut u64, t: *mut Timing) -> i32;
    pub fn pabi_cuda_perft_multi(ctxs: *const *mut Ctx, n_ctx: usize, root: *const RawPosition, depth: u8, split_ply: u32,
                                 nodes: *mut u64, per_ctx_ms: *mut f64) -> i32;
    pub fn pabi_cuda_perft_divide(ctx: *mut Ctx, root: *const RawPosition, depth: u8, moves_out: *mut u16,
                                  nodes_out: *mut u64, n_moves: *mut u32, t: *mut Timing) -> i32;
    pub fn pabi_cuda_generate_moves_batch(ctx: *mut Ctx, positions: *const RawPosition, n: usize, offsets: *mut u32,
                                          moves: *mut u16, moves_cap: usize, t: *mut Timing) -> i32;
    pub fn pabi_cuda_count_moves_batch(ctx: *mut Ctx, positions: *const RawPosition, n: usize, counts: *mut u32,
                                       t: *mut Timing) -> i32;
    pub fn pabi_cuda_in_check_batch(ctx: *mut Ctx, positions: *const RawPosition, n: usize, in_check: *mut u8,
                                    t: *mut Timing) -> i32;
    pub fn pabi_cuda_attack_info_batch(ctx: *mut Ctx, positions: *const RawPosition, n: usize, out: *mut u64, t: *mut Timing) -> i32;
    pub fn pabi_cuda_make_moves_batch(ctx: *mut Ctx, positions: *const RawPosition, moves: *const u16, n: usize,
                                      out: *mut RawPosition, t: *mut Timing) -> i32;
    pub fn pabi_cuda_expand_batch(ctx: *mut Ctx, positions: *const RawPosition, n: usize, offsets: *mut u32,
                                  children: *mut RawPosition, children_cap: usize, t: *mut Timing) -> i32;
    pub fn pabi_cuda_random_playouts(ctx: *mut Ctx, starts: *const RawPosition, seeds: *const u64, n_games: usize, max_plies: u32,
                                     stride: u32, per_game_cap: u32, out: *mut RawPosition, counts: *mut u32, t: *mut Timing) -> i32;
    pub fn pabi_cuda_parse_batch(ctx: *mut Ctx, text: *const c_char, line_offsets: *const u64, n: usize, out: *mut RawPosition,
                                 status: *mut i8, t: *mut Timing) -> i32;
    pub fn pabi_position_pack64(p: *const RawPosition, out: *mut u64);
    pub fn pabi_position_unpack64(words: *const u64, out: *mut RawPosition);
    // Game environment (src/chess/game.rs:21-111 without the Syzygy branch); `Games` is an opaque handle like `Ctx`
    pub fn pabi_cuda_games_create(ctx: *mut Ctx, roots: *const RawPosition, n: usize, history_cap: u32, out: *mut *mut Games) -> i32;
    pub fn pabi_cuda_games_destroy(games: *mut Games);
    pub fn pabi_cuda_games_actions(games: *mut Games, offsets: *mut u32, moves: *mut u16, moves_cap: usize) -> i32;
    pub fn pabi_cuda_games_apply(games: *mut Games, actions: *const u16) -> i32;
    pub fn pabi_cuda_games_result(games: *mut Games, results: *mut u8) -> i32;
    pub fn pabi_cuda_games_positions(games: *mut Games, out: *mut RawPosition) -> i32;
    pub fn pabi_cuda_games_play_random(games: *mut Games, seeds: *const u64, max_plies: u32, results: *mut u8, plies: *mut u32,
                                       t: *mut Timing) -> i32;
    pub fn pabi_cuda_pack_records(ctx: *mut Ctx, positions: *const RawPosition, n: usize, device_records: *mut u64,
                                  stride: usize) -> i32;
    pub fn pabi_cuda_generate_moves_device(ctx: *mut Ctx, device_records: *const u64, stride: usize, n: usize,
                                           device_offsets: *mut u32, device_moves: *mut u16, moves_cap: usize,
                                           total_moves: *mut u64, t: *mut Timing) -> i32;
}

/// Owner of one device context (`pabi_cuda_ctx`).
pub struct CudaContext {
    raw: *mut Ctx,
}

// The context serialises its own stream; it may move between threads but not be shared.
unsafe impl Send for CudaContext {}

impl CudaContext {
    /// `device < 0` selects the current CUDA device; `frontier_capacity == 0` the default.
    pub fn new(device: i32, frontier_capacity: usize) -> anyhow::Result<Self> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { pabi_cuda_create(device, frontier_capacity, &mut raw) };
        anyhow::ensure!(rc == 0, "pabi_cuda_create failed with status {rc} (no CPU fallback)");
        Ok(Self { raw })
    }

    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(pabi_cuda_last_error(self.raw)) }.to_string_lossy().into_owned()
    }

    /// Drop-in for `perft(&Position, depth)` (position.rs:909-924).
    pub fn perft(&mut self, position: &Position, depth: u8) -> u64 {
        let raw = RawPosition::from(position);
        let mut nodes = 0u64;
        let rc = unsafe { pabi_cuda_perft(self.raw, &raw, depth, &mut nodes, std::ptr::null_mut()) };
        assert_eq!(rc, 0, "{}", self.last_error());
        nodes
    }

    /// Drop-in for `Position::generate_moves` (position.rs:317-408) over a slice; returns one
    /// `MoveList` per position in the reference's order.
    pub fn generate_moves(&mut self, positions: &[Position]) -> Vec<MoveList> {
        let raw: Vec<RawPosition> = positions.iter().map(RawPosition::from).collect();
        let mut offsets = vec![0u32; raw.len() + 1];
        let mut moves = vec![0u16; 48 * raw.len().max(1)];
        loop {
            let rc = unsafe {
                pabi_cuda_generate_moves_batch(self.raw, raw.as_ptr(), raw.len(), offsets.as_mut_ptr(), moves.as_mut_ptr(),
                                               moves.len(), std::ptr::null_mut())
            };
            if rc == -2 {
                moves.resize(offsets[raw.len()] as usize, 0);
                continue;
            }
            assert_eq!(rc, 0, "{}", self.last_error());
            break;
        }
        offsets.windows(2).map(|w| moves[w[0] as usize..w[1] as usize].iter().map(|&m| Move::from_raw(m)).collect()).collect()
    }
}

impl Drop for CudaContext {
    fn drop(&mut self) {
        unsafe { pabi_cuda_destroy(self.raw) }
    }
}

impl From<&Position> for RawPosition {
    fn from(p: &Position) -> Self {
        // `Position` would gain `pub(crate) fn to_raw(&self) -> RawPosition` next to its fields
        // (position.rs:45-63); the copy is field for field.
        p.to_raw()
    }
}

/// `Position::from_fen` through the library (same acceptance rules and error text).
pub fn position_from_fen(fen: &str) -> anyhow::Result<RawPosition> {
    let text = CString::new(fen)?;
    let mut out = std::mem::MaybeUninit::<RawPosition>::uninit();
    let mut err = [0 as c_char; 512];
    let rc = unsafe { pabi_position_from_fen(text.as_ptr(), out.as_mut_ptr(), err.as_mut_ptr(), err.len()) };
    if rc != 0 {
        anyhow::bail!("{}", unsafe { CStr::from_ptr(err.as_ptr()) }.to_string_lossy());
    }
    Ok(unsafe { out.assume_init() })
}
