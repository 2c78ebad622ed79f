/*
 * record.h -- conversion between the 112-byte pabi_position_t (the image of
 * pabi's `Position`, position.rs:45-63) and the 64-byte device record of
 * chess.cuh.  Host code.
 */
#pragma once
#include "../../include/pabi_cuda.h"
#include "chess.cuh"

namespace pabi {

/* bitboard.rs:350-357 field order inside pabi_position_t::white / ::black */
enum { F_KING = 0, F_QUEENS, F_ROOKS, F_BISHOPS, F_KNIGHTS, F_PAWNS };

inline Pos pack_record(const pabi_position_t& p) {
  Pos r;
  r.white = p.white[0] | p.white[1] | p.white[2] | p.white[3] | p.white[4] | p.white[5];
  r.pawns = p.white[F_PAWNS] | p.black[F_PAWNS];
  r.knights = p.white[F_KNIGHTS] | p.black[F_KNIGHTS];
  r.bishops = p.white[F_BISHOPS] | p.black[F_BISHOPS];
  r.rooks = p.white[F_ROOKS] | p.black[F_ROOKS];
  r.queens = p.white[F_QUEENS] | p.black[F_QUEENS];
  r.kings = p.white[F_KING] | p.black[F_KING];
  r.meta = make_meta(p.castling & 15, p.en_passant_square > 63 ? NO_SQUARE : p.en_passant_square,
                     p.side_to_move & 1, p.halfmove_clock, p.fullmove_counter);
  return r;
}

inline pabi_position_t unpack_record(const Pos& r, u64 hash) {
  pabi_position_t p;
  const u64 w = r.white;
  p.white[F_KING] = r.kings & w, p.black[F_KING] = r.kings & ~w;
  p.white[F_QUEENS] = r.queens & w, p.black[F_QUEENS] = r.queens & ~w;
  p.white[F_ROOKS] = r.rooks & w, p.black[F_ROOKS] = r.rooks & ~w;
  p.white[F_BISHOPS] = r.bishops & w, p.black[F_BISHOPS] = r.bishops & ~w;
  p.white[F_KNIGHTS] = r.knights & w, p.black[F_KNIGHTS] = r.knights & ~w;
  p.white[F_PAWNS] = r.pawns & w, p.black[F_PAWNS] = r.pawns & ~w;
  p.hash = hash;
  p.fullmove_counter = (uint16_t)meta_fullmove(r.meta);
  p.castling = (uint8_t)meta_castling(r.meta);
  p.side_to_move = (uint8_t)meta_stm(r.meta);
  p.halfmove_clock = (uint8_t)meta_halfmove(r.meta);
  p.en_passant_square = (uint8_t)meta_ep(r.meta);
  p._pad[0] = p._pad[1] = 0;
  return p;
}

} /* namespace pabi */
