/*
 * fen.cuh -- FEN / EPD text to position, one thread per line, for books that are parsed on the
 * device ("GPU-side FEN/EPD batch parser").
 *
 * Same acceptance rules as the host parser (host_api.cpp), i.e. as the reference:
 *   impl TryFrom<&str> for Position   position.rs:809-821  (trim, one optional "fen " / "epd " prefix)
 *   Position::from_fen                position.rs:157-275  (strict single-space split, 4 or 6 fields)
 *   validate                          position.rs:934-1032
 * A thread cannot return the reference's error text, so the result is a status class; a caller that
 * wants the message re-parses the rejected line with pabi_position_parse on the host.  The function
 * is __host__ __device__: tests/native compiles it for the CPU and compares it with the host parser.
 */
#pragma once
#include "chess.cuh"

namespace pabi {

enum FenStatus {
  FEN_OK = 0,
  FEN_BAD_PLACEMENT = -1, /* position.rs:165-213 */
  FEN_BAD_SIDE = -2,      /* :214-217, environment.rs:29-39 */
  FEN_BAD_CASTLING = -3,  /* :218-221, core.rs:634-673 */
  FEN_BAD_EN_PASSANT = -4, /* :222-226, core.rs:286-299 */
  FEN_BAD_HALFMOVE = -5,  /* :227-236 */
  FEN_BAD_FULLMOVE = -6,  /* :237-253 */
  FEN_TRAILING = -7,      /* :255-257 */
  FEN_ILLEGAL = -8        /* validate, :934-1032 */
};

struct ParsedPosition { /* the fields of pabi_position_t */
  u64 white[6], black[6]; /* king, queens, rooks, bishops, knights, pawns (bitboard.rs:350-357) */
  u16 fullmove;
  u8 castling, stm, halfmove, ep;
};

/* the 64-byte device record of a parsed position */
PB_HD Pos record_of(const ParsedPosition& p) {
  Pos r;
  r.white = p.white[0] | p.white[1] | p.white[2] | p.white[3] | p.white[4] | p.white[5];
  r.kings = p.white[0] | p.black[0], r.queens = p.white[1] | p.black[1], r.rooks = p.white[2] | p.black[2];
  r.bishops = p.white[3] | p.black[3], r.knights = p.white[4] | p.black[4], r.pawns = p.white[5] | p.black[5];
  r.meta = make_meta(p.castling, p.ep, p.stm, p.halfmove, p.fullmove);
  return r;
}

struct FenCursor { /* Rust's `split(' ')`: empty pieces are kept */
  const char* s;
  size_t len, at;
  bool done;
  PB_HD bool next(size_t& begin, size_t& end) {
    if (done) return false;
    begin = at;
    while (at < len && s[at] != ' ') ++at;
    end = at;
    if (at >= len) done = true;
    else ++at;
    return true;
  }
};

PB_HD bool fen_is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* Rust `str::parse::<uN>()`: optional '+', at least one digit, no overflow past max */
PB_HD bool fen_number(const char* s, size_t begin, size_t end, u32 max, u32& out) {
  if (begin < end && s[begin] == '+') ++begin;
  if (begin >= end) return false;
  u32 v = 0;
  for (size_t i = begin; i < end; i++) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (u32)(s[i] - '0');
    if (v > max) return false;
  }
  out = v;
  return true;
}

/* position.rs:934-1032 */
template <class TB>
PB_HD bool fen_validate(const ParsedPosition& p, const TB& tb) {
  if (p.fullmove == 0) return false;
  if (popc(p.white[0]) != 1 || popc(p.black[0]) != 1) return false;
  if (popc(p.white[5]) > 8 || popc(p.black[5]) > 8) return false;
  if ((p.white[5] | p.black[5]) & (RANK_1 | RANK_8)) return false;
  const int us = p.stm;
  const u64* our = us == 0 ? p.white : p.black;
  const u64* their = us == 0 ? p.black : p.white;
  u64 occ = 0;
  for (int k = 0; k < 6; k++) occ |= p.white[k] | p.black[k];
  const int king = lsb(our[0]);
  const u64 king_bb = our[0];
  const u64 checkers = (tb.knight(king) & their[4]) | (pawn_attacks(king_bb, us) & their[5]) |
                       (tb.bishop(king, occ) & (their[3] | their[1])) | (tb.rook(king, occ) & (their[2] | their[1]));
  if (popc(checkers) > 2) return false;
  if (p.ep >= NO_SQUARE) return true;
  const int ep = p.ep;
  if ((ep >> 3) != (us == 0 ? 5 : 2)) return false;
  const int pushed = us == 0 ? ep - 8 : ep + 8;
  if (!(their[5] & bit(pushed))) return false;
  if (checkers) {
    if (checkers & (checkers - 1)) return false;
    if (checkers != bit(pushed)) {
      const int origin = us == 0 ? ep + 8 : ep - 8;
      if (!(tb.ray(lsb(checkers), king) & bit(origin))) return false;
    }
  }
  for (u64 b = their[1] | their[3]; b; b &= b - 1) { /* queens and bishops: bishop_ray to our king */
    const int attacker = lsb(b);
    const bool diagonal = ((attacker ^ king) & 7) != 0 && ((attacker ^ king) & 56) != 0;
    const u64 xray = diagonal ? tb.ray(attacker, king) : 0;
    if (popc(xray & occ) == 2 && (xray & bit(attacker)) && (xray & bit(pushed))) return false;
  }
  return true;
}

template <class TB>
PB_HD int parse_position(const char* text, size_t len, const TB& tb, ParsedPosition& p) {
  /* TryFrom<&str>: trim, strip one prefix */
  size_t a = 0, b = len;
  while (a < b && fen_is_space(text[a])) ++a;
  while (b > a && fen_is_space(text[b - 1])) --b;
  if (b - a >= 4 && text[a + 3] == ' ' &&
      ((text[a] == 'f' && text[a + 1] == 'e' && text[a + 2] == 'n') || (text[a] == 'e' && text[a + 1] == 'p' && text[a + 2] == 'd')))
    a += 4;
  const char* s = text + a;
  for (int k = 0; k < 6; k++) p.white[k] = p.black[k] = 0;
  FenCursor parts{s, b - a, 0, false};
  size_t begin, end;

  parts.next(begin, end); /* the placement: split always yields a first piece */
  {
    int rank_id = 8;
    size_t i = begin;
    for (;;) { /* ranks = placement.split('/') */
      if (rank_id == 0) return FEN_BAD_PLACEMENT;
      --rank_id;
      u32 file = 0;
      for (; i < end && s[i] != '/'; i++) {
        const char c = s[i];
        if (file > 8) return FEN_BAD_PLACEMENT;
        if (c == '0') return FEN_BAD_PLACEMENT;
        if (c >= '1' && c <= '9') {
          file += (u32)(c - '0');
          continue;
        }
        int kind;
        switch (c | 0x20) { /* bitboard.rs:350-357 order */
          case 'k': kind = 0; break;
          case 'q': kind = 1; break;
          case 'r': kind = 2; break;
          case 'b': kind = 3; break;
          case 'n': kind = 4; break;
          case 'p': kind = 5; break;
          default: return FEN_BAD_PLACEMENT;
        }
        if (!((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'))) return FEN_BAD_PLACEMENT;
        if (file > 7) return FEN_BAD_PLACEMENT;
        ((c & 0x20) ? p.black : p.white)[kind] |= bit(rank_id * 8 + (int)file);
        ++file;
      }
      if (file != 8) return FEN_BAD_PLACEMENT;
      if (i >= end) break;
      ++i; /* the '/' */
    }
    if (rank_id != 0) return FEN_BAD_PLACEMENT;
  }

  if (!parts.next(begin, end) || end - begin != 1 || (s[begin] != 'w' && s[begin] != 'b')) return FEN_BAD_SIDE;
  p.stm = s[begin] == 'b';

  if (!parts.next(begin, end) || begin == end) return FEN_BAD_CASTLING;
  p.castling = 0;
  if (!(end - begin == 1 && s[begin] == '-')) {
    int last = -1;
    for (size_t i = begin; i < end; i++) {
      const int idx = s[i] == 'K' ? 0 : s[i] == 'Q' ? 1 : s[i] == 'k' ? 2 : s[i] == 'q' ? 3 : -1;
      if (idx <= last) return FEN_BAD_CASTLING;
      p.castling |= (u8)(1 << idx);
      last = idx;
    }
  }

  if (!parts.next(begin, end)) return FEN_BAD_EN_PASSANT;
  if (end - begin == 1 && s[begin] == '-') {
    p.ep = NO_SQUARE;
  } else {
    if (end - begin != 2 || s[begin] < 'a' || s[begin] > 'h' || s[begin + 1] < '1' || s[begin + 1] > '8') return FEN_BAD_EN_PASSANT;
    p.ep = (u8)((s[begin] - 'a') + 8 * (s[begin + 1] - '1'));
  }

  u32 halfmove = 0, fullmove = 1;
  const bool have_half = parts.next(begin, end);
  if (have_half && !fen_number(s, begin, end, 255, halfmove)) return FEN_BAD_HALFMOVE;
  if (parts.next(begin, end)) {
    if (!fen_number(s, begin, end, 65535, fullmove) || fullmove == 0) return FEN_BAD_FULLMOVE;
  } else if (have_half) {
    return FEN_BAD_FULLMOVE;
  }
  if (parts.next(begin, end)) return FEN_TRAILING;
  p.halfmove = (u8)halfmove;
  p.fullmove = (u16)fullmove;
  return fen_validate(p, tb) ? FEN_OK : FEN_ILLEGAL;
}

} /* namespace pabi */
