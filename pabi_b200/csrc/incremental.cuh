/*
 * incremental.cuh -- the "quiet move" fast path of the fused last-two-plies kernel (K2).
 *
 * perft spends its time counting the replies to every move of every parent of the last ply but
 * one.  For most moves the reply count can be derived from a per-parent context instead of being
 * recomputed: if the mover S plays a quiet move whose from- and to-square touch neither a line of
 * the replying side T's king nor the attack set of any T slider, then
 *   - T is not in check and none of its pieces becomes (un)pinned,
 *   - every T knight / bishop / rook / queen has exactly the moves it would have had in the
 *     parent position (the vacated square was a capture target and is now a quiet target, the
 *     occupied square the other way round),
 * so  replies = base(parent) + T's pawn moves + T's king moves, where only the last two depend on
 * the move.  Everything here is in K2's normalised frame: the parent is stored with BLACK (S) to
 * move, T is white.  A move that does not qualify ("other") takes the full path
 * (make_move_kind + count_child_white); the split is decided per target set in the parent's thread,
 * so both kinds of children are processed by full warps.
 *
 * Exactness is tested in tests/test_device_logic_on_host.py: for every node of the reference
 * trees every move classified as quiet-clean has fast count == full count.
 */
#pragma once
#include "chess.cuh"

namespace pabi {

/* per-parent context, carried in the shared-memory parent record: `base` in meta bits 32-47 (the
 * fullmove counter is not needed in K2), `guard` in a side array */
constexpr int META_BASE = 32;

struct TContext {
  u64 reach; /* union of T's slider attack sets plus the pieces that shield T's king from an S slider: a clean move
                neither leaves nor lands on it; ~0 = no clean moves */
  u64 guard; /* reach plus the king guard below: a clean slider move does not land on it */
  u64 zone;  /* squares T's pawn and king moves depend on: a clean move that also stays off it is INERT */
  u32 base;  /* moves of T's knights, bishops, rooks and queens in the parent position */
  u32 total; /* base + T's pawn and king moves in the parent position = the reply count of every inert move */
};

/* king steps + castling of a white-to-move position that is not in check (the `need` / `rem` scan of
 * count_moves_white) */
template <class TB>
PB_HD u32 king_part_white(u64 our, u64 their, u64 occ, u64 pawns, u64 knights, u64 diag_sliders, u64 orth_sliders, int rights, int king,
                          int their_king, const TB& tb) {
  const u64 king_targets = tb.king(king) & ~our;
  u64 castle_walks = 0;
  if (rights) {
    if ((rights & 1) && (occ & 0x60ULL) == 0) castle_walks = 0x60ULL;
    if ((rights & 2) && (occ & 0x0EULL) == 0) castle_walks |= 0x0CULL;
  }
  const u64 need = king_targets | castle_walks;
  if (!need) return 0;
  const u64 occ_nk = occ ^ bit(king);
  u64 rem = need & ~(tb.king(their_king) | pawn_attacks(pawns & their, 1));
  for (Squares b(knights & their & tb.s->knight_zone[king]); b.any();) rem &= ~tb.knight(b.pop());
  for (Squares b(diag_sliders & their & tb.s->diag_zone[king]); b.any() && rem;) {
    const int sq = b.pop();
    if ((tb.s->diag[sq] | tb.s->anti[sq]) & rem) rem &= ~tb.bishop(sq, occ_nk);
  }
  for (Squares b(orth_sliders & their & tb.s->orth_zone[king]); b.any() && rem;) {
    const int sq = b.pop();
    if (tb.s->orth[sq] & rem) rem &= ~tb.rook(sq, occ_nk);
  }
  u32 n = popc(king_targets & rem);
  if (castle_walks) {
    const u64 safe_walks = castle_walks & rem;
    n += ((safe_walks & 0x60ULL) == 0x60ULL) + ((safe_walks & 0x0CULL) == 0x0CULL);
  }
  return n;
}

/* T's pawn moves when nothing is pinned, T is not in check and there is no en passant square
 * (position.rs:1106-1224 set-wise), plus T's king moves */
template <class TB>
PB_HD u32 pawn_king_part_white(u64 W, u64 pawns, u64 knights, u64 bishops, u64 rooks, u64 queens, u64 kings, int rights, int king,
                               int their_king, const TB& tb) {
  const u64 occ = pawns | knights | bishops | rooks | queens | kings;
  const u64 their = occ ^ W;
  u32 cnt = king_part_white(W, their, occ, pawns, knights, bishops | queens, rooks | queens, rights, king, their_king, tb);
  const u64 wp = pawns & W;
  const u64 west = ((wp & ~FILE_A) << 7) & their;
  const u64 east = ((wp & ~FILE_H) << 9) & their;
  const u64 single = (wp << 8) & ~occ;
  const u64 dbl = ((single & RANK_3) << 8) & ~occ;
  cnt += popc(west) + popc(east) + popc(single) + popc(dbl);
  if ((west | east | single) & RANK_8) cnt += 3 * (popc(west & RANK_8) + popc(east & RANK_8) + popc(single & RANK_8));
  return cnt;
}

/* King guard: the squares on the rays from T's king up to and including the SECOND piece of each ray.
 * Only there can a move change who checks or pins: a piece leaving (it is the first or second blocker of
 * the ray) or a slider arriving (it sees the king, or pins the single blocker).  Farther out at least two
 * pieces shield the king whatever happens to a third, and a knight or pawn arriving on a ray changes
 * nothing. */
template <class TB>
PB_HD TContext t_context(const Pos& n, const TB& tb) {
  const u64 occ = n.pawns | n.knights | n.bishops | n.rooks | n.queens | n.kings;
  const u64 W = n.white, B = occ ^ W;
  const int tk = lsb(n.kings & W);
  const u64 s_diag = (n.bishops | n.queens) & B, s_orth = (n.rooks | n.queens) & B;
  const u64 kd = tb.bishop(tk, occ), ko = tb.rook(tk, occ);
  TContext c{~0ULL, ~0ULL, ~0ULL, 0, 0};
  /* T in check (cannot happen when S is to move in a legal position): no fast path */
  if ((tb.knight(tk) & n.knights & B) | (pawn_attacks(n.kings & W, 0) & n.pawns & B) | (kd & s_diag) | (ko & s_orth)) return c;
  /* T's pinned pieces: their restricted mobility goes into `base` (a clean move leaves the pins as they
   * are); a pinned T pawn would need the pin-aware pawn count, so such parents take the full path */
  u64 pins = 0;
  {
    const u64 cand = kd & W;
    if (cand) {
      u64 pinners = tb.bishop(tk, occ ^ cand) & ~kd & s_diag;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), tk) & cand;
    }
  }
  {
    const u64 cand = ko & W;
    if (cand) {
      u64 pinners = tb.rook(tk, occ ^ cand) & ~ko & s_orth;
      for (; pinners; pinners &= pinners - 1) pins |= tb.ray(lsb(pinners), tk) & cand;
    }
  }
  if (pins & n.pawns) return c;
  const u64 not_w = ~W;
  u64 reach = 0;
  u32 base = 0;
  for (Squares b(n.knights & W & ~pins); b.any();) base += popc(tb.knight(b.pop()) & not_w);
  /* sliders hemmed in by T's own pieces see only those: nothing for `reach`, nothing for `base` */
  for (Squares b((n.rooks | n.queens) & W & orth_neighbours(not_w)); b.any();) {
    const int sq = b.pop();
    const u64 a = tb.rook(sq, occ);
    reach |= a;
    base += popc(a & not_w & ((pins >> sq) & 1 ? tb.line(tk, sq) : ~0ULL));
  }
  for (Squares b((n.bishops | n.queens) & W & diag_neighbours(not_w)); b.any();) {
    const int sq = b.pop();
    const u64 a = tb.bishop(sq, occ);
    reach |= a;
    base += popc(a & not_w & ((pins >> sq) & 1 ? tb.line(tk, sq) : ~0ULL));
  }
  c.guard = reach | tb.bishop(tk, occ & ~kd) | tb.rook(tk, occ & ~ko);
  /* Leaving a square of a king line matters only if that uncovers an S slider: the square is one of at most
   * two pieces between T's king and an S slider that moves along that line (with three or more, two still
   * shield the king afterwards). */
  u64 uncover = 0;
  for (Squares b((tb.s->orth[tk] & s_orth) | ((tb.s->diag[tk] | tb.s->anti[tk]) & s_diag)); b.any();) {
    const int sq = b.pop();
    const u64 between = tb.ray(sq, tk) & occ & ~bit(sq);
    if (popc(between) <= 2) uncover |= between;
  }
  /* with a T piece pinned, a knight, pawn or king arriving between it and its pinner would free it:
   * such parents hold every mover to the full guard */
  c.reach = pins ? c.guard : reach | uncover;
  c.base = base;
  /* Inert moves.  T's pawn moves depend on the squares in front of and diagonally ahead of its pawns; its king
   * moves on the king's neighbourhood (and the castling walks) and on every square from which a piece can see
   * one of those squares.  An S move that is clean and whose two squares lie outside this zone leaves the
   * visible set of every such square unchanged, so T has exactly the moves it would have had in the parent
   * position with the move passed: `total`. */
  const u64 wp = n.pawns & W;
  const int rights = meta_castling(n.meta) & 3;
  u64 zone = (wp << 8) | ((wp & RANK_2) << 16) | ((wp & ~FILE_A) << 7) | ((wp & ~FILE_H) << 9);
  u64 watched = tb.king(tk) & not_w; /* squares whose attackers matter */
  if (rights & 1) zone |= 0x60ULL, watched |= (occ & 0x60ULL) ? 0 : 0x60ULL;
  if (rights & 2) zone |= 0x0EULL, watched |= (occ & 0x0EULL) ? 0 : 0x0CULL;
  zone |= watched;
  const u64 occ_nk = occ ^ (n.kings & W);
  for (Squares b(watched); b.any();) {
    const int sq = b.pop();
    zone |= tb.bishop(sq, occ_nk) | tb.rook(sq, occ_nk) | tb.knight(sq);
  }
  c.zone = zone;
  c.total = base + pawn_king_part_white(W, n.pawns, n.knights, n.bishops, n.rooks, n.queens, n.kings, rights, tk, lsb(n.kings & B), tb);
  return c;
}

/* S's legal moves (black to move) as target sets split into clean and other.  Sink:
 *   piece(from, kind, clean_targets, other_targets)   non-pawn moves (no promotion)
 *   pushes(clean_to, other_to, delta)                 pawn pushes, from = to + delta; other_to on rank 1 promote
 *   pawn_capture(from, to)                            promotes when `to` is on rank 1
 *   en_passant(from, to)
 *   castle(from, to)
 * The union of everything reported is exactly generate_moves(n). */
template <class TB, class Sink>
PB_HD void enumerate_split(const Pos& n, const TB& tb, u64 reach, u64 guard, Sink& sink) {
  Analysis a;
  analyze<1>(n, tb, a);
  const u64 W = a.their;
  const u64 tk_bb = n.kings & W;
  const int tk = lsb(tk_bb);
  /* a clean move leaves a square outside `reach`; a slider lands outside `guard`; a knight, pawn or king
   * outside `reach` (arriving on a king ray they change neither checks nor pins) */
  const u64 open = ~guard & ~W;
  const u64 open_leaper = ~reach & ~W;
  {
    const u64 c = (reach >> a.king) & 1 ? 0 : a.safe_king & open_leaper;
    sink.piece(a.king, KIND_KING, c, a.safe_king & ~c);
  }
  if (a.block == 0) return;
  const u64 targets = ~a.our & a.block;
  const u64 movers = a.our;
  const u64 orth_movers = (n.rooks | n.queens) & movers & orth_neighbours(~a.our); /* hemmed-in sliders have no moves */
  const u64 diag_movers = (n.bishops | n.queens) & movers & diag_neighbours(~a.our);

  for (Squares b(n.knights & movers & ~a.pins); b.any();) {
    const int from = b.pop();
    const u64 t = tb.knight(from) & targets;
    const u64 c = (reach >> from) & 1 ? 0 : t & open_leaper & ~tb.knight(tk); /* a knight on KNIGHT[tk] checks */
    sink.piece(from, KIND_KNIGHT, c, t & ~c);
  }
  for (Squares b(orth_movers); b.any();) {
    const int from = b.pop();
    u64 t = tb.rook(from, a.occ) & targets;
    if ((a.pins >> from) & 1) t &= tb.line(a.king, from);
    const u64 c = (reach >> from) & 1 ? 0 : t & open; /* off the king's lines a slider cannot check */
    sink.piece(from, (n.queens >> from) & 1 ? KIND_QUEEN : KIND_ROOK, c, t & ~c);
  }
  for (Squares b(diag_movers); b.any();) {
    const int from = b.pop();
    u64 t = tb.bishop(from, a.occ) & targets;
    if ((a.pins >> from) & 1) t &= tb.line(a.king, from);
    const u64 c = (reach >> from) & 1 ? 0 : t & open;
    sink.piece(from, (n.queens >> from) & 1 ? KIND_QUEEN : KIND_BISHOP, c, t & ~c);
  }

  /* pawns of the mover (black, moving down the board) */
  {
    const PawnSets ps = pawn_sources(n, tb, a);
    {
      const u64 capture_targets = W & a.block;
      const u64 west_sources = ps.west & (capture_targets << 9) & ~FILE_A; /* black captures towards file a: from - 9 */
      const u64 east_sources = ps.east & (capture_targets << 7) & ~FILE_H; /* towards file h: from - 7 */
      for (Squares b(west_sources | east_sources); b.any();) {
        const int from = b.pop();
        const u64 from_bb = bit(from);
        const u64 t = (from_bb & west_sources) >> 9 | (from_bb & east_sources) >> 7;
        for (Squares ts(t); ts.any();) sink.pawn_capture(from, ts.pop());
      }
    }
    {
      const u64 empty = ~a.occ;
      const u64 single_all = (ps.pushers >> 8) & empty;
      const u64 single = single_all & a.block;
      const u64 dbl = ((single_all & RANK_6) >> 8) & empty & a.block;
      const u64 checks = pawn_attacks(tk_bb, 0);            /* squares from which a black pawn attacks T's king */
      const u64 t_pawns = n.pawns & W;
      const u64 beside_t_pawn = ((t_pawns & ~FILE_A) >> 1) | ((t_pawns & ~FILE_H) << 1); /* a double push there creates an en passant square */
      const u64 c1 = single & open_leaper & ~(reach >> 8) & ~RANK_1 & ~checks;
      const u64 c2 = dbl & open_leaper & ~(reach >> 16) & ~checks & ~beside_t_pawn;
      sink.pushes(c1, single & ~c1, 8);
      sink.pushes(c2, dbl & ~c2, 16);
    }
  }
  if (meta_ep(n.meta) < NO_SQUARE)
    for (Squares b(en_passant_sources(n, tb, a)); b.any();) sink.en_passant(b.pop(), meta_ep(n.meta));
  {
    const int c = castle_moves(n, a);
    if (c & 1) sink.castle(60, 62);
    if (c & 2) sink.castle(60, 58);
  }
}

/* Inert moves (clean and off the parent's zone) are never materialised: each of them has TContext::total replies. */
PB_HD u64 inert_targets(u64 zone, int from, u64 clean) { return (zone >> from) & 1 ? 0 : clean & ~zone; }
PB_HD u64 inert_pushes(u64 zone, u64 clean, int delta) { return clean & ~zone & ~(zone >> delta); } /* from = to + delta */

/* (clean, other, inert) move counts of a parent: what K2's DirectEmitter (kernels.cuh) reserves in the pool and what it
 * settles on the spot.  Host-side reference for the tests (tests/native/hostcheck.cpp). */
struct SplitCounter {
  u64 zone;
  u32 clean = 0, other = 0, inert = 0;
  PB_HD void piece(int from, int, u64 c, u64 o) {
    const u32 i = popc(inert_targets(zone, from, c));
    inert += i, clean += popc(c) - i, other += popc(o);
  }
  PB_HD void pushes(u64 c, u64 o, int delta) {
    const u32 i = popc(inert_pushes(zone, c, delta));
    inert += i, clean += popc(c) - i, other += popc(o) + 3 * popc(o & RANK_1);
  }
  PB_HD void pawn_capture(int, int to) { other += to < 8 ? 4 : 1; }
  PB_HD void en_passant(int, int) { ++other; }
  PB_HD void castle(int, int) { ++other; }
};

/* atomicAdd on the device (shared-memory cursor of a thread group); the host build (tests/native/hostcheck.cpp) is
 * single-threaded */
PB_HD u32 fetch_add_u32(u32* p, u32 v) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, v);
#else
  const u32 old = *p;
  *p += v;
  return old;
#endif
}

/* enumerate_split sink of K2 when every parent enumerates its moves ONCE: the slots of a piece's entries are reserved
 * with one shared-memory atomic on a packed cursor (clean entries grow from the front of the pool, the others from its
 * end), the inert moves are counted.  A reservation that does not fit raises `overflow`; the kernel then redoes the range
 * of parents in halves. */
struct DirectEmitter {
  u32* pool;
  u32* cursor;
  u32* overflow;
  u32 cap, tag; /* tag = parent << 16 */
  u64 zone;
  u32 inert = 0;
  PB_HD bool reserve(u32 n_clean, u32 n_other, u32& clean_at, u32& other_at) {
    const u32 old = fetch_add_u32(cursor, n_clean | (n_other << 16));
    const u32 c0 = old & 0xFFFF, o0 = old >> 16;
    if (c0 + n_clean + o0 + n_other > cap) {
      *overflow = 1;
      return false;
    }
    clean_at = c0, other_at = cap - o0 - n_other;
    return true;
  }
  PB_HD void put(u32& at, int from, int to, int promo, int kind) {
    if (PB_IN_BOUNDS(at, cap)) pool[at] = ((u32)kind << 28) | tag | pack_move(from, to, promo);
    ++at;
  }
  PB_HD void other(u32& at, int from, int to, int kind) {
    if (kind == KIND_PAWN && to < 8) { /* black promotes on rank 1 */
      put(at, from, to, PROMO_Q, kind), put(at, from, to, PROMO_R, kind);
      put(at, from, to, PROMO_B, kind), put(at, from, to, PROMO_N, kind);
    } else {
      put(at, from, to, 0, kind);
    }
  }
  PB_HD void piece(int from, int kind, u64 c, u64 o) {
    const u64 rest = c & ~inert_targets(zone, from, c);
    inert += popc(c ^ rest);
    if ((rest | o) == 0) return;
    u32 clean_at, other_at;
    if (!reserve(popc(rest), popc(o), clean_at, other_at)) return;
    for (Squares t(rest); t.any();) put(clean_at, from, t.pop(), 0, kind);
    for (Squares t(o); t.any();) put(other_at, from, t.pop(), 0, kind);
  }
  PB_HD void pushes(u64 c, u64 o, int delta) {
    const u64 rest = c & ~inert_pushes(zone, c, delta);
    inert += popc(c ^ rest);
    if ((rest | o) == 0) return;
    u32 clean_at, other_at;
    if (!reserve(popc(rest), popc(o) + 3 * popc(o & RANK_1), clean_at, other_at)) return;
    for (Squares t(rest); t.any();) {
      const int to = t.pop();
      put(clean_at, to + delta, to, 0, KIND_PAWN);
    }
    for (Squares t(o); t.any();) {
      const int to = t.pop();
      other(other_at, to + delta, to, KIND_PAWN);
    }
  }
  PB_HD void pawn_capture(int from, int to) {
    u32 clean_at, other_at;
    if (reserve(0, to < 8 ? 4 : 1, clean_at, other_at)) other(other_at, from, to, KIND_PAWN);
  }
  PB_HD void en_passant(int from, int to) {
    u32 clean_at, other_at;
    if (reserve(0, 1, clean_at, other_at)) put(other_at, from, to, 0, KIND_PAWN);
  }
  PB_HD void castle(int from, int to) {
    u32 clean_at, other_at;
    if (reserve(0, 1, clean_at, other_at)) put(other_at, from, to, 0, KIND_KING);
  }
};

/* Replies to a clean move: n = the normalised parent (black to move; meta carries base and both king
 * squares), the move is quiet and not a castling move. */
template <class TB>
PB_HD u32 count_child_fast(const Pos& n, const TB& tb, int from, int to, int kind) {
  const u64 m = bit(from) | bit(to);
  u64 pawns = n.pawns, knights = n.knights, bishops = n.bishops, rooks = n.rooks, queens = n.queens, kings = n.kings;
  const int king = (int)(n.meta >> META_OTHER_KING) & 63;
  int their_king = (int)(n.meta >> META_MOVER_KING) & 63;
  if (kind == KIND_PAWN) pawns ^= m;
  else if (kind == KIND_KNIGHT) knights ^= m;
  else if (kind == KIND_BISHOP) bishops ^= m;
  else if (kind == KIND_ROOK) rooks ^= m;
  else if (kind == KIND_QUEEN) queens ^= m;
  else kings ^= m, their_king = to; /* a plain king step (castling is never clean) */
  return ((u32)(n.meta >> META_BASE) & 0xFFFF) +
         pawn_king_part_white(n.white, pawns, knights, bishops, rooks, queens, kings, meta_castling(n.meta) & 3, king, their_king, tb);
}

} /* namespace pabi */
