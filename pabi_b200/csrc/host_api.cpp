/*
 * host_api.cpp -- the text side of pabi's Position API, kept on the host:
 * Position::starting (position.rs:78-91), from_fen (:157-275), TryFrom<&str>
 * (:809-821), Display (:823-860), validate (:934-1032) and Move text
 * conversion (core.rs:105-134).  No move generation happens here; the only
 * chess arithmetic is what validate() needs (who checks the side to move).
 * Error messages are the reference's, which its tests pin by substring
 * (tests/chess.rs:51-130).
 */
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <optional>
#include <string>
#include <string_view>

#include "record.h"
#include "tables.h"

using namespace pabi;

namespace {

struct ParseError {
  std::string message;
};

[[noreturn]] void bail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw ParseError{buf};
}

/* Rust `split(sep)`: empty pieces are kept, an empty input yields one empty piece. */
struct Split {
  std::string_view rest;
  char sep;
  bool done = false;
  std::optional<std::string_view> next() {
    if (done) return std::nullopt;
    const size_t at = rest.find(sep);
    std::string_view part = rest.substr(0, at);
    if (at == std::string_view::npos) done = true;
    else rest.remove_prefix(at + 1);
    return part;
  }
};

/* Rust `str::parse::<u8/u16>()`: optional '+', at least one ASCII digit, no overflow. */
bool parse_number(std::string_view s, unsigned long max, unsigned long& out) {
  if (!s.empty() && s.front() == '+') s.remove_prefix(1);
  if (s.empty()) return false;
  unsigned long v = 0;
  for (char c : s) {
    if (c < '0' || c > '9') return false;
    v = v * 10 + (unsigned long)(c - '0');
    if (v > max) return false;
  }
  out = v;
  return true;
}

constexpr char SYMBOLS[] = "KQRBNP"; /* bitboard.rs:350-357 order */

u64 side_occupancy(const u64 s[6]) { return s[0] | s[1] | s[2] | s[3] | s[4] | s[5]; }

/* Pieces of `their` side that attack `king` (the checkers of attacks.rs:101-200). */
u64 checkers_of(const u64 their[6], int their_colour, int king, u64 occ) {
  const SharedTables& t = host_tables().shared;
  const u64 kb = bit(king);
  u64 c = t.knight[king] & their[F_KNIGHTS];
  c |= pawn_attacks(kb, their_colour ^ 1) & their[F_PAWNS];
  c |= walk_bishop_attacks(king, occ) & (their[F_BISHOPS] | their[F_QUEENS]);
  c |= walk_rook_attacks(king, occ) & (their[F_ROOKS] | their[F_QUEENS]);
  return c;
}

/* position.rs:934-1032 */
void validate(const pabi_position_t& p) {
  if (p.fullmove_counter == 0) bail("fullmove counter cannot be zero");
  if (popc(p.white[F_KING]) != 1) bail("expected 1 white king, got %d", popc(p.white[F_KING]));
  if (popc(p.black[F_KING]) != 1) bail("expected 1 black king, got %d", popc(p.black[F_KING]));
  if (popc(p.white[F_PAWNS]) > 8) bail("expected <= 8 white pawns, got %d", popc(p.white[F_PAWNS]));
  if (popc(p.black[F_PAWNS]) > 8) bail("expected <= 8 black pawns, got %d", popc(p.black[F_PAWNS]));
  if ((p.white[F_PAWNS] | p.black[F_PAWNS]) & (RANK_1 | RANK_8)) bail("pawns can not be placed on backranks");
  const int us = p.side_to_move;
  const u64* our = us == 0 ? p.white : p.black;
  const u64* their = us == 0 ? p.black : p.white;
  const u64 occ = side_occupancy(p.white) | side_occupancy(p.black);
  const int king = lsb(our[F_KING]);
  const u64 checkers = checkers_of(their, us ^ 1, king, occ);
  if (popc(checkers) > 2) bail("expected <= 2 checks, got %d", popc(checkers));
  if (p.en_passant_square >= 64) return;
  const int ep = p.en_passant_square;
  const int expected_rank = us == 0 ? 6 : 3;
  if ((ep >> 3) + 1 != expected_rank)
    bail("expected en passant square to be on rank %d, got %d", expected_rank, (ep >> 3) + 1);
  const int pushed = us == 0 ? ep - 8 : ep + 8;
  if (!(their[F_PAWNS] & bit(pushed))) bail("en passant square is not beyond pushed pawn");
  const u64* rays = host_tables().rays;
  if (checkers) {
    if (popc(checkers) > 1) bail("more than 1 check after double pawn push is impossible");
    if (checkers != bit(pushed)) {
      const int origin = us == 0 ? ep + 8 : ep - 8;
      if (!(rays[lsb(checkers) * 64 + king] & bit(origin)))
        bail("the only possible checks after double pawn push are either discovery targeting the original pawn "
             "square or the pushed pawn itself");
    }
  }
  for (u64 b = their[F_QUEENS] | their[F_BISHOPS]; b; b &= b - 1) {
    const int attacker = lsb(b);
    const bool diagonal = ((attacker ^ king) & 7) != 0 && ((attacker ^ king) & 56) != 0;
    const u64 xray = diagonal ? rays[attacker * 64 + king] : 0; /* bishop_ray, attacks.rs:58-60 */
    if (popc(xray & occ) == 2 && (xray & bit(attacker)) && (xray & bit(pushed)))
      bail("doubly pushed pawn can not be the only blocker on a diagonal");
  }
}

/* position.rs:157-275 */
pabi_position_t from_fen(std::string_view input) {
  pabi_position_t p;
  std::memset(&p, 0, sizeof p);
  Split parts{input, ' '};
  const auto placement = parts.next(); /* never absent: split always yields one piece */
  Split ranks{*placement, '/'};
  int rank_id = 8;
  while (auto rank_fen = ranks.next()) {
    if (rank_id == 0) bail("expected 8 ranks, got %.*s", (int)placement->size(), placement->data());
    --rank_id;
    unsigned file = 0;
    for (unsigned char symbol : *rank_fen) {
      if (file > 8) bail("file exceeded 8");
      if (symbol == '0') bail("increment can not be 0");
      if (symbol >= '1' && symbol <= '9') {
        file += symbol - '0';
        continue;
      }
      const char* at = symbol < 128 && symbol ? std::strchr(SYMBOLS, std::toupper(symbol)) : nullptr;
      if (!at) bail("piece symbol should be in \"PNBRQKpnbrqk\", got '%c'", symbol);
      if (file > 7) bail("file should be within 0..BOARD_WIDTH, got %u", file);
      u64* side = (symbol >= 'a') ? p.black : p.white;
      side[at - SYMBOLS] |= bit(rank_id * 8 + (int)file);
      ++file;
    }
    if (file != 8)
      bail("rank size should be exactly 8, got %.*s of length %u", (int)rank_fen->size(), rank_fen->data(), file);
  }
  if (rank_id != 0) bail("there should be 8 ranks, got %.*s", (int)placement->size(), placement->data());

  const auto side = parts.next();
  if (!side) bail("missing side to move");
  if (*side == "w") p.side_to_move = 0;
  else if (*side == "b") p.side_to_move = 1;
  else bail("color should be 'w' or 'b', got '%.*s'", (int)side->size(), side->data()); /* environment.rs:29-39 */

  const auto castling = parts.next();
  if (!castling) bail("missing castling rights");
  { /* core.rs:634-673: "-" or an ordered, non-empty subset of KQkq */
    int rights = 0;
    bool ok = !castling->empty();
    if (*castling != "-") {
      static const char kRights[] = "KQkq";
      int last = -1;
      for (char c : *castling) {
        const char* at = c ? std::strchr(kRights, c) : nullptr;
        const int idx = at ? (int)(at - kRights) : -1;
        if (idx <= last) ok = false;
        else rights |= 1 << idx, last = idx;
        if (!ok) break;
      }
    }
    if (!ok) bail("unknown castle rights: %.*s", (int)castling->size(), castling->data());
    p.castling = (uint8_t)rights;
  }

  const auto ep = parts.next();
  if (!ep) bail("missing en passant square");
  if (*ep == "-") {
    p.en_passant_square = 64;
  } else { /* core.rs:286-299 */
    if (ep->size() != 2) bail("square should be two-char, got %.*s with %zu chars", (int)ep->size(), ep->data(), ep->size());
    const char f = (*ep)[0], r = (*ep)[1];
    if (f < 'a' || f > 'h') bail("file should be within 'a'..='h', got '%c'", f);
    if (r < '1' || r > '8') bail("rank should be within '1'..='8', got '%c'", r);
    p.en_passant_square = (uint8_t)((f - 'a') + 8 * (r - '1'));
  }

  unsigned long halfmove = 0, fullmove = 1;
  const auto half = parts.next();
  if (half && !parse_number(*half, 255, halfmove)) bail("halfmove clock can not be parsed %.*s", (int)half->size(), half->data());
  const auto full = parts.next();
  if (full) {
    if (!parse_number(*full, 65535, fullmove)) bail("fullmove counter can not be parsed %.*s", (int)full->size(), full->data());
    if (fullmove == 0) bail("fullmove counter can not be 0");
  } else if (half) {
    bail("if halfmove clock is present, fullmove counter must be present");
  }
  if (parts.next()) bail("trailing symbols");
  p.halfmove_clock = (uint8_t)halfmove;
  p.fullmove_counter = (uint16_t)fullmove;
  p.hash = host_compute_hash(pack_record(p)); /* position.rs:266; fixed keys, see chess.cuh */
  try {
    validate(p);
  } catch (ParseError& e) {
    throw ParseError{"illegal position: " + e.message};
  }
  return p;
}

int report(const ParseError& e, char* err, size_t cap) {
  if (err && cap) std::snprintf(err, cap, "%s", e.message.c_str());
  return -1;
}

bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

} /* namespace */

extern "C" {

void pabi_position_starting(pabi_position_t* out) {
  *out = from_fen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1");
}

int pabi_position_from_fen(const char* fen, pabi_position_t* out, char* err, size_t err_cap) {
  if (!fen || !out) return report(ParseError{"null argument"}, err, err_cap);
  try {
    *out = from_fen(fen);
    return 0;
  } catch (const ParseError& e) {
    return report(e, err, err_cap);
  }
}

int pabi_position_parse(const char* text, pabi_position_t* out, char* err, size_t err_cap) {
  if (!text || !out) return report(ParseError{"null argument"}, err, err_cap);
  std::string_view s(text);
  while (!s.empty() && is_space(s.front())) s.remove_prefix(1);
  while (!s.empty() && is_space(s.back())) s.remove_suffix(1);
  if (s.substr(0, 4) == "fen " || s.substr(0, 4) == "epd ") s.remove_prefix(4);
  try {
    *out = from_fen(s);
    return 0;
  } catch (const ParseError& e) {
    return report(e, err, err_cap);
  }
}

int pabi_position_to_fen(const pabi_position_t* p, char* out, size_t cap) {
  if (!p || !out) return -1;
  std::string s;
  for (int rank = 7; rank >= 0; rank--) {
    int empty = 0;
    for (int file = 0; file < 8; file++) {
      const u64 b = bit(rank * 8 + file);
      char c = 0;
      /* position.rs:755-769 looks at white first; bitboard.rs:428-446 at king, pawns, queens, rooks, bishops, knights */
      static const int ORDER[6] = {F_KING, F_PAWNS, F_QUEENS, F_ROOKS, F_BISHOPS, F_KNIGHTS};
      for (int side = 0; side < 2 && !c; side++)
        for (int k : ORDER)
          if ((side == 0 ? p->white : p->black)[k] & b) {
            c = side == 0 ? SYMBOLS[k] : (char)std::tolower(SYMBOLS[k]);
            break;
          }
      if (c) {
        if (empty) s += (char)('0' + empty), empty = 0;
        s += c;
      } else {
        ++empty;
      }
    }
    if (empty) s += (char)('0' + empty);
    if (rank) s += '/';
  }
  s += p->side_to_move == 0 ? " w " : " b ";
  if (p->castling & 15) {
    for (int i = 0; i < 4; i++)
      if (p->castling & (1 << i)) s += "KQkq"[i];
  } else {
    s += '-';
  }
  s += ' ';
  if (p->en_passant_square < 64) {
    s += (char)('a' + (p->en_passant_square & 7));
    s += (char)('1' + (p->en_passant_square >> 3));
  } else {
    s += '-';
  }
  s += ' ' + std::to_string(p->halfmove_clock) + ' ' + std::to_string(p->fullmove_counter);
  if (s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

void pabi_position_pack64(const pabi_position_t* p, uint64_t out[8]) {
  const Pos r = pack_record(*p);
  out[0] = r.white, out[1] = r.pawns, out[2] = r.knights, out[3] = r.bishops;
  out[4] = r.rooks, out[5] = r.queens, out[6] = r.kings, out[7] = r.meta;
}

void pabi_position_unpack64(const uint64_t in[8], pabi_position_t* out) {
  Pos r;
  r.white = in[0], r.pawns = in[1], r.knights = in[2], r.bishops = in[3];
  r.rooks = in[4], r.queens = in[5], r.kings = in[6], r.meta = in[7];
  *out = unpack_record(r, host_compute_hash(r)); /* the record carries no hash: it is recomputed as from_fen does */
}

uint64_t pabi_position_hash(const pabi_position_t* p) { return p ? p->hash : 0; }

uint64_t pabi_position_compute_hash(const pabi_position_t* p) { return p ? host_compute_hash(pack_record(*p)) : 0; }

uint32_t pabi_position_num_pieces(const pabi_position_t* p) {
  if (!p) return 0;
  return (uint32_t)popc(side_occupancy(p->white) | side_occupancy(p->black));
}

int pabi_position_halfmove_clock_expired(const pabi_position_t* p) { return p && p->halfmove_clock >= 100; }

pabi_move_t pabi_move_flip_perspective(pabi_move_t mv) { return (pabi_move_t)(mv ^ (56 | (56 << 6))); }

void pabi_zobrist_keys(uint64_t out[PABI_ZOBRIST_KEYS]) {
  static_assert(PABI_ZOBRIST_KEYS == ZOBRIST_KEYS, "key table layout");
  std::memcpy(out, host_tables().zobrist, sizeof(uint64_t) * ZOBRIST_KEYS);
}

int pabi_move_from_uci(const char* uci, pabi_move_t* out) {
  if (!uci || !out) return -1;
  const size_t n = std::strlen(uci);
  if (n != 4 && n != 5) return -1;
  int sq[2];
  for (int i = 0; i < 2; i++) {
    const char f = uci[2 * i], r = uci[2 * i + 1];
    if (f < 'a' || f > 'h' || r < '1' || r > '8') return -1;
    sq[i] = (f - 'a') + 8 * (r - '1');
  }
  int promo = 0;
  if (n == 5) { /* core.rs:707-717 */
    static const char kPromotions[] = "nbrq";
    const char* at = std::strchr(kPromotions, uci[4]);
    if (!at) return -1;
    promo = (int)(at - kPromotions) + 1;
  }
  *out = pack_move(sq[0], sq[1], promo);
  return 0;
}

int pabi_move_to_uci(pabi_move_t mv, char out[6]) {
  const int from = mv & 63, to = (mv >> 6) & 63, promo = (mv >> 12) & 7;
  int k = 0;
  out[k++] = (char)('a' + (from & 7)), out[k++] = (char)('1' + (from >> 3));
  out[k++] = (char)('a' + (to & 7)), out[k++] = (char)('1' + (to >> 3));
  if (promo >= 1 && promo <= 4) out[k++] = " nbrq"[promo];
  out[k] = 0;
  return k;
}

} /* extern "C" */
