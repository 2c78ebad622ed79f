/*
 * hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles the product's device
 * arithmetic (pabi_b200/csrc/chess.cuh, every function of which is
 * __host__ __device__) for the host CPU, so the CPU test-suite can compare the
 * exact logic the kernels inline with the oracle on millions of positions
 * without a GPU.  Never linked into libpabi_b200.so.
 */
#include "../../pabi_b200/csrc/fen.cuh"
#include "../../pabi_b200/csrc/incremental.cuh"
#include "../../pabi_b200/csrc/record.h"
#include "../../pabi_b200/csrc/tables.h"

#include <algorithm>

using namespace pabi;

static Tables host_tb() {
  const HostTables& h = host_tables();
  Tables t;
  t.s = &h.shared;
  t.g.rook_attacks = h.rook_attacks;
  t.g.rays = h.rays;
  t.g.zobrist = h.zobrist;
  return t;
}

extern "C" {

uint32_t hostcheck_count_moves(const pabi_position_t* p) { return count_moves(pack_record(*p), host_tb()); }

uint32_t hostcheck_generate_moves(const pabi_position_t* p, uint16_t* out) {
  uint32_t k = 0;
  return generate_moves(pack_record(*p), host_tb(), [&](int f, int t, int pr, int) { out[k++] = pack_move(f, t, pr); });
}

void hostcheck_make_move(pabi_position_t* p, uint16_t mv) {
  Pos r = pack_record(*p);
  const uint64_t hash = p->hash ^ hash_delta(r, host_tb(), mv); /* Position::hash as make_move maintains it */
  make_move(r, host_tb(), mv);
  *p = unpack_record(r, hash);
}

uint64_t hostcheck_compute_hash(const pabi_position_t* p) { return compute_hash(pack_record(*p), host_tb()); }

int hostcheck_in_check(const pabi_position_t* p) {
  Analysis a;
  analyze(pack_record(*p), host_tb(), a);
  return a.checkers != 0;
}

void hostcheck_analysis(const pabi_position_t* p, uint64_t out[5]) {
  Analysis a;
  analyze(pack_record(*p), host_tb(), a);
  out[0] = a.attacks, out[1] = a.checkers, out[2] = a.pins, out[3] = a.safe_king, out[4] = a.block;
}

void hostcheck_attack_info(const pabi_position_t* p, uint64_t out[5]) {
  const AttackInfoOut a = attack_info(pack_record(*p), host_tb());
  out[0] = a.attacks, out[1] = a.checkers, out[2] = a.pins, out[3] = a.xrays, out[4] = a.safe_king_squares;
}

/* the device FEN parser (fen.cuh) on the host */
int hostcheck_parse(const char* text, size_t len, pabi_position_t* out) {
  ParsedPosition p;
  const int rc = parse_position(text, len, host_tb(), p);
  __builtin_memset(out, 0, sizeof *out);
  if (rc == 0) {
    for (int k = 0; k < 6; k++) out->white[k] = p.white[k], out->black[k] = p.black[k];
    out->fullmove_counter = p.fullmove, out->castling = p.castling, out->side_to_move = p.stm;
    out->halfmove_clock = p.halfmove, out->en_passant_square = p.ep;
    out->hash = compute_hash(record_of(p), host_tb()); /* as k_parse_fens does */
  }
  return rc;
}

static uint64_t perft_rec(const Pos& p, const Tables& tb, int depth) {
  if (depth == 0) return 1;
  if (depth == 1) return count_moves(p, tb);
  uint16_t moves[256];
  int kinds[256];
  uint32_t k = 0;
  if (depth == 2) { /* the path of k_tile<LEAF>: parent mirrored to black-to-move, children counted as white */
    Pos b = meta_stm(p.meta) == 0 ? mirror(p) : p;
    b.meta = with_king_squares(b);
    generate_moves<1>(b, tb, [&](int f, int t, int pr, int kind) { kinds[k] = kind, moves[k++] = pack_move(f, t, pr); });
    if (k != count_moves(p, tb)) return ~0ULL;
    uint64_t n = 0;
    for (uint32_t i = 0; i < k; i++) {
      Pos c = b;
      make_move_kind<1, true>(c, tb, moves[i], kinds[i]);
      n += count_child_white(c, tb);
    }
    return n;
  }
  generate_moves(p, tb, [&](int f, int t, int pr, int kind) { kinds[k] = kind, moves[k++] = pack_move(f, t, pr); });
  uint64_t n = 0;
  for (uint32_t i = 0; i < k; i++) {
    Pos c = p;
    make_move_kind(c, tb, moves[i], kinds[i]);
    n += perft_rec(c, tb, depth - 1);
  }
  return n;
}
uint64_t hostcheck_perft(const pabi_position_t* p, int depth) { return perft_rec(pack_record(*p), host_tb(), depth); }

/* Walks the perft tree of `root` to `depth` comparing, at every node, the
 * product logic with callbacks into the oracle (passed as function pointers so
 * this file does not link the oracle).  Returns the number of mismatching nodes;
 * first_bad receives the first offending position. */
typedef uint32_t (*oracle_gen_fn)(const pabi_position_t*, uint16_t*);
typedef void (*oracle_make_fn)(pabi_position_t*, uint16_t);

static uint64_t walk_compare(const pabi_position_t& p, int depth, oracle_gen_fn gen, oracle_make_fn mk,
                             pabi_position_t* first_bad, uint64_t* visited) {
  uint16_t want[256], got[256];
  ++*visited;
  const uint32_t nw = gen(&p, want);
  const Tables tb = host_tb();
  const Pos r = pack_record(p);
  uint32_t k = 0;
  int kinds[256];
  const uint32_t ng = generate_moves(r, tb, [&](int f, int t, int pr, int kind) { kinds[k] = kind, got[k++] = pack_move(f, t, pr); });
  uint64_t bad = 0;
  if (ng != nw || count_moves(r, tb) != nw) bad = 1;
  for (uint32_t i = 0; i < nw && !bad; i++)
    if (want[i] != got[i]) bad = 1;
  if (bad) {
    if (first_bad) *first_bad = p;
    return 1;
  }
  if (depth == 0) return 0;
  for (uint32_t i = 0; i < nw; i++) {
    pabi_position_t a = p, b = p;
    mk(&a, want[i]);
    Pos c = r;
    make_move(c, tb, want[i]);
    b = unpack_record(c, p.hash ^ hash_delta(r, tb, want[i])); /* all 112 bytes are compared: the hash is the product's own */
    Pos ck = r; /* the counting path's make_move: same board, rights, ep square, side to move */
    make_move_kind(ck, tb, want[i], kinds[i]);
    Pos cc = r; /* K1's make_move: the same with the clocks, i.e. the whole record */
    make_move_kind<-1, false, true>(cc, tb, want[i], kinds[i]);
    const bool same_kind = ck.white == c.white && ck.pawns == c.pawns && ck.knights == c.knights && ck.bishops == c.bishops &&
                           ck.rooks == c.rooks && ck.queens == c.queens && ck.kings == c.kings &&
                           (ck.meta & 0xFFFF) == (c.meta & 0xFFFF) && __builtin_memcmp(&cc, &c, sizeof c) == 0;
    if (__builtin_memcmp(&a, &b, sizeof a) != 0 || !same_kind) {
      if (first_bad) *first_bad = p;
      return bad + 1;
    }
    bad += walk_compare(a, depth - 1, gen, mk, first_bad, visited);
    if (bad) return bad;
  }
  return bad;
}

uint64_t hostcheck_walk_compare(const pabi_position_t* root, int depth, oracle_gen_fn gen, oracle_make_fn mk,
                                pabi_position_t* first_bad, uint64_t* visited) {
  uint64_t v = 0;
  uint64_t bad = walk_compare(*root, depth, gen, mk, first_bad, &v);
  if (visited) *visited = v;
  return bad;
}

/* incremental.cuh: for every node of the tree as a parent, every move classified clean must have
 * fast count == full count, the split must cover exactly the generated moves, and the counting sink
 * must agree with what is enumerated; every inert move (clean and off the parent's zone) must have exactly
 * TContext::total replies.  out[0] parents, out[1] moves, out[2] clean moves (inert included), out[3] mismatches,
 * out[4] inert moves. */
struct CheckSink {
  const Pos& n;
  const Tables& tb;
  uint64_t* out;
  const TContext& ctx;
  uint32_t clean = 0, other = 0, inert = 0;
  void clean_move(int from, int to, int kind, bool is_inert) {
    ++clean;
    moves[k++] = pack_move(from, to, 0);
    Pos ch = n;
    make_move_kind<1, true>(ch, tb, pack_move(from, to, 0), kind);
    const uint32_t want = count_child_white(ch, tb);
    if (count_child_fast(n, tb, from, to, kind) != want) ++out[3];
    if (is_inert) {
      ++inert;
      if (ctx.total != want) ++out[3];
    }
  }
  uint16_t moves[512];
  uint32_t k = 0;
  void full(int from, int to, int promo, int kind) {
    ++other;
    moves[k++] = pack_move(from, to, promo);
    (void)kind;
  }
  void piece(int from, int kind, u64 c, u64 o) {
    const u64 in = inert_targets(ctx.zone, from, c);
    for (Squares t(c); t.any();) {
      const int to = t.pop();
      clean_move(from, to, kind, (in >> to) & 1);
    }
    for (Squares t(o); t.any();) full(from, t.pop(), 0, kind);
  }
  void pushes(u64 c, u64 o, int delta) {
    const u64 in = inert_pushes(ctx.zone, c, delta);
    for (Squares t(c); t.any();) {
      const int to = t.pop();
      clean_move(to + delta, to, KIND_PAWN, (in >> to) & 1);
    }
    for (Squares t(o); t.any();) {
      const int to = t.pop();
      if (to < 8) for (int pr = 4; pr >= 1; pr--) full(to + delta, to, pr, KIND_PAWN);
      else full(to + delta, to, 0, KIND_PAWN);
    }
  }
  void pawn_capture(int from, int to) {
    if (to < 8) for (int pr = 4; pr >= 1; pr--) full(from, to, pr, KIND_PAWN);
    else full(from, to, 0, KIND_PAWN);
  }
  void en_passant(int from, int to) { full(from, to, 0, KIND_PAWN); }
  void castle(int from, int to) { full(from, to, 0, KIND_KING); }
};

/* what K2's SplitEmitter writes to the pool: every move except the inert ones */
struct PoolSink {
  u64 zone;
  uint32_t k = 0;
  uint32_t entries[1024];
  void put(int from, int to, int promo, int kind, int cls) { entries[k++] = (uint32_t)cls << 31 | (uint32_t)kind << 28 | pack_move(from, to, promo); }
  void other(int from, int to, int kind) {
    if (kind == KIND_PAWN && to < 8) for (int pr = 4; pr >= 1; pr--) put(from, to, pr, kind, 1);
    else put(from, to, 0, kind, 1);
  }
  void piece(int from, int kind, u64 c, u64 o) {
    for (Squares t(c & ~inert_targets(zone, from, c)); t.any();) put(from, t.pop(), 0, kind, 0);
    for (Squares t(o); t.any();) put(from, t.pop(), 0, kind, 1);
  }
  void pushes(u64 c, u64 o, int delta) {
    for (Squares t(c & ~inert_pushes(zone, c, delta)); t.any();) { const int to = t.pop(); put(to + delta, to, 0, KIND_PAWN, 0); }
    for (Squares t(o); t.any();) { const int to = t.pop(); other(to + delta, to, KIND_PAWN); }
  }
  void pawn_capture(int from, int to) { other(from, to, KIND_PAWN); }
  void en_passant(int from, int to) { put(from, to, 0, KIND_PAWN, 1); }
  void castle(int from, int to) { put(from, to, 0, KIND_KING, 1); }
};

static void incremental_walk(const Pos& p, const Tables& tb, int depth, uint64_t* out) {
  Pos n = meta_stm(p.meta) == 0 ? mirror(p) : p;
  const TContext ctx = t_context(n, tb);
  n.meta = (with_king_squares(n) & ~(0xFFFFULL << META_BASE)) | ((u64)ctx.base << META_BASE);
  CheckSink sink{n, tb, out, ctx};
  enumerate_split(n, tb, ctx.reach, ctx.guard, sink);
  SplitCounter counter{ctx.zone};
  enumerate_split(n, tb, ctx.reach, ctx.guard, counter);
  ++out[0], out[1] += sink.k, out[2] += sink.clean, out[4] += sink.inert;
  if (counter.clean + counter.inert != sink.clean || counter.other != sink.other || counter.inert != sink.inert) ++out[3];
  { /* what goes into the pool is what the counter says */
    PoolSink all{ctx.zone};
    enumerate_split(n, tb, ctx.reach, ctx.guard, all);
    if (all.k != counter.clean + counter.other) ++out[3];
    /* K2's own sink (DirectEmitter, incremental.cuh): clean entries from the front of the pool, the others from its end,
     * every entry tagged with the parent, the same set of entries as the PoolSink above, the inert moves counted */
    constexpr u32 CAP = 1024, TAG = 0x155u << 16;
    static thread_local u32 pool[CAP + 2];
    u32 cursor = 0, overflow = 0;
    {
      DirectEmitter emit{pool + 1, &cursor, &overflow, CAP, TAG, ctx.zone};
      pool[0] = pool[CAP + 1] = 0xDEADBEEFu;
      enumerate_split(n, tb, ctx.reach, ctx.guard, emit);
      const u32 nc = cursor & 0xFFFF, no = cursor >> 16;
      if (overflow || nc != counter.clean || no != counter.other || emit.inert != counter.inert || pool[0] != 0xDEADBEEFu ||
          pool[CAP + 1] != 0xDEADBEEFu) {
        ++out[3];
      } else {
        u32 got[1024], want[1024];
        for (u32 i = 0; i < nc; i++) got[i] = pool[1 + i];
        for (u32 i = 0; i < no; i++) got[nc + i] = pool[1 + CAP - no + i] | 0x80000000u;
        bool ok = true;
        for (u32 i = 0; i < nc + no; i++) ok &= (got[i] & 0x0FFF0000u) == TAG, got[i] &= ~TAG;
        for (u32 i = 0; i < all.k; i++) want[i] = all.entries[i];
        std::sort(got, got + nc + no), std::sort(want, want + all.k);
        for (u32 i = 0; i < all.k; i++) ok &= got[i] == want[i];
        if (!ok) ++out[3];
      }
    }
    /* a pool one entry too small: the overflow flag is raised and nothing is written outside it */
    if (all.k > 1) {
      const u32 cap = all.k - 1;
      cursor = 0, overflow = 0;
      DirectEmitter emit{pool + 1, &cursor, &overflow, cap, TAG, ctx.zone};
      pool[0] = pool[cap + 1] = 0xDEADBEEFu;
      enumerate_split(n, tb, ctx.reach, ctx.guard, emit);
      if (!overflow || pool[0] != 0xDEADBEEFu || pool[cap + 1] != 0xDEADBEEFu) ++out[3];
    }
  }
  /* the split covers exactly the legal moves */
  uint16_t want[256];
  uint32_t nw = 0;
  generate_moves<1>(n, tb, [&](int f, int t, int pr, int) { want[nw++] = pack_move(f, t, pr); });
  if (nw != sink.k) ++out[3];
  else {
    std::sort(want, want + nw);
    std::sort(sink.moves, sink.moves + sink.k);
    for (uint32_t i = 0; i < nw; i++)
      if (want[i] != sink.moves[i]) { ++out[3]; break; }
  }
  if (depth == 0) return;
  uint16_t mv[256];
  int kinds[256];
  uint32_t k = 0;
  generate_moves(p, tb, [&](int f, int t, int pr, int kind) { kinds[k] = kind, mv[k++] = pack_move(f, t, pr); });
  for (uint32_t i = 0; i < k; i++) {
    Pos c = p;
    make_move(c, tb, mv[i]);
    incremental_walk(c, tb, depth - 1, out);
  }
}

void hostcheck_incremental(const pabi_position_t* root, int depth, uint64_t out[5]) {
  out[0] = out[1] = out[2] = out[3] = out[4] = 0;
  incremental_walk(pack_record(*root), host_tb(), depth, out);
}

/* Diagnostic for DESIGN.md: why moves miss the fast path.  For every node of the tree as a K2 parent, every legal
 * move falls in one class: out[0] inert, [1] clean, [2] parent ineligible (T in check / pinned T pawn),
 * [3] capture, [4] promotion / en passant / castling, [5] leaves the guard (from on a T slider line or king line),
 * [6] lands on the guard or reach, [7] gives check by knight/pawn or creates an en passant square, [8] parents; [9] / [10] = the part of [5] / [6]
 * that is on a T slider's attack set (the rest is on a king line only). */
static void split_stats_walk(const Pos& p, const Tables& tb, int depth, uint64_t* out) {
  Pos n = meta_stm(p.meta) == 0 ? mirror(p) : p;
  const TContext ctx = t_context(n, tb);
  ++out[8];
  SplitCounter counter{ctx.zone};
  enumerate_split(n, tb, ctx.reach, ctx.guard, counter);
  out[0] += counter.inert, out[1] += counter.clean;
  const u64 W = n.white;
  generate_moves<1>(n, tb, [&](int f, int t, int pr, int kind) {
    bool clean = false;
    { /* is it clean?  ask the splitter through a one-move sink */
      struct One { int f, t; bool hit = false;
        void piece(int from, int, u64 c, u64) { if (from == f && ((c >> t) & 1)) hit = true; }
        void pushes(u64 c, u64, int d) { if (f == t + d && ((c >> t) & 1)) hit = true; }
        void pawn_capture(int, int) {} void en_passant(int, int) {} void castle(int, int) {} } one{f, t};
      enumerate_split(n, tb, ctx.reach, ctx.guard, one);
      clean = one.hit && pr == 0;
    }
    if (clean) return;
    const bool castle = kind == KIND_KING && (f - t == 2 || t - f == 2);
    const bool ep = kind == KIND_PAWN && t == meta_ep(n.meta) && ((f ^ t) & 7);
    if (ctx.reach == ~0ULL) ++out[2];
    else if ((W >> t) & 1) ++out[3];
    else if (pr || ep || castle) ++out[4];
    else if ((ctx.reach >> f) & 1) ++out[5], out[9] += (ctx.reach >> f) & 1;
    else if (((kind == KIND_BISHOP || kind == KIND_ROOK || kind == KIND_QUEEN ? ctx.guard : ctx.reach) >> t) & 1) ++out[6], out[10] += (ctx.reach >> t) & 1;
    else ++out[7];
  });
  if (depth == 0) return;
  uint16_t mv[256];
  uint32_t k = 0;
  generate_moves(p, tb, [&](int f, int t, int pr, int) { mv[k++] = pack_move(f, t, pr); });
  for (uint32_t i = 0; i < k; i++) {
    Pos c = p;
    make_move(c, tb, mv[i]);
    split_stats_walk(c, tb, depth - 1, out);
  }
}
void hostcheck_split_stats(const pabi_position_t* root, int depth, uint64_t out[11]) {
  for (int i = 0; i < 11; i++) out[i] = 0;
  split_stats_walk(pack_record(*root), host_tb(), depth, out);
}

/* table access for tests/test_tables*.py */
const uint64_t* hostcheck_table(const char* name, size_t* count) {
  const HostTables& h = host_tables();
  struct { const char* n; const uint64_t* p; size_t c; } T[] = {
      {"knight", h.shared.knight, 64}, {"king", h.shared.king, 64},
      {"bishop_attacks", h.shared.bishop_attacks, BISHOP_TABLE_SIZE},
      {"rook_attacks", h.rook_attacks, ROOK_TABLE_SIZE}, {"rays", h.rays, 4096},
      {"diag", h.shared.diag, 64}, {"anti", h.shared.anti, 64}};
  for (auto& e : T)
    if (!__builtin_strcmp(e.n, name)) { *count = e.c; return e.p; }
  return nullptr;
}
uint64_t hostcheck_bishop(int sq, uint64_t occ) { return host_tb().bishop(sq, occ); }
uint64_t hostcheck_rook(int sq, uint64_t occ) { return host_tb().rook(sq, occ); }
void hostcheck_slider_entry(int rook, int sq, uint64_t* mask, uint32_t* offset, uint32_t* bits) {
  const SharedTables& s = host_tables().shared;
  *mask = rook ? s.rook[sq].mask : s.bishop[sq].mask;
  const SliderIndex os = rook ? s.rook_os[sq] : s.bishop_os[sq];
  *offset = os.offset, *bits = 32 - os.shift;
}
}
