// bb_rules.cuh — Block Blast rules on 64-bit bitboards, shared by the sm_100a kernels and by
// a host (g++) build used only to fuzz this file against the oracle without a GPU.
//
// Bit convention: bit = row*8 + col (same as the action codec a % 64,
// reference src/environment/block_blast_env.py:114-118).  A piece is its cell mask with the
// top-left of its bounding box at bit 0 plus the mask of anchors that keep it on the board.
//
// What each function restates (file:line in the reference):
//   bb_valid      Board.can_place for all 64 anchors at once      src/game/board.py:71-93, :117-142
//   bb_clear      find_complete_lines + clear_lines               src/game/board.py:144-193
//   bb_holes      Board.count_holes                               src/game/board.py:195-216
//   bb_center     filled cells of rows/cols 2..5                  src/game/board.py:236-243
//   bb_solvable   _can_place_all_pieces (boolean; order is free)  src/game/engine.py:174-238
//   bb_score_gain _calculate_score with post-increment streak     src/game/engine.py:240-312, :419-429
//   bb_reward     _calculate_reward, float64, same op order       src/environment/block_blast_env.py:148-193
//   bb_env_apply  BlockBlastEnv.step + make_move + auto-reset     block_blast_env.py:224-264, engine.py:390-454,
//                                                                 src/environment/wrappers.py:93-108
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#define BB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define BB_HD inline
#define BB_HD_NOINLINE static
#endif

#include "bb_piece_table.inc"

// ---------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------
struct BBTables {
    uint64_t mask[BB_NUM_PIECES + 3];   // piece cells at the origin (padded: slots 37..39 = 0)
    uint64_t inb[BB_NUM_PIECES + 3];    // anchors whose bounding box stays on the board
    uint32_t meta[BB_NUM_PIECES + 3];   // n | h<<4 | w<<8 | maxrow<<12 | maxcol<<16
};

#define BB_META_N(m) ((m) & 0xFu)
#define BB_META_MAXROW(m) (((m) >> 12) & 0xFu)
#define BB_META_MAXCOL(m) (((m) >> 16) & 0xFu)

static const uint64_t BB_HOST_PIECE_MASKS[BB_NUM_PIECES] = BB_PIECE_MASKS;
static const uint64_t BB_HOST_PIECE_INB[BB_NUM_PIECES] = BB_PIECE_INB;
static const uint32_t BB_HOST_PIECE_META[BB_NUM_PIECES] = BB_PIECE_META;

inline void bb_fill_tables(BBTables* t) {
    for (int i = 0; i < BB_NUM_PIECES + 3; ++i) {
        t->mask[i] = i < BB_NUM_PIECES ? BB_HOST_PIECE_MASKS[i] : 0;
        t->inb[i] = i < BB_NUM_PIECES ? BB_HOST_PIECE_INB[i] : 0;
        t->meta[i] = i < BB_NUM_PIECES ? BB_HOST_PIECE_META[i] : 0;
    }
}

// ---------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------
BB_HD int bb_popc(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
BB_HD int bb_ctz(uint64_t x) {   // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
BB_HD uint32_t bb_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#define BB_COL_A 0x0101010101010101ull
#define BB_COL_H 0x8080808080808080ull
#define BB_CENTER 0x00003C3C3C3C0000ull

// ---------------------------------------------------------------------------------------
// board primitives
// ---------------------------------------------------------------------------------------
// All anchors at which the piece (cells pm, in-bounds anchors inb) fits on the EMPTY-cell
// set e = ~board.  board.py:71-93 evaluated for 64 anchors with one shift-AND per block.
BB_HD uint64_t bb_valid(uint64_t e, uint64_t pm, uint64_t inb) {
    uint64_t v = inb;
    while (pm) {                    // one shift-AND per block of the piece (<= 9)
        const int o = bb_ctz(pm);
        v &= e >> o;
        pm &= pm - 1;
    }
    return v;
}

// Full rows / columns are detected on the same board, then all removed (board.py:166-193).
// Returns the cleared board; *lines = rows + cols.
BB_HD uint64_t bb_clear(uint64_t b, int* lines) {
    uint64_t r = b & (b >> 4);
    r &= r >> 2;
    r &= r >> 1;
    r &= BB_COL_A;                  // bit 8*row set iff row full
    uint64_t c = b & (b >> 32);
    c &= c >> 16;
    c &= c >> 8;
    c &= 0xFFull;                   // bit col set iff column full
    *lines = bb_popc(r) + bb_popc(c);
    return b & ~((r * 0xFFull) | (c * BB_COL_A));
}

// true iff b has at least one full row or column
BB_HD bool bb_any_full(uint64_t b) {
    uint64_t r = b & (b >> 4);
    r &= r >> 2;
    r &= r >> 1;
    uint64_t c = b & (b >> 32);
    c &= c >> 16;
    c &= c >> 8;
    return ((r & BB_COL_A) | (c & 0xFFull)) != 0;
}

BB_HD uint64_t bb_clear_only(uint64_t b) {
    int l;
    return bb_clear(b, &l);
}

// board.py:195-216: empty cells whose four neighbours are filled or off-board.
BB_HD int bb_holes(uint64_t b) {
    const uint64_t up = (b << 8) | 0xFFull;                      // neighbour in row-1 (row 0: off-board)
    const uint64_t dn = (b >> 8) | 0xFF00000000000000ull;        // neighbour in row+1
    const uint64_t lf = ((b << 1) & ~BB_COL_A) | BB_COL_A;       // neighbour in col-1
    const uint64_t rt = ((b >> 1) & ~BB_COL_H) | BB_COL_H;       // neighbour in col+1
    return bb_popc(~b & up & dn & lf & rt);
}

BB_HD int bb_center(uint64_t b) { return bb_popc(b & BB_CENTER); }

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  Stream layout documented in philox.py.
// ---------------------------------------------------------------------------------------
struct BBPhilox4 { uint32_t x, y, z, w; };

BB_HD BBPhilox4 bb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = bb_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = bb_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    BBPhilox4 r = {c0, c1, c2, c3};
    return r;
}

#define BB_STREAM_TRIO 0u
#define BB_STREAM_POLICY 1u
#define BB_STREAM_SAMPLE 2u

// candidate trio number `draw` of global env `env_id`: three ids in [0,37), packed in bytes 0..2
BB_HD uint32_t bb_draw_trio(uint64_t seed, uint64_t env_id, uint32_t draw) {
    const BBPhilox4 r = bb_philox((uint32_t)env_id, (uint32_t)(env_id >> 32), draw, BB_STREAM_TRIO,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
    return bb_mulhi(r.x, 37u) | (bb_mulhi(r.y, 37u) << 8) | (bb_mulhi(r.z, 37u) << 16);
}

// ---------------------------------------------------------------------------------------
// trio solvability (engine.py:174-238).  The reference enumerates piece orders and anchors
// depth-first with line clears after every simulated placement and returns a boolean, so any
// sound and complete search gives the identical result.  Facts used to prune:
//   (M) clearing lines only removes cells, so a placement that fits before a clear fits after.
//   (P) hence if the three pieces have pairwise-disjoint placements that all fit on the
//       current board, every order works ("packing"), no simulation needed;
//   (C) conversely a solution that is NOT a packing must complete a line with its first or
//       second placement; a line with `miss` empty cells can only be completed by pieces that
//       can put at least `miss` cells into one row (column): meta maxrow / maxcol.
// ---------------------------------------------------------------------------------------
#ifdef BB_COUNT_WORK
struct BBWork { long long valid_calls, fast_accept, fast_reject, pack_iters, clear_iters, slow; };
static BBWork g_bb_work;
#define BB_WORK(f, n) (g_bb_work.f += (n))
#else
#define BB_WORK(f, n) ((void)0)
#endif

// smallest number of empty cells in any row (*row_miss) and in any column (*col_miss) of b
BB_HD void bb_min_missing(uint64_t b, int* row_miss, int* col_miss) {
    int rmax = 0, cmax = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = bb_popc(b & (0xFFull << (8 * i)));
        const int c = bb_popc(b & (BB_COL_A << i));
        rmax = r > rmax ? r : rmax;
        cmax = c > cmax ? c : cmax;
    }
    *row_miss = 8 - rmax;
    *col_miss = 8 - cmax;
}

struct BBTrio { uint64_t pm[3], inb[3]; uint32_t meta[3]; };

BB_HD void bb_load_trio(const BBTables* T, uint32_t trio, BBTrio* t) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t id = (trio >> (8 * i)) & 0xFFu;
        t->pm[i] = T->mask[id];
        t->inb[i] = T->inb[id];
        t->meta[i] = T->meta[id];
    }
}

// exists an anchor for piece y on b1 such that, afterwards (with clears), piece z still fits?
// "pack2": z fits beside y without needing a clear.
BB_HD bool bb_pack2(uint64_t b, uint64_t pmy, uint64_t inby, uint64_t pmz, uint64_t inbz) {
    uint64_t vy = bb_valid(~b, pmy, inby);
    BB_WORK(valid_calls, 1);
    while (vy) {
        const int a = bb_ctz(vy);
        vy &= vy - 1;
        BB_WORK(pack_iters, 1);
        BB_WORK(valid_calls, 1);
        if (bb_valid(~(b | (pmy << a)), pmz, inbz)) return true;
    }
    return false;
}

// exists a CLEARING placement of x on b after which y fits (one order only)
BB_HD bool bb_clear_then_fit(uint64_t b, uint64_t pmx, uint64_t inbx, uint64_t pmy, uint64_t inby) {
    uint64_t vx = bb_valid(~b, pmx, inbx);
    BB_WORK(valid_calls, 1);
    while (vx) {
        const int a = bb_ctz(vx);
        vx &= vx - 1;
        const uint64_t b1 = b | (pmx << a);
        BB_WORK(clear_iters, 1);
        if (bb_any_full(b1)) {
            BB_WORK(valid_calls, 1);
            if (bb_valid(~bb_clear_only(b1), pmy, inby)) return true;
        }
    }
    return false;
}

// two pieces on board b, any order, clears simulated: engine.py:181-224 at depth 1
BB_HD bool bb_solve2(uint64_t b, const BBTrio* t, int j, int k, bool skip_pack) {
    int rm, cm;
    if (!skip_pack && bb_pack2(b, t->pm[j], t->inb[j], t->pm[k], t->inb[k])) return true;
    bb_min_missing(b, &rm, &cm);
    const int mrj = (int)BB_META_MAXROW(t->meta[j]), mcj = (int)BB_META_MAXCOL(t->meta[j]);
    const int mrk = (int)BB_META_MAXROW(t->meta[k]), mck = (int)BB_META_MAXCOL(t->meta[k]);
    if ((rm <= mrj || cm <= mcj) && bb_clear_then_fit(b, t->pm[j], t->inb[j], t->pm[k], t->inb[k])) return true;
    if ((rm <= mrk || cm <= mck) && bb_clear_then_fit(b, t->pm[k], t->inb[k], t->pm[j], t->inb[j])) return true;
    return false;
}

enum { BB_REJECT = 0, BB_ACCEPT = 1, BB_HARD = 2 };

// Cheap classification of (board, trio): ACCEPT / REJECT when provable in O(1) valid-mask
// evaluations, HARD otherwise.  v[] receives the three valid-anchor masks on b.
BB_HD int bb_solvable_fast(uint64_t b, const BBTrio* t, uint64_t v[3]) {
    const uint64_t e = ~b;
    v[0] = bb_valid(e, t->pm[0], t->inb[0]);
    v[1] = bb_valid(e, t->pm[1], t->inb[1]);
    v[2] = bb_valid(e, t->pm[2], t->inb[2]);
    BB_WORK(valid_calls, 3);
    if (v[0] && v[1] && v[2]) {
        // greedy packing: the piece with the fewest anchors first, lowest anchors
        const int n0 = bb_popc(v[0]), n1 = bb_popc(v[1]), n2 = bb_popc(v[2]);
        int x = 0, y = 1, z = 2;
        if (n1 < n0 && n1 <= n2) { x = 1; y = 0; }
        else if (n2 < n0 && n2 < n1) { x = 2; z = 0; }
        // y before z: fewer anchors first
        const int ny = (y == 0 ? n0 : (y == 1 ? n1 : n2)), nz = (z == 0 ? n0 : (z == 1 ? n1 : n2));
        if (nz < ny) { const int s = y; y = z; z = s; }
        const uint64_t b1 = b | (t->pm[x] << bb_ctz(v[x]));
        const uint64_t vy = bb_valid(~b1, t->pm[y], t->inb[y]);
        BB_WORK(valid_calls, 1);
        if (vy) {
            const uint64_t b2 = b1 | (t->pm[y] << bb_ctz(vy));
            BB_WORK(valid_calls, 1);
            if (bb_valid(~b2, t->pm[z], t->inb[z])) return BB_ACCEPT;
        }
        return BB_HARD;
    }
    // some piece has no anchor now: only a line clear can make room (fact M/C)
    int rm, cm;
    bb_min_missing(b, &rm, &cm);
    // the two largest per-line contributions among the three pieces
    int r0 = (int)BB_META_MAXROW(t->meta[0]), r1 = (int)BB_META_MAXROW(t->meta[1]), r2 = (int)BB_META_MAXROW(t->meta[2]);
    int c0 = (int)BB_META_MAXCOL(t->meta[0]), c1 = (int)BB_META_MAXCOL(t->meta[1]), c2 = (int)BB_META_MAXCOL(t->meta[2]);
    const int rmin = r0 < r1 ? (r0 < r2 ? r0 : r2) : (r1 < r2 ? r1 : r2);
    const int cmin = c0 < c1 ? (c0 < c2 ? c0 : c2) : (c1 < c2 ? c1 : c2);
    const int rtop2 = r0 + r1 + r2 - rmin, ctop2 = c0 + c1 + c2 - cmin;
    if (rm > rtop2 && cm > ctop2) return BB_REJECT;
    return BB_HARD;
}

// Exact search for the cases bb_solvable_fast leaves open.  v[] = valid masks on b.
BB_HD_NOINLINE bool bb_solvable_slow(uint64_t b, const BBTrio* t, const uint64_t v[3]) {
    BB_WORK(slow, 1);
    // Stage A (fact P): a packing of all three.
    if (v[0] && v[1] && v[2]) {
        const int n0 = bb_popc(v[0]), n1 = bb_popc(v[1]), n2 = bb_popc(v[2]);
        int x = 0, y = 1, z = 2;
        if (n1 < n0 && n1 <= n2) { x = 1; y = 0; }
        else if (n2 < n0 && n2 < n1) { x = 2; z = 0; }
        uint64_t vx = v[x];
        while (vx) {
            const int a = bb_ctz(vx);
            vx &= vx - 1;
            const uint64_t b1 = b | (t->pm[x] << a);
            // z must still fit on b1 at all, else no packing through this anchor
            BB_WORK(valid_calls, 1);
            if (!bb_valid(~b1, t->pm[z], t->inb[z])) continue;
            if (bb_pack2(b1, t->pm[y], t->inb[y], t->pm[z], t->inb[z])) return true;
        }
    }
    // Stage B (fact C): a solution must clear a line with its 1st or 2nd placement.
    int rm, cm;
    bb_min_missing(b, &rm, &cm);
    {
        int r0 = (int)BB_META_MAXROW(t->meta[0]), r1 = (int)BB_META_MAXROW(t->meta[1]), r2 = (int)BB_META_MAXROW(t->meta[2]);
        int c0 = (int)BB_META_MAXCOL(t->meta[0]), c1 = (int)BB_META_MAXCOL(t->meta[1]), c2 = (int)BB_META_MAXCOL(t->meta[2]);
        const int rmin = r0 < r1 ? (r0 < r2 ? r0 : r2) : (r1 < r2 ? r1 : r2);
        const int cmin = c0 < c1 ? (c0 < c2 ? c0 : c2) : (c1 < c2 ? c1 : c2);
        if (rm > r0 + r1 + r2 - rmin && cm > c0 + c1 + c2 - cmin) return false;
    }
    for (int i = 0; i < 3; ++i) {
        const int j = i == 0 ? 1 : 0, k = i == 2 ? 1 : 2;
        uint64_t vi = v[i];
        while (vi) {
            const int a = bb_ctz(vi);
            vi &= vi - 1;
            const uint64_t b1 = b | (t->pm[i] << a);
            if (bb_any_full(b1)) {
                if (bb_solve2(bb_clear_only(b1), t, j, k, false)) return true;
            } else {
                // no clear yet: a packing of the other two on b1 would have been found in
                // stage A, so only a clearing second placement can help
                if (bb_solve2(b1, t, j, k, true)) return true;
            }
        }
    }
    return false;
}

BB_HD bool bb_solvable(uint64_t b, const BBTables* T, uint32_t trio) {
    BBTrio t;
    bb_load_trio(T, trio, &t);
    uint64_t v[3];
    const int f = bb_solvable_fast(b, &t, v);
    if (f == BB_ACCEPT) { BB_WORK(fast_accept, 1); return true; }
    if (f == BB_REJECT) { BB_WORK(fast_reject, 1); return false; }
    return bb_solvable_slow(b, &t, v);
}

// ---------------------------------------------------------------------------------------
// per-env state: 48 bytes = three 16-byte words (SoA of uint4 in HBM)
// ---------------------------------------------------------------------------------------
struct BBState {
    uint64_t board;
    uint32_t pieces;       // bytes 0..2 piece ids, byte 3 = used bits (bit i = piece i used)
    uint32_t aux;          // byte 0 prev_holes, byte 1 prev center filled count, byte 2 game over
    int32_t score, streak, moves, lines_total;
    int32_t max_streak, blocks_total;
    uint32_t draw_ctr;     // candidate trios consumed (TRIO stream index)
    uint32_t policy_ctr;   // random-policy steps taken (POLICY stream index)
};

#define BB_USED(s) (((s).pieces >> 24) & 7u)
#define BB_OVER(s) (((s).aux >> 16) & 1u)

struct BBRewardCfg {   // block_blast_env.py:63-71, order = oracle REWARD_KEYS
    double line_clear_base, block_placed, game_over_penalty, hole_penalty, center_bonus,
        combo_multiplier_bonus, survival_bonus;
};

#define BB_FLAG_RESEED_ON_RESET 1u   // reference behaviour when a seed is given (engine.py:137-138)
#define BB_FLAG_NO_AUTO_RESET 2u     // single-env semantics: stay in GAME_OVER (block_blast_env.py)

// exact float64 arithmetic without FMA contraction
BB_HD double bb_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
BB_HD double bb_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

// block_blast_env.py:158-193 (same operation order; result cast to f32 as wrappers.py:105)
BB_HD float bb_reward(const BBRewardCfg& c, int n_blocks, int lines, bool game_over,
                      int holes, int prev_holes, int center, int prev_center) {
    double r = bb_dadd(0.0, bb_dmul((double)n_blocks, c.block_placed));
    r = bb_dadd(r, c.survival_bonus);
    if (lines > 0) {
        const int cmul = lines < 4 ? lines : 4;
        double lr = bb_dmul((double)lines, c.line_clear_base);
        lr = bb_dmul(lr, (double)cmul);
        r = bb_dadd(r, lr);
        if (cmul > 1) r = bb_dadd(r, bb_dmul((double)(cmul - 1), c.combo_multiplier_bonus));
    }
    if (game_over) r = bb_dadd(r, c.game_over_penalty);
    const int d = holes - prev_holes;
    if (d > 0) r = bb_dadd(r, bb_dmul((double)d, c.hole_penalty));
    // openness = 1 - filled/16 is exact, so `openness >= prev` is `filled <= prev_filled`
    if (center <= prev_center) r = bb_dadd(r, bb_dmul(c.center_bonus, 0.1));
    return (float)r;
}

// draw until a solvable trio appears, at most 100 candidates (engine.py:155-172)
BB_HD uint32_t bb_deal(uint64_t board, const BBTables* T, uint64_t seed, uint64_t env_id, uint32_t* draw_ctr) {
    uint32_t trio = 0;
    for (int attempt = 0; attempt < 100; ++attempt) {
        trio = bb_draw_trio(seed, env_id, *draw_ctr);
        *draw_ctr += 1;
        if (bb_solvable(board, T, trio)) break;
    }
    return trio;   // byte 3 = 0: nothing used
}

BB_HD void bb_reset_state(BBState& s, const BBTables* T, uint64_t seed, uint64_t env_id, uint32_t flags) {
    if (flags & BB_FLAG_RESEED_ON_RESET) s.draw_ctr = 0;
    s.board = 0;
    s.score = s.streak = s.moves = s.lines_total = s.max_streak = s.blocks_total = 0;
    s.aux = 0;             // prev_holes = 0, prev center filled = 0 (openness 1.0), not over
    // on an empty board every trio is solvable (checked exhaustively in tests), so the
    // first candidate is always accepted: engine.py:155-172 consumes exactly one draw
    s.pieces = bb_draw_trio(seed, env_id, s.draw_ctr);
    s.draw_ctr += 1;
}

// the three action-mask planes of the current state (engine.py:364-380)
BB_HD void bb_action_mask(const BBState& s, const BBTables* T, uint64_t m[3]) {
    const uint64_t e = ~s.board;
    const uint32_t used = BB_USED(s);
    const bool over = BB_OVER(s);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t id = (s.pieces >> (8 * i)) & 0xFFu;
        const uint64_t v = bb_valid(e, T->mask[id], T->inb[id]);
        m[i] = (((used >> i) & 1u) || over) ? 0ull : v;
    }
}

struct BBStepOut {
    float reward;
    uint32_t terminated;   // 0/1
    uint32_t info;         // bit0 invalid, bits1-3 lines, bits4-7 blocks, bits8-10 combo mult, bits11-17 draws
    int32_t gain;          // score gained by this move
    int32_t ep_score, ep_len;   // valid when terminated
    uint64_t mask[3];      // action mask of the state after the step (after auto-reset)
};

// One env step with the vec-env's auto-reset.  Invalid action: state untouched, reward -10
// (block_blast_env.py:240-245).  The post-step action mask is always produced because the
// game-over test (engine.py:440-441) needs the same three valid masks.
BB_HD void bb_env_apply(BBState& s, int action, const BBTables* T, const BBRewardCfg& cfg,
                        uint64_t seed, uint64_t env_id, uint32_t flags, BBStepOut& o) {
    const uint32_t used = BB_USED(s);
    const int p = action >> 6;                 // action // 64 (negative actions -> p < 0)
    const int a = action & 63;
    bool ok = (action >= 0) && (p < 3) && !((used >> (p & 3)) & 1u) && !BB_OVER(s);
    const uint32_t id = ok ? ((s.pieces >> (8 * p)) & 0xFFu) : 0u;
    const uint64_t pm = T->mask[id];
    if (ok) ok = ((T->inb[id] >> a) & 1ull) && (((pm << a) & s.board) == 0ull);
    o.ep_score = 0; o.ep_len = 0; o.gain = 0;
    if (!ok) {
        o.reward = -10.0f;
        o.terminated = 0;
        o.info = 1u;
        bb_action_mask(s, T, o.mask);
        return;
    }
    const int n = (int)BB_META_N(T->meta[id]);
    int lines;
    s.board = bb_clear(s.board | (pm << a), &lines);
    uint32_t used2 = used | (1u << p);
    s.moves += 1;
    s.blocks_total += n;
    int gain = n;
    if (lines > 0) {
        s.streak += 1;
        s.max_streak = s.max_streak > s.streak ? s.max_streak : s.streak;
        s.lines_total += lines;
        const int cmul = lines < 4 ? lines : 4;
        const int smul = s.streak + 1 < 8 ? s.streak + 1 : 8;
        gain += lines * 80 * cmul * smul;
    } else {
        s.streak = 0;
    }
    s.score += gain;
    uint32_t draws = 0;
    if (used2 == 7u) {
        const uint32_t before = s.draw_ctr;
        s.pieces = bb_deal(s.board, T, seed, env_id, &s.draw_ctr);   // used bits cleared
        draws = s.draw_ctr - before;
    } else {
        s.pieces = (s.pieces & 0x00FFFFFFu) | (used2 << 24);
    }
    // game over iff no unused piece has an anchor (engine.py:440-441); masks double as the obs
    bb_action_mask(s, T, o.mask);
    const bool over = (o.mask[0] | o.mask[1] | o.mask[2]) == 0ull;
    const int h = bb_holes(s.board), ctr = bb_center(s.board);
    o.reward = bb_reward(cfg, n, lines, over, h, (int)(s.aux & 0xFFu), ctr, (int)((s.aux >> 8) & 0xFFu));
    s.aux = (uint32_t)h | ((uint32_t)ctr << 8) | ((over ? 1u : 0u) << 16);
    o.terminated = over ? 1u : 0u;
    o.gain = gain;
    o.info = ((uint32_t)lines << 1) | ((uint32_t)n << 4) | ((uint32_t)(lines > 0 ? (lines < 4 ? lines : 4) : 1) << 8) | (draws << 11);
    if (over) {
        o.ep_score = s.score;
        o.ep_len = s.moves;
        if (!(flags & BB_FLAG_NO_AUTO_RESET)) {
            bb_reset_state(s, T, seed, env_id, flags);
            bb_action_mask(s, T, o.mask);
        }
    }
}

// k-th valid action (piece-major, then bit order = np.where(mask)[0] order,
// block_blast_env.py:313-323) for the fused random-valid policy.
BB_HD int bb_pick_action(const uint64_t m[3], uint32_t word) {
    const int n0 = bb_popc(m[0]), n1 = bb_popc(m[1]), n2 = bb_popc(m[2]);
    const int total = n0 + n1 + n2;
    if (total == 0) return 0;
    int k = (int)bb_mulhi(word, (uint32_t)total);
    int p = 0;
    uint64_t w = m[0];
    if (k >= n0) { k -= n0; p = 1; w = m[1]; if (k >= n1) { k -= n1; p = 2; w = m[2]; } }
    // select the k-th set bit of w
    for (int i = 0; i < k; ++i) w &= w - 1;
    return p * 64 + bb_ctz(w);
}
