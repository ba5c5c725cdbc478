/* CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's Block Blast rules on a cell grid — the same
 * algorithm as oracle/bb_oracle.py, fast enough to check millions of env-steps.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load it, and only as the checker / the CPU baseline.  The product library (libbbgpu.so)
 * shares no code with this file: no bitboards here, cells and loops only.
 *
 * Pinned against the reference by tests/test_oracle_golden.py (golden traces recorded from
 * the unmodified reference by tests/golden/make_golden.py).
 *
 * Citations are file:line into the reference checkout:
 *   src/game/board.py:71-93     fits()            can_place
 *   src/game/board.py:95-115    put()             place_piece
 *   src/game/board.py:117-142   has_anchor()      has_valid_placement
 *   src/game/board.py:144-193   sweep_lines()     find_complete_lines + clear_lines
 *   src/game/board.py:195-216   count_holes()
 *   src/game/board.py:236-243   center_filled()   get_center_openness
 *   src/game/engine.py:155-238  deal()/solvable() _generate_new_pieces + DFS
 *   src/game/engine.py:240-312  score formula     (:261 uses the already-incremented streak)
 *   src/game/engine.py:326-346  legal()           can_place_piece
 *   src/game/engine.py:390-454  move()            make_move
 *   src/environment/block_blast_env.py:104-118  action decode
 *   src/environment/block_blast_env.py:148-193  reward (float64, this operation order)
 *   src/environment/block_blast_env.py:224-264  step (invalid action: -10, state untouched)
 *   src/environment/wrappers.py:75-116          vec step + auto-reset
 *   src/agents/ppo.py:141-169                   GAE
 *
 * Candidate trios are read from a caller-provided stream (one row of 3 piece indices per
 * rng.choice(37,size=3) call, pieces.py:350-355): the reference's numpy PCG64 stream is not
 * reproduced (north_star: "fed the same piece sequences").
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).  -ffp-contract=off is required so
 * the float32/float64 operation order below is what executes.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "bb_piece_cells.h"

#define NB 8

typedef struct {
    signed char grid[NB][NB];
    int trio[3];
    int used[3];
    int64_t score;
    int streak, moves, lines_total, over, max_streak, blocks_total;
    int prev_holes;
    double prev_center;
    /* candidate-trio source */
    const uint8_t *stream;   /* [stream_len][3] */
    int64_t stream_len;
    int64_t cursor;          /* next row to consume */
    int64_t draws;           /* total rows consumed over the env's life */
    int reseed;              /* 1: cursor returns to 0 on every reset (seeded reference path) */
    int exhausted;           /* set when the stream ran out (caller error) */
    double cfg[7];           /* line_clear_base, block_placed, game_over_penalty, hole_penalty,
                                center_bonus, combo_multiplier_bonus, survival_bonus */
    /* last-move record (MoveResult, engine.py:29-41) */
    int last_blocks, last_lines, last_combo_mult, last_rows, last_cols;
    int64_t last_gain;
    int64_t dfs_nodes;       /* instrumentation: _can_place_remaining calls */
} bbo_env;

int64_t bbo_env_size(void) { return (int64_t)sizeof(bbo_env); }

/* ------------------------------------------------------------------ board rules */
static int fits(signed char g[NB][NB], int piece, int row, int col) {
    const bbo_piece *p = &BBO_PIECES[piece];
    for (int k = 0; k < p->n; ++k) {
        int r = row + p->dr[k], c = col + p->dc[k];
        if (r < 0 || r >= NB || c < 0 || c >= NB) return 0;
        if (g[r][c] != 0) return 0;
    }
    return 1;
}

static void put(signed char g[NB][NB], int piece, int row, int col) {
    const bbo_piece *p = &BBO_PIECES[piece];
    for (int k = 0; k < p->n; ++k) g[row + p->dr[k]][col + p->dc[k]] = 1;
}

static int has_anchor(signed char g[NB][NB], int piece) {
    const bbo_piece *p = &BBO_PIECES[piece];
    for (int r = 0; r <= NB - p->h; ++r)
        for (int c = 0; c <= NB - p->w; ++c)
            if (fits(g, piece, r, c)) return 1;
    return 0;
}

/* all full rows and columns are found first, then zeroed */
static void sweep_lines(signed char g[NB][NB], int *n_rows, int *n_cols) {
    int fr[NB], fc[NB], nr = 0, nc = 0;
    for (int r = 0; r < NB; ++r) {
        int full = 1;
        for (int c = 0; c < NB; ++c) if (g[r][c] != 1) { full = 0; break; }
        fr[r] = full; nr += full;
    }
    for (int c = 0; c < NB; ++c) {
        int full = 1;
        for (int r = 0; r < NB; ++r) if (g[r][c] != 1) { full = 0; break; }
        fc[c] = full; nc += full;
    }
    for (int r = 0; r < NB; ++r) if (fr[r]) for (int c = 0; c < NB; ++c) g[r][c] = 0;
    for (int c = 0; c < NB; ++c) if (fc[c]) for (int r = 0; r < NB; ++r) g[r][c] = 0;
    *n_rows = nr; *n_cols = nc;
}

static int count_holes(signed char g[NB][NB]) {
    static const int DR[4] = {-1, 1, 0, 0}, DC[4] = {0, 0, -1, 1};
    int holes = 0;
    for (int r = 0; r < NB; ++r)
        for (int c = 0; c < NB; ++c) {
            if (g[r][c] != 0) continue;
            int blocked = 0;
            for (int k = 0; k < 4; ++k) {
                int rr = r + DR[k], cc = c + DC[k];
                if (rr < 0 || rr >= NB || cc < 0 || cc >= NB || g[rr][cc] == 1) ++blocked;
            }
            if (blocked == 4) ++holes;
        }
    return holes;
}

static int center_filled(signed char g[NB][NB]) {
    int s = 0;
    for (int r = 2; r < 6; ++r) for (int c = 2; c < 6; ++c) s += g[r][c];
    return s;
}

/* ------------------------------------------------------------------ engine */
static int solvable(bbo_env *e, signed char g[NB][NB], const int used[3]) {
    e->dfs_nodes++;
    if (used[0] && used[1] && used[2]) return 1;
    for (int i = 0; i < 3; ++i) {
        if (used[i]) continue;
        int piece = e->trio[i];
        const bbo_piece *p = &BBO_PIECES[piece];
        for (int r = 0; r <= NB - p->h; ++r)
            for (int c = 0; c <= NB - p->w; ++c) {
                if (!fits(g, piece, r, c)) continue;
                signed char g2[NB][NB];
                memcpy(g2, g, sizeof(g2));
                put(g2, piece, r, c);
                int nr, nc;
                sweep_lines(g2, &nr, &nc);
                int u2[3] = {used[0], used[1], used[2]};
                u2[i] = 1;
                if (solvable(e, g2, u2)) return 1;
            }
    }
    return 0;
}

static void deal(bbo_env *e) {
    for (int attempt = 0; attempt < 100; ++attempt) {
        if (e->cursor >= e->stream_len) { e->exhausted = 1; e->cursor = 0; }
        const uint8_t *row = e->stream + 3 * e->cursor;
        e->cursor++; e->draws++;
        e->trio[0] = row[0]; e->trio[1] = row[1]; e->trio[2] = row[2];
        e->used[0] = e->used[1] = e->used[2] = 0;
        const int none[3] = {0, 0, 0};
        if (solvable(e, e->grid, none)) return;
    }
    /* 100 rejections: the last candidate stays */
}

static void zero_game(bbo_env *e) {
    memset(e->grid, 0, sizeof(e->grid));
    e->trio[0] = e->trio[1] = e->trio[2] = 0;
    e->used[0] = e->used[1] = e->used[2] = 0;
    e->score = 0; e->streak = 0; e->moves = 0; e->lines_total = 0; e->over = 0;
    e->max_streak = 0; e->blocks_total = 0;
}

void bbo_env_reset(bbo_env *e) {
    if (e->reseed) e->cursor = 0;          /* engine.py:137-138 re-seed */
    zero_game(e);
    deal(e);
    e->prev_holes = 0;                      /* block_blast_env.py:216-217 */
    e->prev_center = 1.0;
}

void bbo_env_init(bbo_env *e, const uint8_t *stream, int64_t stream_len,
                  const double cfg[7], int reseed) {
    memset(e, 0, sizeof(*e));
    e->stream = stream; e->stream_len = stream_len; e->reseed = reseed;
    for (int k = 0; k < 7; ++k) e->cfg[k] = cfg[k];
    bbo_env_reset(e);
}

static int legal(bbo_env *e, int i, int r, int c) {
    if (i < 0 || i >= 3) return 0;
    if (e->used[i]) return 0;
    if (e->over) return 0;
    return fits(e->grid, e->trio[i], r, c);
}

static int any_move(bbo_env *e) {
    for (int i = 0; i < 3; ++i)
        if (!e->used[i] && has_anchor(e->grid, e->trio[i])) return 1;
    return 0;
}

/* returns 0 if the move was rejected */
static int move(bbo_env *e, int i, int r, int c) {
    if (!legal(e, i, r, c)) return 0;
    int piece = e->trio[i];
    int n = BBO_PIECES[piece].n;
    put(e->grid, piece, r, c);
    e->used[i] = 1;
    e->moves += 1;
    e->blocks_total += n;
    int nr, nc;
    sweep_lines(e->grid, &nr, &nc);
    int lines = nr + nc;
    if (lines > 0) {
        e->streak += 1;
        if (e->streak > e->max_streak) e->max_streak = e->streak;
        e->lines_total += lines;
    } else {
        e->streak = 0;
    }
    int64_t gain = n;
    if (lines > 0) {
        int cm = lines < 4 ? lines : 4;
        int sm = e->streak + 1 < 8 ? e->streak + 1 : 8;
        gain += (int64_t)(lines * NB * 10) * cm * sm;
    }
    e->score += gain;
    if (e->used[0] && e->used[1] && e->used[2]) deal(e);
    if (!any_move(e)) e->over = 1;
    e->last_blocks = n; e->last_lines = lines; e->last_rows = nr; e->last_cols = nc;
    e->last_combo_mult = lines > 0 ? (lines < 4 ? lines : 4) : 1;
    e->last_gain = gain;
    return 1;
}

static double shaped_reward(bbo_env *e) {
    const double *cfg = e->cfg;
    double r = 0.0;
    r += e->last_blocks * cfg[1];
    r += cfg[6];
    if (e->last_lines > 0) {
        double lr = e->last_lines * cfg[0];
        lr *= e->last_combo_mult;
        r += lr;
        if (e->last_combo_mult > 1) r += (e->last_combo_mult - 1) * cfg[5];
    }
    if (e->over) r += cfg[2];
    int h = count_holes(e->grid);
    int d = h - e->prev_holes;
    if (d > 0) r += d * cfg[3];
    e->prev_holes = h;
    double o = 1.0 - (center_filled(e->grid) / 16.0);
    if (o >= e->prev_center) r += cfg[4] * 0.1;
    e->prev_center = o;
    return r;
}

/* single env step, NO auto-reset (block_blast_env.py:224-264) */
void bbo_env_step(bbo_env *e, int action, double *reward, int *terminated, int *invalid) {
    /* Python floor semantics for negative actions: a // 64 < 0 => rejected */
    int i, r, c;
    if (action < 0) { i = -1; r = 0; c = 0; }
    else { i = action / 64; r = (action % 64) / 8; c = action % 8; }
    if (!legal(e, i, r, c)) { *reward = -10.0; *terminated = 0; *invalid = 1; return; }
    move(e, i, r, c);
    *reward = shaped_reward(e);
    *terminated = e->over;
    *invalid = 0;
}

/* packed observation of the current state: board u64 (bit=r*8+c), pieces[4] = 3 ids + used
 * bits, mask[3] (plane of a used piece, or any plane when game over, is 0; engine.py:364-380) */
void bbo_env_export(bbo_env *e, uint64_t *board, uint8_t pieces[4], uint64_t mask[3]) {
    uint64_t b = 0;
    for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c)
        if (e->grid[r][c]) b |= 1ull << (r * 8 + c);
    *board = b;
    pieces[0] = (uint8_t)e->trio[0]; pieces[1] = (uint8_t)e->trio[1]; pieces[2] = (uint8_t)e->trio[2];
    pieces[3] = (uint8_t)(e->used[0] | (e->used[1] << 1) | (e->used[2] << 2));
    for (int i = 0; i < 3; ++i) {
        uint64_t m = 0;
        if (!e->used[i])
            for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c)
                if (fits(e->grid, e->trio[i], r, c)) m |= 1ull << (r * 8 + c);
        mask[i] = m;
    }
}

/* stats[8]: score, streak, moves, lines_total, max_streak, blocks_total, holes, draws */
void bbo_env_stats(bbo_env *e, int64_t stats[8]) {
    stats[0] = e->score; stats[1] = e->streak; stats[2] = e->moves; stats[3] = e->lines_total;
    stats[4] = e->max_streak; stats[5] = e->blocks_total; stats[6] = count_holes(e->grid);
    stats[7] = e->draws;
}

int bbo_env_exhausted(bbo_env *e) { return e->exhausted; }
int64_t bbo_env_dfs_nodes(bbo_env *e) { return e->dfs_nodes; }

void bbo_env_set_board(bbo_env *e, uint64_t board, const uint8_t pieces[4]) {
    for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c)
        e->grid[r][c] = (signed char)((board >> (r * 8 + c)) & 1);
    for (int i = 0; i < 3; ++i) { e->trio[i] = pieces[i]; e->used[i] = (pieces[3] >> i) & 1; }
}

/* solvability of a (board, trio) pair by the reference's DFS; for fuzzing the product's
 * pruned search.  Returns 0/1; *nodes gets the node count. */
int bbo_trio_solvable(uint64_t board, int p0, int p1, int p2, int64_t *nodes) {
    bbo_env e;
    memset(&e, 0, sizeof(e));
    for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c)
        e.grid[r][c] = (signed char)((board >> (r * 8 + c)) & 1);
    e.trio[0] = p0; e.trio[1] = p1; e.trio[2] = p2;
    const int none[3] = {0, 0, 0};
    int ok = solvable(&e, e.grid, none);
    if (nodes) *nodes = e.dfs_nodes;
    return ok;
}

/* Vectorised step with auto-reset (wrappers.py:75-116).  envs is an array of n bbo_env.
 * Outputs describe the state AFTER the step (after the reset when terminated); reward and
 * terminated are the terminal step's.  ep_score/ep_len are written only where terminated
 * (info['final_score'], info['moves']).  All output pointers except rewards/terminated may
 * be NULL.  OpenMP-parallel over envs when built with -fopenmp (the reference itself is a
 * serial loop; threads only matter for the cpu_baseline leg). */
void bbo_vec_step(bbo_env *envs, int64_t n, const int32_t *actions, float *rewards,
                  uint8_t *terminated, uint8_t *invalid, uint64_t *board_out,
                  uint8_t *pieces_out, uint64_t *mask_out, int32_t *ep_score, int32_t *ep_len,
                  int n_threads) {
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t k = 0; k < n; ++k) {
        bbo_env *e = &envs[k];
        double rew; int term, inv;
        bbo_env_step(e, actions[k], &rew, &term, &inv);
        if (term) {
            if (ep_score) ep_score[k] = (int32_t)e->score;
            if (ep_len) ep_len[k] = e->moves;
            bbo_env_reset(e);
        }
        rewards[k] = (float)rew;            /* float64 -> float32, wrappers.py:105 */
        terminated[k] = (uint8_t)term;
        if (invalid) invalid[k] = (uint8_t)inv;
        if (board_out || pieces_out || mask_out) {
            uint64_t b, m[3]; uint8_t pc[4];
            bbo_env_export(e, &b, pc, m);
            if (board_out) board_out[k] = b;
            if (pieces_out) memcpy(pieces_out + 4 * k, pc, 4);
            if (mask_out) { mask_out[3 * k] = m[0]; mask_out[3 * k + 1] = m[1]; mask_out[3 * k + 2] = m[2]; }
        }
    }
}

/* Random-valid-action rollout used as the CPU baseline for config 3: every step each env
 * picks the k-th valid action (piece-major, row, col order = np.where(mask)[0] order,
 * block_blast_env.py:313-323), k = words[step*n + env] % n_valid.  Returns env-steps done. */
int64_t bbo_random_rollout(bbo_env *envs, int64_t n, int64_t n_steps, const uint32_t *words,
                           int64_t *episodes, int64_t *score_sum, int n_threads) {
    int64_t ep = 0, ss = 0;
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1) reduction(+:ep,ss)
#endif
    for (int64_t k = 0; k < n; ++k) {
        bbo_env *e = &envs[k];
        for (int64_t s = 0; s < n_steps; ++s) {
            int valid[192], nv = 0;
            for (int i = 0; i < 3; ++i) {
                if (e->used[i]) continue;
                for (int r = 0; r < NB; ++r) for (int c = 0; c < NB; ++c)
                    if (fits(e->grid, e->trio[i], r, c)) valid[nv++] = i * 64 + r * 8 + c;
            }
            int a = nv ? valid[words[s * n + k] % (uint32_t)nv] : 0;
            double rew; int term, inv;
            bbo_env_step(e, a, &rew, &term, &inv);
            if (term) { ep++; ss += e->score; bbo_env_reset(e); }
        }
    }
    *episodes = ep; *score_sum = ss;
    return n * n_steps;
}

/* GAE, ppo.py:141-169: float32 arrays (T,N) row-major; gamma and gamma*lambda are Python
 * doubles that numpy casts to float32 before the array multiply (NEP 50 weak scalars);
 * every operation rounds to float32 (no FMA: build with -ffp-contract=off). */
void bbo_gae(const float *rewards, const float *values, const float *dones,
             const float *last_values, double gamma, double lam,
             float *adv, float *ret, int64_t T, int64_t N) {
    const float g = (float)gamma;
    const float gl = (float)(gamma * lam);
    for (int64_t k = 0; k < N; ++k) {
        float last = 0.0f;
        for (int64_t t = T - 1; t >= 0; --t) {
            float nnt = 1.0f - dones[t * N + k];
            float nv = (t == T - 1) ? last_values[k] : values[(t + 1) * N + k];
            float x = g * nv;  x = x * nnt;
            float delta = rewards[t * N + k] + x;  delta = delta - values[t * N + k];
            float y = gl * nnt;  y = y * last;
            last = delta + y;
            adv[t * N + k] = last;
            ret[t * N + k] = last + values[t * N + k];
        }
    }
}
