// Host (g++) build of the product's rules header csrc/bb_rules.cuh — TEST INFRASTRUCTURE.
// It exists so that the exact code the sm_100a kernels execute per env can be fuzzed against
// the oracle in a container without a GPU.  The product never loads this library and it is
// not a CPU fallback: bbgpu raises if libbbgpu.so (CUDA) is missing.
#define BB_COUNT_WORK 1
#include "bb_rules.cuh"
#include <string.h>

static BBTables g_tables;
static bool g_init = false;
static void init() { if (!g_init) { bb_fill_tables(&g_tables); g_init = true; } }

extern "C" {

int64_t bbh_state_size() { return (int64_t)sizeof(BBState); }

uint64_t bbh_valid(uint64_t board, int piece) { init(); return bb_valid(~board, bb_piece(&g_tables, (uint32_t)piece)); }
uint64_t bbh_clear(uint64_t board, int* lines) { return bb_clear(board, lines); }
int bbh_holes(uint64_t board) { return bb_holes(board); }
int bbh_center(uint64_t board) { return bb_center(board); }
int bbh_solvable(uint64_t board, int p0, int p1, int p2) {
    init();
    return bb_solvable(board, &g_tables, (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16)) ? 1 : 0;
}
// 0 reject, 1 accept, 2 hard
int bbh_solvable_fast(uint64_t board, int p0, int p1, int p2) {
    init();
    BBPiece P[3] = {bb_piece(&g_tables, (uint32_t)p0), bb_piece(&g_tables, (uint32_t)p1), bb_piece(&g_tables, (uint32_t)p2)};
    BBItem it;
    return bb_classify(board, P[0], P[1], P[2], &it);
}
void bbh_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    BBPhilox4 r = bb_philox(c0, c1, c2, c3, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
uint32_t bbh_draw_trio(uint64_t seed, uint64_t env_id, uint32_t draw) { return bb_draw_trio(seed, env_id, draw); }
int bbh_pick_action(const uint64_t m[3], uint32_t word) { return bb_pick_action(m, word); }

void bbh_reset(BBState* s, int64_t n, uint64_t seed, int64_t env_offset, uint32_t flags) {
    init();
    for (int64_t i = 0; i < n; ++i) {
        memset(&s[i], 0, sizeof(BBState));
        bb_reset_state(s[i], seed, (uint64_t)(env_offset + i), flags);
    }
}

void bbh_masks(const BBState* s, int64_t n, uint64_t* mask /* [n][3] */) {
    init();
    for (int64_t i = 0; i < n; ++i) bb_action_mask(s[i], &g_tables, mask + 3 * i);
}

// actions == NULL: fused random-valid policy (POLICY stream)
void bbh_step(BBState* s, int64_t n, const int32_t* actions, const double cfg[7], uint64_t seed,
              int64_t env_offset, uint32_t flags, float* rewards, uint8_t* terminated,
              uint32_t* info, int32_t* gain, int32_t* ep_score, int32_t* ep_len, uint64_t* mask,
              int32_t* actions_out) {
    init();
    BBRewardCfg c;
    memcpy(&c, cfg, sizeof(c));
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t env_id = (uint64_t)(env_offset + i);
        int a;
        if (actions) a = actions[i];
        else {
            uint64_t m[3];
            bb_action_mask(s[i], &g_tables, m);
            const BBPhilox4 r = bb_philox((uint32_t)env_id, (uint32_t)(env_id >> 32), s[i].policy_ctr,
                                          BB_STREAM_POLICY, (uint32_t)seed, (uint32_t)(seed >> 32));
            s[i].policy_ctr += 1;
            a = bb_pick_action(m, r.x);
        }
        if (actions_out) actions_out[i] = a;
        BBStepOut o;
        bb_env_apply(s[i], a, &g_tables, c, seed, env_id, flags, o);
        rewards[i] = o.reward;
        terminated[i] = (uint8_t)o.terminated;
        if (info) info[i] = o.info;
        if (gain) gain[i] = o.gain;
        if (o.terminated) { if (ep_score) ep_score[i] = o.ep_score; if (ep_len) ep_len[i] = o.ep_len; }
        if (mask) { mask[3 * i] = o.mask[0]; mask[3 * i + 1] = o.mask[1]; mask[3 * i + 2] = o.mask[2]; }
    }
}

void bbh_work(long long out[7]) {
    out[0] = g_bb_work.valid_calls; out[1] = g_bb_work.fast_accept; out[2] = g_bb_work.fast_reject;
    out[3] = g_bb_work.pack_iters; out[4] = g_bb_work.clear_iters; out[5] = g_bb_work.slow;
    out[6] = g_bb_work.branches;
}
// per-branch evaluation of a HARD item, for checking the branch decomposition the kernel uses
int bbh_item_plan(uint64_t board, int p0, int p1, int p2, uint32_t* plan) {
    init();
    BBPiece P[3] = {bb_piece(&g_tables, (uint32_t)p0), bb_piece(&g_tables, (uint32_t)p1), bb_piece(&g_tables, (uint32_t)p2)};
    BBItem it;
    const int cls = bb_classify(board, P[0], P[1], P[2], &it);
    *plan = it.plan;
    return cls;
}
int bbh_item_branch(uint64_t board, int p0, int p1, int p2, uint32_t t) {
    init();
    BBPiece P[3] = {bb_piece(&g_tables, (uint32_t)p0), bb_piece(&g_tables, (uint32_t)p1), bb_piece(&g_tables, (uint32_t)p2)};
    BBItem it;
    bb_classify(board, P[0], P[1], P[2], &it);
    return bb_branch(it, &g_tables, (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16), t) ? 1 : 0;
}
void bbh_work_reset() { memset(&g_bb_work, 0, sizeof(g_bb_work)); }

}  // extern "C"

// debug view of one opened branch (tools only): out = {bb, m0, m1, always, A.meta, B.meta, A.pm, B.pm}
extern "C" void bbh_branch_info(uint64_t board, int p0, int p1, int p2, uint32_t t, uint64_t out[8]) {
    init();
    BBPiece P[3] = {bb_piece(&g_tables, (uint32_t)p0), bb_piece(&g_tables, (uint32_t)p1), bb_piece(&g_tables, (uint32_t)p2)};
    BBItem it;
    bb_classify(board, P[0], P[1], P[2], &it);
    BBBranch br;
    bb_branch_open(br, it, &g_tables, (uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16), t);
    out[0] = br.bb; out[1] = br.m0; out[2] = br.m1; out[3] = br.always; out[4] = br.A.meta; out[5] = br.B.meta;
    out[6] = br.A.pm; out[7] = br.B.pm;
}
