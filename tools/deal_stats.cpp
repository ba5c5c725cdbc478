// Scratch analysis (host build of csrc/bb_rules.cuh): what do the candidate trios of the
// random-policy workload look like to bb_classify, and what would extra cheap tries buy?
//   g++ -O2 -std=c++17 -I block-blast-ai---reinforcement-learning-agent_b200/csrc tools/deal_stats.cpp -o /tmp/deal_stats
#include "bb_rules.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static BBTables T;

int main(int argc, char** argv) {
    bb_fill_tables(&T);
    const int n = 4096, steps = argc > 1 ? atoi(argv[1]) : 300;
    std::vector<BBState> S(n);
    BBRewardCfg cfg = {1.0, 0.01, -1.0, -0.05, 0.02, 0.5, 0.001};
    for (int i = 0; i < n; ++i) { memset(&S[i], 0, sizeof(BBState)); bb_reset_state(S[i], 42, i, 0); }
    long long deals = 0, cands = 0, acc = 0, rej = 0, hard = 0, hard_solv = 0, hard_unsolv = 0;
    long long first_in_A = 0, first_in_B = 0, sum_first = 0, sum_nA = 0, sum_nB = 0, sum_nbr_unsolv = 0;
    long long hist_first[40] = {0};
    long long extra_acc[8] = {0};
    long long nopack_but_solv = 0, f1_rej = 0, f1_wrong = 0, f2_before = 0, f2_after = 0;
    for (int st = 0; st < steps; ++st) {
        for (int i = 0; i < n; ++i) {
            BBState& s = S[i];
            uint64_t m[3];
            bb_action_mask(s, &T, m);
            const BBPhilox4 r = bb_philox((uint32_t)i, 0, s.policy_ctr, BB_STREAM_POLICY, 42, 0);
            s.policy_ctr += 1;
            const int a = bb_pick_action(m, r.x);
            BBStepOut o;
            BBMove mv = bb_env_pre(s, a, &T, o);
            if (!mv.ok) continue;
            uint32_t draws = 0;
            if (mv.needs_deal) {
                deals++;
                uint32_t trio = 0;
                for (int attempt = 0; attempt < 100; ++attempt) {
                    trio = bb_draw_trio(42, i, s.draw_ctr);
                    s.draw_ctr++; draws++;
                    cands++;
                    BBPiece P[3] = {bb_piece(&T, trio & 0xFF), bb_piece(&T, (trio >> 8) & 0xFF), bb_piece(&T, (trio >> 16) & 0xFF)};
                    BBItem it;
                    const int cls = bb_classify(s.board, P[0], P[1], P[2], &it);
                    bool ok;
                    if (cls == BB_ACCEPT) { acc++; ok = true; }
                    else if (cls == BB_REJECT) { rej++; ok = false; }
                    else {
                        hard++;
                        const uint32_t nA = BB_PLAN_NA(it.plan), nB = BB_PLAN_NB(it.plan);
                        sum_nA += nA; sum_nB += nB;
                        int first = -1;
                        for (uint32_t t = 0; t < nA + nB; ++t) if (bb_branch(it, &T, trio, t)) { first = (int)t; break; }
                        ok = first >= 0;
                        {   // filter 1: a piece without anchors must fit once every potentially clearable line is removed
                            const BBLines L = bb_lines(s.board);
                            bool rej1 = false;
                            for (int zi = 0; zi < 3; ++zi) {
                                if (it.v[zi]) continue;
                                int rs = 0, cs = 0;
                                for (int q = 0; q < 3; ++q) if (q != zi) { rs += BB_META_MAXROW(P[q].meta); cs += BB_META_MAXCOL(P[q].meta); }
                                uint64_t rm = rs >= 8 ? ~0ull : bb_rows_within_mask(L, rs);
                                uint64_t cm = cs >= 8 ? ~0ull : (uint64_t)bb_cols_within_bits(L, cs) * BB_COL_A;
                                if (bb_valid(~(s.board & ~(rm | cm)), P[zi]) == 0) rej1 = true;
                            }
                            if (rej1) { f1_rej++; if (ok) f1_wrong++; }
                            // filter 2: useful first placements = clearing alone, or touching a line the two others could finish
                            long nb2 = 0; bool lost = ok;
                            for (int pi = 0; pi < 3; ++pi) {
                                int ro = 0, co = 0;
                                for (int q = 0; q < 3; ++q) if (q != pi) { ro = ro > (int)BB_META_MAXROW(P[q].meta) ? ro : BB_META_MAXROW(P[q].meta); co = co > (int)BB_META_MAXCOL(P[q].meta) ? co : BB_META_MAXCOL(P[q].meta); }
                                int rs = BB_META_MAXROW(P[pi].meta) + ro, cs = BB_META_MAXCOL(P[pi].meta) + co;
                                uint64_t rm = rs >= 8 ? ~0ull : bb_rows_within_mask(L, rs);
                                uint64_t cm = cs >= 8 ? ~0ull : (uint64_t)bb_cols_within_bits(L, cs) * BB_COL_A;
                                uint64_t keep = it.v[pi] & bb_cover(~s.board & (rm | cm), P[pi]);
                                nb2 += bb_popc(keep);
                            }
                            f2_before += nB; f2_after += nb2; (void)lost;
                        }
                        if (ok) {
                            hard_solv++;
                            if ((uint32_t)first < nA) first_in_A++; else first_in_B++;
                            sum_first += first;
                            hist_first[first < 39 ? first : 39]++;
                            // would a full packing search (no clears) have found it?
                            bool pack = false;
                            for (uint32_t t = 0; t < nA && !pack; ++t) if (bb_branch(it, &T, trio, t)) pack = true;
                            if (!pack) nopack_but_solv++;
                            // extra cheap tries: x at its k-th lowest anchor (k = 1..), y lowest/highest
                            const int x = it.plan & 3, y = (it.plan >> 2) & 3, z = (it.plan >> 4) & 3;
                            if (nA) {
                                uint64_t vx = it.v[x];
                                for (int k = 0; k < 8 && vx; ++k) {
                                    const int ax = bb_ctz(vx); vx &= vx - 1;
                                    if (k == 0) continue;
                                    const uint64_t b1 = s.board | (P[x].pm << ax);
                                    const uint64_t vy = bb_valid(~b1, P[y]);
                                    bool hit = false;
                                    if (vy) {
                                        if (bb_valid(~(b1 | (P[y].pm << bb_ctz(vy))), P[z])) hit = true;
                                        else if (bb_valid(~(b1 | (P[y].pm << (63 - __builtin_clzll(vy)))), P[z])) hit = true;
                                    }
                                    if (hit) { for (int q = k; q < 8; ++q) extra_acc[q]++; break; }
                                }
                            }
                        } else { hard_unsolv++; sum_nbr_unsolv += nA + nB; }
                    }
                    if (ok) break;
                }
                s.pieces = trio;
            }
            bb_env_post(s, mv, draws, &T, cfg, 42, i, 0, o);
        }
    }
    printf("deals %lld candidates %lld (%.3f per deal)\n", deals, cands, (double)cands / deals);
    printf("ACCEPT %.4f  REJECT %.4f  HARD %.4f of candidates\n", (double)acc / cands, (double)rej / cands, (double)hard / cands);
    printf("HARD: solvable %.4f unsolvable %.4f | mean nA %.1f nB %.1f | unsolvable mean branches %.1f\n", (double)hard_solv / hard,
           (double)hard_unsolv / hard, (double)sum_nA / hard, (double)sum_nB / hard, hard_unsolv ? (double)sum_nbr_unsolv / hard_unsolv : 0.0);
    printf("HARD solvable: first solving branch in stage A %.4f, stage B %.4f, mean index %.2f; solvable only with clears %.4f\n",
           (double)first_in_A / hard_solv, (double)first_in_B / hard_solv, (double)sum_first / hard_solv, (double)nopack_but_solv / hard_solv);
    printf("filter1 (blocked piece fits after removing all potentially clearable lines): rejects %.4f of HARD (%.4f of unsolvable), wrong %lld\n", (double)f1_rej / hard, (double)f1_rej / hard_unsolv, f1_wrong);
    printf("filter2: stage-B branches %.2f -> %.2f per HARD item\n", (double)f2_before / hard, (double)f2_after / hard);
    printf("first-branch histogram:");
    for (int k = 0; k < 40; ++k) printf(" %lld", hist_first[k]);
    printf("\nextra tries (x at k-th lowest anchor, y lowest/highest): cumulative accept fraction of HARD-solvable:");
    for (int k = 1; k < 8; ++k) printf(" k<=%d: %.3f", k, (double)extra_acc[k] / hard_solv);
    printf("\n");
    return 0;
}
