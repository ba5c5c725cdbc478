// bb_env_kernels.cu — K1: the fused Block Blast env-step kernels for sm_100a.
//
// One thread owns one env.  The env state is 48 B kept as three 16-byte words in three SoA
// arrays (s0/s1/s2), so a warp reads and writes 3 x 512 contiguous bytes with 128-bit
// accesses; all outputs are SoA too (mask planes are plane-major).  The piece tables (37 x
// 20 B) are staged into shared memory once per block because lanes index them with
// different piece ids (constant memory would serialise divergent indices).
//
// Algorithmic HBM bytes per env-step (packed protocol): state 48 R + 48 W,
// action 4, reward 4, terminated 1, mask 24  = 129 B.
//
// Reference path replaced: VectorizedBlockBlastEnv.step (src/environment/wrappers.py:75-116)
// -> BlockBlastEnv.step (src/environment/block_blast_env.py:224-264) -> GameEngine.make_move
// (src/game/engine.py:390-454); rules in bb_rules.cuh.
#include <cuda_runtime.h>
#include "bb_rules.cuh"
#include "bb_kernels.h"

// statically initialised (arrays are padded to 40 entries, the tail is zero)
__constant__ BBTables c_bb_tables = {BB_PIECE_MASKS, BB_PIECE_INB, BB_PIECE_META};

__device__ __forceinline__ void bb_stage_tables(BBTables* sh) {
    // 43 entries of each array; 256 threads: thread k copies entry k of each table
    for (int k = threadIdx.x; k < BB_NUM_PIECES + 3; k += blockDim.x) {
        sh->mask[k] = c_bb_tables.mask[k];
        sh->inb[k] = c_bb_tables.inb[k];
        sh->meta[k] = c_bb_tables.meta[k];
    }
    __syncthreads();
}

__device__ __forceinline__ void bb_load_state(const BBEnvArrays& E, int64_t i, BBState& s) {
    const uint4 a = E.s0[i], b = E.s1[i], c = E.s2[i];
    s.board = (uint64_t)a.x | ((uint64_t)a.y << 32);
    s.pieces = a.z;
    s.aux = a.w;
    s.score = (int32_t)b.x; s.streak = (int32_t)b.y; s.moves = (int32_t)b.z; s.lines_total = (int32_t)b.w;
    s.max_streak = (int32_t)c.x; s.blocks_total = (int32_t)c.y; s.draw_ctr = c.z; s.policy_ctr = c.w;
}

__device__ __forceinline__ void bb_store_state(const BBEnvArrays& E, int64_t i, const BBState& s) {
    E.s0[i] = make_uint4((uint32_t)s.board, (uint32_t)(s.board >> 32), s.pieces, s.aux);
    E.s1[i] = make_uint4((uint32_t)s.score, (uint32_t)s.streak, (uint32_t)s.moves, (uint32_t)s.lines_total);
    E.s2[i] = make_uint4((uint32_t)s.max_streak, (uint32_t)s.blocks_total, s.draw_ctr, s.policy_ctr);
}

// ---------------------------------------------------------------------------------------
// K1: one env step per thread; RANDOM fuses the uniform-random-valid policy and may run
// n_steps back to back with the state held in registers.
// ---------------------------------------------------------------------------------------
template <bool RANDOM>
__global__ void __launch_bounds__(BB_STEP_THREADS)
bb_step_kernel(BBEnvArrays E, BBRewardCfg cfg, const int32_t* __restrict__ actions, int n_steps,
               int32_t* __restrict__ actions_out, float* __restrict__ rewards,
               uint8_t* __restrict__ terminated, uint64_t* __restrict__ mask_out,
               int32_t* __restrict__ ep_score, int32_t* __restrict__ ep_len,
               uint32_t* __restrict__ info_out, unsigned long long* __restrict__ stats) {
    __shared__ BBTables T;
    bb_stage_tables(&T);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < E.n;
    unsigned long long st_eps = 0, st_score = 0, st_len = 0;
    if (live) {
        BBState s;
        bb_load_state(E, i, s);
        const uint64_t env_id = (uint64_t)(E.env_offset + i);
        BBStepOut o;
        int action = 0;
        if (RANDOM) {
            bb_action_mask(s, &T, o.mask);
            for (int step = 0; step < n_steps; ++step) {
                const BBPhilox4 r = bb_philox((uint32_t)env_id, (uint32_t)(env_id >> 32), s.policy_ctr,
                                              BB_STREAM_POLICY, (uint32_t)E.seed, (uint32_t)(E.seed >> 32));
                s.policy_ctr += 1;
                action = bb_pick_action(o.mask, r.x);
                bb_env_apply(s, action, &T, cfg, E.seed, env_id, E.flags, o);
                if (o.terminated) { st_eps += 1; st_score += (unsigned)o.ep_score; st_len += (unsigned)o.ep_len; }
            }
        } else {
            action = actions[i];
            bb_env_apply(s, action, &T, cfg, E.seed, env_id, E.flags, o);
        }
        bb_store_state(E, i, s);
        if (RANDOM && actions_out) actions_out[i] = action;
        if (rewards) rewards[i] = o.reward;
        if (terminated) terminated[i] = (uint8_t)o.terminated;
        if (mask_out) {
            mask_out[i] = o.mask[0];
            mask_out[E.n + i] = o.mask[1];
            mask_out[2 * E.n + i] = o.mask[2];
        }
        if (o.terminated) {
            if (ep_score) ep_score[i] = o.ep_score;
            if (ep_len) ep_len[i] = o.ep_len;
        }
        if (info_out) info_out[i] = o.info;
    }
    if (RANDOM && stats) {
        // warp-reduce, one atomic per warp per counter
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            st_eps += __shfl_down_sync(0xffffffffu, st_eps, d);
            st_score += __shfl_down_sync(0xffffffffu, st_score, d);
            st_len += __shfl_down_sync(0xffffffffu, st_len, d);
        }
        const unsigned long long nlive = __popc(__ballot_sync(0xffffffffu, live));
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&stats[0], nlive * (unsigned long long)n_steps);
            if (st_eps) {
                atomicAdd(&stats[1], st_eps);
                atomicAdd(&stats[2], st_score);
                atomicAdd(&stats[3], st_len);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// reset (VectorizedBlockBlastEnv.reset, wrappers.py:53-73) and observe
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BB_STEP_THREADS)
bb_reset_kernel(BBEnvArrays E, const uint8_t* __restrict__ reset_mask, uint64_t* __restrict__ mask_out) {
    __shared__ BBTables T;
    bb_stage_tables(&T);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    BBState s;
    bb_load_state(E, i, s);
    if (!reset_mask || reset_mask[i]) {
        bb_reset_state(s, &T, E.seed, (uint64_t)(E.env_offset + i), E.flags);
        bb_store_state(E, i, s);
    }
    if (mask_out) {
        uint64_t m[3];
        bb_action_mask(s, &T, m);
        mask_out[i] = m[0];
        mask_out[E.n + i] = m[1];
        mask_out[2 * E.n + i] = m[2];
    }
}

__global__ void __launch_bounds__(BB_STEP_THREADS)
bb_observe_kernel(BBEnvArrays E, uint64_t* __restrict__ board_out, uint32_t* __restrict__ pieces_out,
                  uint64_t* __restrict__ mask_out) {
    __shared__ BBTables T;
    bb_stage_tables(&T);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    BBState s;
    bb_load_state(E, i, s);
    if (board_out) board_out[i] = s.board;
    if (pieces_out) pieces_out[i] = s.pieces;
    if (mask_out) {
        uint64_t m[3];
        bb_action_mask(s, &T, m);
        mask_out[i] = m[0];
        mask_out[E.n + i] = m[1];
        mask_out[2 * E.n + i] = m[2];
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static inline unsigned bb_grid(int64_t n) { return (unsigned)((n + BB_STEP_THREADS - 1) / BB_STEP_THREADS); }

cudaError_t bb_launch_step(const BBEnvArrays& E, const BBRewardCfg& cfg, const int32_t* actions,
                           float* rewards, uint8_t* terminated, uint64_t* mask_out, int32_t* ep_score,
                           int32_t* ep_len, uint32_t* info_out, cudaStream_t stream) {
    bb_step_kernel<false><<<bb_grid(E.n), BB_STEP_THREADS, 0, stream>>>(
        E, cfg, actions, 1, nullptr, rewards, terminated, mask_out, ep_score, ep_len, info_out, nullptr);
    return cudaGetLastError();
}

cudaError_t bb_launch_step_random(const BBEnvArrays& E, const BBRewardCfg& cfg, int n_steps,
                                  int32_t* actions_out, float* rewards, uint8_t* terminated,
                                  uint64_t* mask_out, unsigned long long* stats, cudaStream_t stream) {
    bb_step_kernel<true><<<bb_grid(E.n), BB_STEP_THREADS, 0, stream>>>(
        E, cfg, nullptr, n_steps, actions_out, rewards, terminated, mask_out, nullptr, nullptr, nullptr, stats);
    return cudaGetLastError();
}

cudaError_t bb_launch_reset(const BBEnvArrays& E, const uint8_t* reset_mask, uint64_t* mask_out, cudaStream_t stream) {
    bb_reset_kernel<<<bb_grid(E.n), BB_STEP_THREADS, 0, stream>>>(E, reset_mask, mask_out);
    return cudaGetLastError();
}

cudaError_t bb_launch_observe(const BBEnvArrays& E, uint64_t* board_out, uint32_t* pieces_out,
                              uint64_t* mask_out, cudaStream_t stream) {
    bb_observe_kernel<<<bb_grid(E.n), BB_STEP_THREADS, 0, stream>>>(E, board_out, pieces_out, mask_out);
    return cudaGetLastError();
}
