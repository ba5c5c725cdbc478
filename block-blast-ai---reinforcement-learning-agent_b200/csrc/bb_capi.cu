// bb_capi.cu — the extern "C" boundary declared in include/bbgpu.h.
// Plain pointers and sizes only; errors become return codes + a thread-local message.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "../../include/bbgpu.h"
#include "bb_kernels.h"

static thread_local char g_err[512] = "";

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
#define BB_CUDA(call, what) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(-2, what, e__); } while (0)

struct bb_env {
    BBEnvArrays arr;
    BBRewardCfg cfg;
    int device;
    // lazily allocated device staging for the host-buffer entry point
    int32_t* d_actions;
    float* d_rewards;
    uint8_t* d_terminated;
    uint64_t* d_board;
    uint32_t* d_pieces;
    uint64_t* d_mask;
    int32_t* d_ep_score;
    int32_t* d_ep_len;
    uint32_t* d_info;
    uint8_t* d_block;      // the eight result arrays above live in this one allocation (bb_env_host_layout)
};

static const double BB_DEFAULT_CFG[7] = {1.0, 0.01, -1.0, -0.05, 0.02, 0.5, 0.001};

extern "C" {

static int ensure_staging(bb_env* e);

int bb_version(void) { return BB_ABI_VERSION; }
const char* bb_last_error(void) { return g_err; }

int bb_piece_table(uint64_t* masks37, uint64_t* inb37, uint8_t* nblk37) {
    for (int i = 0; i < BB_NUM_PIECES; ++i) {
        if (masks37) masks37[i] = BB_HOST_PIECE_MASKS[i];
        if (inb37) inb37[i] = BB_HOST_PIECE_INB[i];
        if (nblk37) nblk37[i] = (uint8_t)BB_META_N(BB_HOST_PIECE_META[i]);
    }
    return 0;
}

int bb_env_create(bb_env** out, int64_t n_envs, uint64_t seed, int64_t global_env_offset,
                  const double reward_cfg[7], uint32_t flags) {
    if (!out) return fail(-1, "bb_env_create: out is NULL");
    *out = nullptr;
    if (n_envs <= 0) return fail(-1, "bb_env_create: n_envs must be > 0");
    if (global_env_offset < 0) return fail(-1, "bb_env_create: negative global_env_offset");
    bb_env* e = new (std::nothrow) bb_env();
    if (!e) return fail(-3, "bb_env_create: out of host memory");
    memset(e, 0, sizeof(*e));
    BB_CUDA(cudaGetDevice(&e->device), "cudaGetDevice");
    e->arr.n = n_envs;
    e->arr.env_offset = global_env_offset;
    e->arr.seed = seed;
    e->arr.flags = flags;
    const double* c = reward_cfg ? reward_cfg : BB_DEFAULT_CFG;
    e->cfg.line_clear_base = c[0]; e->cfg.block_placed = c[1]; e->cfg.game_over_penalty = c[2];
    e->cfg.hole_penalty = c[3]; e->cfg.center_bonus = c[4]; e->cfg.combo_multiplier_bonus = c[5];
    e->cfg.survival_bonus = c[6];
    cudaError_t err;
    const size_t bytes = (size_t)n_envs * sizeof(uint4);
    if ((err = cudaMalloc(&e->arr.s0, bytes)) != cudaSuccess || (err = cudaMalloc(&e->arr.s1, bytes)) != cudaSuccess ||
        (err = cudaMalloc(&e->arr.s2, bytes)) != cudaSuccess) {
        bb_env_destroy(e);
        return fail(-2, "bb_env_create: cudaMalloc state", err);
    }
    cudaMemset(e->arr.s0, 0, bytes);
    cudaMemset(e->arr.s1, 0, bytes);
    cudaMemset(e->arr.s2, 0, bytes);
    // the first reset (the reference constructs each GameEngine with a dealt trio)
    err = bb_launch_reset(e->arr, nullptr, nullptr, 0);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        bb_env_destroy(e);
        return fail(-2, "bb_env_create: reset kernel", err);
    }
    *out = e;
    return 0;
}

int bb_env_destroy(bb_env* e) {
    if (!e) return 0;
    cudaFree(e->arr.s0); cudaFree(e->arr.s1); cudaFree(e->arr.s2);
    cudaFree(e->d_actions); cudaFree(e->d_block);
    delete e;
    return 0;
}

int64_t bb_env_num_envs(const bb_env* e) { return e ? e->arr.n : -1; }

int bb_env_set_episode_end_buffer(bb_env* e, void* records) {
    if (!e) return fail(-1, "bb_env_set_episode_end_buffer: env is NULL");
    e->arr.ep_end = (BBEpisodeEnd*)records;
    return 0;
}

int bb_env_reset(bb_env* e, const uint8_t* reset_mask, uint64_t* mask_out, void* stream) {
    if (!e) return fail(-1, "bb_env_reset: env is NULL");
    BB_CUDA(bb_launch_reset(e->arr, reset_mask, mask_out, (cudaStream_t)stream), "bb_env_reset launch");
    return 0;
}

int bb_env_step(bb_env* e, const int32_t* actions, float* rewards, uint8_t* terminated,
                uint64_t* mask_out, int32_t* ep_score, int32_t* ep_len, uint32_t* info_out, void* stream) {
    if (!e) return fail(-1, "bb_env_step: env is NULL");
    if (!actions || !rewards || !terminated) return fail(-1, "bb_env_step: actions/rewards/terminated are required");
    BB_CUDA(bb_launch_step(e->arr, e->cfg, actions, rewards, terminated, mask_out, ep_score, ep_len, info_out,
                           (cudaStream_t)stream), "bb_env_step launch");
    return 0;
}

int bb_env_step_random(bb_env* e, int32_t n_steps, int32_t* actions_out, float* rewards, uint8_t* terminated,
                       uint64_t* mask_out, uint64_t* stats, void* stream) {
    if (!e) return fail(-1, "bb_env_step_random: env is NULL");
    if (n_steps < 1) return fail(-1, "bb_env_step_random: n_steps must be >= 1");
    BB_CUDA(bb_launch_step_random(e->arr, e->cfg, n_steps, 0, actions_out, rewards, terminated, mask_out,
                                  (unsigned long long*)stats, (cudaStream_t)stream), "bb_env_step_random launch");
    return 0;
}

int bb_env_rollout_random(bb_env* e, int32_t n_steps, int32_t* actions_out, float* rewards, uint8_t* terminated,
                          uint64_t* mask_out, uint64_t* stats, void* stream) {
    if (!e) return fail(-1, "bb_env_rollout_random: env is NULL");
    if (n_steps < 1) return fail(-1, "bb_env_rollout_random: n_steps must be >= 1");
    BB_CUDA(bb_launch_step_random(e->arr, e->cfg, n_steps, 1, actions_out, rewards, terminated, mask_out,
                                  (unsigned long long*)stats, (cudaStream_t)stream), "bb_env_rollout_random launch");
    return 0;
}

int bb_env_observe(bb_env* e, uint64_t* board_out, uint32_t* pieces_out, uint64_t* mask_out, void* stream) {
    if (!e) return fail(-1, "bb_env_observe: env is NULL");
    BB_CUDA(bb_launch_observe(e->arr, board_out, pieces_out, mask_out, (cudaStream_t)stream), "bb_env_observe launch");
    return 0;
}

int bb_env_sample_valid_actions(bb_env* e, uint64_t call_counter, int32_t* actions_out, int32_t* h_actions_out, void* stream) {
    if (!e) return fail(-1, "bb_env_sample_valid_actions: env is NULL");
    if (!actions_out && !h_actions_out) return fail(-1, "bb_env_sample_valid_actions: no output given");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* d = actions_out;
    if (!d) {
        if (int rc = ensure_staging(e)) return rc;
        d = e->d_actions;
    }
    BB_CUDA(bb_launch_sample_valid(e->arr, call_counter, d, s), "bb_env_sample_valid_actions launch");
    if (h_actions_out) {
        BB_CUDA(cudaMemcpyAsync(h_actions_out, d, e->arr.n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H actions");
        BB_CUDA(cudaStreamSynchronize(s), "bb_env_sample_valid_actions sync");
    }
    return 0;
}

int bb_env_get_state(bb_env* e, void* host_records, void* stream) {
    if (!e || !host_records) return fail(-1, "bb_env_get_state: NULL argument");
    const int64_t n = e->arr.n;
    uint4* tmp = (uint4*)malloc((size_t)n * 3 * sizeof(uint4));
    if (!tmp) return fail(-3, "bb_env_get_state: out of host memory");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemcpyAsync(tmp, e->arr.s0, n * sizeof(uint4), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(tmp + n, e->arr.s1, n * sizeof(uint4), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(tmp + 2 * n, e->arr.s2, n * sizeof(uint4), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) { free(tmp); return fail(-2, "bb_env_get_state copy", err); }
    uint4* rec = (uint4*)host_records;
    for (int64_t i = 0; i < n; ++i) { rec[3 * i] = tmp[i]; rec[3 * i + 1] = tmp[n + i]; rec[3 * i + 2] = tmp[2 * n + i]; }
    free(tmp);
    return 0;
}

int bb_env_set_state(bb_env* e, const void* host_records, void* stream) {
    if (!e || !host_records) return fail(-1, "bb_env_set_state: NULL argument");
    const int64_t n = e->arr.n;
    uint4* tmp = (uint4*)malloc((size_t)n * 3 * sizeof(uint4));
    if (!tmp) return fail(-3, "bb_env_set_state: out of host memory");
    const uint4* rec = (const uint4*)host_records;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t p = rec[3 * i].z;
        if ((p & 0xFF) >= BB_NUM_PIECES || ((p >> 8) & 0xFF) >= BB_NUM_PIECES || ((p >> 16) & 0xFF) >= BB_NUM_PIECES) {
            free(tmp);
            return fail(-1, "bb_env_set_state: piece index out of range");
        }
        tmp[i] = rec[3 * i]; tmp[n + i] = rec[3 * i + 1]; tmp[2 * n + i] = rec[3 * i + 2];
    }
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemcpyAsync(e->arr.s0, tmp, n * sizeof(uint4), cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->arr.s1, tmp + n, n * sizeof(uint4), cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(e->arr.s2, tmp + 2 * n, n * sizeof(uint4), cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    free(tmp);
    if (err != cudaSuccess) return fail(-2, "bb_env_set_state copy", err);
    return 0;
}

// Result arrays of the host-buffer step, as byte offsets into ONE block of `total` bytes:
// [0] mask u64[3][n], [1] board u64[n], [2] rewards f32[n], [3] pieces u32[n], [4] ep_score i32[n],
// [5] ep_len i32[n], [6] info u32[n], [7] terminated u8[n].  A caller whose eight host arrays sit
// in one pinned block at these offsets gets them with a single device-to-host copy.
int bb_env_host_layout(int64_t n_envs, int64_t offsets8[8], int64_t* total_bytes) {
    if (n_envs <= 0 || !offsets8 || !total_bytes) return fail(-1, "bb_env_host_layout: bad argument");
    const int64_t n = n_envs;
    offsets8[0] = 0; offsets8[1] = 24 * n; offsets8[2] = 32 * n; offsets8[3] = 36 * n; offsets8[4] = 40 * n;
    offsets8[5] = 44 * n; offsets8[6] = 48 * n; offsets8[7] = 52 * n;
    *total_bytes = 53 * n;
    return 0;
}

static int ensure_staging(bb_env* e) {
    if (e->d_actions) return 0;
    const int64_t n = e->arr.n;
    int64_t off[8], total;
    bb_env_host_layout(n, off, &total);
    cudaError_t err = cudaMalloc(&e->d_actions, n * sizeof(int32_t));
    if (err == cudaSuccess) err = cudaMalloc(&e->d_block, (size_t)total);
    if (err != cudaSuccess) return fail(-2, "bb_env_step_host: cudaMalloc staging", err);
    cudaMemset(e->d_block, 0, (size_t)total);
    e->d_mask = (uint64_t*)(e->d_block + off[0]);
    e->d_board = (uint64_t*)(e->d_block + off[1]);
    e->d_rewards = (float*)(e->d_block + off[2]);
    e->d_pieces = (uint32_t*)(e->d_block + off[3]);
    e->d_ep_score = (int32_t*)(e->d_block + off[4]);
    e->d_ep_len = (int32_t*)(e->d_block + off[5]);
    e->d_info = (uint32_t*)(e->d_block + off[6]);
    e->d_terminated = e->d_block + off[7];
    return 0;
}

int bb_env_step_host(bb_env* e, const int32_t* h_actions, float* h_rewards, uint8_t* h_terminated,
                     uint64_t* h_board, uint32_t* h_pieces, uint64_t* h_mask, int32_t* h_ep_score,
                     int32_t* h_ep_len, uint32_t* h_info, void* stream) {
    if (!e) return fail(-1, "bb_env_step_host: env is NULL");
    if (!h_actions || !h_rewards || !h_terminated) return fail(-1, "bb_env_step_host: actions/rewards/terminated are required");
    if (int rc = ensure_staging(e)) return rc;
    const int64_t n = e->arr.n;
    cudaStream_t s = (cudaStream_t)stream;
    BB_CUDA(cudaMemcpyAsync(e->d_actions, h_actions, n * sizeof(int32_t), cudaMemcpyHostToDevice, s), "H2D actions");
    // the step kernel writes the packed next observation (board, pieces) itself
    BB_CUDA(bb_launch_step_range(e->arr, e->cfg, 0, n, e->d_actions, e->d_rewards, e->d_terminated,
                                 h_mask ? e->d_mask : nullptr, h_ep_score ? e->d_ep_score : nullptr,
                                 h_ep_len ? e->d_ep_len : nullptr, h_info ? e->d_info : nullptr,
                                 h_board ? e->d_board : nullptr, h_pieces ? e->d_pieces : nullptr, s),
            "bb_env_step_host launch");
    int64_t off[8], total;
    bb_env_host_layout(n, off, &total);
    uint8_t* hb = (uint8_t*)h_mask;
    const bool one_block = hb && (uint8_t*)h_board == hb + off[1] && (uint8_t*)h_rewards == hb + off[2] &&
                           (uint8_t*)h_pieces == hb + off[3] && (uint8_t*)h_ep_score == hb + off[4] &&
                           (uint8_t*)h_ep_len == hb + off[5] && (uint8_t*)h_info == hb + off[6] && h_terminated == hb + off[7];
    if (one_block) {
        // one 53 B/env transfer instead of eight calls (each costs several microseconds of driver time)
        BB_CUDA(cudaMemcpyAsync(hb, e->d_block, (size_t)total, cudaMemcpyDeviceToHost, s), "D2H results");
    } else {
        BB_CUDA(cudaMemcpyAsync(h_rewards, e->d_rewards, n * sizeof(float), cudaMemcpyDeviceToHost, s), "D2H rewards");
        BB_CUDA(cudaMemcpyAsync(h_terminated, e->d_terminated, n, cudaMemcpyDeviceToHost, s), "D2H terminated");
        if (h_board) BB_CUDA(cudaMemcpyAsync(h_board, e->d_board, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s), "D2H board");
        if (h_pieces) BB_CUDA(cudaMemcpyAsync(h_pieces, e->d_pieces, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s), "D2H pieces");
        if (h_mask) BB_CUDA(cudaMemcpyAsync(h_mask, e->d_mask, 3 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s), "D2H mask");
        if (h_ep_score) BB_CUDA(cudaMemcpyAsync(h_ep_score, e->d_ep_score, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H ep_score");
        if (h_ep_len) BB_CUDA(cudaMemcpyAsync(h_ep_len, e->d_ep_len, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s), "D2H ep_len");
        if (h_info) BB_CUDA(cudaMemcpyAsync(h_info, e->d_info, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s), "D2H info");
    }
    BB_CUDA(cudaStreamSynchronize(s), "bb_env_step_host sync");
    return 0;
}

int bb_unpack_obs(const uint64_t* board, const uint32_t* pieces, const uint64_t* mask, int64_t mask_stride,
                  void* obs_nchw, int obs_dtype, void* mask_dense, int mask_dtype, int64_t n, void* stream) {
    if (n < 0) return fail(-1, "bb_unpack_obs: negative n");
    if (obs_nchw && obs_dtype != BB_F32 && obs_dtype != BB_BF16) return fail(-1, "bb_unpack_obs: obs_dtype must be BB_F32 or BB_BF16");
    if (mask_dense && mask_dtype != BB_F32 && mask_dtype != BB_U8) return fail(-1, "bb_unpack_obs: mask_dtype must be BB_F32 or BB_U8");
    if (n == 0) return 0;
    if (obs_nchw && (!board || !pieces)) return fail(-1, "bb_unpack_obs: board/pieces required for obs");
    if (obs_nchw && obs_dtype != BB_F32 && obs_dtype != BB_BF16) return fail(-1, "bb_unpack_obs: obs_dtype must be BB_F32 or BB_BF16");
    if (mask_dense && !mask) return fail(-1, "bb_unpack_obs: mask required for mask_dense");
    if (mask_dense && mask_dtype != BB_F32 && mask_dtype != BB_U8) return fail(-1, "bb_unpack_obs: mask_dtype must be BB_F32 or BB_U8");
    BB_CUDA(bb_launch_unpack_obs(board, pieces, mask, mask_stride, obs_nchw, obs_dtype, mask_dense, mask_dtype, n,
                                 (cudaStream_t)stream), "bb_unpack_obs launch");
    return 0;
}

int bb_masked_sample(const void* logits, int logits_dtype, const uint64_t* mask, int64_t mask_stride,
                     uint64_t seed, uint64_t call_counter, int mode, int32_t* action, float* logp,
                     float* entropy, int64_t n, void* stream) {
    if (n < 0) return fail(-1, "bb_masked_sample: negative n");
    if (logits_dtype != BB_F32 && logits_dtype != BB_BF16) return fail(-1, "bb_masked_sample: logits_dtype must be BB_F32 or BB_BF16");
    if (mode < 0 || mode > 2) return fail(-1, "bb_masked_sample: mode must be 0, 1 or 2");
    if (n == 0) return 0;
    if (!logits || !mask || !action) return fail(-1, "bb_masked_sample: logits/mask/action are required");
    if (logits_dtype != BB_F32 && logits_dtype != BB_BF16) return fail(-1, "bb_masked_sample: logits_dtype must be BB_F32 or BB_BF16");
    if (mode < 0 || mode > 2) return fail(-1, "bb_masked_sample: mode must be 0, 1 or 2");
    BB_CUDA(bb_launch_masked_sample(logits, logits_dtype, mask, mask_stride, seed, call_counter, mode, action, logp,
                                    entropy, n, (cudaStream_t)stream), "bb_masked_sample launch");
    return 0;
}

int bb_masked_head_backward(const void* logits, int logits_dtype, const uint64_t* mask, int64_t mask_stride,
                            const int32_t* action, const float* grad_logp, const float* grad_entropy,
                            void* grad_logits, int64_t n, void* stream) {
    if (n < 0) return fail(-1, "bb_masked_head_backward: negative n");
    if (logits_dtype != BB_F32 && logits_dtype != BB_BF16) return fail(-1, "bb_masked_head_backward: logits_dtype must be BB_F32 or BB_BF16");
    if (n == 0) return 0;
    if (!logits || !mask || !action || !grad_logp || !grad_logits) return fail(-1, "bb_masked_head_backward: NULL array");
    BB_CUDA(bb_launch_masked_head_bwd(logits, logits_dtype, mask, mask_stride, action, grad_logp, grad_entropy, grad_logits,
                                      n, (cudaStream_t)stream), "bb_masked_head_backward launch");
    return 0;
}

int bb_ppo_loss(const void* logits, int logits_dtype, const uint64_t* mask, int64_t mask_stride,
                const int32_t* action, const float* old_logp, const float* advantages, const float* returns,
                const float* values, double clip_epsilon, double value_coef, double entropy_coef,
                void* grad_logits, float* grad_values, double* sums5, int64_t n, void* stream) {
    if (n < 0) return fail(-1, "bb_ppo_loss: negative n");
    if (logits_dtype != BB_F32 && logits_dtype != BB_BF16) return fail(-1, "bb_ppo_loss: logits_dtype must be BB_F32 or BB_BF16");
    if (n == 0) return 0;
    if (!logits || !mask || !action || !old_logp || !advantages || !returns || !values || !grad_logits || !grad_values || !sums5)
        return fail(-1, "bb_ppo_loss: NULL array");
    BB_CUDA(bb_launch_ppo_loss(logits, logits_dtype, mask, mask_stride, action, old_logp, advantages, returns, values,
                               (float)clip_epsilon, (float)value_coef, (float)entropy_coef, grad_logits, grad_values,
                               sums5, n, (cudaStream_t)stream), "bb_ppo_loss launch");
    return 0;
}

int64_t bb_bn_workspace_size(int channels) {
    if (channels <= 0 || channels % 8 != 0 || channels > 2048) return -1;
    return (int64_t)bb_bn_workspace_floats(channels);
}

static int bn_check(int64_t rows, int channels) {
    if (rows < 0) return fail(-1, "bb_bn_relu: negative rows");
    if (channels <= 0 || channels % 8 != 0 || channels > 2048) return fail(-1, "bb_bn_relu: channels must be a multiple of 8 in [8, 2048]");
    return 0;
}

int bb_bn_relu_forward(const void* x, const void* skip, const float* gamma, const float* beta, const float* pre_bias,
                       float* running_mean,
                       float* running_var, double momentum, double eps, int training, void* y, float* save_mean,
                       float* save_rstd, float* workspace, int64_t rows, int channels, void* stream) {
    if (bn_check(rows, channels)) return -1;
    if (rows == 0) return 0;
    if (!x || !gamma || !beta || !y || !workspace) return fail(-1, "bb_bn_relu_forward: NULL array");
    if (training && (!save_mean || !save_rstd)) return fail(-1, "bb_bn_relu_forward: training needs save_mean / save_rstd");
    if (!training && (!running_mean || !running_var)) return fail(-1, "bb_bn_relu_forward: eval needs the running statistics");
    BB_CUDA(bb_launch_bn_relu_fwd(x, skip, gamma, beta, pre_bias, running_mean, running_var, (float)momentum, (float)eps, training, y,
                                  save_mean, save_rstd, workspace, rows, channels, (cudaStream_t)stream), "bb_bn_relu_forward launch");
    return 0;
}

int bb_bn_relu_backward(const void* x, const void* y, const void* grad_y, const float* gamma, const float* save_mean,
                        const float* save_rstd, void* grad_x, void* grad_skip, float* grad_gamma, float* grad_beta,
                        float* workspace, int64_t rows, int channels, void* stream) {
    if (bn_check(rows, channels)) return -1;
    if (rows == 0) return 0;
    if (!x || !y || !grad_y || !gamma || !save_mean || !save_rstd || !grad_x || !grad_gamma || !grad_beta || !workspace)
        return fail(-1, "bb_bn_relu_backward: NULL array");
    BB_CUDA(bb_launch_bn_relu_bwd(x, y, grad_y, gamma, save_mean, save_rstd, grad_x, grad_skip, grad_gamma, grad_beta, workspace,
                                  rows, channels, (cudaStream_t)stream), "bb_bn_relu_backward launch");
    return 0;
}

int bb_gae(const float* rewards, const float* values, const float* dones, const float* last_values,
           double gamma, double lam, float* adv, float* ret, double* moments, int64_t T, int64_t N, void* stream) {
    if (T < 0 || N < 0) return fail(-1, "bb_gae: negative size");
    if (T == 0 || N == 0) return 0;                 // empty rollout: nothing to do (pointers may be NULL)
    if (!rewards || !values || !dones || !last_values || !adv || !ret) return fail(-1, "bb_gae: NULL array");
    // numpy casts the Python doubles gamma and gamma*lambda to float32 before the array multiply
    BB_CUDA(bb_launch_gae(rewards, values, dones, last_values, (float)gamma, (float)(gamma * lam), adv, ret, moments,
                          T, N, (cudaStream_t)stream), "bb_gae launch");
    return 0;
}

}  // extern "C"
