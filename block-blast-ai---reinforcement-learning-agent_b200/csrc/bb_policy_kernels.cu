// bb_policy_kernels.cu — K2 (packed obs -> dense NCHW planes / dense mask) and K3 (fused
// -inf masking + softmax + categorical sample + log-prob + masked entropy) for sm_100a.
//
// K2 replaces engine.get_observation + Piece.to_mask (src/game/engine.py:489-507,
// src/game/pieces.py:39-45) and the torch.cat at src/models/network.py:152-158:
// 36 B in, 1,024 B (f32) / 512 B (bf16) out per env; a warp writes one env's 4x8x8 block as
// one contiguous run of 128-bit stores.
//
// K3 replaces src/models/network.py:172-262 (mask -> softmax -> Categorical sample/log_prob ->
// masked entropy): 8 lanes per row, 768 B of logits read once with 128-bit loads (128 B
// contiguous per row group per instruction), everything else in registers / 3-step butterflies.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include "bb_kernels.h"

__constant__ uint64_t c_piece_cells[BB_NUM_PIECES + 3] = BB_PIECE_MASKS;

// ------------------------------------------------------------------------------------ K2
// One item = 16 output bytes, so consecutive lanes always store consecutive 16-byte pieces
// (full 32-byte sectors per instruction): bf16 item = (env, channel, row) -> 8 values,
// f32 item = (env, channel, row, half) -> 4 values.
template <bool BF16>
__device__ __forceinline__ uint64_t bb_obs_plane(const uint64_t* __restrict__ board, const uint32_t* __restrict__ pieces,
                                                 const uint64_t* cells, int64_t t, const int64_t* __restrict__ index = nullptr) {
    int64_t env = t >> (BF16 ? 5 : 6);
    if (index) env = __ldg(index + env);          // minibatch gather: output row -> rollout-buffer sample
    const int ch = (int)(t >> (BF16 ? 3 : 4)) & 3;
    if (ch == 0) return __ldg(board + env);
    const uint32_t pw = __ldg(pieces + env);
    const uint32_t id = (pw >> (8 * (ch - 1))) & 0xFFu;
    const bool used = (pw >> (24 + ch - 1)) & 1u;
    return used ? 0ull : cells[id < BB_NUM_PIECES ? id : BB_NUM_PIECES];
}

// split != NULL (f32 only): the reference's separate arrays — board planes f32[n][8][8] at `obs`, piece
// planes f32[n][3][8][8] at `split` (engine.get_observation, engine.py:489-507) — instead of one
// [n][4][8][8] block
template <bool BF16>
__device__ __forceinline__ void bb_unpack_obs_item(void* __restrict__ obs, int64_t t, uint64_t plane_bits,
                                                   void* __restrict__ split = nullptr) {
    if (!BF16 && split) {
        const int64_t env = t >> 6;
        const int ch = (int)(t >> 4) & 3, q = (int)t & 15;
        const uint32_t bits = (uint32_t)(plane_bits >> (4 * q)) & 0xFu;
        float4 v;
        v.x = (bits & 1u) ? 1.f : 0.f; v.y = (bits & 2u) ? 1.f : 0.f;
        v.z = (bits & 4u) ? 1.f : 0.f; v.w = (bits & 8u) ? 1.f : 0.f;
        if (ch == 0) __stcs(reinterpret_cast<float4*>(obs) + env * 16 + q, v);
        else __stcs(reinterpret_cast<float4*>(split) + env * 48 + (ch - 1) * 16 + q, v);
        return;
    }
    if (BF16) {
        const uint32_t bits = (uint32_t)(plane_bits >> (8 * ((int)t & 7))) & 0xFFu;
        uint32_t w[4];                              // bf16 1.0 = 0x3F80
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (((bits >> (2 * k)) & 1u) ? 0x3F80u : 0u) | (((bits >> (2 * k + 1)) & 1u) ? 0x3F800000u : 0u);
        __stcs(reinterpret_cast<uint4*>(obs) + t, make_uint4(w[0], w[1], w[2], w[3]));   // streaming: written once
    } else {
        const uint32_t bits = (uint32_t)(plane_bits >> (4 * ((int)t & 15))) & 0xFu;   // row*8 + half*4
        float4 v;
        v.x = (bits & 1u) ? 1.f : 0.f; v.y = (bits & 2u) ? 1.f : 0.f;
        v.z = (bits & 4u) ? 1.f : 0.f; v.w = (bits & 8u) ? 1.f : 0.f;
        __stcs(reinterpret_cast<float4*>(obs) + t, v);
    }
}

// Grid-stride: a few thousand long-lived blocks instead of one short block per 8 envs (CTA
// launch rate, not HBM, bounded the one-shot version); two independent items per iteration.
template <bool BF16>
__global__ void __launch_bounds__(256)
bb_unpack_obs_kernel(const uint64_t* __restrict__ board, const uint32_t* __restrict__ pieces,
                     void* __restrict__ obs, int64_t n, void* __restrict__ split, const int64_t* __restrict__ index) {
    __shared__ uint64_t cells[BB_NUM_PIECES + 3];
    for (int k = threadIdx.x; k < BB_NUM_PIECES + 3; k += blockDim.x) cells[k] = c_piece_cells[k];
    __syncthreads();
    const int64_t total = n * (BF16 ? 32 : 64);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t + 3 * stride < total; t += 4 * stride) {
        const uint64_t p0 = bb_obs_plane<BF16>(board, pieces, cells, t, index);
        const uint64_t p1 = bb_obs_plane<BF16>(board, pieces, cells, t + stride, index);
        const uint64_t p2 = bb_obs_plane<BF16>(board, pieces, cells, t + 2 * stride, index);
        const uint64_t p3 = bb_obs_plane<BF16>(board, pieces, cells, t + 3 * stride, index);
        bb_unpack_obs_item<BF16>(obs, t, p0, split);
        bb_unpack_obs_item<BF16>(obs, t + stride, p1, split);
        bb_unpack_obs_item<BF16>(obs, t + 2 * stride, p2, split);
        bb_unpack_obs_item<BF16>(obs, t + 3 * stride, p3, split);
    }
    for (; t < total; t += stride) bb_unpack_obs_item<BF16>(obs, t, bb_obs_plane<BF16>(board, pieces, cells, t, index), split);
}

template <bool F32>
__global__ void __launch_bounds__(256)
bb_unpack_mask_kernel(const uint64_t* __restrict__ mask, int64_t stride, void* __restrict__ dense, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (env, group of 16 actions)
    const int64_t env = t / 12;
    if (env >= n) return;
    const int g = (int)(t - env * 12);
    const uint64_t w = __ldg(mask + (int64_t)(g >> 2) * stride + env);
    const uint32_t bits = (uint32_t)(w >> (16 * (g & 3))) & 0xFFFFu;
    if (F32) {
        float4* o = reinterpret_cast<float4*>(dense) + 4 * t;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 v;
            v.x = ((bits >> (4 * k)) & 1u) ? 1.f : 0.f;
            v.y = ((bits >> (4 * k + 1)) & 1u) ? 1.f : 0.f;
            v.z = ((bits >> (4 * k + 2)) & 1u) ? 1.f : 0.f;
            v.w = ((bits >> (4 * k + 3)) & 1u) ? 1.f : 0.f;
            o[k] = v;
        }
    } else {
        uint32_t w4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t nib = (bits >> (4 * k)) & 0xFu;
            w4[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
        }
        reinterpret_cast<uint4*>(dense)[t] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

static void bb_launch_obs_planes(const uint64_t* board, const uint32_t* pieces, void* obs, int obs_dtype, int64_t n,
                                 void* split, const int64_t* index, cudaStream_t stream) {
    const int64_t threads = n * (obs_dtype == 1 ? 32 : 64);
    int64_t blocks = (threads + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;        // grid-stride: 8 resident blocks per SM x 2 waves
    const unsigned grid = (unsigned)blocks;
    if (obs_dtype == 1) bb_unpack_obs_kernel<true><<<grid, 256, 0, stream>>>(board, pieces, obs, n, nullptr, index);
    else bb_unpack_obs_kernel<false><<<grid, 256, 0, stream>>>(board, pieces, obs, n, split, index);
}

cudaError_t bb_launch_unpack_obs(const uint64_t* board, const uint32_t* pieces, const uint64_t* mask,
                                 int64_t mask_stride, void* obs_nchw, int obs_dtype, void* mask_dense,
                                 int mask_dtype, int64_t n, cudaStream_t stream, void* pieces_split) {
    if (n <= 0) return cudaSuccess;
    if (obs_nchw) bb_launch_obs_planes(board, pieces, obs_nchw, obs_dtype, n, pieces_split, nullptr, stream);
    if (mask_dense) {
        const int64_t threads = n * 12;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        if (mask_dtype == 0) bb_unpack_mask_kernel<true><<<grid, 256, 0, stream>>>(mask, mask_stride, mask_dense, n);
        else bb_unpack_mask_kernel<false><<<grid, 256, 0, stream>>>(mask, mask_stride, mask_dense, n);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ minibatch gather
// RolloutBuffer.get_samples (src/agents/ppo.py:171-213): the reference gathers seven arrays of
// 1.8 KB rows with numpy fancy indexing and uploads them.  Here sample index[b] of the packed
// device-resident buffer (board u64, pieces u32, mask planes [T][3][N], action, log-prob,
// advantage, return; T*N samples, flat index t*N + e) becomes row b of the minibatch: the obs
// planes are expanded by the K2 kernel reading through the index, the scalars and the mask planes
// [3][B] by one thread per row; advantages are normalised on the way, (a - mean) / (std + 1e-8) in
// float32 like ppo.py:196.
__global__ void __launch_bounds__(256)
bb_gather_rows_kernel(const int64_t* __restrict__ index, int64_t B, int64_t N, const uint64_t* __restrict__ mask,
                      const int32_t* __restrict__ action, const float* __restrict__ logp, const float* __restrict__ adv,
                      const float* __restrict__ ret, const float* __restrict__ mean_std, uint64_t* __restrict__ mask_out,
                      int32_t* __restrict__ action_out, float* __restrict__ logp_out, float* __restrict__ adv_out,
                      float* __restrict__ ret_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t j = index[b];
    const int64_t t = j / N, e = j - t * N;
    const uint64_t* m = mask + t * 3 * N + e;
    mask_out[b] = m[0]; mask_out[B + b] = m[N]; mask_out[2 * B + b] = m[2 * N];
    action_out[b] = action[j];
    logp_out[b] = logp[j];
    ret_out[b] = ret[j];
    const float a = adv[j];
    adv_out[b] = mean_std ? __fdiv_rn(__fsub_rn(a, mean_std[0]), __fadd_rn(mean_std[1], 1e-8f)) : a;
}

cudaError_t bb_launch_gather_minibatch(const int64_t* index, int64_t B, int64_t N, const uint64_t* board,
                                       const uint32_t* pieces, const uint64_t* mask, const int32_t* action,
                                       const float* logp, const float* adv, const float* ret, const float* mean_std,
                                       void* obs, int obs_dtype, uint64_t* mask_out, int32_t* action_out, float* logp_out,
                                       float* adv_out, float* ret_out, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    bb_launch_obs_planes(board, pieces, obs, obs_dtype, B, nullptr, index, stream);
    bb_gather_rows_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(index, B, N, mask, action, logp, adv, ret, mean_std,
                                                                          mask_out, action_out, logp_out, adv_out, ret_out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ K3
// 8 lanes per row, 4 rows per warp.  Lane l of a row group owns the 24 actions
//     a(l,k,c) = 32k + 4l + c        k = 0..5, c = 0..3
// i.e. the k-th 128-bit load of the group is one contiguous 128 B (f32) run.  Everything else
// is per-lane arithmetic plus 3-step butterflies inside the 8-lane group:
//   p_i = exp(z_i - m) / S            (exp2 on pre-scaled inputs)
//   log_prob = log(clamp(p_a / sum(p), eps, 1-eps))          torch Categorical (network.py:213-225)
//   entropy  = log S - sum_i e_i (z_i - m) / S               == -sum q log q (network.py:246-260);
//              the reference's clamps at 1e-10 change it by < 1e-8
// Sampling is the inverse CDF of u * sum(p) over the order (lane, k, c) — a fixed permutation
// of the actions, so the draw is an exact categorical sample; philox.py documents the order.
#define K3_LOG2E 1.4426950408889634f
#define K3_DMIN (-150.0f)     // exp2(-150 * log2e) == 0 in float32: masked and hopeless actions alike

__device__ __forceinline__ float k3_ex2(float x) {      // 2^x, one MUFU (rel. error 2^-22)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float grp_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float grp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

template <bool BF16>
__global__ void __launch_bounds__(128)
bb_masked_sample_kernel(const void* __restrict__ logits, const uint64_t* __restrict__ mask, int64_t stride,
                        uint64_t seed, uint64_t call_counter, int mode, int32_t* __restrict__ action,
                        float* __restrict__ logp_out, float* __restrict__ ent_out, int64_t n, int64_t row_offset,
                        const uint64_t* __restrict__ counter_dev) {
    const int lane = threadIdx.x & 31;
    const int l = lane & 7;                       // lane inside the row group
    const int64_t row_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = row_raw < n;
    const int64_t row = live ? row_raw : n - 1;   // out-of-range groups shadow the last row, never store

    float z[24];
    uint32_t mb = 0;                              // bit (4k + c) = mask of a(l,k,c)
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint64_t w = __ldg(mask + (int64_t)p * stride + row);
        mb |= ((uint32_t)(w >> (4 * l)) & 0xFu) << (8 * p);
        mb |= ((uint32_t)(w >> (32 + 4 * l)) & 0xFu) << (8 * p + 4);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float4 v;
        if (BF16) {
            const uint2 h = reinterpret_cast<const uint2*>(logits)[row * 48 + k * 8 + l];
            v.x = __uint_as_float(h.x << 16); v.y = __uint_as_float(h.x & 0xFFFF0000u);
            v.z = __uint_as_float(h.y << 16); v.w = __uint_as_float(h.y & 0xFFFF0000u);
        } else {
            v = reinterpret_cast<const float4*>(logits)[row * 48 + k * 8 + l];
        }
        z[4 * k + 0] = v.x; z[4 * k + 1] = v.y; z[4 * k + 2] = v.z; z[4 * k + 3] = v.w;
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        z[i] = ((mb >> i) & 1u) ? z[i] : -INFINITY;               // -inf masking (network.py:175-180)
        m = fmaxf(m, z[i]);
    }
    m = grp_max(m);
    const bool any_valid = m > -INFINITY;
    const float mm = any_valid ? m : 0.f;
    float s = 0.f, sz = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const float d = fmaxf(z[i] - mm, K3_DMIN);                 // <= 0; clamped so that 0 * d stays 0
        const float e = k3_ex2(d * K3_LOG2E);                      // exactly 0 for masked actions
        s += e;
        sz = fmaf(e, d, sz);
        z[i] = e;                                                  // z now holds exp(z - m)
    }
    s = grp_sum(s);
    sz = grp_sum(sz);
    const float inv = any_valid ? __frcp_rn(s) : 0.f;
    float ps = 0.f;                                                // sum of the rounded probabilities
#pragma unroll
    for (int i = 0; i < 24; ++i) { z[i] *= inv; ps += z[i]; }     // z now holds p
    const float lane_tot = ps;
    ps = grp_sum(ps);

    int act = 0;
    float pa = 0.f;
    if (mode == 2) {
        act = action[row];
        // owner lane / slot of action a: k = a / 32, l = (a % 32) / 4, c = a % 4
        float mine = 0.f;
        const int slot = ((act >> 5) << 2) | (act & 3);
        const bool own = act >= 0 && act < 192 && ((act & 31) >> 2) == l;
#pragma unroll
        for (int i = 0; i < 24; ++i) mine += (own && i == slot) ? z[i] : 0.f;
        pa = grp_sum(mine);
    } else if (mode == 1) {
        // argmax of probs, lowest action index on ties (torch.argmax)
        float best = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            const int idx = 32 * (i >> 2) + 4 * l + (i & 3);
            if (z[i] > best || (z[i] == best && idx < bi)) { best = z[i]; bi = idx; }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        act = any_valid ? bi : 0;
        pa = any_valid ? best : 0.f;
    } else {
        // keyed by the GLOBAL row id (env shards on several GPUs draw independent noise) and by a call
        // counter that may live on the device (a CUDA graph replays the launch with a new value)
        const uint64_t grow = (uint64_t)(row + row_offset);
        const uint64_t ctr = call_counter + (counter_dev ? *counter_dev : 0ull);
        const BBPhilox4 r = bb_philox((uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)ctr,
                                      BB_STREAM_SAMPLE, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
        const float t = u * ps;
        // exclusive prefix of the lane totals inside the group
        float incl = lane_tot;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const float o = __shfl_up_sync(0xffffffffu, incl, d, 8);
            if (l >= d) incl += o;
        }
        const float excl = incl - lane_tot;
        // first lane whose inclusive prefix exceeds t (and that has any probability mass)
        const unsigned hit = (__ballot_sync(0xffffffffu, incl > t && lane_tot > 0.f) >> (lane & 24)) & 0xFFu;
        const unsigned nz = (__ballot_sync(0xffffffffu, lane_tot > 0.f) >> (lane & 24)) & 0xFFu;
        // rounding can leave t >= total: fall back to the last lane with mass
        const int L = hit ? (__ffs((int)hit) - 1) : (nz ? (31 - __clz((int)nz)) : 0);
        // inside lane L: first element whose running sum exceeds t, else its last non-zero one;
        // first the 4-action chunk, then the action inside it
        float c4[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) c4[k] = (z[4 * k] + z[4 * k + 1]) + (z[4 * k + 2] + z[4 * k + 3]);
        float run = excl;
        int kc = -1, klast = 0;
        float runc = excl, runlast = excl;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (c4[k] > 0.f) { klast = k; runlast = run; if (kc < 0 && run + c4[k] > t) { kc = k; runc = run; } }
            run += c4[k];
        }
        if (kc < 0) { kc = klast; runc = runlast; }
        float e4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            e4[c] = z[c];
#pragma unroll
            for (int k = 1; k < 6; ++k) e4[c] = kc == k ? z[4 * k + c] : e4[c];
        }
        int pick = -1, lastnz = 0;
        float ppick = 0.f, plast = 0.f, r2 = runc;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            r2 += e4[c];
            if (e4[c] > 0.f) { lastnz = c; plast = e4[c]; if (pick < 0 && r2 > t) { pick = c; ppick = e4[c]; } }
        }
        if (pick < 0) { pick = lastnz; ppick = plast; }
        pick += 4 * kc;
        const int idx = 32 * (pick >> 2) + 4 * l + (pick & 3);
        const int src = (lane & 24) | L;
        act = __shfl_sync(0xffffffffu, idx, src);
        pa = __shfl_sync(0xffffffffu, ppick, src);
        if (!any_valid) { act = 0; pa = 0.f; }
    }
    if (live && l == 0) {
        if (mode != 2) action[row] = act;
        if (logp_out) {
            const float eps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
            const float pn = any_valid ? pa / ps : 1.f;
            logp_out[row] = logf(fminf(fmaxf(pn, eps), 1.f - eps));
        }
        if (ent_out) ent_out[row] = any_valid ? (logf(s) - sz * inv) : 0.f;
    }
}

cudaError_t bb_launch_masked_sample(const void* logits, int logits_dtype, const uint64_t* mask,
                                    int64_t mask_stride, uint64_t seed, uint64_t call_counter, int mode,
                                    int32_t* action, float* logp, float* entropy, int64_t n, cudaStream_t stream,
                                    int64_t row_offset, const uint64_t* counter_dev) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n * 8 + 127) / 128);
    if (logits_dtype == 1)
        bb_masked_sample_kernel<true><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, seed, call_counter, mode, action, logp, entropy, n, row_offset, counter_dev);
    else
        bb_masked_sample_kernel<false><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, seed, call_counter, mode, action, logp, entropy, n, row_offset, counter_dev);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ K3 backward
// Gradient of (log_prob[action], masked entropy) w.r.t. the raw logits, for the PPO update
// (autograd of src/models/network.py:210-262 as used at src/agents/ppo.py:366-392):
//   d log_prob / d z_i = [eps <= p_a <= 1-eps] (delta_ia - p_i)        (Categorical clamps p)
//   d entropy  / d z_i = -p_i (log p_i + H)
// for valid actions i, 0 for masked ones.  Same 8-lanes-per-row layout as the forward; p is
// recomputed from the logits (768 B read + 768 B written per row, nothing saved between passes).
//
// LOSS = true turns the same pass into the whole PPO loss tail (src/agents/ppo.py:366-395 +
// network.py:210-262): with old log-probs, advantages, returns and the value
// head's output it forms ratio / clipped surrogate / value MSE / entropy bonus for the row,
// writes d(loss)/d(logits) and d(loss)/d(value) of
//     loss = mean(-min(r A, clip(r) A)) + value_coef mean((v - R)^2) - entropy_coef mean(H)
// directly (the loss is a mean with constant coefficients, so no second pass is needed), and adds
// the row's terms to five double sums: -min(..), (v-R)^2, H, (r-1) - log r, [|r-1| > clip].
struct BBPpoLossArgs {
    const float* old_logp; const float* adv; const float* ret; const float* value;
    float* dvalue; double* sums;
    float clip, value_coef, entropy_coef, inv_n;
};

template <bool BF16, bool LOSS>
__global__ void __launch_bounds__(128)
bb_masked_head_bwd_kernel(const void* __restrict__ logits, const uint64_t* __restrict__ mask, int64_t stride,
                          const int32_t* __restrict__ action, const float* __restrict__ g_logp,
                          const float* __restrict__ g_ent, void* __restrict__ dlogits, int64_t n, BBPpoLossArgs L) {
    const int lane = threadIdx.x & 31;
    const int l = lane & 7;
    const int64_t row_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = row_raw < n;
    const int64_t row = live ? row_raw : n - 1;
    float z[24];
    uint32_t mb = 0;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const uint64_t w = __ldg(mask + (int64_t)p * stride + row);
        mb |= ((uint32_t)(w >> (4 * l)) & 0xFu) << (8 * p);
        mb |= ((uint32_t)(w >> (32 + 4 * l)) & 0xFu) << (8 * p + 4);
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        float4 v;
        if (BF16) {
            const uint2 h = reinterpret_cast<const uint2*>(logits)[row * 48 + k * 8 + l];
            v.x = __uint_as_float(h.x << 16); v.y = __uint_as_float(h.x & 0xFFFF0000u);
            v.z = __uint_as_float(h.y << 16); v.w = __uint_as_float(h.y & 0xFFFF0000u);
        } else {
            v = reinterpret_cast<const float4*>(logits)[row * 48 + k * 8 + l];
        }
        z[4 * k + 0] = v.x; z[4 * k + 1] = v.y; z[4 * k + 2] = v.z; z[4 * k + 3] = v.w;
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 24; ++i) m = fmaxf(m, ((mb >> i) & 1u) ? z[i] : -INFINITY);
    m = grp_max(m);
    const bool any_valid = m > -INFINITY;
    float s = 0.f, sz = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const bool ok = (mb >> i) & 1u;
        const float d = ok ? z[i] - m : 0.f;
        const float e = ok ? exp2f(d * K3_LOG2E) : 0.f;
        s += e;
        sz = fmaf(e, d, sz);
        z[i] = d;                                   // keep z - m (0 where masked)
    }
    s = grp_sum(s);
    sz = grp_sum(sz);
    const float inv = any_valid ? __frcp_rn(s) : 0.f;
    const float logS = any_valid ? logf(s) : 0.f;
    const float H = logS - sz * inv;                // masked entropy
    const int act = action[row];
    const int slot = ((act >> 5) << 2) | (act & 3);
    const bool own = act >= 0 && act < 192 && ((act & 31) >> 2) == l;
    // p_a for the clamp indicator
    float mine = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) mine += (own && i == slot && ((mb >> i) & 1u)) ? exp2f(z[i] * K3_LOG2E) * inv : 0.f;
    const float pa = grp_sum(mine);
    const float eps = 1.1920928955078125e-07f;
    float g1, g2;
    if (LOSS) {
        const float logp = any_valid ? logf(fminf(fmaxf(pa, eps), 1.f - eps)) : 0.f;
        const float lr = logp - L.old_logp[row];
        const float ratio = expf(lr);
        const float A = L.adv[row];
        const float s1 = ratio * A, s2 = fminf(fmaxf(ratio, 1.f - L.clip), 1.f + L.clip) * A;
        // d min(s1, s2) / d ratio: A where the unclipped term is the minimum (ties included: inside
        // the clip range both terms are the same function), 0 where the clipped constant is
        const bool through = s1 < s2 || (ratio >= 1.f - L.clip && ratio <= 1.f + L.clip);
        const float dv = L.value[row] - L.ret[row];
        g1 = (through && pa >= eps && pa <= 1.f - eps) ? -A * ratio * L.inv_n : 0.f;
        g2 = -L.entropy_coef * L.inv_n;
        // per-row metric terms: lane 0 of each 8-lane group holds them, summed per warp, one
        // atomic per warp and sum
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
        if (live && l == 0) {
            L.dvalue[row] = 2.f * L.value_coef * dv * L.inv_n;
            t0 = -fminf(s1, s2); t1 = dv * dv; t2 = H; t3 = (ratio - 1.f) - lr;
            t4 = fabsf(ratio - 1.f) > L.clip ? 1.f : 0.f;
        }
#pragma unroll
        for (int d = 8; d < 32; d <<= 1) {
            t0 += __shfl_xor_sync(0xffffffffu, t0, d); t1 += __shfl_xor_sync(0xffffffffu, t1, d);
            t2 += __shfl_xor_sync(0xffffffffu, t2, d); t3 += __shfl_xor_sync(0xffffffffu, t3, d);
            t4 += __shfl_xor_sync(0xffffffffu, t4, d);
        }
        if (lane == 0) {
            atomicAdd(&L.sums[0], (double)t0); atomicAdd(&L.sums[1], (double)t1); atomicAdd(&L.sums[2], (double)t2);
            atomicAdd(&L.sums[3], (double)t3); atomicAdd(&L.sums[4], (double)t4);
        }
    } else {
        g1 = (pa >= eps && pa <= 1.f - eps) ? g_logp[row] : 0.f;
        g2 = g_ent ? g_ent[row] : 0.f;
    }
    float out[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const bool ok = (mb >> i) & 1u;
        const float p = ok ? exp2f(z[i] * K3_LOG2E) * inv : 0.f;
        const float logp = z[i] - logS;             // log p_i for valid i
        float g = -g1 * p - g2 * p * (logp + H);
        if (own && i == slot) g += g1;
        out[i] = ok ? g : 0.f;
    }
    if (!live) return;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (BF16) {
            const __nv_bfloat162 a = __floats2bfloat162_rn(out[4 * k], out[4 * k + 1]);
            const __nv_bfloat162 b = __floats2bfloat162_rn(out[4 * k + 2], out[4 * k + 3]);
            uint2 h;
            h.x = *reinterpret_cast<const uint32_t*>(&a);
            h.y = *reinterpret_cast<const uint32_t*>(&b);
            reinterpret_cast<uint2*>(dlogits)[row * 48 + k * 8 + l] = h;
        } else {
            reinterpret_cast<float4*>(dlogits)[row * 48 + k * 8 + l] =
                make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        }
    }
}

cudaError_t bb_launch_masked_head_bwd(const void* logits, int dtype, const uint64_t* mask, int64_t mask_stride,
                                      const int32_t* action, const float* g_logp, const float* g_ent,
                                      void* dlogits, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n * 8 + 127) / 128);
    const BBPpoLossArgs none = {};
    if (dtype == 1)
        bb_masked_head_bwd_kernel<true, false><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, action, g_logp, g_ent, dlogits, n, none);
    else
        bb_masked_head_bwd_kernel<false, false><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, action, g_logp, g_ent, dlogits, n, none);
    return cudaGetLastError();
}

cudaError_t bb_launch_ppo_loss(const void* logits, int dtype, const uint64_t* mask, int64_t mask_stride,
                               const int32_t* action, const float* old_logp, const float* adv, const float* ret,
                               const float* value, float clip, float value_coef, float entropy_coef,
                               void* dlogits, float* dvalue, double* sums, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n * 8 + 127) / 128);
    const BBPpoLossArgs L = {old_logp, adv, ret, value, dvalue, sums, clip, value_coef, entropy_coef, 1.0f / (float)n};
    if (dtype == 1)
        bb_masked_head_bwd_kernel<true, true><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, action, nullptr, nullptr, dlogits, n, L);
    else
        bb_masked_head_bwd_kernel<false, true><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, action, nullptr, nullptr, dlogits, n, L);
    return cudaGetLastError();
}
