// bb_policy_kernels.cu — K2 (packed obs -> dense NCHW planes / dense mask) and K3 (fused
// -inf masking + softmax + categorical sample + log-prob + masked entropy) for sm_100a.
//
// K2 replaces engine.get_observation + Piece.to_mask (src/game/engine.py:489-507,
// src/game/pieces.py:39-45) and the torch.cat at src/models/network.py:152-158:
// 36 B in, 1,024 B (f32) / 512 B (bf16) out per env; a warp writes one env's 4x8x8 block as
// one contiguous run of 128-bit stores.
//
// K3 replaces src/models/network.py:172-262 (mask -> softmax -> Categorical sample/log_prob ->
// masked entropy): one warp per row, 768 B of logits read once with 8-byte loads (256 B
// contiguous per warp instruction), everything else in registers / shuffles.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include "bb_kernels.h"

__constant__ uint64_t c_piece_cells[BB_NUM_PIECES + 3] = BB_PIECE_MASKS;

// ------------------------------------------------------------------------------------ K2
template <bool BF16>
__global__ void __launch_bounds__(256)
bb_unpack_obs_kernel(const uint64_t* __restrict__ board, const uint32_t* __restrict__ pieces,
                     void* __restrict__ obs, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (env, channel, row)
    const int64_t env = t >> 5;
    if (env >= n) return;
    const int ch = (int)(t >> 3) & 3, row = (int)t & 7;
    uint64_t plane;
    if (ch == 0) {
        plane = __ldg(board + env);
    } else {
        const uint32_t pw = __ldg(pieces + env);
        const uint32_t id = (pw >> (8 * (ch - 1))) & 0xFFu;
        const bool used = (pw >> (24 + ch - 1)) & 1u;
        plane = used ? 0ull : c_piece_cells[id < BB_NUM_PIECES ? id : BB_NUM_PIECES];
    }
    const uint32_t bits = (uint32_t)(plane >> (8 * row)) & 0xFFu;
    if (BF16) {
        // bf16 1.0 = 0x3F80
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (((bits >> (2 * k)) & 1u) ? 0x3F80u : 0u) | (((bits >> (2 * k + 1)) & 1u) ? 0x3F800000u : 0u);
        reinterpret_cast<uint4*>(obs)[t] = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        float4 lo, hi;
        lo.x = (bits & 1u) ? 1.f : 0.f;  lo.y = (bits & 2u) ? 1.f : 0.f;
        lo.z = (bits & 4u) ? 1.f : 0.f;  lo.w = (bits & 8u) ? 1.f : 0.f;
        hi.x = (bits & 16u) ? 1.f : 0.f; hi.y = (bits & 32u) ? 1.f : 0.f;
        hi.z = (bits & 64u) ? 1.f : 0.f; hi.w = (bits & 128u) ? 1.f : 0.f;
        float4* o = reinterpret_cast<float4*>(obs) + 2 * t;
        o[0] = lo;
        o[1] = hi;
    }
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

cudaError_t bb_launch_unpack_obs(const uint64_t* board, const uint32_t* pieces, const uint64_t* mask,
                                 int64_t mask_stride, void* obs_nchw, int obs_dtype, void* mask_dense,
                                 int mask_dtype, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (obs_nchw) {
        const int64_t threads = n * 32;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        if (obs_dtype == 1) bb_unpack_obs_kernel<true><<<grid, 256, 0, stream>>>(board, pieces, obs_nchw, n);
        else bb_unpack_obs_kernel<false><<<grid, 256, 0, stream>>>(board, pieces, obs_nchw, n);
    }
    if (mask_dense) {
        const int64_t threads = n * 12;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        if (mask_dtype == 0) bb_unpack_mask_kernel<true><<<grid, 256, 0, stream>>>(mask, mask_stride, mask_dense, n);
        else bb_unpack_mask_kernel<false><<<grid, 256, 0, stream>>>(mask, mask_stride, mask_dense, n);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ K3
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// lane l of the warp owns actions 64k + 2l and 64k + 2l + 1 for planes k = 0,1,2
template <bool BF16>
__global__ void __launch_bounds__(128)
bb_masked_sample_kernel(const void* __restrict__ logits, const uint64_t* __restrict__ mask, int64_t stride,
                        uint64_t seed, uint64_t call_counter, int mode, int32_t* __restrict__ action,
                        float* __restrict__ logp_out, float* __restrict__ ent_out, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;   // whole warp exits together
    float z[3][2];
    uint32_t mb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint64_t w = __ldg(mask + (int64_t)k * stride + row);
        mb[k] = (uint32_t)(w >> (2 * lane)) & 3u;
        float2 v;
        if (BF16) {
            const __nv_bfloat162 h = reinterpret_cast<const __nv_bfloat162*>(logits)[row * 96 + k * 32 + lane];
            v = __bfloat1622float2(h);
        } else {
            v = reinterpret_cast<const float2*>(logits)[row * 96 + k * 32 + lane];
        }
        z[k][0] = (mb[k] & 1u) ? v.x : -INFINITY;
        z[k][1] = (mb[k] & 2u) ? v.y : -INFINITY;
    }
    float m = fmaxf(fmaxf(fmaxf(z[0][0], z[0][1]), fmaxf(z[1][0], z[1][1])), fmaxf(z[2][0], z[2][1]));
    m = warp_max(m);
    const bool any_valid = m > -INFINITY;
    float p[3][2];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        p[k][0] = (mb[k] & 1u) ? expf(z[k][0] - m) : 0.f;
        p[k][1] = (mb[k] & 2u) ? expf(z[k][1] - m) : 0.f;
        s += p[k][0] + p[k][1];
    }
    s = warp_sum(s);
    const float inv = any_valid ? 1.f / s : 0.f;
    float ps = 0.f;   // sum of probabilities (Categorical re-normalises by it)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        p[k][0] = __fdiv_rn(p[k][0], s);
        p[k][1] = __fdiv_rn(p[k][1], s);
        if (!any_valid) { p[k][0] = 0.f; p[k][1] = 0.f; }
        ps += p[k][0] + p[k][1];
    }
    (void)inv;
    ps = warp_sum(ps);

    int act = 0;
    if (mode == 2) {
        act = action[row];
    } else if (mode == 1) {
        // argmax of probs, first index on ties (torch.argmax)
        float best = -1.f; int bi = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int idx = 64 * k + 2 * lane + e;
                if (p[k][e] > best || (p[k][e] == best && idx < bi)) { best = p[k][e]; bi = idx; }
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        act = any_valid ? bi : 0;
    } else {
        // inverse CDF in action-index order on u * sum(p)
        const BBPhilox4 r = bb_philox((uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)call_counter,
                                      BB_STREAM_SAMPLE, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);
        const float t = u * ps;
        float base = 0.f;
        int found = -1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float c = p[k][0] + p[k][1];
            const float incl = warp_incl_scan(c, lane);
            const float tot = __shfl_sync(0xffffffffu, incl, 31);
            const unsigned hit = __ballot_sync(0xffffffffu, (base + incl > t) && (c > 0.f));
            if (found < 0 && hit) {
                const int L = __ffs(hit) - 1;
                const float excl = base + incl - c;
                int e = ((excl + p[k][0] > t) && (p[k][0] > 0.f)) ? 0 : ((p[k][1] > 0.f) ? 1 : 0);
                e = __shfl_sync(0xffffffffu, e, L);
                found = 64 * k + 2 * L + e;
            }
            base += tot;
        }
        if (found < 0) {
            // t >= total by rounding: take the last action with non-zero probability
#pragma unroll
            for (int k = 2; k >= 0; --k) {
                const unsigned nz = __ballot_sync(0xffffffffu, (p[k][0] > 0.f) || (p[k][1] > 0.f));
                if (found < 0 && nz) {
                    const int L = 31 - __clz(nz);
                    int e = (p[k][1] > 0.f) ? 1 : 0;
                    e = __shfl_sync(0xffffffffu, e, L);
                    found = 64 * k + 2 * L + e;
                }
            }
        }
        act = found < 0 ? 0 : found;
    }

    // probability of the chosen action
    const int ak = (act >> 6), al = (act & 63) >> 1, ae = act & 1;
    float pa = 0.f;
    if (act >= 0 && act < 192) {
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k == ak) mine = ae ? p[k][1] : p[k][0];
        pa = __shfl_sync(0xffffffffu, mine, al);
    }
    // entropy over the valid actions (network.py:246-260)
    float ent = 0.f;
    if (ent_out) {
        const float den = fmaxf(ps, 1e-10f);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if ((mb[k] >> e) & 1u) {
                    const float q = __fdiv_rn(p[k][e], den);
                    ent -= q * logf(fmaxf(q, 1e-10f));
                }
        ent = warp_sum(ent);
    }
    if (lane == 0) {
        if (mode != 2) action[row] = act;
        if (logp_out) {
            const float eps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
            const float pn = any_valid ? __fdiv_rn(pa, ps) : 1.f;
            logp_out[row] = logf(fminf(fmaxf(pn, eps), 1.f - eps));
        }
        if (ent_out) ent_out[row] = ent;
    }
}

cudaError_t bb_launch_masked_sample(const void* logits, int logits_dtype, const uint64_t* mask,
                                    int64_t mask_stride, uint64_t seed, uint64_t call_counter, int mode,
                                    int32_t* action, float* logp, float* entropy, int64_t n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n * 32 + 127) / 128);
    if (logits_dtype == 1)
        bb_masked_sample_kernel<true><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, seed, call_counter, mode, action, logp, entropy, n);
    else
        bb_masked_sample_kernel<false><<<grid, 128, 0, stream>>>(logits, mask, mask_stride, seed, call_counter, mode, action, logp, entropy, n);
    return cudaGetLastError();
}
