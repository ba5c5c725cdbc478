// bb_gae_kernels.cu — K4: reverse-scan GAE / returns for sm_100a.
//
// Replaces RolloutBuffer.compute_returns_and_advantages (src/agents/ppo.py:141-169): for
// t = T-1..0:  nnt = 1 - done[t];  delta = r[t] + gamma*V[t+1]*nnt - V[t];
//              A[t] = delta + gamma*lambda*nnt*A[t+1];   R[t] = A[t] + V[t]
// in float32 with the reference's operation order and no FMA contraction, so the result is
// bit-identical to numpy's.  Sequential in t, parallel over envs: one thread owns 4 adjacent
// envs and moves them with 128-bit loads/stores (a warp touches 512 contiguous bytes per
// array per timestep); U timesteps are loaded ahead of the dependent chain to keep
// 3*U 16-byte loads in flight per thread.  20 B of HBM traffic per sample.
// Also accumulates sum(A) and sum(A^2) in float64 for the whole-buffer advantage
// normalisation (ppo.py:196).
#include <cuda_runtime.h>
#include "bb_kernels.h"

struct f4 { float v[4]; };

template <int V> struct Vec;
template <> struct Vec<4> {
    static __device__ __forceinline__ f4 ld(const float* p) { const float4 x = *reinterpret_cast<const float4*>(p); f4 r = {{x.x, x.y, x.z, x.w}}; return r; }
    static __device__ __forceinline__ void st(float* p, const f4& a) { *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
};
template <> struct Vec<1> {
    static __device__ __forceinline__ f4 ld(const float* p) { f4 r = {{*p, 0.f, 0.f, 0.f}}; return r; }
    static __device__ __forceinline__ void st(float* p, const f4& a) { *p = a.v[0]; }
};

template <int V, int U>
__global__ void __launch_bounds__(128)
bb_gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
              const float* __restrict__ dones, const float* __restrict__ last_values,
              float g, float gl, float* __restrict__ adv, float* __restrict__ ret,
              double* __restrict__ moments, int64_t T, int64_t N) {
    const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    double s1 = 0.0, s2 = 0.0;
    if (col < N) {
        f4 last = {{0.f, 0.f, 0.f, 0.f}};
        f4 nv = Vec<V>::ld(last_values + col);
        int64_t t = T - 1;
        while (t >= 0) {
            f4 rr[U], vv[U], dd[U];
            const int cnt = t + 1 < U ? (int)(t + 1) : U;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u < cnt) {
                    const int64_t off = (t - u) * N + col;
                    rr[u] = Vec<V>::ld(rewards + off);
                    vv[u] = Vec<V>::ld(values + off);
                    dd[u] = Vec<V>::ld(dones + off);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u < cnt) {
                    f4 a, r;
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float nnt = __fsub_rn(1.0f, dd[u].v[k]);
                        const float x = __fmul_rn(__fmul_rn(g, nv.v[k]), nnt);
                        const float delta = __fsub_rn(__fadd_rn(rr[u].v[k], x), vv[u].v[k]);
                        const float y = __fmul_rn(__fmul_rn(gl, nnt), last.v[k]);
                        last.v[k] = __fadd_rn(delta, y);
                        a.v[k] = last.v[k];
                        r.v[k] = __fadd_rn(last.v[k], vv[u].v[k]);
                        s1 += (double)last.v[k];
                        s2 += (double)last.v[k] * (double)last.v[k];
                    }
                    const int64_t off = (t - u) * N + col;
                    Vec<V>::st(adv + off, a);
                    Vec<V>::st(ret + off, r);
                    nv = vv[u];
                }
            t -= cnt;
        }
    }
    if (moments) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s1 += __shfl_down_sync(0xffffffffu, s1, d);
            s2 += __shfl_down_sync(0xffffffffu, s2, d);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&moments[0], s1);
            atomicAdd(&moments[1], s2);
        }
    }
}

cudaError_t bb_launch_gae(const float* rewards, const float* values, const float* dones,
                          const float* last_values, float gamma, float gamma_lam, float* adv, float* ret,
                          double* moments, int64_t T, int64_t N, cudaStream_t stream) {
    if (T <= 0 || N <= 0) return cudaSuccess;
    const bool aligned = ((N & 3) == 0) &&
        ((((uintptr_t)rewards | (uintptr_t)values | (uintptr_t)dones | (uintptr_t)last_values |
           (uintptr_t)adv | (uintptr_t)ret) & 15) == 0);
    if (aligned) {
        const int64_t threads = N / 4;
        bb_gae_kernel<4, 4><<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(
            rewards, values, dones, last_values, gamma, gamma_lam, adv, ret, moments, T, N);
    } else {
        bb_gae_kernel<1, 8><<<(unsigned)((N + 127) / 128), 128, 0, stream>>>(
            rewards, values, dones, last_values, gamma, gamma_lam, adv, ret, moments, T, N);
    }
    return cudaGetLastError();
}
