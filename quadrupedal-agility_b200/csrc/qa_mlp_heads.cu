// K20 / K21: the narrow output layers of the actor, the critic and the estimator on the CUDA cores, full fp32.
//
//   actor_head    Linear(128 -> 12)   bbc/rsl_rl/modules/actor_critic.py:118-119
//   critic_head   Linear(128 -> 1)    bbc/rsl_rl/modules/actor_critic.py:128-129
//   estimator.4   Linear(64 -> 4)     bbc/rsl_rl/modules/estimator.py:24-33
//
// These layers have 1..16 output columns: a tensor-core tile would be >= 87 % padding and the layer is a pure stream over
// the (M, 128) hidden activation -- HBM bound at 4 B/flop.  They therefore do not go through K7; they run as row-streaming
// kernels that also absorb their neighbours:
//   K20 qa_head_fwd   y = h W^T + b
//   K21 qa_head_bwd   gz_prev = (gz W) * act'(h)   (the gradient w.r.t. the PREVIOUS layer's pre-activation, ELU'/ReLU'
//                     recovered from its output h), dW += gz^T h, db += colsum(gz), db_prev += colsum(gz_prev)
//                     -- what autograd does in 5 kernels (mm, mm, sum, elu_backward, sum) in ONE pass over h.
// (Round 2 tried a lane = ROW formulation -- 32-row tiles transposed through shared memory, weights broadcast from shared
// memory, no shuffles: tools/bench_heads.py measured 12.3 / 22.7 us fwd / bwd for the 12-wide actor head at M = 24576 against
// 20.4 / 25.3 us here, but 23 us against 8.8 us for the 7 x 256 discriminator heads at M = 4096 and 17-19 us for every backward
// regardless of M: one tile per warp is a serial shared-memory-latency chain that 12 warps per SM cannot hide.  Not kept.)
// Lane = 4 consecutive hidden columns (float4); a row of 128 columns is one warp-wide 512 B access, a row of 64 columns half
// a warp (two rows per pass).  W sits in shared memory as [n][Kh].
#include "qa_b200.h"
#include "qa_common.cuh"

#define HD_THREADS 256
#define HD_MAXN 16
#define HD_MAXK 256

// Forward.  LPR lanes share a row (LPR = min(32, Kh / 4)); a lane owns KV = Kh / (4 * LPR) float4 column groups (2 for
// Kh = 256).  Each warp pass handles 4 * (32 / LPR) rows: the 4 row loads of a lane are issued back to back (memory-level
// parallelism: the kernel is a pure stream over h), then each row's NT partial dot products are summed across its lanes by
// xor butterflies -- level by level over all NT values, so the NT shuffles of a level are independent and pipeline.
template <int NT, int KV, int LPR>
__global__ void __launch_bounds__(HD_THREADS) k_head_fwd(const __grid_constant__ QaHeadFwdArgs a) {
    __shared__ __align__(16) float s_w[NT * HD_MAXK];
    __shared__ float s_b[NT];
    constexpr int Kh = LPR * KV * 4;
    const int N = a.N;
    for (int i = threadIdx.x; i < NT * Kh; i += HD_THREADS) {
        const int n = i / Kh, k = i - n * Kh;
        s_w[i] = n < N ? __ldg(a.w + (size_t)n * a.w_pitch + k) : 0.f;
    }
    if (threadIdx.x < NT) s_b[threadIdx.x] = (threadIdx.x < N && a.bias != nullptr) ? __ldg(a.bias + threadIdx.x) : 0.f;
    __syncthreads();
    constexpr int RPP = 32 / LPR;                  // rows per lane-group pass
    constexpr int U = 4;                           // rows in flight per lane
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, rsel = lane / LPR;
    const long long warps_total = (long long)gridDim.x * (HD_THREADS / 32);
    const long long stride = warps_total * RPP * U;
    for (long long r0 = ((long long)blockIdx.x * (HD_THREADS / 32) + warp) * RPP * U; r0 < a.M; r0 += stride) {
        float4 h4[U][KV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * RPP + rsel;
#pragma unroll
            for (int v = 0; v < KV; ++v)
                h4[u][v] = r < a.M ? *reinterpret_cast<const float4*>(a.h + r * a.h_pitch + (v * LPR + sub) * 4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * RPP + rsel;
            float p[NT];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                float acc = 0.f;
#pragma unroll
                for (int v = 0; v < KV; ++v) {
                    const float4 w4 = *reinterpret_cast<const float4*>(s_w + n * Kh + (v * LPR + sub) * 4);
                    acc += h4[u][v].x * w4.x + h4[u][v].y * w4.y + h4[u][v].z * w4.z + h4[u][v].w * w4.w;
                }
                p[n] = acc;
            }
#pragma unroll
            for (int o = LPR >> 1; o > 0; o >>= 1) {
#pragma unroll
                for (int n = 0; n < NT; ++n) p[n] += __shfl_xor_sync(QA_FULL, p[n], o);
            }
            if (r < a.M) {
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    if (n < N && (n % LPR) == sub) a.y[r * a.y_pitch + n] = p[n] + s_b[n];
            }
        }
    }
}

template <int NT>
static void launch_head_fwd(const QaHeadFwdArgs* a, cudaStream_t s) {
    const int lpr = a->Kh >= 128 ? 32 : a->Kh / 4;
    const int rpp = 32 / lpr;
    long long blocks = (a->M + 8LL * rpp * 4 - 1) / (8LL * rpp * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
    const unsigned g = (unsigned)blocks;
    switch (a->Kh) {
        case 256: k_head_fwd<NT, 2, 32><<<g, HD_THREADS, 0, s>>>(*a); break;
        case 128: k_head_fwd<NT, 1, 32><<<g, HD_THREADS, 0, s>>>(*a); break;
        case 64: k_head_fwd<NT, 1, 16><<<g, HD_THREADS, 0, s>>>(*a); break;
        default: k_head_fwd<NT, 1, 8><<<g, HD_THREADS, 0, s>>>(*a); break;
    }
}

extern "C" int qa_head_fwd(const QaHeadFwdArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->M == 0) return 0;
    QA_CHECK_PTR(a->h);
    QA_CHECK_PTR(a->w);
    QA_CHECK_PTR(a->y);
    if (a->M < 0 || a->N <= 0 || a->N > HD_MAXN) return QA_EINVAL;
    if (a->Kh != 32 && a->Kh != 64 && a->Kh != 128 && a->Kh != 256) return QA_EINVAL;
    if ((reinterpret_cast<uintptr_t>(a->h) & 15u) || (a->h_pitch & 3) || a->h_pitch < a->Kh || a->w_pitch < a->Kh || a->y_pitch < a->N)
        return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->N == 1) launch_head_fwd<1>(a, s);
    else if (a->N <= 4) launch_head_fwd<4>(a, s);
    else if (a->N <= 8) launch_head_fwd<8>(a, s);
    else if (a->N <= 12) launch_head_fwd<12>(a, s);
    else launch_head_fwd<16>(a, s);
    QA_LAUNCH_RET();
}

// Backward.  Every lane keeps dW[n][4 columns] partial sums over the rows it visits; the block folds them in shared memory
// (shared atomics) and leaves with one global atomicAdd per element.
template <int NT>
__global__ void __launch_bounds__(HD_THREADS, 2) k_head_bwd(const __grid_constant__ QaHeadBwdArgs a) {
    __shared__ __align__(16) float s_w[NT * 128];
    __shared__ __align__(16) float s_dw[NT * 128];
    __shared__ float s_dbp[128];
    __shared__ float s_db[NT];
    const int Kh = a.Kh, N = a.N;
    for (int i = threadIdx.x; i < NT * Kh; i += HD_THREADS) {
        const int n = i / Kh, k = i - n * Kh;
        s_w[i] = n < N ? __ldg(a.w + (size_t)n * a.w_pitch + k) : 0.f;
        s_dw[i] = 0.f;
    }
    if (threadIdx.x < Kh) s_dbp[threadIdx.x] = 0.f;
    if (threadIdx.x < NT) s_db[threadIdx.x] = 0.f;
    __syncthreads();
    const int lpr = Kh >> 2, rpp = 32 / lpr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % lpr, rsel = lane / lpr;
    const long long warps_total = (long long)gridDim.x * (HD_THREADS / 32);
    float4 dw[NT];
    float dbn[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) dw[n] = make_float4(0.f, 0.f, 0.f, 0.f), dbn[n] = 0.f;
    float4 dbp = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = NT <= 4 ? 2 : 1;             // rows in flight per lane (register budget: 2 blocks of 256 threads per SM)
    for (long long r0 = ((long long)blockIdx.x * (HD_THREADS / 32) + warp) * rpp * U; r0 < a.M; r0 += warps_total * rpp * U) {
        float4 h4[U];
        float g[U][NT];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rpp + rsel;
            ok[u] = r < a.M;
            h4[u] = ok[u] ? *reinterpret_cast<const float4*>(a.h + r * a.h_pitch + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < NT; ++n)        // same address across the row's lanes: broadcast
                g[u][n] = (ok[u] && n < N) ? __ldg(a.gz + r * a.gz_pitch + n) * a.gz_scale : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = r0 + u * rpp + rsel;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const float4 w4 = *reinterpret_cast<const float4*>(s_w + n * Kh + sub * 4);
                o.x += g[u][n] * w4.x, o.y += g[u][n] * w4.y, o.z += g[u][n] * w4.z, o.w += g[u][n] * w4.w;
                dw[n].x += g[u][n] * h4[u].x, dw[n].y += g[u][n] * h4[u].y, dw[n].z += g[u][n] * h4[u].z, dw[n].w += g[u][n] * h4[u].w;
                dbn[n] += g[u][n];
            }
            const float4 h = h4[u];
            if (a.act == 1) {                                                      // ELU'(z) from the output: 1 if h > 0 else h + 1
                o.x = h.x > 0.f ? o.x : o.x * (h.x + 1.0f), o.y = h.y > 0.f ? o.y : o.y * (h.y + 1.0f);
                o.z = h.z > 0.f ? o.z : o.z * (h.z + 1.0f), o.w = h.w > 0.f ? o.w : o.w * (h.w + 1.0f);
            } else if (a.act == 2) {
                o.x = h.x > 0.f ? o.x : 0.f, o.y = h.y > 0.f ? o.y : 0.f, o.z = h.z > 0.f ? o.z : 0.f, o.w = h.w > 0.f ? o.w : 0.f;
            }
            if (ok[u] && a.gz_prev != nullptr) *reinterpret_cast<float4*>(a.gz_prev + r * a.gz_prev_pitch + sub * 4) = o;
            dbp.x += o.x, dbp.y += o.y, dbp.z += o.z, dbp.w += o.w;
        }
    }
    // fold: lanes that share the columns (the rpp row groups of a warp, the 8 warps) meet in shared memory
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (n < N) {
            float* d = s_dw + n * Kh + sub * 4;
            atomicAdd(d + 0, dw[n].x), atomicAdd(d + 1, dw[n].y), atomicAdd(d + 2, dw[n].z), atomicAdd(d + 3, dw[n].w);
            if (sub == 0) atomicAdd(s_db + n, dbn[n]);
        }
    }
    atomicAdd(s_dbp + sub * 4 + 0, dbp.x), atomicAdd(s_dbp + sub * 4 + 1, dbp.y);
    atomicAdd(s_dbp + sub * 4 + 2, dbp.z), atomicAdd(s_dbp + sub * 4 + 3, dbp.w);
    __syncthreads();
    if (a.dw != nullptr)
        for (int i = threadIdx.x; i < N * Kh; i += HD_THREADS) {
            const int n = i / Kh, k = i - n * Kh;
            atomicAdd(a.dw + (size_t)n * a.dw_pitch + k, s_dw[i]);
        }
    if (a.db != nullptr && threadIdx.x < N) atomicAdd(a.db + threadIdx.x, s_db[threadIdx.x]);
    if (a.db_prev != nullptr && threadIdx.x < Kh) atomicAdd(a.db_prev + threadIdx.x, s_dbp[threadIdx.x]);
}

extern "C" int qa_head_bwd(const QaHeadBwdArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->M == 0) return 0;
    QA_CHECK_PTR(a->gz);
    QA_CHECK_PTR(a->h);
    QA_CHECK_PTR(a->w);
    if (a->M < 0 || a->N <= 0 || a->N > HD_MAXN || a->act < 0 || a->act > 2) return QA_EINVAL;
    if (a->Kh != 32 && a->Kh != 64 && a->Kh != 128) return QA_EINVAL;
    if ((reinterpret_cast<uintptr_t>(a->h) & 15u) || (a->h_pitch & 3) || a->h_pitch < a->Kh || a->w_pitch < a->Kh || a->gz_pitch < a->N)
        return QA_EINVAL;
    if (a->gz_prev != nullptr && ((reinterpret_cast<uintptr_t>(a->gz_prev) & 15u) || (a->gz_prev_pitch & 3) || a->gz_prev_pitch < a->Kh))
        return QA_EINVAL;
    if (a->dw != nullptr && a->dw_pitch < a->Kh) return QA_EINVAL;
    const int rpp = 32 / (a->Kh >> 2);
    long long blocks = (a->M + 8LL * rpp * 4 - 1) / (8LL * rpp * 4);        // >= 4 passes per warp: the fold is amortised
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (a->N == 1) k_head_bwd<1><<<(unsigned)blocks, HD_THREADS, 0, s>>>(*a);
    else if (a->N <= 4) k_head_bwd<4><<<(unsigned)blocks, HD_THREADS, 0, s>>>(*a);
    else if (a->N <= 12) k_head_bwd<12><<<(unsigned)blocks, HD_THREADS, 0, s>>>(*a);
    else k_head_bwd<16><<<(unsigned)blocks, HD_THREADS, 0, s>>>(*a);
    QA_LAUNCH_RET();
}


// ------------------------------------------------------------------------------------------------------------------------
// K22 qa_policy_sample -- `Normal(mean, std).sample()`, `log_prob(actions).sum(-1)` and the transition's storage writes of
//     SSInfoGAIL.act (bbc/rsl_rl/algorithms/gail.py:186-196, modules/actor_critic.py:189-197) in one launch:
//       actions = mu + std * n,  logp = sum_j -(a_j - mu_j)^2 / (2 std_j^2) - log std_j - log sqrt(2 pi)
//     n is either supplied (parity mode: the reference consumes torch's generator) or drawn in-kernel: Philox4x32-10 keyed by
//     (seed; env, site, step) + Box-Muller, the same counter layout as the env kernels' production stream.
// ------------------------------------------------------------------------------------------------------------------------
#define SITE_ACT0 48
__global__ void __launch_bounds__(256) k_policy_sample(const __grid_constant__ QaPolicySampleArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    const int A = a.A;
    unsigned long long step = a.rng_step;
    if (a.step_state != nullptr) step = (unsigned long long)(*reinterpret_cast<const volatile long long*>(a.step_state)) + 1ull;
    float logp = 0.f;
    for (int j0 = 0; j0 < A; j0 += 4) {
        float n4[4];
        if (a.noise != nullptr) {
#pragma unroll
            for (int e = 0; e < 4; ++e) n4[e] = j0 + e < A ? a.noise[i * A + j0 + e] : 0.f;
        } else {
            const Philox4 r = philox4x32_10((uint32_t)i, SITE_ACT0 + (j0 >> 2), (uint32_t)step, (uint32_t)(step >> 32),
                                            (uint32_t)a.rng_seed, (uint32_t)(a.rng_seed >> 32));
            // Box-Muller on (0, 1] uniforms: two pairs -> four standard normals
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float u1 = ((float)(r.v[2 * e] >> 8) + 1.0f) * (1.0f / 16777216.0f);
                const float u2 = (float)(r.v[2 * e + 1] >> 8) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(u1));
                float sn, cs;
                sincospif(2.0f * u2, &sn, &cs);
                n4[2 * e] = rad * cs;
                n4[2 * e + 1] = rad * sn;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            if (j < A) {
                const float mu = a.mu[i * a.mu_pitch + j], sd = a.std[j];
                const float act = mu + sd * n4[e];
                const float d = act - mu;
                logp += -(d * d) / (2.f * (sd * sd)) - logf(sd) - 0.9189385332046727f;
                a.actions[i * A + j] = act;
                if (a.actions_st != nullptr) a.actions_st[i * A + j] = act;
                if (a.mu_st != nullptr) a.mu_st[i * A + j] = mu;
                if (a.sigma_st != nullptr) a.sigma_st[i * A + j] = sd;
            }
        }
    }
    if (a.logp != nullptr) a.logp[i] = logp;
    if (a.logp_st != nullptr) a.logp_st[i] = logp;
}

extern "C" int qa_policy_sample(const QaPolicySampleArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->M == 0) return 0;
    QA_CHECK_PTR(a->mu);
    QA_CHECK_PTR(a->std);
    QA_CHECK_PTR(a->actions);
    if (a->M < 0 || a->A <= 0 || a->A > 64 || a->mu_pitch < a->A) return QA_EINVAL;
    k_policy_sample<<<(unsigned)((a->M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}
