// K5: GAE returns + normalised advantages (sm_100a).
//
// Replaces RolloutStorage.compute_returns, bbc/rsl_rl/storage/rollout_storage.py:97-111
// (identical in tsc/rsl_rl/storage/rollout_storage.py:102-116):
//     delta_t = r_t + (1-d_t) gamma V_{t+1} - V_t
//     A_t     = delta_t + (1-d_t) gamma lam A_{t+1}          (reverse recurrence over the horizon)
//     R_t     = A_t + V_t
//     adv     = (A - mean(A)) / (std_unbiased(A) + 1e-8)     over all T*N samples
//
// The recurrence A_t = b_t + a_t A_{t+1} is a composition of affine maps, i.e. a scan.  A CTA owns a
// tile of 32 envs x T steps: rows are loaded coalesced (lane = env), transposed through shared memory,
// then every warp runs a reverse Kogge-Stone scan over the horizon with lane = time step (T <= 32, five
// shuffle rounds on the (a, b) pair), and the tile is written back coalesced.  The grid-wide mean / std
// is accumulated in fp64 (per-CTA shuffle reduction + one atomic pair per CTA) and applied by a second
// tiny kernel.  Algorithmic traffic: 9 B read + 8 B written per sample in pass 1, 4 B + 4 B in pass 2.
#include "qa_b200.h"
#include "qa_common.cuh"

#define GAE_TILE 32
#define GAE_THREADS 256
#define GAE_MAXT 32

struct GaeWorkspace {
    double sum;
    double sumsq;
};

__global__ void __launch_bounds__(GAE_THREADS) k_gae_scan(QaGaeArgs g) {
    __shared__ float s_a[GAE_TILE][GAE_MAXT + 1];     // [env][t] coefficient a_t, later A_t
    __shared__ float s_b[GAE_TILE][GAE_MAXT + 1];     // [env][t] delta_t
    __shared__ float s_v[GAE_TILE][GAE_MAXT + 1];     // [env][t] V_t
    __shared__ double s_red[2][GAE_THREADS / 32];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = GAE_THREADS / 32;
    const int n0 = blockIdx.x * GAE_TILE;
    const int T = g.num_steps, N = g.num_envs;
    const int n = n0 + lane;
    const bool ok = n < N;

    // phase 1: coalesced row loads (lane = env), one time step per warp iteration
    for (int t = wid; t < T; t += nw) {
        float r = 0.f, v = 0.f, vn = 0.f, nt = 0.f;
        if (ok) {
            const size_t i = (size_t)t * N + n;
            r = g.rewards[i];
            v = g.values[i];
            vn = (t == T - 1) ? g.last_values[n] : g.values[i + N];
            nt = 1.0f - (float)g.dones[i];
        }
        s_b[lane][t] = r + nt * g.gamma * vn - v;
        s_a[lane][t] = nt * g.gamma * g.lam;        // ((1-d) gamma) lam, the reference's rounding order
        s_v[lane][t] = v;
    }
    __syncthreads();

    // phase 2: reverse inclusive scan over the horizon, lane = t; maps compose as
    //   (a, b) o (a', b') = (a a', b + a b')   [apply the later step first]
    double psum = 0.0, psq = 0.0;
    for (int el = wid; el < GAE_TILE; el += nw) {
        float a = (lane < T) ? s_a[el][lane] : 0.f;
        float b = (lane < T) ? s_b[el][lane] : 0.f;
#pragma unroll
        for (int o = 1; o < GAE_MAXT; o <<= 1) {
            const float a2 = __shfl_down_sync(QA_FULL, a, o);
            const float b2 = __shfl_down_sync(QA_FULL, b, o);
            if (lane + o < GAE_MAXT) {
                b = b + a * b2;
                a = a * a2;
            }
        }
        // b is now A_t (A_T = 0)
        if (lane < T) {
            s_a[el][lane] = b;
            if (n0 + el < N) {
                psum += (double)b;
                psq += (double)b * (double)b;
            }
        }
    }
    __syncthreads();

    // phase 3: coalesced write-back of returns and raw advantages
    for (int t = wid; t < T; t += nw) {
        if (ok) {
            const size_t i = (size_t)t * N + n;
            const float adv = s_a[lane][t];
            g.returns[i] = adv + s_v[lane][t];
            g.advantages[i] = adv;
        }
    }

    // grid-wide moments
    psum = warp_sum_d(psum);
    psq = warp_sum_d(psq);
    if (lane == 0) {
        s_red[0][wid] = psum;
        s_red[1][wid] = psq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < nw; ++k) {
            s += s_red[0][k];
            q += s_red[1][k];
        }
        GaeWorkspace* ws = reinterpret_cast<GaeWorkspace*>(g.workspace);
        atomicAdd(&ws->sum, s);
        atomicAdd(&ws->sumsq, q);
    }
}

__global__ void __launch_bounds__(256) k_gae_normalise(QaGaeArgs g) {
    const GaeWorkspace* ws = reinterpret_cast<const GaeWorkspace*>(g.workspace);
    const double cnt = (double)g.num_steps * (double)g.num_envs;
    const double mean = ws->sum / cnt;
    double var = (ws->sumsq - cnt * mean * mean) / (cnt - 1.0);     // torch.std(): unbiased
    var = var > 0.0 ? var : 0.0;
    const float meanf = (float)mean;
    const float denom = (float)sqrt(var) + 1e-8f;
    const size_t total = (size_t)g.num_steps * g.num_envs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        g.advantages[i] = (g.advantages[i] - meanf) / denom;
}

extern "C" int qa_gae(const QaGaeArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->num_envs == 0 && g->num_steps > 0) return 0;
    QA_CHECK_PTR(g->rewards);
    QA_CHECK_PTR(g->values);
    QA_CHECK_PTR(g->dones);
    QA_CHECK_PTR(g->last_values);
    QA_CHECK_PTR(g->returns);
    QA_CHECK_PTR(g->advantages);
    QA_CHECK_PTR(g->workspace);
    if (g->num_envs < 0 || g->num_steps <= 0) return QA_EINVAL;
    if (g->num_steps > GAE_MAXT) return QA_ERANGE;
    if (g->num_envs == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t err = cudaMemsetAsync(g->workspace, 0, sizeof(GaeWorkspace), s);
    if (err != cudaSuccess) return (int)err;
    const int grid = (g->num_envs + GAE_TILE - 1) / GAE_TILE;
    k_gae_scan<<<grid, GAE_THREADS, 0, s>>>(*g);
    err = cudaGetLastError();
    if (err != cudaSuccess) return (int)err;
    const size_t total = (size_t)g->num_steps * g->num_envs;
    int grid2 = (int)((total + 255) / 256);
    if (grid2 > 148 * 8) grid2 = 148 * 8;
    k_gae_normalise<<<grid2, 256, 0, s>>>(*g);
    QA_LAUNCH_RET();
}
