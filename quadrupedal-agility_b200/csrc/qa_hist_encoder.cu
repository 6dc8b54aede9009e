// K11: StateHistoryEncoder forward, fused (tsteps = 10) -- bbc/rsl_rl/modules/actor_critic.py:9-59:
//   per time step  Linear(57 -> 30) + ELU
//   Conv1d(30 -> 20, k = 4, s = 2) + ELU      (length 10 -> 4)
//   Conv1d(20 -> 10, k = 2, s = 1) + ELU      (length 4 -> 3)
//   Flatten (channel-major) -> Linear(30 -> 29) + ELU
// The reference runs this as 1 GEMM on (10 M, 57), two cuDNN convolutions with layout transposes and 1 GEMM, i.e. ~8
// launches that stream a (10 M, 30) intermediate through HBM.  Here one block owns 32 samples: all weights (5.4 K floats)
// live in shared memory, each time step's 32 x 57 input tile is staged with coalesced loads (the next tile is prefetched
// into registers while the current one is consumed), and a thread owns 2 samples x 4 output channels of every stage so
// that each 16-byte weight load from shared memory feeds 8 (stage 0, 3) to 32 (conv 1) FMAs -- the kernel is bound by the
// shared-memory pipe otherwise; intermediates never leave shared memory and their per-sample strides are odd so that the
// samples of a warp fall in different banks.  Forward only: in the PPO update
// the history latent is computed under inference_mode (gail.py:349-351), and the rollout only needs the forward.
#include "qa_b200.h"
#include "qa_common.cuh"

#define HE_S 32                 // samples per block
#define HE_THREADS 128          // thread = (sample pair sp = tid / 8, channel quad q = tid % 8)
#define HE_T 10
#define HE_IN 57
#define HE_C0 30
#define HE_C1 20
#define HE_L1 4
#define HE_C2 10
#define HE_L2 3
#define HE_OUT 29
// per-sample strides chosen odd (mod 32) so that the 4 sample pairs of a warp hit different shared-memory banks
#define PROJ_S (HE_T * 32 + 1)
#define C1_S (HE_L1 * HE_C1 + 1)
#define C2_S 33
#define HE_XR ((HE_S * HE_IN + HE_THREADS - 1) / HE_THREADS)      // 15 prefetch registers per thread

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

struct HeSmem {
    float w0[HE_IN][32];                    // [i][c]   (c padded to 32)
    float w1[HE_C0 * 4][HE_C1];             // [(c,k)][o]
    float w2[HE_C1 * 2][12];                // [(c,k)][o] (o padded to 12)
    float w3[HE_C0][32];                    // [i][j]   (j padded to 32)
    float b0[32], b1[HE_C1], b2[12], b3[32];
    float x[HE_S][HE_IN];                   // current time step's inputs
    float proj[HE_S * PROJ_S];              // [s][t*32 + c]
    float c1[HE_S * C1_S];                  // [s][p*20 + o]
    float c2[HE_S * C2_S];                  // [s][o*3 + p]  (nn.Flatten of (channels, length))
};

__global__ void __launch_bounds__(HE_THREADS) k_hist_encoder(QaHistEncArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeSmem& S = *reinterpret_cast<HeSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int sp = tid >> 3, q = tid & 7;
    const int s0 = 2 * sp, s1 = 2 * sp + 1;
    const long long m0 = (long long)blockIdx.x * HE_S;

    // ---- weights -> shared memory: coalesced global reads, transposed so a thread's output channels are contiguous ----
    for (int i = tid; i < HE_C0 * HE_IN; i += HE_THREADS) {
        const int c = i / HE_IN, ii = i - c * HE_IN;
        S.w0[ii][c] = a.w0[(size_t)c * a.w0_pitch + ii];
    }
    for (int i = tid; i < HE_IN * 2; i += HE_THREADS) S.w0[i >> 1][HE_C0 + (i & 1)] = 0.f;
    for (int i = tid; i < HE_C1 * HE_C0 * 4; i += HE_THREADS) {
        const int o = i / (HE_C0 * 4), ck = i - o * (HE_C0 * 4);
        S.w1[ck][o] = a.w1[i];
    }
    for (int i = tid; i < HE_C2 * HE_C1 * 2; i += HE_THREADS) {
        const int o = i / (HE_C1 * 2), ck = i - o * (HE_C1 * 2);
        S.w2[ck][o] = a.w2[i];
    }
    for (int i = tid; i < HE_C1 * 2 * 2; i += HE_THREADS) S.w2[i >> 1][HE_C2 + (i & 1)] = 0.f;
    for (int i = tid; i < HE_OUT * HE_C0; i += HE_THREADS) {
        const int j = i / HE_C0, ii = i - j * HE_C0;
        S.w3[ii][j] = a.w3[(size_t)j * a.w3_pitch + ii];
    }
    for (int i = tid; i < HE_C0 * 3; i += HE_THREADS) S.w3[i / 3][HE_OUT + i % 3] = 0.f;
    if (tid < 32) {
        S.b0[tid] = tid < HE_C0 ? a.b0[tid] : 0.f;
        S.b3[tid] = tid < HE_OUT ? a.b3[tid] : 0.f;
        if (tid < HE_C1) S.b1[tid] = a.b1[tid];
        if (tid < 12) S.b2[tid] = tid < HE_C2 ? a.b2[tid] : 0.f;
    }

    // ---- stage 0: per-step projection 57 -> 30 (+ELU); the next step's input tile is prefetched into registers ----------
    float xr[HE_XR];
    auto prefetch = [&](int t) {
#pragma unroll
        for (int k = 0; k < HE_XR; ++k) {
            const int i = tid + k * HE_THREADS;
            float v = 0.f;
            if (i < HE_S * HE_IN) {
                const int ss = i / HE_IN, ii = i - ss * HE_IN;
                const long long m = m0 + ss;
                if (m < a.M) v = __ldg(a.hist + (size_t)m * a.hist_pitch + t * HE_IN + ii);
            }
            xr[k] = v;
        }
    };
    prefetch(0);
    for (int t = 0; t < HE_T; ++t) {
        __syncthreads();                                             // weights ready / previous tile consumed
#pragma unroll
        for (int k = 0; k < HE_XR; ++k) {
            const int i = tid + k * HE_THREADS;
            if (i < HE_S * HE_IN) (&S.x[0][0])[i] = xr[k];
        }
        __syncthreads();
        if (t + 1 < HE_T) prefetch(t + 1);
        const float4 bb = *reinterpret_cast<const float4*>(&S.b0[q * 4]);
        float a0[4] = {bb.x, bb.y, bb.z, bb.w}, a1[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll 19
        for (int ii = 0; ii < HE_IN; ++ii) {
            const float x0 = S.x[s0][ii], x1 = S.x[s1][ii];
            const float4 w = *reinterpret_cast<const float4*>(&S.w0[ii][q * 4]);
            a0[0] += w.x * x0, a0[1] += w.y * x0, a0[2] += w.z * x0, a0[3] += w.w * x0;
            a1[0] += w.x * x1, a1[1] += w.y * x1, a1[2] += w.z * x1, a1[3] += w.w * x1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            S.proj[s0 * PROJ_S + t * 32 + q * 4 + j] = elu1(a0[j]);
            S.proj[s1 * PROJ_S + t * 32 + q * 4 + j] = elu1(a1[j]);
        }
    }
    __syncthreads();

    // ---- stage 1: Conv1d(30 -> 20, k4, s2), 4 output positions: quads q < 5 own channels q*4 .. q*4+3 ------------------
    if (q < HE_C1 / 4) {
        float acc[2][HE_L1][4];
        const float4 bb = *reinterpret_cast<const float4*>(&S.b1[q * 4]);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int p = 0; p < HE_L1; ++p) acc[e][p][0] = bb.x, acc[e][p][1] = bb.y, acc[e][p][2] = bb.z, acc[e][p][3] = bb.w;
#pragma unroll 2
        for (int c = 0; c < HE_C0; ++c) {
            float v0[HE_T], v1[HE_T];
#pragma unroll
            for (int t = 0; t < HE_T; ++t) {
                v0[t] = S.proj[s0 * PROJ_S + t * 32 + c];
                v1[t] = S.proj[s1 * PROJ_S + t * 32 + c];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 w = *reinterpret_cast<const float4*>(&S.w1[c * 4 + k][q * 4]);
#pragma unroll
                for (int p = 0; p < HE_L1; ++p) {
                    const float u0 = v0[2 * p + k], u1 = v1[2 * p + k];
                    acc[0][p][0] += w.x * u0, acc[0][p][1] += w.y * u0, acc[0][p][2] += w.z * u0, acc[0][p][3] += w.w * u0;
                    acc[1][p][0] += w.x * u1, acc[1][p][1] += w.y * u1, acc[1][p][2] += w.z * u1, acc[1][p][3] += w.w * u1;
                }
            }
        }
#pragma unroll
        for (int p = 0; p < HE_L1; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                S.c1[s0 * C1_S + p * HE_C1 + q * 4 + j] = elu1(acc[0][p][j]);
                S.c1[s1 * C1_S + p * HE_C1 + q * 4 + j] = elu1(acc[1][p][j]);
            }
    }
    __syncthreads();

    // ---- stage 2: Conv1d(20 -> 10, k2, s1), 3 output positions: quads q < 5 own channels 2q, 2q+1 -----------------------
    if (q < HE_C2 / 2) {
        float acc[2][HE_L2][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int p = 0; p < HE_L2; ++p) acc[e][p][0] = S.b2[2 * q], acc[e][p][1] = S.b2[2 * q + 1];
#pragma unroll 4
        for (int c = 0; c < HE_C1; ++c) {
            float v0[HE_L1], v1[HE_L1];
#pragma unroll
            for (int p = 0; p < HE_L1; ++p) {
                v0[p] = S.c1[s0 * C1_S + p * HE_C1 + c];
                v1[p] = S.c1[s1 * C1_S + p * HE_C1 + c];
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float2 w = *reinterpret_cast<const float2*>(&S.w2[c * 2 + k][2 * q]);
#pragma unroll
                for (int p = 0; p < HE_L2; ++p) {
                    acc[0][p][0] += w.x * v0[p + k], acc[0][p][1] += w.y * v0[p + k];
                    acc[1][p][0] += w.x * v1[p + k], acc[1][p][1] += w.y * v1[p + k];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < HE_L2; ++p)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                S.c2[s0 * C2_S + (2 * q + j) * HE_L2 + p] = elu1(acc[0][p][j]);
                S.c2[s1 * C2_S + (2 * q + j) * HE_L2 + p] = elu1(acc[1][p][j]);
            }
    }
    __syncthreads();

    // ---- stage 3: Linear(30 -> 29) + ELU: quad q owns outputs q*4 .. q*4+3 -------------------------------------------------
    {
        const float4 bb = *reinterpret_cast<const float4*>(&S.b3[q * 4]);
        float a0[4] = {bb.x, bb.y, bb.z, bb.w}, a1[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll 10
        for (int ii = 0; ii < HE_C0; ++ii) {
            const float x0 = S.c2[s0 * C2_S + ii], x1 = S.c2[s1 * C2_S + ii];
            const float4 w = *reinterpret_cast<const float4*>(&S.w3[ii][q * 4]);
            a0[0] += w.x * x0, a0[1] += w.y * x0, a0[2] += w.z * x0, a0[3] += w.w * x0;
            a1[0] += w.x * x1, a1[1] += w.y * x1, a1[2] += w.z * x1, a1[3] += w.w * x1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = q * 4 + j;
            if (o < HE_OUT) {
                if (m0 + s0 < a.M) a.out[(size_t)(m0 + s0) * a.out_pitch + o] = elu1(a0[j]);
                if (m0 + s1 < a.M) a.out[(size_t)(m0 + s1) * a.out_pitch + o] = elu1(a1[j]);
            }
        }
    }
}

extern "C" int qa_hist_encoder_fwd(const QaHistEncArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->M == 0) return 0;
    QA_CHECK_PTR(a->hist);
    QA_CHECK_PTR(a->w0);
    QA_CHECK_PTR(a->b0);
    QA_CHECK_PTR(a->w1);
    QA_CHECK_PTR(a->b1);
    QA_CHECK_PTR(a->w2);
    QA_CHECK_PTR(a->b2);
    QA_CHECK_PTR(a->w3);
    QA_CHECK_PTR(a->b3);
    QA_CHECK_PTR(a->out);
    if (a->M < 0 || a->hist_pitch < HE_T * HE_IN || a->w0_pitch < HE_IN || a->w3_pitch < HE_C0 || a->out_pitch < HE_OUT)
        return QA_EINVAL;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_hist_encoder, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeSmem));
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const unsigned blocks = (unsigned)((a->M + HE_S - 1) / HE_S);
    k_hist_encoder<<<blocks, HE_THREADS, sizeof(HeSmem), (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}
