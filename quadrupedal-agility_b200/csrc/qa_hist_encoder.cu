// K11: StateHistoryEncoder forward, fused (tsteps = 10) -- bbc/rsl_rl/modules/actor_critic.py:9-59:
//   per time step  Linear(57 -> 30) + ELU
//   Conv1d(30 -> 20, k = 4, s = 2) + ELU      (length 10 -> 4)
//   Conv1d(20 -> 10, k = 2, s = 1) + ELU      (length 4 -> 3)
//   Flatten (channel-major) -> Linear(30 -> 29) + ELU
// The reference runs this as 1 GEMM on (10 M, 57), two cuDNN convolutions with layout transposes and 1 GEMM, i.e. ~8
// launches that stream a (10 M, 30) intermediate through HBM.  Here one block owns 32 samples: all weights (5.4 K floats)
// live in shared memory, each time step's 32 x 57 input tile is staged with coalesced loads, and 4 threads per sample
// split the output channels of every stage; intermediates never leave shared memory.  Forward only: in the PPO update
// the history latent is computed under inference_mode (gail.py:349-351), and the rollout only needs the forward.
#include "qa_b200.h"
#include "qa_common.cuh"

#define HE_S 32                 // samples per block
#define HE_Q 4                  // threads per sample
#define HE_T 10
#define HE_IN 57
#define HE_C0 30
#define HE_C1 20
#define HE_L1 4
#define HE_C2 10
#define HE_L2 3
#define HE_OUT 29

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

struct HeSmem {
    float w0[HE_IN][32];                    // [i][c] (c padded to 32): thread q owns channels q*8 .. q*8+7
    float b0[32];
    float w1[HE_C0 * 4][HE_C1];             // [(c,k)][o]
    float b1[HE_C1];
    float w2[HE_C1 * 2][12];                // [(c,k)][o] (o padded to 12)
    float b2[12];
    float w3[HE_C0][32];                    // [i][j] (j padded to 32)
    float b3[32];
    float x[HE_S][HE_IN];                   // current time step's inputs
    float proj[HE_S][HE_T][32];             // [s][t][c]
    float c1[HE_S][HE_L1][HE_C1];           // [s][p][o]
    float c2[HE_S][32];                     // flattened [o*3+p]
};

__global__ void __launch_bounds__(HE_S* HE_Q) k_hist_encoder(QaHistEncArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeSmem& S = *reinterpret_cast<HeSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int s = tid / HE_Q, q = tid % HE_Q;
    const long long m0 = (long long)blockIdx.x * HE_S;

    // ---- weights -> shared memory (transposed so that a thread's output channels are contiguous) ----------
    for (int i = tid; i < HE_IN * 32; i += HE_S * HE_Q) {
        const int ii = i / 32, c = i % 32;
        S.w0[ii][c] = c < HE_C0 ? a.w0[(size_t)c * a.w0_pitch + ii] : 0.f;
    }
    for (int i = tid; i < HE_C0 * 4 * HE_C1; i += HE_S * HE_Q) {
        const int ck = i / HE_C1, o = i % HE_C1;                      // ck = c*4 + k
        S.w1[ck][o] = a.w1[(size_t)o * (HE_C0 * 4) + ck];
    }
    for (int i = tid; i < HE_C1 * 2 * 12; i += HE_S * HE_Q) {
        const int ck = i / 12, o = i % 12;                            // ck = c*2 + k
        S.w2[ck][o] = o < HE_C2 ? a.w2[(size_t)o * (HE_C1 * 2) + ck] : 0.f;
    }
    for (int i = tid; i < HE_C0 * 32; i += HE_S * HE_Q) {
        const int ii = i / 32, j = i % 32;
        S.w3[ii][j] = j < HE_OUT ? a.w3[(size_t)j * a.w3_pitch + ii] : 0.f;
    }
    if (tid < 32) {
        S.b0[tid] = tid < HE_C0 ? a.b0[tid] : 0.f;
        S.b3[tid] = tid < HE_OUT ? a.b3[tid] : 0.f;
        if (tid < HE_C1) S.b1[tid] = a.b1[tid];
        if (tid < 12) S.b2[tid] = tid < HE_C2 ? a.b2[tid] : 0.f;
    }

    // ---- stage 0: per-step projection 57 -> 30 (+ELU) ---------------------------------------------------------
    for (int t = 0; t < HE_T; ++t) {
        __syncthreads();                                             // weights ready / previous tile consumed
        for (int i = tid; i < HE_S * HE_IN; i += HE_S * HE_Q) {
            const int ss = i / HE_IN, ii = i % HE_IN;
            const long long m = m0 + ss;
            S.x[ss][ii] = m < a.M ? a.hist[(size_t)m * a.hist_pitch + t * HE_IN + ii] : 0.f;
        }
        __syncthreads();
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = S.b0[q * 8 + j];
#pragma unroll 3
        for (int ii = 0; ii < HE_IN; ++ii) {
            const float xv = S.x[s][ii];
            const float4 wa = *reinterpret_cast<const float4*>(&S.w0[ii][q * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&S.w0[ii][q * 8 + 4]);
            acc[0] += wa.x * xv, acc[1] += wa.y * xv, acc[2] += wa.z * xv, acc[3] += wa.w * xv;
            acc[4] += wb.x * xv, acc[5] += wb.y * xv, acc[6] += wb.z * xv, acc[7] += wb.w * xv;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) S.proj[s][t][q * 8 + j] = elu1(acc[j]);
    }
    __syncthreads();

    // ---- stage 1: Conv1d(30 -> 20, k4, s2): thread q owns output channels q*5 .. q*5+4 ---------------------------
    for (int p = 0; p < HE_L1; ++p) {
        float acc[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[j] = S.b1[q * 5 + j];
        for (int c = 0; c < HE_C0; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v = S.proj[s][2 * p + k][c];
                const float* w = &S.w1[c * 4 + k][q * 5];
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[j] += w[j] * v;
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) S.c1[s][p][q * 5 + j] = elu1(acc[j]);
    }
    __syncthreads();

    // ---- stage 2: Conv1d(20 -> 10, k2, s1): thread q owns output channels q*3 .. (padded to 12) -----------------
    for (int p = 0; p < HE_L2; ++p) {
        float acc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] = S.b2[q * 3 + j];
        for (int c = 0; c < HE_C1; ++c) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float v = S.c1[s][p + k][c];
                const float* w = &S.w2[c * 2 + k][q * 3];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[j] += w[j] * v;
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int o = q * 3 + j;
            if (o < HE_C2) S.c2[s][o * HE_L2 + p] = elu1(acc[j]);       // nn.Flatten of (channels, length)
        }
    }
    __syncthreads();

    // ---- stage 3: Linear(30 -> 29) + ELU: thread q owns outputs q*8 .. q*8+7 ---------------------------------------
    {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = S.b3[q * 8 + j];
        for (int ii = 0; ii < HE_C0; ++ii) {
            const float v = S.c2[s][ii];
            const float4 wa = *reinterpret_cast<const float4*>(&S.w3[ii][q * 8]);
            const float4 wb = *reinterpret_cast<const float4*>(&S.w3[ii][q * 8 + 4]);
            acc[0] += wa.x * v, acc[1] += wa.y * v, acc[2] += wa.z * v, acc[3] += wa.w * v;
            acc[4] += wb.x * v, acc[5] += wb.y * v, acc[6] += wb.z * v, acc[7] += wb.w * v;
        }
        const long long m = m0 + s;
        if (m < a.M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int o = q * 8 + j;
                if (o < HE_OUT) a.out[(size_t)m * a.out_pitch + o] = elu1(acc[j]);
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
    k_hist_encoder<<<blocks, HE_S * HE_Q, sizeof(HeSmem), (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}
