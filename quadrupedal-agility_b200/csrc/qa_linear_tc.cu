// K7: Y = act(X W^T + b) on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// The one dense-contraction site of the hot path: every nn.Linear of ActorCritic / Estimator / Discriminator
// (bbc/rsl_rl/modules/actor_critic.py:96-129, modules/estimator.py:24-33, algorithms/discriminator.py:36-46, 64-69),
// with the bias add and the ELU / ReLU that follow each of them in the reference fused into the epilogue.
//
// Operands stay fp32 in HBM and are consumed as TF32 (kind::tf32: fp32 bit patterns, 10-bit mantissa used, fp32
// accumulate in TMEM) -- no conversion pass.  X is (M,K) row-major and W is (N,K) row-major (PyTorch's Linear layout),
// i.e. both are K-major UMMA operands.
//
// One CTA computes a 128 x BN tile of Y:
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of the 128x32 (A) and BNx32 (B) fp32 boxes, 128-byte
//               swizzle, into a STAGES-deep shared-memory ring; full/empty mbarriers.  Out-of-bounds rows / K columns
//               are zero-filled by TMA, so M, N, K need no padding (only 16-byte row pitches).
//   warp 1      allocates TMEM (BN fp32 columns) and issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) from
//               one elected lane, 4 MMAs per 32-float K block; tcgen05.commit releases ring slots and finally signals
//               the epilogue.
//   warps 2..5  epilogue: tcgen05.ld 32x32b (lane = output row) -> bias + activation in registers -> 16-byte global
//               stores of the row segment.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "qa_b200.h"
#include "qa_common.cuh"
#include "qa_k2_common.cuh"   // mbarrier helpers

#define TC_BM 128
#define TC_BK 32              // floats per K block = 128 bytes = one swizzle span
#define TC_STAGES 3          // 3 x (16 KB A + <=16 KB B) = 96 KB -> two CTAs per SM: one's epilogue overlaps the other's main loop
#define TC_THREADS 192

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) { mbar_expect_tx(bar, bytes); }

__device__ __forceinline__ void tma_load_2d(void* sdst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            (unsigned)__cvta_generic_to_shared(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle: rows at 128 B pitch, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(const void* smem) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem);
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows x 128 B, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
    return d;
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// UMMA shared-memory descriptor, MN-major 32-bit operand.  For tf32 the only MN-major layout the tensor core accepts
// is SWIZZLE_128B_BASE32B: 128-byte rows, 4-row atoms, the 32-byte chunk index XORed with (row & 3) -- what TMA writes
// with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The tile is a set of TMA boxes of 32 floats (the contiguous MN direction)
// x 32 reduction rows = 4096 B each: 4-row reduction atoms are 512 B apart (SBO), consecutive 32-float MN chunks are
// 4096 B apart (LBO).  `kgroup` selects the 8 reduction rows of one K=8 MMA.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(const void* smem, int kgroup) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)kgroup * 1024u;
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;                // leading byte offset: next 32-float MN chunk
    d |= (uint64_t)(512 >> 4) << 32;                 // stride byte offset: next 4-row reduction atom
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                          // layout type SWIZZLE_128B_BASE32B
    return d;
}

// A general tile problem  OUT[Mo, No] (+)= A[Mo, Kred] * B[No, Kred]^T  with each operand either K-major (reduction index
// contiguous in memory) or MN-major (output index contiguous).
struct TcProblem {
    int Mo, No, Kred;
    int kb_per_split;                   // reduction blocks (of 32) per blockIdx.z
    float* out;
    long long out_pitch;
    const float* bias;                  // store epilogue only
    int act;
};

template <int BN>
struct TcSmem {
    float a[TC_STAGES][TC_BM * TC_BK];      // 16 KB per stage, 1024 B aligned
    float b[TC_STAGES][BN * TC_BK];
    uint64_t full[TC_STAGES];
    uint64_t empty[TC_STAGES];
    uint64_t tmem_full;
    uint32_t tmem_base;
};

enum { EPI_STORE = 0, EPI_ATOMIC = 1 };

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ TcProblem g) {
    extern __shared__ unsigned char smem_raw[];
    // 1024 B alignment for the 128 B swizzle atoms
    TcSmem<BN>& S = *reinterpret_cast<TcSmem<BN>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
    const int total_kb = (g.Kred + TC_BK - 1) / TC_BK;
    const int kb0 = blockIdx.z * g.kb_per_split;
    const int num_kb = min(g.kb_per_split, total_kb - kb0);
    constexpr unsigned TMEM_COLS = BN < 32 ? 32 : BN;
    constexpr unsigned STAGE_BYTES = (TC_BM + BN) * TC_BK * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        mbar_init(&S.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {                       // TMEM allocation is a warp-wide instruction
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (unsigned)__cvta_generic_to_shared(&S.tmem_base)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = S.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && num_kb > 0) {
            for (int i = 0; i < num_kb; ++i) {
                const int kb = kb0 + i;
                const int s = i % TC_STAGES;
                const unsigned ph = (i / TC_STAGES) & 1;
                mbar_wait(&S.empty[s], ph ^ 1);                      // slot free (passes immediately the first time)
                mbar_arrive_expect_tx(&S.full[s], STAGE_BYTES);
                if (A_MN) {
#pragma unroll
                    for (int c = 0; c < TC_BM / 32; ++c)
                        tma_load_2d(S.a[s] + c * 1024, &map_a, m0 + c * 32, kb * TC_BK, &S.full[s]);
                } else {
                    tma_load_2d(S.a[s], &map_a, kb * TC_BK, m0, &S.full[s]);
                }
                if (B_MN) {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)
                        tma_load_2d(S.b[s] + c * 1024, &map_b, n0 + c * 32, kb * TC_BK, &S.full[s]);
                } else {
                    tma_load_2d(S.b[s], &map_b, kb * TC_BK, n0, &S.full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && num_kb > 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % TC_STAGES;
                const unsigned ph = (i / TC_STAGES) & 1;
                mbar_wait(&S.full[s], ph);                           // TMA bytes have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {                // UMMA_K = 8 tf32
                    const uint64_t adesc = A_MN ? umma_desc_mnmajor_sw128(S.a[s], k) : umma_desc_kmajor_sw128(S.a[s]) + 2 * k;
                    const uint64_t bdesc = B_MN ? umma_desc_mnmajor_sw128(S.b[s], k) : umma_desc_kmajor_sw128(S.b[s]) + 2 * k;
                    umma_tf32(tmem_d, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&S.empty[s]);                            // frees the slot when these MMAs retire
            }
            umma_commit(&S.tmem_full);                               // accumulator complete
        }
    } else if (num_kb > 0) {
        // ===== epilogue: warps 2..5 own TMEM lane quadrants (warp % 4) =====
        const int q = warp & 3;
        mbar_wait(&S.tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // The smem ring is idle once tmem_full has fired (every MMA that read it has retired): reuse it as a
        // per-warp 32x33 transpose buffer so that the global accesses are full 128-byte lines (lane = column).
        float* stg = reinterpret_cast<float*>(&S.a[0][0]) + q * (32 * 33);
        constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CH) {
            if (n0 + c0 >= g.No) break;
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if (CH == 32) tmem_ld32(taddr, r);
            else tmem_ld16(taddr, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < CH; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);      // lane = row
            __syncwarp();
            const int col = n0 + c0 + lane;                                             // lane = column from here on
            if (lane < CH && col < g.No) {
                const int rows = min(32, g.Mo - (m0 + q * 32));
                float* yp = g.out + (size_t)(m0 + q * 32) * g.out_pitch + col;
                if (EPI == EPI_ATOMIC) {
                    for (int rr = 0; rr < rows; ++rr) atomicAdd(yp + (size_t)rr * g.out_pitch, stg[rr * 33 + lane]);
                } else {
                    const float bcol = g.bias != nullptr ? __ldg(g.bias + col) : 0.f;   // one bias load per lane per chunk
                    if (g.act == 1) {
                        for (int rr = 0; rr < rows; ++rr) {
                            const float x = stg[rr * 33 + lane] + bcol;
                            yp[(size_t)rr * g.out_pitch] = x > 0.f ? x : expm1f(x);     // ELU(alpha = 1)
                        }
                    } else if (g.act == 2) {
                        for (int rr = 0; rr < rows; ++rr) yp[(size_t)rr * g.out_pitch] = fmaxf(stg[rr * 33 + lane] + bcol, 0.f);
                    } else {
                        for (int rr = 0; rr < rows; ++rr) yp[(size_t)rr * g.out_pitch] = stg[rr * 33 + lane] + bcol;
                    }
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D fp32 tensor (rows, cols) with `pitch` floats per row; box = (box_rows, 32 floats), 128 B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t pitch, int box_rows,
                    bool mn_major) {
    PFN_encodeTiled enc = get_encode();
    if (enc == nullptr) return QA_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : QA_EINVAL;
}

// Operand description for the launcher: a row-major 2-D tensor (rows, cols, pitch).  K-major operand: rows = output index,
// cols = reduction index.  MN-major operand: rows = reduction index, cols = output index.
struct TcOperand {
    const float* base;
    int64_t rows, cols, pitch;
};

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const TcOperand& A, const TcOperand& B, const TcProblem& prob, int splits, cudaStream_t stream) {
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A.base, A.rows, A.cols, A.pitch, A_MN ? 32 : TC_BM, A_MN);
    if (rc) return rc;
    rc = make_map(&mb, B.base, B.rows, B.cols, B.pitch, B_MN ? 32 : BN, B_MN);
    if (rc) return rc;
    const size_t smem = sizeof(TcSmem<BN>) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tf32<BN, A_MN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((prob.Mo + TC_BM - 1) / TC_BM, (prob.No + BN - 1) / BN, splits);
    k_gemm_tf32<BN, A_MN, B_MN, EPI><<<grid, TC_THREADS, smem, stream>>>(ma, mb, prob);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

static bool tma_ok(const void* p, int64_t pitch) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (pitch & 3) == 0; }

extern "C" int qa_linear_fwd(const QaLinearArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->M == 0) return 0;
    QA_CHECK_PTR(g->x);
    QA_CHECK_PTR(g->w);
    QA_CHECK_PTR(g->y);
    if (g->M < 0 || g->N <= 0 || g->K <= 0) return QA_EINVAL;
    if (g->act < 0 || g->act > 2) return QA_EINVAL;
    // TMA constraints: 16 B aligned bases and row pitches
    if (!tma_ok(g->x, g->x_pitch) || !tma_ok(g->w, g->w_pitch) || g->x_pitch < g->K || g->w_pitch < g->K || g->y_pitch < g->N)
        return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const TcOperand A{g->x, g->M, g->K, g->x_pitch}, B{g->w, g->N, g->K, g->w_pitch};
    TcProblem p{g->M, g->N, g->K, (g->K + TC_BK - 1) / TC_BK, g->y, g->y_pitch, g->bias, g->act};
    const int n = g->N;
    if (n <= 16) return launch_gemm<16, false, false, EPI_STORE>(A, B, p, 1, s);
    if (n <= 32) return launch_gemm<32, false, false, EPI_STORE>(A, B, p, 1, s);
    if (n <= 64) return launch_gemm<64, false, false, EPI_STORE>(A, B, p, 1, s);
    return launch_gemm<128, false, false, EPI_STORE>(A, B, p, 1, s);
}

// Backward of y = x W^T (+ b):  dx = gz W   (A = gz K-major, B = W MN-major, reduction over N)
//                               dw += gz^T x (A = gz MN-major, B = x MN-major, reduction over M, split-K + atomics)
extern "C" int qa_linear_bwd(const QaLinearBwdArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->M == 0) return 0;
    QA_CHECK_PTR(g->gz);
    if (g->M < 0 || g->N <= 0 || g->K <= 0) return QA_EINVAL;
    if (!tma_ok(g->gz, g->gz_pitch) || g->gz_pitch < g->N) return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = 0;
    if (g->dx != nullptr) {
        QA_CHECK_PTR(g->w);
        if (!tma_ok(g->w, g->w_pitch) || g->w_pitch < g->K || g->dx_pitch < g->K) return QA_EINVAL;
        const TcOperand A{g->gz, g->M, g->N, g->gz_pitch}, B{g->w, g->N, g->K, g->w_pitch};
        TcProblem p{g->M, g->K, g->N, (g->N + TC_BK - 1) / TC_BK, g->dx, g->dx_pitch, nullptr, 0};
        if (g->K <= 32) rc = launch_gemm<32, false, true, EPI_STORE>(A, B, p, 1, s);
        else if (g->K <= 64) rc = launch_gemm<64, false, true, EPI_STORE>(A, B, p, 1, s);
        else rc = launch_gemm<128, false, true, EPI_STORE>(A, B, p, 1, s);
        if (rc) return rc;
    }
    if (g->dw != nullptr) {
        QA_CHECK_PTR(g->x);
        if (!tma_ok(g->x, g->x_pitch) || g->x_pitch < g->K || g->dw_pitch < g->K) return QA_EINVAL;
        const TcOperand A{g->gz, g->M, g->N, g->gz_pitch}, B{g->x, g->M, g->K, g->x_pitch};
        const int total_kb = (g->M + TC_BK - 1) / TC_BK;
        const int bn = g->K <= 32 ? 32 : (g->K <= 64 ? 64 : 128);
        const int tiles = ((g->N + TC_BM - 1) / TC_BM) * ((g->K + bn - 1) / bn);
        int splits = (2 * 148 + tiles - 1) / tiles;
        if (splits > total_kb) splits = total_kb;
        if (splits < 1) splits = 1;
        const int per = (total_kb + splits - 1) / splits;
        splits = (total_kb + per - 1) / per;
        TcProblem p{g->N, g->K, g->M, per, g->dw, g->dw_pitch, nullptr, 0};
        if (bn == 32) rc = launch_gemm<32, true, true, EPI_ATOMIC>(A, B, p, splits, s);
        else if (bn == 64) rc = launch_gemm<64, true, true, EPI_ATOMIC>(A, B, p, splits, s);
        else rc = launch_gemm<128, true, true, EPI_ATOMIC>(A, B, p, splits, s);
    }
    return rc;
}
