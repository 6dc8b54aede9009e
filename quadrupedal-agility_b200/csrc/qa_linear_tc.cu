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
//   warps 2..9  epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks): tcgen05.ld 32x32b (lane = output
//               row) -> bias + activation in registers -> swizzled 32x32 slab in shared memory -> TMA store
//               (cp.async.bulk.tensor, or cp.reduce...add for split-K partial sums).
// The kernel is persistent (grid = min(#tiles, #SMs)) with two TMEM accumulators, so the epilogue of one tile overlaps
// the main loop of the next.
//
// CG = 2 (large M): the two CTAs of a cluster -- the two SMs of a TPC -- compute one 256 x BN tile with
// tcgen05.mma.cta_group::2.  Each CTA loads ITS 128 rows of A and HALF of the B tile (BN/2 rows), the leader CTA (cluster rank 0)
// issues the MMAs, which read both halves of B through the pair's shared memory; each CTA's TMEM holds its own 128 x BN
// accumulator and each CTA runs its own epilogue.  Per stage a CTA pulls (128 + BN/2) x 128 B out of L2 instead of
// (128 + BN) x 128 B for the same number of FLOPs: fp32 operands make this kernel L2->SM bound (43.7 FLOP per operand byte at
// 128 x 256, measured ~8 TB/s chip-wide => ~350 TFLOP/s; profiles/r1_ncu_k7_and_trainer_kernels.txt), so halving the B traffic
// is worth 1.5x.  Protocol: every TMA load of the pair completes on the LEADER's full barrier (expect_tx = both CTAs' bytes);
// the leader's tcgen05.commit multicasts the slot-free and accumulator-ready arrivals to both CTAs; both CTAs' epilogue warps
// arrive on the leader's accumulator-free barrier.
// (Tried and dropped: cp.async.bulk.prefetch.tensor of the A stream 12 K blocks ahead -- 3-20 % SLOWER on every layer shape,
// profiles/r2_k7_microbench_prefetch_on_rejected.txt vs r2_k7_microbench.txt: the prefetches compete with the demand loads for the same L2 request slots.)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "qa_b200.h"
#include "qa_common.cuh"
#include "qa_k2_common.cuh"   // mbarrier helpers

#define TC_BM 128
#define TC_BK 32              // floats per K block = 128 bytes = one swizzle span
#define TC_STAGES 3          // 3 x (16 KB A + <=16 KB B) = 96 KB -> two CTAs per SM: one's epilogue overlaps the other's main loop
#define TC_NEPI 8             // epilogue warps: two per TMEM lane quadrant, alternating 32-column chunks
#define TC_THREADS (64 + 32 * TC_NEPI)

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) { mbar_expect_tx(bar, bytes); }

// ---- cluster helpers (CG = 2) ----
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// (not .aligned: the role loops leave the lanes of warps 0 and 1 diverged when they get here)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// shared::cluster address of `p`'s twin in the CTA of rank `rank`
__device__ __forceinline__ unsigned mapa_shared(const void* p, unsigned rank) {
    unsigned a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// mbarrier wait that traps instead of hanging the device when a protocol error leaves it unsatisfied (~2 s)
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0, spins = 0;
    long long t0 = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && (++spins & 1023u) == 0u) {
            long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000LL) __trap();
        }
    }
}
// TMA load issued by either CTA of a pair; the bytes complete on the mbarrier at cluster address `bar_cluster` (the leader's)
__device__ __forceinline__ void tma_load_2d_cg2(void* sdst, const CUtensorMap* map, int c0, int c1, unsigned bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            (unsigned)__cvta_generic_to_shared(sdst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_cluster)
        : "memory");
}

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
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n, int m = TC_BM) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

__device__ __forceinline__ void umma_tf32_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive (when all MMAs issued so far have retired) on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// exp(x) - 1 for x <= 0 through the hardware ex2 (flush-to-zero form: no denormal range fix-up around the MUFU)
__device__ __forceinline__ float elu_neg(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y - 1.f;
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
    // EPI_ACTBWD (dx of layer L fused with the activation backward of layer L-1): out = acc * act'(yprev), db += colsum(out)
    const float* yprev;
    long long yprev_pitch;
    float* db;
    // offsets added to the contiguous-dimension TMA coordinate of each operand: a column window of a wider row-major tensor
    // (e.g. the 29 privileged-latent lanes at column 61 of the 671-wide observation row) is read / written in place
    int a_c0, b_c0, out_c0;
};

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_ACTBWD = 2 };

template <int BN, int STAGES, int EPI, int CG = 1>
struct TcSmem {
    float a[STAGES][TC_BM * TC_BK];         // 16 KB per stage, 1024 B aligned
    float b[STAGES][(BN / CG) * TC_BK];     // CG = 2: this CTA's half of the B tile
    float stg[TC_NEPI][2][32 * 32];         // per-epilogue-warp, double-buffered 32 x 32 output slabs (TMA store source)
    // EPI_ACTBWD: per-epilogue-warp 32 x 32 slab of the previous layer's OUTPUT, prefetched by TMA one chunk ahead (the loads do
    // not depend on the accumulator, so their latency hides behind the main loop); the slab is lifted into registers as soon as
    // it lands, which frees the buffer for the next request
    float ybuf[EPI == EPI_ACTBWD ? TC_NEPI : 1][EPI == EPI_ACTBWD ? 32 * 32 : 4];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t ybar[TC_NEPI];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* ssrc, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* ssrc, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// Persistent, warp-specialised tile loop.  grid = CG * min(#work items, #SMs / CG); a work item is (m tile of 128 * CG rows,
// n tile, K split) and belongs to one CTA (CG = 1) or one CTA pair (CG = 2).  Two TMEM accumulators (2 x BN columns) let the
// epilogue of item i overlap the main loop of item i+1; the shared-memory ring keeps streaming across items.
template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_yp,
            const __grid_constant__ TcProblem g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // 1024 B alignment for the 128 B swizzle atoms (the dynamic shared window starts at the same offset in both CTAs of a pair).
    // The pad is ADDED to the __shared__ array: a pointer rebuilt from an integer is a generic pointer to the compiler, and
    // every epilogue access then becomes a generic LD.E / ST.E that cannot be reordered against the others
    // (profiles/r2_k7_epilogue_generic_smem.txt: the epilogue warps were the bottleneck of every layer but the 671-wide one)
    using Smem = TcSmem<BN, STAGES, EPI, CG>;
    const unsigned pad = (1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw + pad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned rank = CG == 2 ? cluster_ctarank() : 0u;          // 0 = leader (issues the MMAs)
    const int unit = blockIdx.x / CG, units = gridDim.x / CG;        // persistent work unit: CTA or CTA pair
    constexpr int BM_EFF = TC_BM * CG, BN_LOAD = BN / CG;
    const int tiles_m = (g.Mo + BM_EFF - 1) / BM_EFF, tiles_n = (g.No + BN - 1) / BN;
    const int total_kb = (g.Kred + TC_BK - 1) / TC_BK;
    const int splits = (total_kb + g.kb_per_split - 1) / g.kb_per_split;
    const int total_items = tiles_m * tiles_n * splits;
    constexpr unsigned TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
    constexpr unsigned STAGE_BYTES = CG * (TC_BM + BN_LOAD) * TC_BK * 4;   // what lands on the (leader's) full barrier

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&S.tmem_full[a], 1);
            mbar_init(&S.tmem_empty[a], TC_NEPI * CG);               // one arrive per epilogue warp (of both CTAs)
        }
        for (int e = 0; e < TC_NEPI; ++e) mbar_init(&S.ybar[e], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {                       // TMEM allocation is a warp-wide instruction (CG = 2: the same warp of both CTAs)
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&S.tmem_base)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             (unsigned)__cvta_generic_to_shared(&S.tmem_base)),
                         "r"(TMEM_COLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();       // the peer's barriers are initialised before anything remote touches them
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = S.tmem_base;

    // work item -> (split, m tile, n tile); n fastest so that CTAs running concurrently share the A rows in L2.
    // m0 = first row of THIS CTA's 128-row slab; nb0 = first B row this CTA loads.
    auto decode = [&](int w, int& m0, int& n0, int& kb0, int& nkb) {
        const int tn = w % tiles_n;
        const int tm = (w / tiles_n) % tiles_m;
        const int sp = w / (tiles_n * tiles_m);
        m0 = tm * BM_EFF + (int)rank * TC_BM;
        n0 = tn * BN;
        kb0 = sp * g.kb_per_split;
        nkb = min(g.kb_per_split, total_kb - kb0);
    };
    auto wait = [&](uint64_t* bar, unsigned ph) {
        if (CG == 2) mbar_wait_guard(bar, ph);
        else mbar_wait(bar, ph);
    };

    if (warp == 0) {
        // ===== TMA producer (every CTA loads its own operand slices) =====
        if (lane == 0) {
            unsigned it = 0;
            for (int w = unit; w < total_items; w += units) {
                int m0, n0, kb0, nkb;
                decode(w, m0, n0, kb0, nkb);
                const int nb0 = n0 + (int)rank * BN_LOAD;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int kb = kb0 + i;
                    const int s = it % STAGES;
                    const unsigned ph = (it / STAGES) & 1;
                    wait(&S.empty[s], ph ^ 1);                       // slot free (passes immediately the first time)
                    if (CG == 2) {
                        const unsigned fb = mapa_shared(&S.full[s], 0);
                        if (rank == 0) mbar_arrive_expect_tx(&S.full[s], STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int c = 0; c < TC_BM / 32; ++c)
                                tma_load_2d_cg2(S.a[s] + c * 1024, &map_a, g.a_c0 + m0 + c * 32, kb * TC_BK, fb);
                        } else {
                            tma_load_2d_cg2(S.a[s], &map_a, g.a_c0 + kb * TC_BK, m0, fb);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int c = 0; c < BN_LOAD / 32; ++c)
                                tma_load_2d_cg2(S.b[s] + c * 1024, &map_b, g.b_c0 + nb0 + c * 32, kb * TC_BK, fb);
                        } else {
                            tma_load_2d_cg2(S.b[s], &map_b, g.b_c0 + kb * TC_BK, nb0, fb);
                        }
                    } else {
                        mbar_arrive_expect_tx(&S.full[s], STAGE_BYTES);
                        if (A_MN) {
#pragma unroll
                            for (int c = 0; c < TC_BM / 32; ++c)
                                tma_load_2d(S.a[s] + c * 1024, &map_a, g.a_c0 + m0 + c * 32, kb * TC_BK, &S.full[s]);
                        } else {
                            tma_load_2d(S.a[s], &map_a, g.a_c0 + kb * TC_BK, m0, &S.full[s]);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int c = 0; c < BN / 32; ++c)
                                tma_load_2d(S.b[s] + c * 1024, &map_b, g.b_c0 + n0 + c * 32, kb * TC_BK, &S.full[s]);
                        } else {
                            tma_load_2d(S.b[s], &map_b, g.b_c0 + kb * TC_BK, n0, &S.full[s]);
                        }
                    }
                }
            }
            if (CG == 2) {
                // drain: the leader's last slot-free arrivals are multicast into THIS CTA's barriers -- do not leave before
                // they have all landed (the wait each slot's next use would have done)
                for (int j = 0; j < STAGES; ++j, ++it) wait(&S.empty[it % STAGES], ((it / STAGES) & 1) ^ 1);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (CG = 2: the leader CTA only) =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(BN, BM_EFF) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
            unsigned it = 0, t = 0;
            for (int w = unit; w < total_items; w += units) {
                int m0, n0, kb0, nkb;
                decode(w, m0, n0, kb0, nkb);
                if (nkb <= 0) continue;
                const unsigned acc = t & 1, aph = (t >> 1) & 1;
                wait(&S.tmem_empty[acc], aph ^ 1);                   // epilogue(s) have drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_d + acc * BN;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES;
                    const unsigned ph = (it / STAGES) & 1;
                    wait(&S.full[s], ph);                            // TMA bytes (of both CTAs) have landed
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {            // UMMA_K = 8 tf32
                        const uint64_t adesc = A_MN ? umma_desc_mnmajor_sw128(S.a[s], k) : umma_desc_kmajor_sw128(S.a[s]) + 2 * k;
                        const uint64_t bdesc = B_MN ? umma_desc_mnmajor_sw128(S.b[s], k) : umma_desc_kmajor_sw128(S.b[s]) + 2 * k;
                        if (CG == 2) umma_tf32_cg2(d, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
                        else umma_tf32(d, adesc, bdesc, idesc, (i > 0 || k > 0) ? 1u : 0u);
                    }
                    if (CG == 2) umma_commit_cg2(&S.empty[s]);       // frees the slot (in both CTAs) when these MMAs retire
                    else umma_commit(&S.empty[s]);
                }
                if (CG == 2) umma_commit_cg2(&S.tmem_full[acc]);     // accumulator complete (both CTAs' epilogues)
                else umma_commit(&S.tmem_full[acc]);
                ++t;
            }
        }
    } else {
        // ===== epilogue: warps 2..9.  A warp can only read the TMEM lane quadrant (warp % 4), i.e. 32 output rows; the two warps
        // of a quadrant take the even / the odd 32-column chunks of every tile =====
        // TMEM -> registers (tcgen05.ld, lane = row, 32 columns) -> bias + activation -> 128-byte-swizzled 32 x 32
        // slab in shared memory -> TMA store (or TMA reduce-add for split-K).  TMA clips rows >= Mo / columns >= No.
        const int e = warp - 2, q = warp & 3, h = e >> 2;
        constexpr int CH = BN >= 32 ? 32 : 16;
        float* ybuf = S.ybuf[EPI == EPI_ACTBWD ? e : 0];
        unsigned t = 0, chunk = 0;                        // tiles seen / own chunks processed
        auto n_chunks = [&](int w) {                      // chunks of work item w (all warps)
            int m0, n0, kb0, nkb;
            decode(w, m0, n0, kb0, nkb);
            if (nkb <= 0) return 0;
            const int cols = min(BN, g.No - n0);
            return (cols + CH - 1) / CH;
        };
        // EPI_ACTBWD: the previous layer's output slab of this warp's NEXT chunk is requested (TMA, same 128 B swizzle as the
        // output slab) as soon as the current one has been lifted into registers; the very first request is issued before the
        // accumulator wait.
        auto y_request = [&](int w, int ci) {
            int m0, n0, kb0, nkb;
            decode(w, m0, n0, kb0, nkb);
            mbar_arrive_expect_tx(&S.ybar[e], 32 * 32 * 4);
            tma_load_2d(ybuf, &map_yp, n0 + ci * CH, m0 + q * 32, &S.ybar[e]);
        };
        auto y_request_after = [&](int w, int ci) {       // this warp's chunk after (w, ci), if any
            if (ci + 2 < n_chunks(w)) {
                y_request(w, ci + 2);
                return;
            }
            int w2 = w + units;
            while (w2 < total_items && n_chunks(w2) <= h) w2 += units;
            if (w2 < total_items) y_request(w2, h);
        };
        if (EPI == EPI_ACTBWD && lane == 0) {
            int w = unit;
            while (w < total_items && n_chunks(w) <= h) w += units;
            if (w < total_items) y_request(w, h);
        }
        const unsigned te_leader = CG == 2 ? mapa_shared(&S.tmem_empty[0], 0) : 0u;
        const bool bias_vec = (reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0;      // n0, c0 are multiples of 16 floats
        for (int w = unit; w < total_items; w += units) {
            int m0, n0, kb0, nkb;
            decode(w, m0, n0, kb0, nkb);
            if (nkb <= 0) continue;
            const int nch = n_chunks(w);
            const unsigned acc = t & 1, aph = (t >> 1) & 1;
            wait(&S.tmem_full[acc], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int ci = h; ci < nch; ci += 2, ++chunk) {
                const int c0 = ci * CH;
                float* buf = S.stg[e][chunk & 1];
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // slab of 2 chunks ago is free
                __syncwarp();
                uint32_t r[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
                if (CH == 32) tmem_ld32(taddr, r);
                else tmem_ld16(taddr, r);
                // side operands into registers up front (the compiler will not move a shared load above the swizzled slab
                // stores below): the bias slice while the TMEM load is in flight, the y slab once it has landed
                float4 bv[CH / 4], yv4[CH / 4];
                if (EPI == EPI_STORE) {
                    // bias slice straight from global memory (<= 1 KB per tile: L1 resident after the first touch; every lane
                    // reads the same address, one broadcast transaction per load)
                    const float* bp = g.bias + n0 + c0;
                    if (g.bias != nullptr && bias_vec && n0 + c0 + CH <= g.No) {
#pragma unroll
                        for (int j4 = 0; j4 < CH / 4; ++j4) bv[j4] = __ldg(reinterpret_cast<const float4*>(bp) + j4);
                    } else {
#pragma unroll
                        for (int j4 = 0; j4 < CH / 4; ++j4) {
                            float t4[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                t4[k] = (g.bias != nullptr && n0 + c0 + j4 * 4 + k < g.No) ? __ldg(bp + j4 * 4 + k) : 0.f;
                            bv[j4] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                        }
                    }
                }
                if (EPI == EPI_ACTBWD) {
                    wait(&S.ybar[e], chunk & 1);
#pragma unroll
                    for (int j4 = 0; j4 < CH / 4; ++j4)
                        yv4[j4] = *reinterpret_cast<const float4*>(ybuf + lane * 32 + ((j4 ^ (lane & 7)) << 2));
                    __syncwarp();                                    // every lane holds its row: the buffer may be refilled
                    if (lane == 0) y_request_after(w, ci);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (EPI == EPI_STORE && (g.out_c0 & 3) != 0) {
                    // output window that does not start on a 16-byte boundary (e.g. the 29 latent lanes at column 61 of the
                    // actor's input row): TMA needs 16-byte aligned box starts, so these few columns leave through plain stores
                    const int row = m0 + q * 32 + lane;
                    if (row < g.Mo) {
                        float* o = g.out + (size_t)row * g.out_pitch + g.out_c0 + n0 + c0;
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            const float4 b4 = bv[j >> 2];
                            float x = __uint_as_float(r[j]) + ((j & 3) == 0 ? b4.x : ((j & 3) == 1 ? b4.y : ((j & 3) == 2 ? b4.z : b4.w)));
                            if (g.act == 1) x = x > 0.f ? x : elu_neg(x);
                            else if (g.act == 2) x = fmaxf(x, 0.f);
                            if (n0 + c0 + j < g.No) o[j] = x;
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int j4 = 0; j4 < CH / 4; ++j4) {
                    float v[4];
                    const float b4[4] = {bv[j4].x, bv[j4].y, bv[j4].z, bv[j4].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float x = __uint_as_float(r[j4 * 4 + k]);
                        if (EPI == EPI_STORE) {
                            x += b4[k];
                            if (g.act == 1) x = x > 0.f ? x : elu_neg(x);                   // ELU(alpha = 1)
                            else if (g.act == 2) x = fmaxf(x, 0.f);                         // ReLU
                        }
                        v[k] = x;
                    }
                    const int sw = lane * 32 + ((j4 ^ (lane & 7)) << 2);                // SWIZZLE_128B position of (row, chunk j4)
                    if (EPI == EPI_ACTBWD) {
                        // gradient w.r.t. the previous layer's pre-activation: multiply by act'(z) recovered from its OUTPUT
                        // (rows >= Mo / columns >= No of the slab are TMA zero fill; the store clips them anyway)
                        const float yv[4] = {yv4[j4].x, yv4[j4].y, yv4[j4].z, yv4[j4].w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (g.act == 1) v[k] = yv[k] > 0.f ? v[k] : v[k] * (yv[k] + 1.0f);      // ELU'
                            else if (g.act == 2) v[k] = yv[k] > 0.f ? v[k] : 0.f;                   // ReLU'
                        }
                    }
                    *reinterpret_cast<float4*>(buf + sw) = make_float4(v[0], v[1], v[2], v[3]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (EPI == EPI_ACTBWD && g.db != nullptr) {
                    // bias gradient of the previous layer: column sums of the slab (rows >= Mo hold exact zeros)
                    const int col = n0 + c0 + lane;
                    float sum = 0.f;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) sum += buf[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))];
                    if (col < g.No) atomicAdd(g.db + col, sum);
                }
                if (lane == 0) {
                    if (EPI == EPI_ATOMIC) tma_reduce_add_2d(&map_y, buf, g.out_c0 + n0 + c0, m0 + q * 32);
                    else tma_store_2d(&map_y, buf, g.out_c0 + n0 + c0, m0 + q * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {                                         // this warp's TMEM reads of the accumulator are done
                if (CG == 2) mbar_arrive_cluster(te_leader + acc * (unsigned)sizeof(uint64_t));
                else mbar_arrive(&S.tmem_empty[acc]);
            }
            ++t;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync_all();       // neither CTA of the pair retires while the other may still reach into it
    else __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
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
                    bool mn_major, int box_cols = TC_BK) {
    PFN_encodeTiled enc = get_encode();
    if (enc == nullptr) return QA_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : QA_EINVAL;
}

// Operand description for the launcher: a row-major 2-D tensor (rows, cols, pitch).  K-major operand: rows = output index,
// cols = reduction index.  MN-major operand: rows = reduction index, cols = output index.  `c0` = first column of the window
// inside a wider row (the TMA map then spans columns [0, c0 + cols) so that reads past the window are zero fill).
struct TcOperand {
    const float* base;
    int64_t rows, cols, pitch;
    int c0;
};

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ring depth: as many stages as fit next to the epilogue buffers in 227 KB (<= 8)
template <int BN, int EPI, int CG>
constexpr int tc_stages() {
    constexpr size_t fixed = sizeof(TcSmem<BN, 1, EPI, CG>) - (size_t)(TC_BM + BN / CG) * TC_BK * 4;
    constexpr size_t per = (size_t)(TC_BM + BN / CG) * TC_BK * 4 + 16;
    constexpr size_t n = (227 * 1024 - 1024 - fixed) / per;
    return n > 8 ? 8 : (int)n;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int CG>
static int launch_gemm(const TcOperand& A, const TcOperand& B, TcProblem prob, int splits, cudaStream_t stream) {
    constexpr int STAGES = tc_stages<BN, EPI, CG>();
    static_assert(STAGES >= 3 || (EPI == EPI_ACTBWD && CG == 1 && BN == 256 && STAGES >= 2), "ring depth");
    static_assert(sizeof(TcSmem<BN, STAGES, EPI, CG>) + 1024 <= 227 * 1024, "shared-memory budget");
    static_assert(CG == 1 || (BN >= 64 && (BN / CG) % 32 == 0), "a CTA pair splits the B tile into two halves of whole 32-row chunks");
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A.base, A.rows, A.c0 + A.cols, A.pitch, A_MN ? 32 : TC_BM, A_MN);
    if (rc) return rc;
    rc = make_map(&mb, B.base, B.rows, B.c0 + B.cols, B.pitch, B_MN ? 32 : BN / CG, B_MN);
    if (rc) return rc;
    prob.a_c0 = A.c0;
    prob.b_c0 = B.c0;
    CUtensorMap my, myp;                                             // output slabs: 32 rows x 32 columns, 128 B swizzle
    rc = make_map(&my, prob.out, prob.Mo, prob.out_c0 + prob.No, prob.out_pitch, 32, false, 32);
    if (rc) return rc;
    myp = my;
    if (EPI == EPI_ACTBWD) {
        rc = make_map(&myp, prob.yprev, prob.Mo, prob.No, prob.yprev_pitch, 32, false, 32);
        if (rc) return rc;
    }
    const size_t smem = sizeof(TcSmem<BN, STAGES, EPI, CG>) + 1024;
    auto kern = k_gemm_tf32<BN, STAGES, A_MN, B_MN, EPI, CG>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const int items = ((prob.Mo + TC_BM * CG - 1) / (TC_BM * CG)) * ((prob.No + BN - 1) / BN) * splits;
    const int units = num_sms() / CG;
    const int grid = CG * (items < units ? items : units);
    if (CG == 1) {
        kern<<<grid, TC_THREADS, smem, stream>>>(ma, mb, my, myp, prob);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CG, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mb, my, myp, prob);
        if (e != cudaSuccess) return (int)e;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// Tile width: the widest BN that still yields enough work items to occupy the chip (wide tiles read A fewer times).
static int pick_bn(int Mo, int No, int splits, bool mn_major_b) {
    const int tiles_m = (Mo + TC_BM - 1) / TC_BM;
    const int cands[5] = {256, 128, 64, 32, 16};
    // the epilogue stores 32-column slabs: a 16-wide tile is only legal when it is the single n tile (No <= 16)
    const int smallest = (mn_major_b || No > 16) ? 32 : 16;
    int need = 16;
    while (need < No && need < 256) need <<= 1;                      // smallest power of two >= No (capped)
    if (need < smallest) need = smallest;
    int best = need;
    for (int i = 0; i < 5; ++i) {
        const int bn = cands[i];
        if (bn > need || bn < smallest) continue;
        best = bn;
        if (tiles_m * ((No + bn - 1) / bn) * splits >= 120) break;   // enough items for 148 SMs
    }
    return best;
}

// QA_TC_CG2=0 switches the CTA-pair variant off (A/B measurements)
static bool cg2_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("QA_TC_CG2");
        on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

// Tall problems (the PPO minibatch: M = 24576): CTA pairs.  The kernel is L2->SM bound, so the cost of a tiling is
// (rounds of the persistent loop) x (operand bytes one CTA pulls per K block) ~ ceil(items / units) x (128 + BN / CG); returns
// the BN of the cheapest pair tiling, or 0 when a single-CTA tiling is at least as cheap / the problem is not tall.
static int pick_bn_cg2(int Mo, int No, int bn1) {
    if (!cg2_enabled() || Mo < 8192 || No < 33) return 0;
    const int units2 = num_sms() / 2;
    auto rounds = [](int items, int units) { return (items + units - 1) / units; };
    const long long cost1 = (long long)rounds(((Mo + 127) / 128) * ((No + bn1 - 1) / bn1), num_sms()) * (128 + bn1);
    int best = 0;
    long long best_cost = cost1;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        if (bn >= 2 * No && bn > 64) continue;                       // more than half of the tile would be padding
        const long long c = (long long)rounds(((Mo + 255) / 256) * ((No + bn - 1) / bn), units2) * (128 + bn / 2);
        if (c < best_cost) best_cost = c, best = bn;
    }
    return best;
}

template <bool A_MN, bool B_MN, int EPI>
static int dispatch_bn(int bn, const TcOperand& A, const TcOperand& B, const TcProblem& p, int splits, cudaStream_t s, int cg = 1) {
    if (cg == 2) {
        switch (bn) {
            case 256: return launch_gemm<256, A_MN, B_MN, EPI, 2>(A, B, p, splits, s);
            case 128: return launch_gemm<128, A_MN, B_MN, EPI, 2>(A, B, p, splits, s);
            case 64: return launch_gemm<64, A_MN, B_MN, EPI, 2>(A, B, p, splits, s);
            default: return QA_EINVAL;
        }
    }
    switch (bn) {
        case 256: return launch_gemm<256, A_MN, B_MN, EPI, 1>(A, B, p, splits, s);
        case 128: return launch_gemm<128, A_MN, B_MN, EPI, 1>(A, B, p, splits, s);
        case 64: return launch_gemm<64, A_MN, B_MN, EPI, 1>(A, B, p, splits, s);
        case 32: return launch_gemm<32, A_MN, B_MN, EPI, 1>(A, B, p, splits, s);
        default: return launch_gemm<16, A_MN, B_MN, EPI, 1>(A, B, p, splits, s);
    }
}

static bool tma_ok(const void* p, int64_t pitch) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (pitch & 3) == 0; }

static TcProblem make_problem(int Mo, int No, int Kred, int kb_per_split, float* out, int64_t out_pitch, int out_c0) {
    TcProblem p{};
    p.Mo = Mo, p.No = No, p.Kred = Kred, p.kb_per_split = kb_per_split;
    p.out = out, p.out_pitch = out_pitch, p.out_c0 = out_c0;
    return p;
}

extern "C" int qa_linear_fwd(const QaLinearArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->M == 0) return 0;
    QA_CHECK_PTR(g->x);
    QA_CHECK_PTR(g->w);
    QA_CHECK_PTR(g->y);
    if (g->M < 0 || g->N <= 0 || g->K <= 0) return QA_EINVAL;
    if (g->act < 0 || g->act > 2 || g->x_col0 < 0 || g->y_col0 < 0) return QA_EINVAL;
    if (g->x_col0 & 3) return QA_EINVAL;          // TMA box starts are 16-byte aligned; y_col0 may be odd (plain-store epilogue)
    // TMA constraints: 16 B aligned bases and row pitches
    if (!tma_ok(g->x, g->x_pitch) || !tma_ok(g->w, g->w_pitch) || !tma_ok(g->y, g->y_pitch) || g->x_pitch < g->x_col0 + g->K ||
        g->w_pitch < g->K || g->y_pitch < g->y_col0 + g->N)
        return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const TcOperand A{g->x, g->M, g->K, g->x_pitch, g->x_col0}, B{g->w, g->N, g->K, g->w_pitch, 0};
    TcProblem p = make_problem(g->M, g->N, g->K, (g->K + TC_BK - 1) / TC_BK, g->y, g->y_pitch, g->y_col0);
    p.bias = g->bias, p.act = g->act;
    const int bn1 = pick_bn(g->M, g->N, 1, false);
    const int bn2 = (g->y_col0 & 3) == 0 ? pick_bn_cg2(g->M, g->N, bn1) : 0;
    if (bn2) return dispatch_bn<false, false, EPI_STORE>(bn2, A, B, p, 1, s, 2);
    return dispatch_bn<false, false, EPI_STORE>(bn1, A, B, p, 1, s);
}

// Backward of y = x W^T (+ b):  dx = gz W   (A = gz K-major, B = W MN-major, reduction over N)
//                               dw += gz^T x (A = gz MN-major, B = x MN-major, reduction over M, split-K + atomics)
extern "C" int qa_linear_bwd(const QaLinearBwdArgs* g, void* stream) {
    QA_CHECK_PTR(g);
    if (g->M == 0) return 0;
    QA_CHECK_PTR(g->gz);
    if (g->M < 0 || g->N <= 0 || g->K <= 0 || g->x_col0 < 0 || g->w_col0 < 0) return QA_EINVAL;
    if ((g->x_col0 & 3) || (g->w_col0 & 3)) return QA_EINVAL;      // TMA box starts are 16-byte aligned
    if (!tma_ok(g->gz, g->gz_pitch) || g->gz_pitch < g->N) return QA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = 0;
    if (g->dx != nullptr) {
        // dx (M, K) = gz (M, N) W[:, w_col0 : w_col0 + K]
        QA_CHECK_PTR(g->w);
        if (!tma_ok(g->w, g->w_pitch) || !tma_ok(g->dx, g->dx_pitch) || g->w_pitch < g->w_col0 + g->K || g->dx_pitch < g->K)
            return QA_EINVAL;
        const TcOperand A{g->gz, g->M, g->N, g->gz_pitch, 0}, B{g->w, g->N, g->K, g->w_pitch, g->w_col0};
        TcProblem p = make_problem(g->M, g->K, g->N, (g->N + TC_BK - 1) / TC_BK, g->dx, g->dx_pitch, 0);
        if (g->act_prev != 0) {
            // fused with the previous layer's activation backward (+ its bias gradient); y_prev is read through TMA
            QA_CHECK_PTR(g->y_prev);
            if (g->act_prev < 0 || g->act_prev > 2 || g->y_prev_pitch < g->K || !tma_ok(g->y_prev, g->y_prev_pitch)) return QA_EINVAL;
            if (g->db_prev != nullptr && !g->db_accumulate) {
                cudaError_t e = cudaMemsetAsync(g->db_prev, 0, sizeof(float) * g->K, s);
                if (e != cudaSuccess) return (int)e;
            }
            p.act = g->act_prev, p.yprev = g->y_prev, p.yprev_pitch = g->y_prev_pitch, p.db = g->db_prev;
            const int bn1 = pick_bn(g->M, g->K, 1, true), bn2 = pick_bn_cg2(g->M, g->K, bn1);
            rc = bn2 ? dispatch_bn<false, true, EPI_ACTBWD>(bn2, A, B, p, 1, s, 2) : dispatch_bn<false, true, EPI_ACTBWD>(bn1, A, B, p, 1, s);
        } else {
            const int bn1 = pick_bn(g->M, g->K, 1, true), bn2 = pick_bn_cg2(g->M, g->K, bn1);
            rc = bn2 ? dispatch_bn<false, true, EPI_STORE>(bn2, A, B, p, 1, s, 2) : dispatch_bn<false, true, EPI_STORE>(bn1, A, B, p, 1, s);
        }
        if (rc) return rc;
    }
    if (g->dw != nullptr) {
        // dw (N, K) += gz^T x[:, x_col0 : x_col0 + K]
        QA_CHECK_PTR(g->x);
        if (!tma_ok(g->x, g->x_pitch) || !tma_ok(g->dw, g->dw_pitch) || g->x_pitch < g->x_col0 + g->K || g->dw_pitch < g->K)
            return QA_EINVAL;
        const TcOperand A{g->gz, g->M, g->N, g->gz_pitch, 0}, B{g->x, g->M, g->K, g->x_pitch, g->x_col0};
        const int total_kb = (g->M + TC_BK - 1) / TC_BK;
        // CTA pairs when the weight has >= 256 output features (the MMA's M) and the reduction is long; wide n tiles (fewer
        // operand bytes per FLOP) as far as the in-features go
        const int cg = (cg2_enabled() && g->N >= 256 && g->M >= 8192 && g->K > 32) ? 2 : 1;
        const int bn = g->K <= 32 ? 32 : (g->K <= 64 ? 64 : ((g->K <= 160 || !cg2_enabled()) ? 128 : 256));
        const int tiles = ((g->N + TC_BM * cg - 1) / (TC_BM * cg)) * ((g->K + bn - 1) / bn);
        const int units = num_sms() / cg;
        int splits = cg == 2 ? (2 * units) / tiles : (2 * units + tiles - 1) / tiles;      // ~2 items per CTA (pair)
        if (splits > total_kb) splits = total_kb;
        if (splits < 1) splits = 1;
        const int per = (total_kb + splits - 1) / splits;
        splits = (total_kb + per - 1) / per;
        TcProblem p = make_problem(g->N, g->K, g->M, per, g->dw, g->dw_pitch, 0);
        rc = dispatch_bn<true, true, EPI_ATOMIC>(bn, A, B, p, splits, s, cg);
    }
    return rc;
}
