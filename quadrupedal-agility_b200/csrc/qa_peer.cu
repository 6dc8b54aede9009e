// K31: in-place all-reduce(SUM) of the PPO gradient arena over NVLink peer memory, fused with the gradient-norm pass of K8.
//
// env shards, ONE all-reduce of the PPO gradients per optimiser step (2.96 MB: actor-critic + estimator gradients +
// the KL scalar).  Over NCCL that collective costs ~55 us per step at 2 ranks -- launch and protocol latency, not bandwidth -- in
// a 570 us minibatch step, and the fused clip + Adam (K8) that consumes it then needs a separate sum-of-squares launch per
// optimiser.  Here every rank maps the other ranks' arenas (cudaIpc) and one kernel does, per rank r of W:
//
//   barrier 1   (flag exchange over peer memory) every rank's backward pass has finished -- this kernel is stream ordered
//               behind it, so a rank that has ENTERED the kernel has complete gradients;
//   reduce      rank r owns slice r (n / W elements): it reads that slice from all W arenas in rank order 0..W-1, sums in
//               fp32, and writes the sum back into slice r of ALL W arenas (peer stores).  Slices are disjoint, so one rank's
//               write-back never touches what another rank is reading; every element is summed by exactly one rank in one
//               order => all ranks hold bit-identical sums;
//   norm        while it holds the sums, the owner accumulates sum(g^2) per optimiser segment (K8's clip_grad_norm_ input);
//               the partial norms are exchanged through the control blocks and added in one fixed order by the CTA that takes
//               the last ticket (bit-identical on every rank again);
//   barrier 2   all write-backs are visible everywhere before anything downstream (K13, K8, next step's memset) runs.
//
// Flags are monotonic epochs (no reset, safe under CUDA-graph replay); one flag row per CTA, so CTA b of rank r pairs with
// CTA b of the other ranks.  Spins are bounded: a protocol error traps after ~2 s instead of hanging the box.
#include <stdint.h>

#include "qa_b200.h"
#include "qa_common.cuh"

#define PA_CTAS 96            // x 512 threads x 2 float4 per thread: one rank's slice of the 2.96 MB arena in ONE round of loads
#define PA_THREADS 512

// layout of one rank's control block (uint32 words): [PA_CTAS][QA_PEER_MAX_RANKS] barrier flags, then
// [PA_CTAS][QA_PEER_MAX_RANKS][2] float partial norms, then [PA_CTAS] local epochs, then the local last-CTA ticket
#define PA_FLAG(b, r) ((b) * QA_PEER_MAX_RANKS + (r))
#define PA_NORM(b, r, k) (PA_CTAS * QA_PEER_MAX_RANKS + ((b) * QA_PEER_MAX_RANKS + (r)) * 2 + (k))
#define PA_EPOCH(b) (PA_CTAS * QA_PEER_MAX_RANKS * 3 + (b))
#define PA_TICKET (PA_CTAS * QA_PEER_MAX_RANKS * 3 + PA_CTAS)
#define PA_CTRL_WORDS (PA_CTAS * QA_PEER_MAX_RANKS * 3 + PA_CTAS + 4)

// optional phase trace (tools/k2_trace.py --build compiles with -DQA_PEER_TRACE; never in the product build): %globaltimer
// stamps of CTA 0 / thread 0
#ifdef QA_PEER_TRACE
__device__ long long g_peer_trace[8];
extern "C" int qa_peer_trace_dump(long long* host) { return (int)cudaMemcpyFromSymbol(host, g_peer_trace, sizeof(long long) * 8); }
#define PSTAMP(k)                                                          \
    do {                                                                   \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                         \
            long long t_;                                                  \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));         \
            g_peer_trace[k] = t_;                                          \
        }                                                                  \
    } while (0)
#else
#define PSTAMP(k) do { } while (0)
#endif

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_sys(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CTA-wide barrier with the same CTA of every other rank.  Thread q < W signals rank q and waits for rank q's signal.
// `publish`: this CTA has written peer memory that the other side reads after the barrier.  Then EVERY thread fences its own
// stores at system scope first (in parallel: one NVLink round trip; a single thread's fence after the __syncthreads measured
// 8 us, profiles/r2_k31_phase_trace.txt) and the flag goes out relaxed behind the CTA barrier.  Without `publish` (barrier 1:
// what the peers read was written by earlier kernels of this stream) no fence is needed at all.
__device__ __forceinline__ void peer_barrier(const QaPeerAllreduceArgs& a, unsigned value, bool publish) {
    if (publish) asm volatile("fence.acq_rel.sys;" ::: "memory");     // release is all that is needed (not membar.sys = fence.sc.sys)
    __syncthreads();
    const int q = threadIdx.x;
    if (q < a.world_size) {
        st_relaxed_sys(a.ctrl[q] + PA_FLAG(blockIdx.x, a.rank), value);
        const unsigned* mine = a.ctrl[a.rank] + PA_FLAG(blockIdx.x, q);
        long long t0 = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(mine) - value) < 0) {               // monotonic epochs, wrap safe
            if ((++spins & 4095u) == 0u) {
                long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 2000000000LL) __trap();
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PA_THREADS) k_peer_allreduce(const __grid_constant__ QaPeerAllreduceArgs a) {
    __shared__ float s_red[2][PA_THREADS / 32];
    const int W = a.world_size, r = a.rank;
    unsigned* my_ctrl = a.ctrl[r];
    const unsigned epoch = my_ctrl[PA_EPOCH(blockIdx.x)];               // written only by this CTA (thread 0, at the end)
    PSTAMP(0);
    peer_barrier(a, epoch + 1u, false);
    PSTAMP(1);
    // slice r, in float4 units; this CTA's share of it
    const long long n4 = a.n / 4;                                       // n % 4 == 0 (checked at launch)
    const long long per = (n4 + W - 1) / W;
    const long long lo = (long long)r * per, hi = min(n4, lo + per);
    float sq0 = 0.f, sq1 = 0.f;
    constexpr int U = 2;                                                // float4s per thread in flight per peer: NVLink round trips are
    const long long stride = (long long)PA_CTAS * PA_THREADS;           // ~2-3 us, so the loads of a whole pass are issued before any use
    for (long long i0 = lo + (long long)blockIdx.x * PA_THREADS + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            // .cg: peer lines must not be served from this SM's L1 (a previous step's copy of the same addresses)
            s[u] = i < hi ? __ldcg(reinterpret_cast<const float4*>(a.arena[0]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = 1; p < W; ++p) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + u * stride;
                v[u] = i < hi ? __ldcg(reinterpret_cast<const float4*>(a.arena[p]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) s[u].x += v[u].x, s[u].y += v[u].y, s[u].z += v[u].z, s[u].w += v[u].w;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= hi) continue;
            for (int p = 0; p < W; ++p) reinterpret_cast<float4*>(a.arena[p])[i] = s[u];
            const float c4[4] = {s[u].x, s[u].y, s[u].z, s[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long idx = i * 4 + k;
                if (idx < a.seg_split) sq0 += c4[k] * c4[k];
                else if (idx < a.norm_end) sq1 += c4[k] * c4[k];
            }
        }
    }
    PSTAMP(2);
    // CTA partial norms -> every rank's control block (row of this CTA, column of this rank)
    sq0 = warp_sum(sq0), sq1 = warp_sum(sq1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[0][warp] = sq0, s_red[1][warp] = sq1;
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < PA_THREADS / 32; ++k) t += s_red[threadIdx.x][k];
        for (int p = 0; p < W; ++p) reinterpret_cast<float*>(a.ctrl[p])[PA_NORM(blockIdx.x, r, threadIdx.x)] = t;
    }
    PSTAMP(3);
    peer_barrier(a, epoch + 2u, true);
    PSTAMP(4);
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        my_ctrl[PA_EPOCH(blockIdx.x)] = epoch + 2u;
        __threadfence();
        s_last = (atomicAdd(my_ctrl + PA_TICKET, 1u) == PA_CTAS - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 64) {
        // every CTA of this rank has passed barrier 2, i.e. the partial norms of all CTAs of all ranks have landed: warp k adds
        // the PA_CTAS x W partials of segment k -- lane = CTA rows lane, lane + 32, ..., ranks in order, then a fixed xor
        // butterfly: the same order on every rank, fp64 like K8's own norm pass
        __threadfence();
        const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const volatile float* c = reinterpret_cast<const volatile float*>(my_ctrl);
        double t = 0.0;
        for (int b = lane; b < PA_CTAS; b += 32)
            for (int p = 0; p < W; ++p) t += (double)c[PA_NORM(b, p, k)];
        t = warp_sum_d(t);
        if (lane == 0 && a.sumsq_out[k] != nullptr) {
            *a.sumsq_out[k] = t * (double)a.grad_scale * (double)a.grad_scale;
            if (a.step_inc[k] != nullptr) *a.step_inc[k] += 1;
        }
        if (threadIdx.x == 1 && a.scale_index >= 0) a.arena[r][a.scale_index] *= a.grad_scale;
        if (threadIdx.x == 0) my_ctrl[PA_TICKET] = 0u;
    }
}

extern "C" int qa_peer_ctrl_bytes(void) { return (int)(PA_CTRL_WORDS * sizeof(unsigned)); }

extern "C" int qa_peer_allreduce(const QaPeerAllreduceArgs* a, void* stream) {
    QA_CHECK_PTR(a);
    if (a->world_size < 2 || a->world_size > QA_PEER_MAX_RANKS || a->rank < 0 || a->rank >= a->world_size) return QA_EINVAL;
    if (a->n <= 0 || (a->n & 3) || a->seg_split < 0 || a->seg_split > a->n || a->norm_end < a->seg_split || a->norm_end > a->n)
        return QA_EINVAL;
    for (int p = 0; p < a->world_size; ++p) {
        QA_CHECK_PTR(a->arena[p]);
        QA_CHECK_PTR(a->ctrl[p]);
        if (reinterpret_cast<uintptr_t>(a->arena[p]) & 15u) return QA_EINVAL;
    }
    if (a->scale_index >= a->n) return QA_EINVAL;
    k_peer_allreduce<<<PA_CTAS, PA_THREADS, 0, (cudaStream_t)stream>>>(*a);
    QA_LAUNCH_RET();
}

// ---- IPC plumbing (cudaMalloc'ed, exportable buffers) ------------------------------------------------------------------------
extern "C" int qa_ipc_alloc(void** ptr, uint64_t bytes) {
    QA_CHECK_PTR(ptr);
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*ptr, 0, bytes);
    return e == cudaSuccess ? 0 : (int)e;
}
extern "C" int qa_ipc_free(void* ptr) { return (int)cudaFree(ptr); }
extern "C" int qa_ipc_get_handle(const void* ptr, uint8_t* handle64) {
    QA_CHECK_PTR(ptr);
    QA_CHECK_PTR(handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64, &h, 64);
    return 0;
}
extern "C" int qa_ipc_open_handle(const uint8_t* handle64, void** ptr) {
    QA_CHECK_PTR(handle64);
    QA_CHECK_PTR(ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int qa_ipc_close_handle(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
