// K31: in-place all-reduce(SUM) of the PPO gradient arena over NVLink peer memory, fused with the gradient-norm pass of K8.
//
// env shards, ONE all-reduce of the PPO gradients per optimiser step (2.96 MB: actor-critic + estimator gradients +
// the KL scalar).  Over NCCL that collective costs 33 us (2 ranks) to 56 us (8 ranks) per step -- launch and protocol latency,
// not bandwidth -- in a 570 us minibatch step, and the fused clip + Adam (K8) that consumes it then needs a separate
// sum-of-squares launch per optimiser.  Here every rank maps the other ranks' arenas and control blocks (cudaIpc) and one kernel
// does, per rank r of W:
//
//   barrier     (flag exchange over peer memory, no fence) every rank's backward pass has finished -- this kernel is stream
//               ordered behind it, so a rank that has ENTERED the kernel has complete gradients;
//   reduce      rank r owns slice r (n / W elements): it reads that slice from all W arenas in rank order 0..W-1 (peer loads,
//               every load of the pass in flight), sums in fp32, stores the sum into its own arena and PUSHES it to every
//               peer as flagged words -- {value, epoch, value, epoch} per 16-byte store into the peer's staging area (the
//               scheme of NCCL's LL protocol: the 8-byte halves are written atomically, so a reader that sees the epoch sees
//               the value).  Every element is summed by exactly one rank in one order => bit-identical replicas;
//   poll        every rank spins on ITS staging words of the other W - 1 slices and unpacks them into its arena.  The data is
//               its own synchronisation: no second barrier and no system-scope fence (which costs ~7 us on this platform
//               whatever it has to drain: profiles/r2_k31_phase_trace.txt).  Receiving slice q also proves that rank q has
//               finished reading this rank's raw gradients, so the arena may be overwritten as soon as the kernel ends;
//   norm        while it holds the sums, the owner accumulates sum(g^2) per optimiser segment (K8's clip_grad_norm_ input);
//               the CTA partials travel as flagged words too and are added in one fixed order by the CTA that takes the last
//               ticket (bit-identical on every rank again).
//
// Epochs are monotonic call counts (no reset, safe under CUDA-graph replay; a rank cannot enter call e + 1, and overwrite
// staging words of call e, before every peer has signalled the barrier of call e + 1, i.e. has left call e).  One flag row per
// CTA, so CTA b of rank r pairs with CTA b of the other ranks.  Spins are bounded: a protocol error traps after ~2 s.
#include <stdint.h>
#include <stdlib.h>

#include "qa_b200.h"
#include "qa_common.cuh"

#define PA_CTAS 96            // x 512 threads x 2 float4 per thread: one rank's slice of the 2.96 MB arena in ONE round of loads
#define PA_THREADS 512

// layout of one rank's control block (uint32 words): [PA_CTAS][QA_PEER_MAX_RANKS] barrier flags, then
// [PA_CTAS][QA_PEER_MAX_RANKS][2] flagged partial norms {float, epoch}, then [PA_CTAS] local epochs, the local last-CTA ticket,
// and from PA_STAGE_WORD on the staging area: one {value, epoch} pair per arena element
#define PA_FLAG(b, r) ((b) * QA_PEER_MAX_RANKS + (r))
#define PA_NORM(b, r, k) (PA_CTAS * QA_PEER_MAX_RANKS + (((b) * QA_PEER_MAX_RANKS + (r)) * 2 + (k)) * 2)
#define PA_EPOCH(b) (PA_CTAS * QA_PEER_MAX_RANKS * 5 + (b))
#define PA_TICKET (PA_CTAS * QA_PEER_MAX_RANKS * 5 + PA_CTAS)
#define PA_STAGE_WORD (((PA_CTAS * QA_PEER_MAX_RANKS * 5 + PA_CTAS + 4) + 3) / 4 * 4)

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
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_guard(unsigned& spins, long long& t0) {
    if ((++spins & 4095u) == 0u) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        else if (t - t0 > 2000000000LL) __trap();
    }
}

// CTA-wide barrier with the same CTA of every other rank: thread q < W signals rank q and waits for rank q's signal.  Nothing is
// published through it (what the peers read afterwards was written by earlier kernels of this stream): no fence.
__device__ __forceinline__ void peer_barrier(const QaPeerAllreduceArgs& a, unsigned value) {
    __syncthreads();
    const int q = threadIdx.x;
    if (q < a.world_size) {
        st_relaxed_sys(a.ctrl[q] + PA_FLAG(blockIdx.x, a.rank), value);
        const unsigned* mine = a.ctrl[a.rank] + PA_FLAG(blockIdx.x, q);
        long long t0 = 0;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(mine) - value) < 0) spin_guard(spins, t0);   // monotonic epochs, wrap safe
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PA_THREADS) k_peer_allreduce(const __grid_constant__ QaPeerAllreduceArgs a) {
    __shared__ float s_red[2][PA_THREADS / 32];
    __shared__ unsigned s_last;
    const int W = a.world_size, r = a.rank;
    unsigned* my_ctrl = a.ctrl[r];
    const unsigned e = my_ctrl[PA_EPOCH(blockIdx.x)] + 1u;              // this call's epoch (the word is written only by this CTA)
    PSTAMP(0);
    peer_barrier(a, e);
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
            reinterpret_cast<float4*>(a.arena[r])[i] = s[u];
            const uint4 w0 = make_uint4(__float_as_uint(s[u].x), e, __float_as_uint(s[u].y), e);
            const uint4 w1 = make_uint4(__float_as_uint(s[u].z), e, __float_as_uint(s[u].w), e);
            for (int dp = 1; dp < W; ++dp) {
                const int p = (r + dp) % W;
                uint4* st = reinterpret_cast<uint4*>(a.ctrl[p] + PA_STAGE_WORD) + i * 2;
                st_volatile_v4(st, w0);
                st_volatile_v4(st + 1, w1);
            }
            const float c4[4] = {s[u].x, s[u].y, s[u].z, s[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long idx = i * 4 + k;
                if (idx < a.seg_split) sq0 += c4[k] * c4[k];
                else if (idx < a.norm_end) sq1 += c4[k] * c4[k];
            }
        }
    }
    // CTA partial norms -> every rank's control block (row of this CTA, column of this rank) as flagged words
    sq0 = warp_sum(sq0), sq1 = warp_sum(sq1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[0][warp] = sq0, s_red[1][warp] = sq1;
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < PA_THREADS / 32; ++k) t += s_red[threadIdx.x][k];
        const unsigned long long word = ((unsigned long long)e << 32) | (unsigned long long)__float_as_uint(t);
        for (int p = 0; p < W; ++p)
            *reinterpret_cast<volatile unsigned long long*>(a.ctrl[p] + PA_NORM(blockIdx.x, r, threadIdx.x)) = word;
    }
    PSTAMP(2);
    // poll: the reduced slices of the other ranks arrive in this rank's staging area
    const uint4* stage = reinterpret_cast<const uint4*>(my_ctrl + PA_STAGE_WORD);
    for (int dq = 1; dq < W; ++dq) {
        const int q = (r + W - dq) % W;                                 // the peer whose first target this rank was comes first
        const long long qlo = (long long)q * per, qhi = min(n4, qlo + per);
        for (long long i = qlo + (long long)blockIdx.x * PA_THREADS + threadIdx.x; i < qhi; i += stride) {
            uint4 w0, w1;
            long long t0 = 0;
            unsigned spins = 0;
            while (true) {
                w0 = ld_volatile_v4(stage + i * 2);
                w1 = ld_volatile_v4(stage + i * 2 + 1);
                if (w0.y == e && w0.w == e && w1.y == e && w1.w == e) break;
                spin_guard(spins, t0);
            }
            reinterpret_cast<float4*>(a.arena[r])[i] =
                make_float4(__uint_as_float(w0.x), __uint_as_float(w0.z), __uint_as_float(w1.x), __uint_as_float(w1.z));
        }
    }
    PSTAMP(3);
    __syncthreads();
    if (threadIdx.x == 0) {
        my_ctrl[PA_EPOCH(blockIdx.x)] = e;
        __threadfence();
        s_last = (atomicAdd(my_ctrl + PA_TICKET, 1u) == PA_CTAS - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 64) {
        // warp k adds the PA_CTAS x W partials of segment k as they arrive -- lane = CTA rows lane, lane + 32, ..., ranks in
        // order, then a fixed xor butterfly: the same order on every rank, fp64 like K8's own norm pass
        __threadfence();
        const int k = threadIdx.x >> 5;
        double t = 0.0;
        for (int b = lane; b < PA_CTAS; b += 32)
            for (int p = 0; p < W; ++p) {
                const volatile unsigned long long* slot = reinterpret_cast<const volatile unsigned long long*>(my_ctrl + PA_NORM(b, p, k));
                unsigned long long word;
                long long t0 = 0;
                unsigned spins = 0;
                while ((unsigned)((word = *slot) >> 32) != e) spin_guard(spins, t0);
                t += (double)__uint_as_float((unsigned)word);
            }
        t = warp_sum_d(t);
        if (lane == 0 && a.sumsq_out[k] != nullptr) {
            *a.sumsq_out[k] = t * (double)a.grad_scale * (double)a.grad_scale;
            if (a.step_inc[k] != nullptr) *a.step_inc[k] += 1;
        }
        if (threadIdx.x == 1 && a.scale_index >= 0) a.arena[r][a.scale_index] *= a.grad_scale;
        if (threadIdx.x == 0) my_ctrl[PA_TICKET] = 0u;
    }
    PSTAMP(4);
}

// ------------------------------------------------------------------------------------------------------------------------------
// Variant for more than two ranks: the owner writes the plain sums into slice r of ALL arenas and a second flag barrier behind a
// system-scope release publishes them.  The flagged-word version above doubles the write-back bytes and pushes them to W - 1
// peers: at 8 ranks it measured 43.1 us per call against 34.1 us for this one (profiles/r2_k31_phase_trace.txt); at 2 ranks
// 17.0 against 20.8.  Epochs advance by two per call here.
// ------------------------------------------------------------------------------------------------------------------------------
// CTA-wide barrier with the same CTA of every other rank.  Thread q < W signals rank q and waits for rank q's signal.
// `publish`: this CTA has written peer memory that the other side reads after the barrier.  Then EVERY thread fences its own
// stores at system scope first (in parallel: one NVLink round trip; a single thread's fence after the __syncthreads measured
// 8 us, profiles/r2_k31_phase_trace.txt) and the flag goes out relaxed behind the CTA barrier.  Without `publish` (barrier 1:
// what the peers read was written by earlier kernels of this stream) no fence is needed at all.
__device__ __forceinline__ void peer_barrier_fence(const QaPeerAllreduceArgs& a, unsigned value, bool publish) {
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

__global__ void __launch_bounds__(PA_THREADS) k_peer_allreduce_fence(const __grid_constant__ QaPeerAllreduceArgs a) {
    __shared__ float s_red[2][PA_THREADS / 32];
    const int W = a.world_size, r = a.rank;
    unsigned* my_ctrl = a.ctrl[r];
    const unsigned epoch = my_ctrl[PA_EPOCH(blockIdx.x)];               // written only by this CTA (thread 0, at the end)
    peer_barrier_fence(a, epoch + 1u, false);
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
    // CTA partial norms -> every rank's control block (row of this CTA, column of this rank)
    sq0 = warp_sum(sq0), sq1 = warp_sum(sq1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[0][warp] = sq0, s_red[1][warp] = sq1;
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < PA_THREADS / 32; ++k) t += s_red[threadIdx.x][k];
        for (int p = 0; p < W; ++p) reinterpret_cast<float*>(a.ctrl[p])[PA_NORM(blockIdx.x, r, threadIdx.x)] = t;      // (low word of the slot)
    }
    peer_barrier_fence(a, epoch + 2u, true);
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

// bytes of one rank's control block for an arena of n floats: flags + flagged norms + epochs + ticket + staging (8 B per element)
extern "C" long long qa_peer_ctrl_bytes(long long n) { return (long long)PA_STAGE_WORD * 4 + (n > 0 ? n : 0) * 8; }

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
    static int ll_max = -1;                                          // QA_PEER_LL_MAX: largest world size served by the flagged-word version
    if (ll_max < 0) {
        const char* e = getenv("QA_PEER_LL_MAX");
        ll_max = e != nullptr ? atoi(e) : 2;
    }
    if (a->world_size <= ll_max) k_peer_allreduce<<<PA_CTAS, PA_THREADS, 0, (cudaStream_t)stream>>>(*a);
    else k_peer_allreduce_fence<<<PA_CTAS, PA_THREADS, 0, (cudaStream_t)stream>>>(*a);
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
