// comm.cu — receiver-bin exchange between the GPUs of one node over peer memory (NVLink / NVSwitch), without NCCL.
//
// The multi-GPU form of the path shards primary rays (or pulses) over the GPUs; the only cross-GPU step is the commutative
// aggregation of the per-(receiver, path) bins (aggregation.cu:56-69 summed over all rays).  The bins are small (C4: 5,832
// bins x 48 B = 280 KB), so the exchange is latency, not bandwidth: two NCCL all-reduces (SUM fp64 + MIN int64) cost
// ~0.1 ms of a 2.7 ms pulse.  Here every rank owns an exchange block in its own HBM that its peers map (CUDA IPC between
// processes, plain peer access between the threads of one process):
//
//   header   arrived[q]  : sequence number rank q has published — written BY rank q into every peer's block
//   payload  two halves (sequence parity): this rank's sums[n][5] and mins[n]
//
//   k_xchg_publish   copies the rank's accumulators into its own payload half; the last CTA to finish issues a system-scope
//                    fence and stores the sequence number into arrived[rank] of every peer's header (remote 8-byte stores)
//   k_xchg_reduce    every CTA waits (acquire loads on its OWN header, bounded) until all ranks have published, then reads
//                    the peers' payload halves over NVLink and reduces them in rank order — every GPU forms the same sums
//                    in the same order, so the result is bitwise identical on all ranks and from run to run (NCCL's ring
//                    order is neither) — straight into the engine's accumulators, where myKernel2 / emission follow
//
// A half is rewritten two exchanges later; by then every peer has published the exchange in between, which it does only
// after its own k_xchg_reduce of this one has finished (stream order) — so no reader can still be on the old data.
// Like a collective, every rank must call rts_comm_allreduce_bins once per pulse.  A rank that never arrives makes the
// others give up after ~2 s and report RTS_ERR_STATE at the next collection instead of hanging the GPU.
#include "engine.h"
#include <cstring>

namespace {

constexpr unsigned XCHG_HEADER_BYTES = 512;
struct XchgHeader {
    unsigned long long arrived[RTS_MAX_WORLD];
    unsigned long long ctas_done;      // k_xchg_publish: CTAs finished in this launch
    unsigned long long timed_out;      // k_xchg_reduce gave up waiting at least once
    unsigned long long wait_ns, reduce_ns, calls;   // CTA 0 of k_xchg_reduce: time spent waiting for the peers' flags / in the whole kernel (rts_comm_stats)
};
static_assert(sizeof(XchgHeader) <= XCHG_HEADER_BYTES, "exchange header");

struct XchgPeers { char *block[RTS_MAX_WORLD]; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// words = n_bins * 6 u64 words: sums (5 per bin) then mins
__global__ void __launch_bounds__(256) k_xchg_publish(const unsigned long long *__restrict__ sums, const unsigned long long *__restrict__ mins,
                                                       unsigned long long n_bins, XchgPeers peers, uint32_t rank, uint32_t world,
                                                       unsigned long long half_words, unsigned long long seq)
{
    unsigned long long *mine = reinterpret_cast<unsigned long long *>(peers.block[rank] + XCHG_HEADER_BYTES) + (seq & 1ull) * half_words;
    const unsigned long long n_sum = n_bins * 5ull;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sum + n_bins;
         i += (unsigned long long)gridDim.x * blockDim.x)
        mine[i] = i < n_sum ? sums[i] : mins[i - n_sum];
    __syncthreads();
    if (threadIdx.x == 0) {
        XchgHeader *h = reinterpret_cast<XchgHeader *>(peers.block[rank]);
        __threadfence();            // cumulative: the CTA's stores (ordered before this by the barrier) before the count
        if (atomicAdd(&h->ctas_done, 1ull) == (unsigned long long)gridDim.x - 1ull) {   // every CTA's copy is done and fenced
            h->ctas_done = 0;
            __threadfence_system();
            for (uint32_t q = 0; q < world; q++)
                st_release_sys(&reinterpret_cast<XchgHeader *>(peers.block[q])->arrived[rank], seq);
        }
    }
}

__global__ void __launch_bounds__(256) k_xchg_reduce(double *__restrict__ sums, unsigned long long *__restrict__ mins, unsigned long long n_bins,
                                                      XchgPeers peers, uint32_t rank, uint32_t world, unsigned long long half_words,
                                                      unsigned long long seq, long long spin_clocks)
{
    XchgHeader *h = reinterpret_cast<XchgHeader *>(peers.block[rank]);
    __shared__ int s_ok;
    const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
    const unsigned long long t_in = stamp ? global_ns() : 0ull;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(&h->arrived[threadIdx.x]) < seq) {
            if (clock64() - t0 > spin_clocks) { s_ok = 0; h->timed_out = 1ull; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();
    const unsigned long long t_go = stamp ? global_ns() : 0ull;
    if (!s_ok) return;              // a peer never published: leave the local accumulators as they are, the host reports it
    const unsigned long long n_sum = n_bins * 5ull;
    const unsigned long long half = (seq & 1ull) * half_words;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_sum + n_bins;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        // all ranks' words first — independent loads, in flight together over NVLink (summing as they arrive would
        // serialise eight ~1.5 us round trips) — then the reduction in rank order: identical bits on every GPU
        unsigned long long v[RTS_MAX_WORLD];
#pragma unroll
        for (uint32_t q = 0; q < RTS_MAX_WORLD; q++)
            if (q < world) v[q] = ld_relaxed_sys(reinterpret_cast<const unsigned long long *>(peers.block[q] + XCHG_HEADER_BYTES) + half + i);
        if (i < n_sum) {
            double acc = 0.0;
#pragma unroll
            for (uint32_t q = 0; q < RTS_MAX_WORLD; q++)
                if (q < world) acc += __longlong_as_double((long long)v[q]);
            sums[i] = acc;
        } else {
            unsigned long long m = ~0ull;
#pragma unroll
            for (uint32_t q = 0; q < RTS_MAX_WORLD; q++)
                if (q < world) m = min(m, v[q]);
            mins[i - n_sum] = m;
        }
    }
    if (stamp) { h->wait_ns += t_go - t_in; h->reduce_ns += global_ns() - t_in; h->calls += 1ull; }
}

} // namespace

extern "C" int rts_comm_create(rts_engine *e, uint32_t rank, uint32_t world, uint64_t max_bins)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (world == 0 || world > RTS_MAX_WORLD || rank >= world) return rts_fail(RTS_ERR_ARG, "rank %u / world %u (at most %u GPUs)", rank, world, RTS_MAX_WORLD);
    if (max_bins == 0 || max_bins > (1ull << 24)) return rts_fail(RTS_ERR_ARG, "max_bins must be 1..2^24");
    RTS_CUDA(cudaSetDevice(e->device));
    rts_comm_destroy(e);
    Comm &c = e->comm;
    c.rank = rank; c.world = world; c.max_bins = max_bins;
    c.half_words = max_bins * 6ull;
    c.bytes = XCHG_HEADER_BYTES + 2ull * c.half_words * sizeof(unsigned long long);
    RTS_CUDA(cudaMalloc(&c.block, c.bytes));
    RTS_CUDA(cudaMemset(c.block, 0, c.bytes));
    RTS_CUDA(cudaDeviceSynchronize());
    c.peer[rank] = c.block;
    c.seq = 0;
    c.clock_khz = 1965000;
    cudaDeviceGetAttribute(&c.clock_khz, cudaDevAttrClockRate, e->device);
    c.connected = world == 1;
    return RTS_OK;
}

extern "C" int rts_comm_ipc_handle(rts_engine *e, void *handle64)
{
    if (!e || !handle64) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->comm.block) return rts_fail(RTS_ERR_STATE, "rts_comm_create first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    RTS_CUDA(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    RTS_CUDA(cudaIpcGetMemHandle(&h, e->comm.block));
    memcpy(handle64, &h, 64);
    return RTS_OK;
}

extern "C" int rts_comm_local_ptr(rts_engine *e, void **ptr)
{
    if (!e || !ptr) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->comm.block) return rts_fail(RTS_ERR_STATE, "rts_comm_create first");
    *ptr = e->comm.block;
    return RTS_OK;
}

extern "C" int rts_comm_connect_ipc(rts_engine *e, const void *handles)
{
    if (!e || !handles) return rts_fail(RTS_ERR_ARG, "NULL argument");
    Comm &c = e->comm;
    if (!c.block) return rts_fail(RTS_ERR_STATE, "rts_comm_create first");
    RTS_CUDA(cudaSetDevice(e->device));
    for (uint32_t q = 0; q < c.world; q++) {
        if (q == c.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + 64 * (size_t)q, 64);
        void *p = nullptr;
        cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            cudaGetLastError();
            return rts_fail(RTS_ERR_CUDA, "cudaIpcOpenMemHandle of rank %u's exchange block failed: %s", q, cudaGetErrorString(err));
        }
        c.peer[q] = static_cast<char *>(p);
        c.ipc_opened[q] = true;
    }
    c.connected = true;
    return RTS_OK;
}

extern "C" int rts_comm_connect_ptrs(rts_engine *e, void *const *ptrs)
{
    if (!e || !ptrs) return rts_fail(RTS_ERR_ARG, "NULL argument");
    Comm &c = e->comm;
    if (!c.block) return rts_fail(RTS_ERR_STATE, "rts_comm_create first");
    for (uint32_t q = 0; q < c.world; q++) {
        if (q == c.rank) continue;
        if (!ptrs[q]) return rts_fail(RTS_ERR_ARG, "peer %u: NULL block", q);
        c.peer[q] = static_cast<char *>(ptrs[q]);
    }
    c.connected = true;
    return RTS_OK;
}

// Enqueue the exchange of the last pulse's (dense) bins behind it on the engine's stream, then their finalisation.
extern "C" int rts_comm_allreduce_bins(rts_engine *e)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    NvtxRange nvtx_range("rts:comm_allreduce_bins (peer memory)");
    Comm &c = e->comm;
    if (!c.block || !c.connected) return rts_fail(RTS_ERR_STATE, "exchange not set up (rts_comm_create / rts_comm_connect_*)");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    if (e->bins_hashed) return rts_fail(RTS_ERR_STATE, "the last pulse used the sparse bin table: exchange it with rts_bins_compact_device / rts_bins_load_compact");
    const uint64_t nb = e->n_bins_dense;
    if (nb > c.max_bins) return rts_fail(RTS_ERR_CAPACITY, "%llu bins > the %llu the exchange block was created for", (unsigned long long)nb, (unsigned long long)c.max_bins);
    RTS_CUDA(cudaSetDevice(e->device));
    c.seq++;
    if (nb) {
        XchgPeers P;
        memset(&P, 0, sizeof(P));
        for (uint32_t q = 0; q < c.world; q++) P.block[q] = c.peer[q];
        const unsigned long long words = nb * 6ull;
        const int grid = (int)std::min<unsigned long long>((words + 255ull) / 256ull, (unsigned long long)e->num_sms * 4ull);
        const long long spin = (long long)c.clock_khz * 2000ll;      // ~2 s of SM clocks
        k_xchg_publish<<<grid, 256, 0, e->stream>>>(reinterpret_cast<const unsigned long long *>(e->d_bin_sums), e->d_bin_mins, nb, P, c.rank, c.world,
                                                    c.half_words, c.seq);
        k_xchg_reduce<<<grid, 256, 0, e->stream>>>(e->d_bin_sums, e->d_bin_mins, nb, P, c.rank, c.world, c.half_words, c.seq, spin);
        RTS_CUDA(cudaGetLastError());
        e->launches += 2;
        // the time-out flag travels with the pulse's read-back
        RTS_CUDA(cudaMemcpyAsync(&e->h_rb->comm_timed_out, c.block + offsetof(XchgHeader, timed_out), sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        c.check_timeout = true;
    }
    e->bins_finalised = true;
    return agg_emit_bins_async(e);
}

// {exchanges, mean ns CTA 0 of k_xchg_reduce waited for the peers' flags, mean ns it spent in the kernel}: how much of an
// exchange is waiting for the slowest rank, how much the transfer.  Waits for the stream.
extern "C" int rts_comm_stats(rts_engine *e, double out[3])
{
    if (!e || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    Comm &c = e->comm;
    if (!c.block) return rts_fail(RTS_ERR_STATE, "rts_comm_create first");
    RTS_CUDA(cudaSetDevice(e->device));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    XchgHeader h;
    RTS_CUDA(cudaMemcpy(&h, c.block, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = (double)h.calls;
    out[1] = h.calls ? (double)h.wait_ns / (double)h.calls : 0.0;
    out[2] = h.calls ? (double)h.reduce_ns / (double)h.calls : 0.0;
    return RTS_OK;
}

extern "C" int rts_comm_destroy(rts_engine *e)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    Comm &c = e->comm;
    if (!c.block) return RTS_OK;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (uint32_t q = 0; q < RTS_MAX_WORLD; q++) {
        if (c.ipc_opened[q] && c.peer[q]) cudaIpcCloseMemHandle(c.peer[q]);
        c.peer[q] = nullptr; c.ipc_opened[q] = false;
    }
    cudaFree(c.block);
    c = Comm();
    return RTS_OK;
}
