// aggregate.cu — receiver-bin finalisation and the drop-in ray aggregator (sm_100a).
//
// Replaces the reference's aggregation.cu:
//   myKernel1 (aggregation.cu:32-74)   O(R^2 * D) all-pairs path match      -> O(R * D) hash grouping
//   myKernel2 (aggregation.cu:79-97)   mean voltage squared, mean delay/phase/Doppler
//   rs::kernel_wrapper (:103-184)      same signature, same array contract (api.cu exports it)
// and the unique-path step of ray_tracer.cpp:1289-1294 for the fused bins.
#include "engine.h"
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <algorithm>
#include <cstring>
#include <vector>
#include <math_constants.h>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace cg = cooperative_groups;

namespace {

inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// ---- record-mode defaults (ray_tracer.cu:227-240, ray_tracer.cpp:854-868) ----
__global__ void k_fill_records(rts_ray_record *res, unsigned long long n)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // 144 bytes = 9 x 16: zero everything, then the two non-zero defaults
    double2 *p = reinterpret_cast<double2 *>(res + i);
#pragma unroll
    for (int k = 0; k < 9; k++) p[k] = make_double2(0.0, 0.0);
    res[i].refrIndex[0] = 1; res[i].refrIndex[1] = 1;
    res[i].received = -1;
}
__global__ void k_fill_i32(int32_t *p, unsigned long long n, int32_t v)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_f64(double *p, unsigned long long n, double v)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- fused bins: finalise + compact ----
// One thread per dense bin.  Direct-ray rule (aggregation.cu:56): a direct ray sums over every
// ray of its receiver, so the direct bin (key 0 = all columns -1) reports the receiver totals.
__global__ void k_rx_totals(const double *__restrict__ sums, const unsigned long long *__restrict__ mins,
                            unsigned long long bins_per_rx, uint32_t n_rx, double *__restrict__ rx_sums,
                            unsigned long long *__restrict__ rx_mins)
{
    // grid.y = receiver; block-stride over its bins; fp64 block reduction then one atomic set per block
    const uint32_t rx = blockIdx.y;
    if (rx >= n_rx) return;
    double acc[5] = {0, 0, 0, 0, 0};
    unsigned long long mn = ~0ull;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < bins_per_rx;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long g = rx * bins_per_rx + b;
        const double n = sums[g * 5];
        if (n > 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) acc[k] += sums[g * 5 + k];
            mn = min(mn, mins[g]);
        }
    }
    cg::thread_block_tile<32> w = cg::tiled_partition<32>(cg::this_thread_block());
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k] = cg::reduce(w, acc[k], cg::plus<double>());
    mn = cg::reduce(w, mn, cg::less<unsigned long long>());
    if (w.thread_rank() == 0 && acc[0] > 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) atomicAdd(rx_sums + rx * 5 + k, acc[k]);
        atomicMin(rx_mins + rx, mn);
    }
}

__global__ void k_emit_bins(const double *__restrict__ sums, const unsigned long long *__restrict__ mins,
                            unsigned long long n_bins, unsigned long long bins_per_rx, uint32_t B, uint32_t D,
                            const double *__restrict__ rx_sums, const unsigned long long *__restrict__ rx_mins,
                            rts_bin *__restrict__ out, uint32_t *__restrict__ out_count, uint32_t cap)
{
    const unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_bins) return;
    if (!(sums[g * 5] > 0)) return;
    const uint32_t rx = (uint32_t)(g / bins_per_rx);
    unsigned long long key = g % bins_per_rx;
    rts_bin b;
    b.rx = (int32_t)rx;
    const bool direct = key == 0;
    for (uint32_t c = 0; c < RTS_MAX_DEPTH; c++) {
        b.path[c] = c < D ? (int32_t)(key % B) - 1 : -1;
        if (c < D) key /= B;
    }
    b.direct = direct ? 1 : 0;
    const double *s = direct ? rx_sums + rx * 5 : sums + g * 5;
    b.npath = s[0]; b.sum_sqrt_power = s[1]; b.sum_delay = s[2]; b.sum_phase = s[3]; b.sum_doppler = s[4];
    b.min_slot = direct ? rx_mins[rx] : mins[g];
    b.own_min_slot = mins[g];
    // myKernel2 (aggregation.cu:86-92)
    b.power = pow(b.sum_sqrt_power / b.npath, 2);
    b.delay = b.sum_delay / b.npath;
    b.phase = b.sum_phase / b.npath;
    b.doppler = b.sum_doppler / b.npath;
    const uint32_t at = atomicAdd(out_count, 1u);
    if (at < cap) out[at] = b;
}

// ---- sparse bins: the same three steps over the list of occupied slots ----
__global__ void k_hash_clear(unsigned long long *__restrict__ keys, double *__restrict__ sums, unsigned long long *__restrict__ mins,
                             const uint32_t *__restrict__ used, const unsigned *__restrict__ count)
{
    const unsigned n = *count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = used[i];
        keys[s] = ~0ull;
#pragma unroll
        for (int k = 0; k < 5; k++) sums[(size_t)s * 5 + k] = 0.0;
        mins[s] = 0x7f7f7f7f7f7f7f7full;
    }
}
__global__ void k_hash_rx_totals(const unsigned long long *__restrict__ keys, const double *__restrict__ sums,
                                 const unsigned long long *__restrict__ mins, const uint32_t *__restrict__ used,
                                 const unsigned *__restrict__ count, unsigned long long bins_per_rx, uint32_t n_rx,
                                 double *__restrict__ rx_sums, unsigned long long *__restrict__ rx_mins)
{
    const unsigned n = *count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t s = used[i];
        const uint32_t rx = (uint32_t)(keys[s] / bins_per_rx);
        if (rx >= n_rx || !(sums[(size_t)s * 5] > 0)) continue;
#pragma unroll
        for (int k = 0; k < 5; k++) atomicAdd(rx_sums + rx * 5 + k, sums[(size_t)s * 5 + k]);
        atomicMin(rx_mins + rx, mins[s]);
    }
}
__global__ void k_hash_emit(const unsigned long long *__restrict__ keys, const double *__restrict__ sums,
                            const unsigned long long *__restrict__ mins, const uint32_t *__restrict__ used,
                            const unsigned *__restrict__ count, unsigned long long bins_per_rx, uint32_t B, uint32_t D,
                            const double *__restrict__ rx_sums, const unsigned long long *__restrict__ rx_mins,
                            rts_bin *__restrict__ out, uint32_t *__restrict__ out_count, uint32_t cap)
{
    const unsigned n = *count;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t g = used[i];
        if (!(sums[(size_t)g * 5] > 0)) continue;
        const uint32_t rx = (uint32_t)(keys[g] / bins_per_rx);
        unsigned long long key = keys[g] % bins_per_rx;
        rts_bin b;
        b.rx = (int32_t)rx;
        const bool direct = key == 0;
        for (uint32_t c = 0; c < RTS_MAX_DEPTH; c++) {
            b.path[c] = c < D ? (int32_t)(key % B) - 1 : -1;
            if (c < D) key /= B;
        }
        b.direct = direct ? 1 : 0;
        const double *s = direct ? rx_sums + rx * 5 : sums + (size_t)g * 5;
        b.npath = s[0]; b.sum_sqrt_power = s[1]; b.sum_delay = s[2]; b.sum_phase = s[3]; b.sum_doppler = s[4];
        b.min_slot = direct ? rx_mins[rx] : mins[g];
        b.own_min_slot = mins[g];
        b.power = pow(b.sum_sqrt_power / b.npath, 2);
        b.delay = b.sum_delay / b.npath;
        b.phase = b.sum_phase / b.npath;
        b.doppler = b.sum_doppler / b.npath;
        const uint32_t at = atomicAdd(out_count, 1u);
        if (at < cap) out[at] = b;
    }
}

// occupied slots -> compact arrays (keys, five sums, representative slot), for an exchange between GPUs
__global__ void k_hash_compact(const unsigned long long *__restrict__ keys, const double *__restrict__ sums,
                               const unsigned long long *__restrict__ mins, const uint32_t *__restrict__ used, unsigned n,
                               unsigned long long *__restrict__ o_keys, double *__restrict__ o_sums, unsigned long long *__restrict__ o_mins)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = used[i];
    o_keys[i] = keys[s];
#pragma unroll
    for (int k = 0; k < 5; k++) o_sums[(size_t)i * 5 + k] = sums[(size_t)s * 5 + k];
    o_mins[i] = mins[s];
}
// compact arrays (distinct keys) -> the cleared table
__global__ void k_hash_load(const unsigned long long *__restrict__ i_keys, const double *__restrict__ i_sums,
                            const unsigned long long *__restrict__ i_mins, unsigned n, unsigned long long *__restrict__ keys,
                            double *__restrict__ sums, unsigned long long *__restrict__ mins, uint32_t *__restrict__ used,
                            unsigned *__restrict__ count, unsigned long long mask, unsigned long long *overflow)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long bin = i_keys[i];
    unsigned long long h = bin * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    unsigned long long at = h & mask, probes = 0;
    for (;; at = (at + 1ull) & mask) {
        const unsigned long long cur = atomicCAS(keys + at, ~0ull, bin);
        if (cur == ~0ull) { used[atomicAdd(count, 1u)] = (uint32_t)at; break; }
        if (cur == bin) break;
        if (++probes > mask) { atomicAdd(overflow, 1ull); return; }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) atomicAdd(sums + (size_t)at * 5 + k, i_sums[(size_t)i * 5 + k]);
    atomicMin(mins + at, i_mins[i]);
}

// ---- received rays, compacted (first half of the host loop ray_tracer.cpp:1190-1258) ----
struct IsReceived {
    const rts_ray_record *res;
    __device__ bool operator()(unsigned i) const { return res[i].received >= 0; }
};
// one thread per 16-byte piece of a record (9 per ray); rows and angles by the first piece's thread
__global__ void k_gather_received(const unsigned *__restrict__ idx, unsigned n, uint32_t D, const rts_ray_record *__restrict__ res,
                                  const int32_t *__restrict__ ti, const double *__restrict__ rcs, rts_ray_record *__restrict__ o_res,
                                  int32_t *__restrict__ o_ti, double *__restrict__ o_rcs, unsigned long long *__restrict__ o_slot)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned k = (unsigned)(t / 9), piece = (unsigned)(t % 9);
    if (k >= n) return;
    const unsigned src = idx[k];
    reinterpret_cast<double2 *>(o_res + k)[piece] = reinterpret_cast<const double2 *>(res + src)[piece];
    if (piece == 0) {
        o_slot[k] = src;
        for (uint32_t c = 0; c < D; c++) {
            o_ti[(size_t)k * D + c] = ti[(size_t)src * D + c];
            o_rcs[((size_t)k * D + c) * 2] = rcs[((size_t)src * D + c) * 2];
            o_rcs[((size_t)k * D + c) * 2 + 1] = rcs[((size_t)src * D + c) * 2 + 1];
        }
    }
}

// ---- the records of one shard, compacted: ray k of the shard (launch index begin + k*stride) has slot s at k + s*n ----
__global__ void k_gather_shard(unsigned long long n, uint32_t M, unsigned long long R3, unsigned long long begin, unsigned long long stride,
                               uint32_t D, uint32_t W, const rts_ray_record *__restrict__ res, const int32_t *__restrict__ ti,
                               const double *__restrict__ rcs, const int32_t *__restrict__ tp, rts_ray_record *__restrict__ o_res,
                               int32_t *__restrict__ o_ti, double *__restrict__ o_rcs, int32_t *__restrict__ o_tp)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long item = t / 9;
    const unsigned piece = (unsigned)(t % 9);
    if (item >= n * M) return;
    const unsigned long long k = item % n, s = item / n;
    const unsigned long long src = begin + k * stride + s * R3;
    if (o_res) reinterpret_cast<double2 *>(o_res + item)[piece] = reinterpret_cast<const double2 *>(res + src)[piece];
    if (piece == 0) {
        for (uint32_t c = 0; c < D; c++) {
            if (o_ti) o_ti[item * D + c] = ti[src * D + c];
            if (o_rcs) { o_rcs[(item * D + c) * 2] = rcs[(src * D + c) * 2]; o_rcs[(item * D + c) * 2 + 1] = rcs[(src * D + c) * 2 + 1]; }
        }
        if (o_tp) for (uint32_t c = 0; c < W; c++) o_tp[item * W + c] = tp[src * W + c];
    }
}

// ---- drop-in aggregator (rs::kernel_wrapper) ----
// Open-addressing table keyed by (receiver, path row); a slot stores the index of the ray that
// claimed it and later arrivals compare rows against that representative.
__device__ __forceinline__ unsigned long long row_hash(const rts_ray_record *res, const int32_t *rows, uint32_t i, uint32_t D)
{
    unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(uint32_t)res[i].received;
    for (uint32_t k = 0; k < D; k++) {
        h ^= (unsigned long long)(uint32_t)rows[(size_t)i * D + k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    h ^= h >> 31; h *= 0x7fb5d329728ea185ull; h ^= h >> 27;
    return h;
}
__device__ __forceinline__ bool same_group(const rts_ray_record *res, const int32_t *rows, uint32_t a, uint32_t b, uint32_t D)
{
    if (res[a].received != res[b].received) return false;
    for (uint32_t k = 0; k < D; k++)
        if (rows[(size_t)a * D + k] != rows[(size_t)b * D + k]) return false;
    return true;
}

__global__ void k_group(const rts_ray_record *__restrict__ res, const int32_t *__restrict__ rows, uint32_t R, uint32_t D,
                        uint32_t *table, uint32_t table_mask, uint32_t *__restrict__ slot_of)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    uint32_t s = (uint32_t)row_hash(res, rows, i, D) & table_mask;
    for (;;) {
        uint32_t cur = table[s];
        if (cur == 0xffffffffu) {
            const uint32_t prev = atomicCAS(table + s, 0xffffffffu, i);
            cur = prev == 0xffffffffu ? i : prev;
        }
        if (cur == i || same_group(res, rows, cur, i, D)) { slot_of[i] = s; return; }
        s = (s + 1) & table_mask;
    }
}

// per-ray terms of myKernel1 (aggregation.cu:59-69) summed per group and per receiver
__global__ void k_accumulate(const rts_ray_record *__restrict__ res, uint32_t R, const uint32_t *__restrict__ slot_of,
                             double cspeed, double carrier, double *grp_sums, uint32_t *grp_min, double *rx_sums,
                             uint32_t *rx_min, uint32_t n_rx_slots)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const double delay = (res[i].rayLength) / cspeed;
    const double phase = -fmod(delay * 2 * M_PI * carrier, 2 * M_PI);
    const double amp = sqrt(res[i].power);
    const double dop = res[i].doppler;
    const uint32_t s = slot_of[i];
    const uint32_t rx = (uint32_t)res[i].received;
    cg::coalesced_group g = cg::coalesced_threads();
    {
        auto part = cg::labeled_partition(g, s);
        const double a0 = cg::reduce(part, 1.0, cg::plus<double>());
        const double a1 = cg::reduce(part, amp, cg::plus<double>());
        const double a2 = cg::reduce(part, delay, cg::plus<double>());
        const double a3 = cg::reduce(part, phase, cg::plus<double>());
        const double a4 = cg::reduce(part, dop, cg::plus<double>());
        const uint32_t mn = cg::reduce(part, i, cg::less<uint32_t>());
        if (part.thread_rank() == 0) {
            double *b = grp_sums + (size_t)s * 5;
            atomicAdd(b, a0); atomicAdd(b + 1, a1); atomicAdd(b + 2, a2); atomicAdd(b + 3, a3); atomicAdd(b + 4, a4);
            atomicMin(grp_min + s, mn);
        }
    }
    if (rx < n_rx_slots) {
        // a group of the threads that are here: records with received < 0 (which the reference's kernel tolerates) skip
        // this branch, and a partition of the outer group would wait for them
        cg::coalesced_group g2 = cg::coalesced_threads();
        auto part = cg::labeled_partition(g2, rx);
        const double a0 = cg::reduce(part, 1.0, cg::plus<double>());
        const double a1 = cg::reduce(part, amp, cg::plus<double>());
        const double a2 = cg::reduce(part, delay, cg::plus<double>());
        const double a3 = cg::reduce(part, phase, cg::plus<double>());
        const double a4 = cg::reduce(part, dop, cg::plus<double>());
        const uint32_t mn = cg::reduce(part, i, cg::less<uint32_t>());
        if (part.thread_rank() == 0) {
            double *b = rx_sums + (size_t)rx * 5;
            atomicAdd(b, a0); atomicAdd(b + 1, a1); atomicAdd(b + 2, a2); atomicAdd(b + 3, a3); atomicAdd(b + 4, a4);
            atomicMin(rx_min + rx, mn);
        }
    }
}

// write-out in the reference's array contract + myKernel2
__global__ void k_scatter(rts_ray_record *res, uint32_t R, const uint32_t *__restrict__ slot_of,
                          const double *__restrict__ grp_sums, const uint32_t *__restrict__ grp_min,
                          const double *__restrict__ rx_sums, const uint32_t *__restrict__ rx_min, uint32_t n_rx_slots,
                          double *npath, double *power, double *doppler, double *delay, double *phase, int32_t *pathMatch)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const bool direct = (res[i].reflDepth == 0) && (res[i].refrDepth == 0);
    const uint32_t rx = (uint32_t)res[i].received;
    const double *s = (direct && rx < n_rx_slots) ? rx_sums + (size_t)rx * 5 : grp_sums + (size_t)slot_of[i] * 5;
    const uint32_t mn = (direct && rx < n_rx_slots) ? rx_min[rx] : grp_min[slot_of[i]];
    npath[i] += s[0];
    power[i] += s[1];
    delay[i] += s[2];
    phase[i] += s[3];
    doppler[i] += s[4];
    if ((int32_t)mn < pathMatch[i]) pathMatch[i] = (int32_t)mn;
    if (npath[i] > 0) {
        res[i].power = pow(power[i] / npath[i], 2);
        delay[i] /= npath[i];
        phase[i] /= npath[i];
        res[i].doppler = doppler[i] / npath[i];
    }
}

} // namespace

int agg_fill_records(rts_engine *e, uint64_t ray_total, uint32_t D, uint32_t W)
{
    const unsigned bs = 256;
    { k_fill_records<<<blocks_for(ray_total, bs), bs, 0, e->stream>>>(e->d_results, ray_total); e->launches++; }
    if (D) {
        { k_fill_i32<<<blocks_for(ray_total * D, bs), bs, 0, e->stream>>>(e->d_targ_intersect, ray_total * D, -1); e->launches++; }
        { k_fill_f64<<<blocks_for(ray_total * D * 2, bs), bs, 0, e->stream>>>(e->d_rcs_angle, ray_total * D * 2, -1000000.0); e->launches++; }
    }
    { k_fill_i32<<<blocks_for(ray_total * W, bs), bs, 0, e->stream>>>(e->d_tri_path, ray_total * W, -1); e->launches++; }
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

static void sort_bins(rts_bin *out, uint32_t got)
{
    std::sort(out, out + got, [](const rts_bin &a, const rts_bin &b) {
        if (a.rx != b.rx) return a.rx < b.rx;
        for (uint32_t c = 0; c < RTS_MAX_DEPTH; c++)
            if (a.path[c] != b.path[c]) return a.path[c] < b.path[c];
        return false;
    });
}

// The sparse bin table of a pulse: allocated on first use, afterwards cleared by walking the previous pulse's list of
// occupied slots — clearing, like emission, costs what is occupied, not the table size.
int agg_hash_prepare(rts_engine *e, uint64_t slots)
{
    cudaStream_t st = e->stream;
    if (e->hash_alloc < slots) {
        void **ptrs[] = {(void **)&e->d_hash_keys, (void **)&e->d_hash_used, (void **)&e->d_hash_count};
        for (void **p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
        e->hash_alloc = 0; e->hash_ready = false;
        RTS_CUDA(cudaMalloc(&e->d_hash_keys, sizeof(unsigned long long) * slots));
        RTS_CUDA(cudaMalloc(&e->d_hash_used, sizeof(uint32_t) * slots));
        RTS_CUDA(cudaMalloc(&e->d_hash_count, sizeof(unsigned)));
        e->hash_alloc = slots;
    }
    if (e->hash_ready) {
        k_hash_clear<<<64, 256, 0, st>>>(e->d_hash_keys, e->d_bin_sums, e->d_bin_mins, e->d_hash_used, e->d_hash_count);
        e->launches++;
    } else {
        RTS_CUDA(cudaMemsetAsync(e->d_hash_keys, 0xff, sizeof(unsigned long long) * slots, st));
        RTS_CUDA(cudaMemsetAsync(e->d_bin_sums, 0, sizeof(double) * 5 * slots, st));
        RTS_CUDA(cudaMemsetAsync(e->d_bin_mins, 0x7f, sizeof(unsigned long long) * slots, st));
    }
    RTS_CUDA(cudaMemsetAsync(e->d_hash_count, 0, sizeof(unsigned), st));
    e->hash_ready = true;
    return RTS_OK;
}

// Sparse bins, multi-GPU: this rank's occupied bins as compact device arrays (the count comes back to the host: the caller
// sizes its all-gather with it), and the way back — the table refilled from the merged arrays of all ranks.
int agg_hash_compact(rts_engine *e, void **keys, void **sums, void **mins, uint32_t *n)
{
    cudaStream_t st = e->stream;
    unsigned count = 0;
    RTS_CUDA(cudaMemcpyAsync(&count, e->d_hash_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    RTS_CUDA(cudaStreamSynchronize(st));
    if (e->compact_alloc < count) {
        void **ptrs[] = {(void **)&e->d_ckeys, (void **)&e->d_csums, (void **)&e->d_cmins};
        for (void **p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
        const size_t cap = std::max<size_t>(1024, (size_t)count * 2);
        RTS_CUDA(cudaMalloc(&e->d_ckeys, sizeof(unsigned long long) * cap));
        RTS_CUDA(cudaMalloc(&e->d_csums, sizeof(double) * 5 * cap));
        RTS_CUDA(cudaMalloc(&e->d_cmins, sizeof(unsigned long long) * cap));
        e->compact_alloc = cap;
    }
    if (count) {
        k_hash_compact<<<blocks_for(count, 256), 256, 0, st>>>(e->d_hash_keys, e->d_bin_sums, e->d_bin_mins, e->d_hash_used, count,
                                                              e->d_ckeys, e->d_csums, e->d_cmins);
        e->launches++;
        RTS_CUDA(cudaGetLastError());
    }
    if (keys) *keys = e->d_ckeys;
    if (sums) *sums = e->d_csums;
    if (mins) *mins = e->d_cmins;
    if (n) *n = count;
    return RTS_OK;
}

int agg_hash_load(rts_engine *e, const void *keys, const void *sums, const void *mins, uint32_t n)
{
    cudaStream_t st = e->stream;
    int rc = agg_hash_prepare(e, e->n_bins_dense);
    if (rc) return rc;
    if (n > e->n_bins_dense) return rts_fail(RTS_ERR_CAPACITY, "%u merged bins exceed the sparse table's %llu slots (option hash_log2)", n, (unsigned long long)e->n_bins_dense);
    if (n) {
        k_hash_load<<<blocks_for(n, 256), 256, 0, st>>>((const unsigned long long *)keys, (const double *)sums, (const unsigned long long *)mins, n,
                                                       e->d_hash_keys, e->d_bin_sums, e->d_bin_mins, e->d_hash_used, e->d_hash_count,
                                                       e->n_bins_dense - 1ull, reinterpret_cast<unsigned long long *>(e->d_counters) + 9);
        e->launches++;
        RTS_CUDA(cudaGetLastError());
    }
    return RTS_OK;
}

// Everything a pulse starts from, cleared by ONE launch instead of eight cudaMemsetAsync calls (each a few microseconds of
// stream time that nothing overlaps): the dense bin table (sums = 0, representative slots = 0x7f..7f, see api.cu), the
// counters, the per-wave segment counts, the first batch's queue / work counters, and what the emission at the end of
// the pulse accumulates into (receiver totals, emitted-bin count).
struct ClearArgs {
    unsigned long long *sums; unsigned long long n_sums;      // fp64 zeros as words
    unsigned long long *mins; unsigned long long n_mins;
    unsigned long long *zero[4]; unsigned n_zero[4];          // counters, wave_segs, counts, rx_sums: words to zero
    unsigned long long *rx_mins; unsigned n_rx_mins;           // ~0
    uint32_t *bins_out_count;
};
__global__ void k_pulse_clear(const ClearArgs A)
{
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = tid; i < A.n_sums; i += step) A.sums[i] = 0ull;
    for (unsigned long long i = tid; i < A.n_mins; i += step) A.mins[i] = 0x7f7f7f7f7f7f7f7full;
#pragma unroll
    for (int k = 0; k < 4; k++)
        for (unsigned long long i = tid; i < A.n_zero[k]; i += step) A.zero[k][i] = 0ull;
    for (unsigned long long i = tid; i < A.n_rx_mins; i += step) A.rx_mins[i] = ~0ull;
    if (tid == 0 && A.bins_out_count) *A.bins_out_count = 0u;
}

// dense: also clear the bin table (the sparse table is cleared through its list of occupied slots, agg_hash_prepare)
int agg_pulse_clear(rts_engine *e, bool dense_bins, uint64_t n_bins, uint32_t n_rx)
{
    if (!e->d_rx_sums) {
        RTS_CUDA(cudaMalloc(&e->d_rx_sums, sizeof(double) * 5 * RTS_MAX_RX));
        RTS_CUDA(cudaMalloc(&e->d_rx_mins, sizeof(unsigned long long) * RTS_MAX_RX));
    }
    if (!e->d_bins_out_count) RTS_CUDA(cudaMalloc(&e->d_bins_out_count, sizeof(uint32_t)));
    ClearArgs A;
    memset(&A, 0, sizeof(A));
    if (dense_bins) {
        A.sums = reinterpret_cast<unsigned long long *>(e->d_bin_sums); A.n_sums = n_bins * 5ull;
        A.mins = e->d_bin_mins; A.n_mins = n_bins;
    }
    A.zero[0] = reinterpret_cast<unsigned long long *>(e->d_counters); A.n_zero[0] = (unsigned)(sizeof(Counters) / 8);
    A.zero[1] = e->d_wave_segs; A.n_zero[1] = 32;
    A.zero[2] = e->d_counts; A.n_zero[2] = 96;
    A.zero[3] = reinterpret_cast<unsigned long long *>(e->d_rx_sums); A.n_zero[3] = 5u * n_rx;
    A.rx_mins = e->d_rx_mins; A.n_rx_mins = n_rx;
    A.bins_out_count = e->d_bins_out_count;
    const unsigned long long words = std::max<unsigned long long>(A.n_sums, 256ull);
    const unsigned grid = (unsigned)std::min<unsigned long long>((words + 255ull) / 256ull, (unsigned long long)e->num_sms * 8ull);
    k_pulse_clear<<<grid, 256, 0, e->stream>>>(A);
    RTS_CUDA(cudaGetLastError());
    e->launches++;
    e->emit_precleared = true;
    return RTS_OK;
}

// receiver totals (direct-ray rule) + myKernel2 + compaction of the non-empty bins into d_bins_out, enqueued
static int enqueue_emit(rts_engine *e, uint32_t cap)
{
    const uint64_t nb = e->n_bins_dense;
    const uint32_t n_rx = e->last_nrx, B = e->last_B, D = e->last_D;
    const uint64_t per_rx = e->bins_hashed ? e->bins_per_rx : nb / n_rx;
    if (!e->d_rx_sums) {
        RTS_CUDA(cudaMalloc(&e->d_rx_sums, sizeof(double) * 5 * RTS_MAX_RX));
        RTS_CUDA(cudaMalloc(&e->d_rx_mins, sizeof(unsigned long long) * RTS_MAX_RX));
    }
    const bool precleared = e->emit_precleared;      // the pulse's clear kernel did it (first emission of the pulse only)
    e->emit_precleared = false;
    if (!precleared) {
        RTS_CUDA(cudaMemsetAsync(e->d_rx_sums, 0, sizeof(double) * 5 * n_rx, e->stream));
        RTS_CUDA(cudaMemsetAsync(e->d_rx_mins, 0xff, sizeof(unsigned long long) * n_rx, e->stream));
    }
    if (e->bins_hashed) {
        k_hash_rx_totals<<<64, 256, 0, e->stream>>>(e->d_hash_keys, e->d_bin_sums, e->d_bin_mins, e->d_hash_used, e->d_hash_count, per_rx, n_rx,
                                                    e->d_rx_sums, e->d_rx_mins);
        e->launches++;
    } else {
        dim3 grid((unsigned)std::min<uint64_t>(64, (per_rx + 255) / 256), n_rx);
        k_rx_totals<<<grid, 256, 0, e->stream>>>(e->d_bin_sums, e->d_bin_mins, per_rx, n_rx, e->d_rx_sums, e->d_rx_mins);
        e->launches++;
    }
    const uint64_t want = std::max<uint64_t>(cap, 1);
    if (e->bins_out_alloc < want) {
        if (e->d_bins_out) cudaFree(e->d_bins_out);
        e->d_bins_out = nullptr;
        RTS_CUDA(cudaMalloc(&e->d_bins_out, sizeof(rts_bin) * want));
        e->bins_out_alloc = want;
    }
    if (!e->d_bins_out_count) RTS_CUDA(cudaMalloc(&e->d_bins_out_count, sizeof(uint32_t)));
    if (!precleared) RTS_CUDA(cudaMemsetAsync(e->d_bins_out_count, 0, sizeof(uint32_t), e->stream));
    if (e->bins_hashed) {
        k_hash_emit<<<64, 256, 0, e->stream>>>(e->d_hash_keys, e->d_bin_sums, e->d_bin_mins, e->d_hash_used, e->d_hash_count, per_rx, B, D,
                                               e->d_rx_sums, e->d_rx_mins, e->d_bins_out, e->d_bins_out_count, cap);
        e->launches++;
    } else {
        k_emit_bins<<<blocks_for(nb, 256), 256, 0, e->stream>>>(e->d_bin_sums, e->d_bin_mins, nb, per_rx, B, D, e->d_rx_sums, e->d_rx_mins,
                                                             e->d_bins_out, e->d_bins_out_count, cap);
        e->launches++;
    }
    RTS_CUDA(cudaGetLastError());
    return RTS_OK;
}

// Emit the bins right behind the pulse (or the caller's reduction) and bring them to pinned host memory with the
// rest of the read-back, so that rts_get_bins costs one wait instead of three.  Small tables only.
int agg_emit_bins_async(rts_engine *e)
{
    const bool again = e->bins_eager;   // this pulse's bins are being emitted a second time (rts_finalise_bins after a reduction): same block
    e->bins_eager = false;
    const uint64_t nb = e->n_bins_dense;
    if (!nb || !e->last_nrx || (!e->bins_hashed && nb > (1ull << 18))) return RTS_OK;
    if (!e->h_bins_buf[0]) {   // two pinned blocks, their counts and events (engine.h)
        bool ok = cudaMallocHost((void **)&e->h_bins_count, sizeof(uint32_t) * 2) == cudaSuccess;
        for (int k = 0; k < 2 && ok; k++)
            ok = cudaMallocHost((void **)&e->h_bins_buf[k], sizeof(rts_bin) * RTS_EAGER_BINS) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->bins_ev[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            for (int k = 0; k < 2; k++) { if (e->h_bins_buf[k]) cudaFreeHost(e->h_bins_buf[k]); e->h_bins_buf[k] = nullptr; }
            if (e->h_bins_count) cudaFreeHost(e->h_bins_count);
            e->h_bins_count = nullptr; e->h_bins = nullptr;
            cudaGetLastError();
            return RTS_OK;
        }
    }
    if (!again) e->bins_slot ^= 1;   // the other block: the previous pulse's may not have been read yet (rts_get_bins_previous)
    e->h_bins = e->h_bins_buf[e->bins_slot];
    int rc = enqueue_emit(e, RTS_EAGER_BINS);
    if (rc) return rc;
    RTS_CUDA(cudaMemcpyAsync(e->h_bins_count + e->bins_slot, e->d_bins_out_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaMemcpyAsync(e->h_bins, e->d_bins_out, sizeof(rts_bin) * RTS_EAGER_BINS, cudaMemcpyDeviceToHost, e->stream));
    cudaEventRecord(e->bins_ev[e->bins_slot], e->stream);
    e->bins_eager = true;
    return RTS_OK;
}

int agg_collect_bins(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n)
{
    const uint64_t nb = e->n_bins_dense;
    if (!nb || !e->last_nrx) { if (n) *n = 0; return RTS_OK; }
    if (e->bins_eager) {   // already in pinned memory (the caller has waited for the stream)
        RTS_CUDA(cudaStreamSynchronize(e->stream));
        const uint32_t count = e->h_bins_count[e->bins_slot];
        if (count <= RTS_EAGER_BINS) {
            const uint32_t got = std::min(count, cap);
            if (out && got) {
                std::vector<rts_bin> all(e->h_bins, e->h_bins + count);
                sort_bins(all.data(), count);
                memcpy(out, all.data(), sizeof(rts_bin) * got);
            }
            if (n) *n = count;
            e->stats.n_bins = count;
            return RTS_OK;
        }
    }
    int rc = enqueue_emit(e, cap);
    if (rc) return rc;
    uint32_t count = 0;
    RTS_CUDA(cudaMemcpyAsync(&count, e->d_bins_out_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    const uint32_t got = std::min(count, cap);
    if (out && got) {
        RTS_CUDA(cudaMemcpy(out, e->d_bins_out, sizeof(rts_bin) * got, cudaMemcpyDeviceToHost));
        sort_bins(out, got);
    }
    if (n) *n = count;
    e->stats.n_bins = count;
    return RTS_OK;
}

// The bins of the pulse before the last one, from the pinned block that pulse filled (engine.h: h_bins_buf).
int agg_collect_bins_previous(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n)
{
    if (!e->prev_bins_eager || !e->h_bins_buf[0])
        return rts_fail(RTS_ERR_STATE, "the pulse before the last one left no bins in pinned memory (it must have been traced with RTS_OUT_BINS, finalised, with a small bin table)");
    RTS_CUDA(cudaEventSynchronize(e->bins_ev[e->prev_bins_slot]));
    const uint32_t count = e->h_bins_count[e->prev_bins_slot];
    if (count > RTS_EAGER_BINS)
        return rts_fail(RTS_ERR_CAPACITY, "the pulse before the last one produced %u bins, more than the %u kept in pinned memory: read them with rts_get_bins before enqueuing the next pulse", count, RTS_EAGER_BINS);
    const uint32_t got = std::min(count, cap);
    if (out && got) {
        std::vector<rts_bin> all(e->h_bins_buf[e->prev_bins_slot], e->h_bins_buf[e->prev_bins_slot] + count);
        sort_bins(all.data(), count);
        memcpy(out, all.data(), sizeof(rts_bin) * got);
    }
    if (n) *n = count;
    return RTS_OK;
}

int agg_kernel_wrapper(rts_engine *e, rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t R, uint32_t D,
                       double cspeed, double carrier, double *npath, double *power, double *doppler, double *delay,
                       double *phase, int32_t *path_match)
{
    if (R == 0) return RTS_OK;
    cudaStream_t st = e->stream;
    uint32_t table_size = 64;
    while (table_size < 2ull * R) table_size <<= 1;
    // receivers are small non-negative ints; size the per-receiver table from the data
    int32_t max_rx = 0;
    for (uint32_t i = 0; i < R; i++) max_rx = std::max(max_rx, rx_results[i].received);
    const uint32_t n_rx_slots = (uint32_t)max_rx + 1;

    rts_ray_record *d_res = nullptr;
    int32_t *d_rows = nullptr, *d_pm = nullptr;
    uint32_t *d_table = nullptr, *d_slot = nullptr, *d_gmin = nullptr, *d_rmin = nullptr;
    double *d_gs = nullptr, *d_rs = nullptr, *d_acc = nullptr; // d_acc: npath,power,doppler,delay,phase [5*R]
    int rc = RTS_OK;
#define AGG_CUDA(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            rc = rts_fail(RTS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            goto done;                                                                                          \
        }                                                                                                       \
    } while (0)
    AGG_CUDA(cudaMalloc(&d_res, sizeof(rts_ray_record) * (size_t)R));
    AGG_CUDA(cudaMalloc(&d_rows, sizeof(int32_t) * std::max<size_t>(1, (size_t)R * D)));
    AGG_CUDA(cudaMalloc(&d_pm, sizeof(int32_t) * (size_t)R));
    AGG_CUDA(cudaMalloc(&d_table, sizeof(uint32_t) * (size_t)table_size));
    AGG_CUDA(cudaMalloc(&d_slot, sizeof(uint32_t) * (size_t)R));
    AGG_CUDA(cudaMalloc(&d_gmin, sizeof(uint32_t) * (size_t)table_size));
    AGG_CUDA(cudaMalloc(&d_rmin, sizeof(uint32_t) * (size_t)n_rx_slots));
    AGG_CUDA(cudaMalloc(&d_gs, sizeof(double) * 5 * (size_t)table_size));
    AGG_CUDA(cudaMalloc(&d_rs, sizeof(double) * 5 * (size_t)n_rx_slots));
    AGG_CUDA(cudaMalloc(&d_acc, sizeof(double) * 5 * (size_t)R));
    AGG_CUDA(cudaMemcpyAsync(d_res, rx_results, sizeof(rts_ray_record) * (size_t)R, cudaMemcpyHostToDevice, st));
    if (D) AGG_CUDA(cudaMemcpyAsync(d_rows, rx_intersects, sizeof(int32_t) * (size_t)R * D, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_pm, path_match, sizeof(int32_t) * (size_t)R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_acc + 0 * (size_t)R, npath, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_acc + 1 * (size_t)R, power, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_acc + 2 * (size_t)R, doppler, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_acc + 3 * (size_t)R, delay, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemcpyAsync(d_acc + 4 * (size_t)R, phase, sizeof(double) * R, cudaMemcpyHostToDevice, st));
    AGG_CUDA(cudaMemsetAsync(d_table, 0xff, sizeof(uint32_t) * (size_t)table_size, st));
    AGG_CUDA(cudaMemsetAsync(d_gmin, 0xff, sizeof(uint32_t) * (size_t)table_size, st));
    AGG_CUDA(cudaMemsetAsync(d_rmin, 0xff, sizeof(uint32_t) * (size_t)n_rx_slots, st));
    AGG_CUDA(cudaMemsetAsync(d_gs, 0, sizeof(double) * 5 * (size_t)table_size, st));
    AGG_CUDA(cudaMemsetAsync(d_rs, 0, sizeof(double) * 5 * (size_t)n_rx_slots, st));
    {
        const unsigned bs = 256, gb = blocks_for(R, bs);
        { k_group<<<gb, bs, 0, st>>>(d_res, d_rows, R, D, d_table, table_size - 1, d_slot); e->launches++; }
        { k_accumulate<<<gb, bs, 0, st>>>(d_res, R, d_slot, cspeed, carrier, d_gs, d_gmin, d_rs, d_rmin, n_rx_slots); e->launches++; }
        { k_scatter<<<gb, bs, 0, st>>>(d_res, R, d_slot, d_gs, d_gmin, d_rs, d_rmin, n_rx_slots, d_acc + 0 * (size_t)R,
                                     d_acc + 1 * (size_t)R, d_acc + 2 * (size_t)R, d_acc + 3 * (size_t)R,
                                     d_acc + 4 * (size_t)R, d_pm); e->launches++; }
        AGG_CUDA(cudaGetLastError());
    }
    AGG_CUDA(cudaMemcpyAsync(rx_results, d_res, sizeof(rts_ray_record) * (size_t)R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(npath, d_acc + 0 * (size_t)R, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(power, d_acc + 1 * (size_t)R, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(doppler, d_acc + 2 * (size_t)R, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(delay, d_acc + 3 * (size_t)R, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(phase, d_acc + 4 * (size_t)R, sizeof(double) * R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaMemcpyAsync(path_match, d_pm, sizeof(int32_t) * (size_t)R, cudaMemcpyDeviceToHost, st));
    AGG_CUDA(cudaStreamSynchronize(st));
done:
#undef AGG_CUDA
    cudaFree(d_res); cudaFree(d_rows); cudaFree(d_pm); cudaFree(d_table); cudaFree(d_slot); cudaFree(d_gmin);
    cudaFree(d_rmin); cudaFree(d_gs); cudaFree(d_rs); cudaFree(d_acc);
    return rc;
}

// The last RTS_OUT_RECORDS pulse's records for the rays of its shard only (a strided sample or one rank's share of a
// large launch), gathered on the device: the caller never touches the launch-sized arrays.
int agg_get_records_shard(rts_engine *e, uint64_t *n_shard, rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle,
                          int32_t *tri_path)
{
    const uint64_t n = e->last_n_primary;
    const uint32_t M = e->last_sizes.slots, D = e->last_sizes.depth_total, W = e->last_sizes.tri_cols;
    if (n_shard) *n_shard = n;
    if (!n || (!results && !targ_intersect && !rcs_angle && !tri_path)) return RTS_OK;
    cudaStream_t st = e->stream;
    rts_ray_record *o_res = nullptr;
    int32_t *o_ti = nullptr, *o_tp = nullptr;
    double *o_rcs = nullptr;
    int rc = RTS_OK;
    const size_t items = (size_t)n * M;
#define SH_CUDA(call)                                                                                           \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            rc = rts_fail(RTS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            goto done;                                                                                          \
        }                                                                                                       \
    } while (0)
    if (results) SH_CUDA(cudaMalloc(&o_res, sizeof(rts_ray_record) * items));
    if (targ_intersect && D) SH_CUDA(cudaMalloc(&o_ti, sizeof(int32_t) * items * D));
    if (rcs_angle && D) SH_CUDA(cudaMalloc(&o_rcs, sizeof(double) * 2 * items * D));
    if (tri_path) SH_CUDA(cudaMalloc(&o_tp, sizeof(int32_t) * items * W));
    { k_gather_shard<<<blocks_for((uint64_t)items * 9, 256), 256, 0, st>>>(n, M, e->last_sizes.rays, e->last_begin, e->last_stride, D, W, e->d_results,
                                                                         e->d_targ_intersect, e->d_rcs_angle, e->d_tri_path, o_res, o_ti, o_rcs, o_tp); e->launches++; }
    SH_CUDA(cudaGetLastError());
    if (o_res) SH_CUDA(cudaMemcpyAsync(results, o_res, sizeof(rts_ray_record) * items, cudaMemcpyDeviceToHost, st));
    if (o_ti) SH_CUDA(cudaMemcpyAsync(targ_intersect, o_ti, sizeof(int32_t) * items * D, cudaMemcpyDeviceToHost, st));
    if (o_rcs) SH_CUDA(cudaMemcpyAsync(rcs_angle, o_rcs, sizeof(double) * 2 * items * D, cudaMemcpyDeviceToHost, st));
    if (o_tp) SH_CUDA(cudaMemcpyAsync(tri_path, o_tp, sizeof(int32_t) * items * W, cudaMemcpyDeviceToHost, st));
    SH_CUDA(cudaStreamSynchronize(st));
done:
#undef SH_CUDA
    cudaFree(o_res); cudaFree(o_ti); cudaFree(o_rcs); cudaFree(o_tp);
    return rc;
}

// Received rays of the last RTS_OUT_RECORDS pulse in result-slot order (ray_tracer.cpp:1190-1221 without the
// callbacks): stream compaction of the slot indices, then a gather of record, path row and RCS-angle row.
int agg_get_received(rts_engine *e, uint64_t cap, uint64_t *n, uint64_t *slots, rts_ray_record *results, int32_t *targ_intersect,
                     double *rcs_angle)
{
    const uint64_t total = e->last_sizes.ray_total;
    const uint32_t D = e->last_sizes.depth_total;
    if (total >= (1ull << 31)) return rts_fail(RTS_ERR_CAPACITY, "%llu result slots exceed 2^31", (unsigned long long)total);
    cudaStream_t st = e->stream;
    unsigned *d_idx = nullptr, *d_num = nullptr;
    void *d_tmp = nullptr;
    rts_ray_record *o_res = nullptr;
    int32_t *o_ti = nullptr;
    double *o_rcs = nullptr;
    unsigned long long *o_slot = nullptr;
    int rc = RTS_OK;
    unsigned count = 0;
    size_t tmp_bytes = 0;
    const IsReceived pred{e->d_results};
    thrust::counting_iterator<unsigned> first(0u);
#define REC_CUDA(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (call);                                                                                \
        if (_e != cudaSuccess) {                                                                                \
            rc = rts_fail(RTS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            goto done;                                                                                          \
        }                                                                                                       \
    } while (0)
    REC_CUDA(cudaMalloc(&d_idx, sizeof(unsigned) * std::max<uint64_t>(1, total)));
    REC_CUDA(cudaMalloc(&d_num, sizeof(unsigned)));
    REC_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, first, d_idx, d_num, (int)total, pred, st));
    REC_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(16, tmp_bytes)));
    REC_CUDA(cub::DeviceSelect::If(d_tmp, tmp_bytes, first, d_idx, d_num, (int)total, pred, st));
    e->launches += 2;
    REC_CUDA(cudaMemcpyAsync(&count, d_num, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    REC_CUDA(cudaStreamSynchronize(st));
    if (n) *n = count;
    if (count && cap) {
        const unsigned take = (unsigned)std::min<uint64_t>(count, cap);
        const size_t Dn = std::max<uint32_t>(1, D);
        REC_CUDA(cudaMalloc(&o_res, sizeof(rts_ray_record) * take));
        REC_CUDA(cudaMalloc(&o_ti, sizeof(int32_t) * take * Dn));
        REC_CUDA(cudaMalloc(&o_rcs, sizeof(double) * 2 * take * Dn));
        REC_CUDA(cudaMalloc(&o_slot, sizeof(unsigned long long) * take));
        { k_gather_received<<<blocks_for((uint64_t)take * 9, 256), 256, 0, st>>>(d_idx, take, D, e->d_results, e->d_targ_intersect,
                                                                             e->d_rcs_angle, o_res, o_ti, o_rcs, o_slot); e->launches++; }
        REC_CUDA(cudaGetLastError());
        if (results) REC_CUDA(cudaMemcpyAsync(results, o_res, sizeof(rts_ray_record) * take, cudaMemcpyDeviceToHost, st));
        if (targ_intersect && D) REC_CUDA(cudaMemcpyAsync(targ_intersect, o_ti, sizeof(int32_t) * take * D, cudaMemcpyDeviceToHost, st));
        if (rcs_angle && D) REC_CUDA(cudaMemcpyAsync(rcs_angle, o_rcs, sizeof(double) * 2 * take * D, cudaMemcpyDeviceToHost, st));
        if (slots) REC_CUDA(cudaMemcpyAsync(slots, o_slot, sizeof(unsigned long long) * take, cudaMemcpyDeviceToHost, st));
        REC_CUDA(cudaStreamSynchronize(st));
    }
done:
#undef REC_CUDA
    cudaFree(d_idx); cudaFree(d_num); cudaFree(d_tmp); cudaFree(o_res); cudaFree(o_ti); cudaFree(o_rcs); cudaFree(o_slot);
    return rc;
}
