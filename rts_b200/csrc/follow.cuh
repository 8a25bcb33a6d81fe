// follow.cuh — the projected primary wave's shading pass with the first reflection traced in place (included inside
// trace.cu's anonymous namespace, after split.cuh).
//
// k_primary_shade writes one 88-byte ray state per lit pixel (1.2 GB for the benchmark's 13.5 M first reflections) that
// the second wave reads straight back: the two kernels were 0.50 + 1.37 ms of a 2.78 ms pulse, the first bound by those
// DRAM writes, the second by traversal latency with DRAM idle.  Here the warp that shades a pixel's primary hit keeps the
// reflected ray in registers and walks the BVH with it at once (closest_hit's rtTrace, normal_shader.cu:332, as the
// reference has it: recursion, not a queue).  Only what survives that first reflection is queued — 0.2 % of the rays on
// the terrain benchmark — so the next launch is a thin, chained wave.  Refracted children go to the queue as always.
//
// Work distribution: the cost per pixel now varies (sky pixels end at once, lit pixels traverse), so warps claim 32
// pixels at a time from the wave's work counter like k_wave instead of being dealt a static stride.
// Results are bit-identical to the two-kernel form: same shade / traverse / miss code, same order per ray.
template <bool RECORDS>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_WAVE_MIN_BLOCKS) k_primary_follow(const __grid_constant__ WaveParams P)
{
    if (!raster_on(P)) return;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_in = (unsigned)P.n_primary;
    Local L = {0, 0, 0, 0, 0};
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)n_in);
        atomicAdd(&P.counters->segments, (unsigned long long)n_in);
    }
    unsigned *work = reinterpret_cast<unsigned *>(P.work_counter);
    unsigned followed = 0;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_in) break;
        const unsigned rel = base + lane;
        if (rel >= n_in) continue;
        const unsigned long long hit = __ldcs(P.hits + rel);
        Ray r;
        r.ox = P.origin[0]; r.oy = P.origin[1]; r.oz = P.origin[2];
        r.dx = __ldcs(P.dirs[0] + rel); r.dy = __ldcs(P.dirs[1] + rel); r.dz = __ldcs(P.dirs[2] + rel);
        r.meta = m_make(0, 0, 0, false, true, 0);
        r.len = 0; r.pw = 0; r.dop = 0; r.fx = 0; r.fy = 0; r.fz = 0; r.n0 = 1; r.n1 = 1;
        r.key = 0; r.ray = (uint32_t)(P.ray_begin + (P.batch_base + rel) * P.ray_stride);
        bool follow = false;
        if (hit != ~0ull) {
            HitRec h;
            h.pos = P.hits_resolved ? (int)((uint32_t)hit & 0x7fffffffu) : (int)__ldg(P.leaf_of_tri + (uint32_t)hit);
            h.t = __uint_as_float((unsigned)(hit >> 32)); h.id = 0;   // shade() takes the id from the record
            L.a += C_HIT;
            follow = shade<RECORDS>(P, r, h, L, true);
        } else {
            const int received = miss<RECORDS>(P, r, L);
            if (received >= 0) {
                L.a += C_CAPTURED;
                if (P.flags & RTS_OUT_BINS) accumulate_bin(P, r, received);
            }
        }
        if (!follow) continue;
        // the first reflection, in place (one step: what it hits is queued for the next launch)
        followed++;
        HitRec h;
        unsigned nn = 0, nt = 0;
        traverse<false>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, L.overflow);
        if (h.pos >= 0) {
            L.a += C_HIT;
            shade<RECORDS>(P, r, h, L, false);
        } else {
            const int received = miss<RECORDS>(P, r, L);
            if (received >= 0) {
                L.a += C_CAPTURED;
                if (P.flags & RTS_OUT_BINS) accumulate_bin(P, r, received);
            }
        }
    }
    {   // segments traced in place belong to this launch
        const unsigned x = __reduce_add_sync(0xffffffffu, followed);
        if (lane == 0 && x) {
            atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)x);
            atomicAdd(&P.counters->segments, (unsigned long long)x);
        }
    }
    unsigned long long *c = reinterpret_cast<unsigned long long *>(P.counters);
    const unsigned f[7] = {(unsigned)(L.a & 0x1fffff), (unsigned)((L.a >> 21) & 0x1fffff), (unsigned)(L.a >> 42),
                           (unsigned)(L.b & 0x1fffff), (unsigned)((L.b >> 21) & 0x1fffff), (unsigned)(L.b >> 42), L.overflow};
    const int slot[7] = {1, 2, 3, 4, 5, 6, 9};
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const unsigned x = __reduce_add_sync(0xffffffffu, f[k]);
        if (lane == 0 && x) atomicAdd(c + slot[k], (unsigned long long)x);
    }
}
