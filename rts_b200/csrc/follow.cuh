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
//
// Straggler rays.  A first reflection that skims the terrain for kilometres visits thousands of nodes where the average
// ray visits 25: alone it takes ~130 us of dependent fetches.  In a warp it holds 31 finished lanes; claimed near the end
// of the launch it holds the whole GPU (measured: +70 us on 1.67 ms for rank 1's share of a 2-GPU launch, same instruction
// and byte counts as rank 0's — pure tail).  So traverse() gives such a ray up after RTS_FOLLOW_BUDGET leaf visits / pops,
// and the warp then walks the tree for it TOGETHER (coop_traverse): one stack shared by the 32 lanes in a per-warp
// scratch area, every lane pops its own node, tests both children (or the leaf's triangles, fp64, the same code as
// traverse()), the lanes agree on the closest hit so far, and push what they overlap.  Same boxes, same triangle test,
// same (t, id) order of the closest hit — only the visiting order differs, which the result does not depend on.
struct SlabRay { u64 inv_xy, noi_xy, ainv_xy, e_xy, inv_zz, noi_zz, ainv_zz, e_zz; };

// the per-ray constants of the conservative fp32 slab test (see traverse(): same arithmetic, same bound)
__device__ __forceinline__ void slab_setup(const WaveParams &P, const d3 &o, const d3 &dir, SlabRay &S)
{
    const float oo[3] = {(float)o.x, (float)o.y, (float)o.z}, dd[3] = {(float)dir.x, (float)dir.y, (float)dir.z};
    float inv[3], noi[3], ainv[3], E[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (!(fabsf(dd[a]) >= 1e-20f)) {
            inv[a] = 0.f; noi[a] = 0.f; ainv[a] = 0.f; E[a] = CUDART_INF_F;
        } else {
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd[a]));
            inv[a] = r;
            noi[a] = -(oo[a] * r);
            ainv[a] = fabsf(r);
            E[a] = 9.5367431640625e-07f * ((P.scene_abs[a] + fabsf(oo[a])) * ainv[a]) + 1e-30f;
        }
    }
    S.inv_xy = pk2(inv[0], inv[1]); S.noi_xy = pk2(noi[0], noi[1]); S.ainv_xy = pk2(ainv[0], ainv[1]); S.e_xy = pk2(E[0], E[1]);
    S.inv_zz = pk2(inv[2], inv[2]); S.noi_zz = pk2(noi[2], noi[2]); S.ainv_zz = pk2(ainv[2], ainv[2]); S.e_zz = pk2(E[2], E[2]);
}

// Closest hit of ONE ray (the same o, dir in every lane), found by the 32 lanes of a converged warp together.
// stk: this warp's RTS_COOP_STACK ints of scratch.  The result is the same in every lane.
__device__ __noinline__ void coop_traverse(const WaveParams &P, int *__restrict__ stk, const d3 o, const d3 dir, float tmin_f, HitRec &best,
                                           unsigned &stack_ovf)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SENT = 0x7fffffff;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    best.pos = -1; best.t = RT_DEFAULT_MAX_F; best.id = 0xffffffffu;
    SlabRay S;
    slab_setup(P, o, dir, S);
    const double tmin_d = (double)tmin_f, tmax_d = (double)RT_DEFAULT_MAX_F;
    float best_pad = CUDART_INF_F;
    unsigned sp = 1;
    if (lane == 0) stk[0] = P.root_ref;
    __syncwarp();
    while (sp) {
        const unsigned n = min(sp, 32u);
        const int cur = lane < n ? stk[sp - 1u - lane] : SENT;
        sp -= n;
        __syncwarp();
        int r0 = SENT, r1 = SENT;                                  // what this lane pushes
        float ct = RT_DEFAULT_MAX_F; int cpos = -1; uint32_t cid = 0xffffffffu;   // this lane's candidate hit
        if ((unsigned)cur < (unsigned)SENT) {
            const ulonglong2 *np = reinterpret_cast<const ulonglong2 *>(P.nodes + cur);
            const ulonglong2 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
            const int2 refs = __ldg(reinterpret_cast<const int2 *>(np + 3));
            const u64 T0 = fma2(q0.x, S.inv_xy, S.noi_xy), H0 = fma2(q0.y, S.ainv_xy, S.e_xy);
            const u64 T1 = fma2(q1.x, S.inv_xy, S.noi_xy), H1 = fma2(q1.y, S.ainv_xy, S.e_xy);
            const u64 Tz = fma2(q2.x, S.inv_zz, S.noi_zz), Hz = fma2(q2.y, S.ainv_zz, S.e_zz);
            float n0x, n0y, f0x, f0y, n1x, n1y, f1x, f1y, nz0, nz1, fz0, fz1;
            upk2(sub2(T0, H0), n0x, n0y); upk2(add2(T0, H0), f0x, f0y);
            upk2(sub2(T1, H1), n1x, n1y); upk2(add2(T1, H1), f1x, f1y);
            upk2(sub2(Tz, Hz), nz0, nz1); upk2(add2(Tz, Hz), fz0, fz1);
            const float tn0 = fmaxf(fmaxf(n0x, n0y), nz0), tf0 = fminf(fminf(f0x, f0y), fz0);
            const float tn1 = fmaxf(fmaxf(n1x, n1y), nz1), tf1 = fminf(fminf(f1x, f1y), fz1);
            if (fmaxf(tn0, 0.f) <= fminf(tf0, best_pad)) r0 = refs.x;
            if (fmaxf(tn1, 0.f) <= fminf(tf1, best_pad)) r1 = refs.y;
        } else if (cur < 0) {
            const int code = ~cur;
            const int first = code >> 3, cnt = (code & 7) + 1;
            for (int k = 0; k < cnt; k++) {
                const Tri T = load_tri(P.trirec, first + k);
                double t;
                if (tri_accept(T, o, dir, tmin_d, tmax_d, t)) {
                    const float tf = (float)t;
                    if (tf > tmin_f && (tf < ct || (tf == ct && T.id < cid))) { ct = tf; cpos = first + k; cid = T.id; }
                }
            }
        }
        // the closest candidate of this round in the (t, id) order of traverse(): positive fp32 bit patterns order like the values
        const unsigned tb = __float_as_uint(ct);
        const unsigned tb_min = __reduce_min_sync(FULL, tb);
        if (tb_min != __float_as_uint(RT_DEFAULT_MAX_F)) {
            const unsigned id_min = __reduce_min_sync(FULL, tb == tb_min ? cid : 0xffffffffu);
            const float t_min = __uint_as_float(tb_min);
            if (t_min < best.t || (t_min == best.t && id_min < best.id)) {
                const unsigned owner = __ffs(__ballot_sync(FULL, tb == tb_min && cid == id_min)) - 1u;
                best.t = t_min; best.id = id_min; best.pos = __shfl_sync(FULL, cpos, owner);
                best_pad = t_min * 1.000001f;
            }
        }
        const unsigned b0 = __ballot_sync(FULL, r0 != SENT), b1 = __ballot_sync(FULL, r1 != SENT);
        const unsigned total = __popc(b0) + __popc(b1);
        if (sp + total > RTS_COOP_STACK) { stack_ovf++; break; }    // never seen; reported as RTS_ERR_CAPACITY like every dropped state
        if (r0 != SENT) stk[sp + __popc(b0 & lt)] = r0;
        if (r1 != SENT) stk[sp + __popc(b0) + __popc(b1 & lt)] = r1;
        sp += total;
        __syncwarp();
    }
}

// Per-warp scratch (global memory, L2-resident): the shared stack of coop_traverse, then one parked ray state per lane.
constexpr size_t COOP_WARP_BYTES = sizeof(int) * RTS_COOP_STACK + sizeof(Ray) * 32;

// The rays of this warp that traverse() gave up on, parked in `slots` by their lanes: walked by the whole warp one after
// the other, then shaded by their own lane.  Called with the warp converged.  Out of line: it runs for a few rays per
// launch and must not cost the loop around it any registers.
template <bool RECORDS, bool TABLES>
__device__ __noinline__ void finish_stragglers(const WaveParams &P, int *stk, const Ray *slots, bool pending, Local &L, unsigned &followed)
{
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    for (unsigned todo = __ballot_sync(FULL, pending); todo; todo &= todo - 1u) {
        const int src = __ffs(todo) - 1;
        Ray r = slots[src];                       // the same address in every lane: one broadcast load
        HitRec h;
        coop_traverse(P, stk, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, L.overflow);
        if ((int)lane == src) {
            followed++;
            if (h.pos >= 0) {
                L.a += C_HIT;
                shade<RECORDS, TABLES>(P, r, h, L, false);
            } else {
                const int received = miss<RECORDS>(P, r, L);
                if (received >= 0) {
                    L.a += C_CAPTURED;
                    if (P.flags & RTS_OUT_BINS) accumulate_bin<TABLES>(P, r, received);
                }
            }
        }
        __syncwarp();
    }
}

template <bool RECORDS, bool TABLES = false>
#ifndef RTS_FOLLOW_MIN_BLOCKS
#define RTS_FOLLOW_MIN_BLOCKS 7      // 72 registers (400 B of spills, outside the walk): 6 / 7 / 8 CTAs per SM: 1.631 / 1.601 / 1.615 ms
#endif
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_FOLLOW_MIN_BLOCKS) k_primary_follow(const __grid_constant__ WaveParams P)
{
    if (!raster_on(P)) return;
    constexpr unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_in = (unsigned)P.n_primary;
    Local L = {0, 0, 0, 0, 0};
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)n_in);
        atomicAdd(&P.counters->segments, (unsigned long long)n_in);
    }
    unsigned *work = reinterpret_cast<unsigned *>(P.work_counter);
    unsigned followed = 0;
    bins_smem_init(P);
    bool pending = false;            // this lane parked a straggler ray in its slot during the last round
    for (;;) {
        // the warp is converged here: stragglers of the last round first
        if (__any_sync(FULL, pending)) {
            char *scratch = reinterpret_cast<char *>(P.coop_stacks) + COOP_WARP_BYTES * ((size_t)blockIdx.x * (RTS_WAVE_BLOCK / 32) + (threadIdx.x >> 5));
            finish_stragglers<RECORDS, TABLES>(P, reinterpret_cast<int *>(scratch), reinterpret_cast<const Ray *>(scratch + sizeof(int) * RTS_COOP_STACK), pending, L, followed);
            pending = false;
        }
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(FULL, base, 0);
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
            follow = shade<RECORDS, TABLES>(P, r, h, L, true);
        } else {
            const int received = miss<RECORDS>(P, r, L);
            if (received >= 0) {
                L.a += C_CAPTURED;
                if (P.flags & RTS_OUT_BINS) accumulate_bin<TABLES>(P, r, received);
            }
        }
        if (!follow) continue;
        // the first reflection, in place (one step: what it hits is queued for the next launch)
        HitRec h;
        unsigned nn = 0, nt = 0;
        if (!traverse<false, false, true>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, L.overflow)) {
            // a straggler: park the state, the warp finishes it together at the top of the loop
            char *scratch = reinterpret_cast<char *>(P.coop_stacks) + COOP_WARP_BYTES * ((size_t)blockIdx.x * (RTS_WAVE_BLOCK / 32) + (threadIdx.x >> 5));
            reinterpret_cast<Ray *>(scratch + sizeof(int) * RTS_COOP_STACK)[lane] = r;
            pending = true;
            continue;
        }
        followed++;
        if (h.pos >= 0) {
            L.a += C_HIT;
            shade<RECORDS, TABLES>(P, r, h, L, false);
        } else {
            const int received = miss<RECORDS>(P, r, L);
            if (received >= 0) {
                L.a += C_CAPTURED;
                if (P.flags & RTS_OUT_BINS) accumulate_bin<TABLES>(P, r, received);
            }
        }
    }
    bins_smem_flush(P);
    {   // segments traced in place belong to this launch
        const unsigned x = __reduce_add_sync(FULL, followed);
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
        const unsigned x = __reduce_add_sync(FULL, f[k]);
        if (lane == 0 && x) atomicAdd(c + slot[k], (unsigned long long)x);
    }
}
