// coherent.cuh — kept first-reflection hits (included inside trace.cu's anonymous namespace, after raster.cuh).
//
// With a staring transmitter over a scene that is mostly static, almost every ray of the second wave (the first
// reflection of a primary ray) is, bit for bit, the ray it was in the previous pulse: same primary ray, same static
// triangle, same shading inputs.  What such a ray hits next can only change through the targets that move.  Since
//     closest hit over all triangles = min over { closest hit among the static ones, closest hit among the moving ones }
// (both by the (fp32 t, triangle id) order of traverse()), the static part is kept per launch-grid pixel from the
// pulse it was first computed in, and a pulse only has to look at the moving targets:
//   k_target_boxes / k_mover_nodes   bounding box of every moving target for this pulse, packed two per BvhNode
//   k_wave1_fill (first pulse only)  closest STATIC hit of every flagged first reflection -> w1_static[pixel]
//   k_wave1_kept                     flagged rays: conservative slab test against the moving targets' boxes (same packed
//                                    arithmetic and error bound as the node test of traverse()); none touched -> the kept
//                                    hit is the answer, shade / miss as usual; any touched -> the queue index goes
//                                    on a to-do list with all the unflagged rays, and the ordinary k_wave launched
//                                    right behind traces just that list in full
// A ray carries the flag (M_COH in its queue word) only when its primary hit equals the kept static primary hit
// (raster.cuh: k_raster_resolve), i.e. when it provably is the same ray.  The kept data is tied to the launch geometry,
// the scene, the moving set, the tree (leaf positions) and the shading switches that steer the ray (interpolation,
// depths); anything else changing means a refill.  Results are bit-identical to tracing every ray.

__device__ __forceinline__ unsigned c_f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float c_ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

struct MoverIds { uint32_t n; uint32_t id[32]; };

// w1_static entry of a pixel whose first reflection was not flagged in the fill pulse (memset pattern 0xfe; a kept
// hit has positive fp32 t bits in its upper word, a kept miss is ~0)
constexpr unsigned long long W1_UNKNOWN = 0xfefefefefefefefeull;

// box[k] = empty for the moving targets (ordered-uint encoding: lo = +inf, hi = -inf)
__global__ void k_target_box_init(unsigned *__restrict__ box, MoverIds M)
{
    const unsigned i = threadIdx.x;
    if (i < M.n * 6) box[6 * M.id[i / 6] + i % 6] = (i % 6) < 3 ? 0xff800000u : 0x007fffffu;
}
// union of the leaf boxes (the reference's bound program, already rounded outward) of every moving triangle, per target
__global__ void k_target_boxes(const uint32_t *__restrict__ list, uint32_t count, const float *__restrict__ tri_box,
                               const uint32_t *__restrict__ tri_target, unsigned *__restrict__ box)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t id = list[i];
    const float *b = tri_box + 6 * (size_t)id;
    unsigned *t = box + 6 * (size_t)tri_target[id];
#pragma unroll
    for (int a = 0; a < 3; a++) { atomicMin(t + a, c_f2ord(b[a])); atomicMax(t + 3 + a, c_f2ord(b[3 + a])); }
}
// two target boxes per node, centre / half-extent rounded up exactly as k_pack does; a missing partner never hits
__global__ void k_mover_nodes(const unsigned *__restrict__ box, MoverIds M, BvhNode *__restrict__ nodes)
{
    const unsigned p = threadIdx.x;
    if (p >= (M.n + 1) / 2) return;
    float cc[2][3], hh[2][3];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const unsigned k = 2 * p + c;
#pragma unroll
        for (int a = 0; a < 3; a++) { cc[c][a] = 0.f; hh[c][a] = -1.0e30f; }
        if (k < M.n) {
            const unsigned *b = box + 6 * (size_t)M.id[k];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float lo = c_ord2f(b[a]), hi = c_ord2f(b[3 + a]);
                if (lo <= hi) {
                    const float c0 = 0.5f * lo + 0.5f * hi;
                    const double d = fmax((double)hi - (double)c0, (double)c0 - (double)lo);
                    cc[c][a] = c0;
                    hh[c][a] = nextafterf(__double2float_ru(d), CUDART_INF_F);
                }
            }
        }
    }
    BvhNode nd;
    nd.c0x = cc[0][0]; nd.c0y = cc[0][1]; nd.h0x = hh[0][0]; nd.h0y = hh[0][1];
    nd.c1x = cc[1][0]; nd.c1y = cc[1][1]; nd.h1x = hh[1][0]; nd.h1y = hh[1][1];
    nd.c0z = cc[0][2]; nd.c1z = cc[1][2]; nd.h0z = hh[0][2]; nd.h1z = hh[1][2];
    nd.ref0 = 0; nd.ref1 = 0; nd.pad[0] = nd.pad[1] = 0;
    nodes[p] = nd;
}

// shard-local pixel of a launch index
__device__ __forceinline__ unsigned w1_pixel(const WaveParams &P, uint32_t ray)
{
    const unsigned long long off = (unsigned long long)ray - P.ray_begin;
    return (unsigned)((P.ray_stride > 1 ? off / P.ray_stride : off) - P.batch_base);
}

// Does the ray come near any moving target?  The node test of traverse(), same constants and error bound, over the
// packed target boxes; no distance limit.
__device__ __forceinline__ bool near_movers(const WaveParams &P, const d3 &o, const d3 &dir)
{
    if (P.n_mover_nodes == 0) return false;
    u64 inv_xy, noi_xy, ainv_xy, e_xy, inv_zz, noi_zz, ainv_zz, e_zz;
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
        inv_xy = pk2(inv[0], inv[1]); noi_xy = pk2(noi[0], noi[1]); ainv_xy = pk2(ainv[0], ainv[1]); e_xy = pk2(E[0], E[1]);
        inv_zz = pk2(inv[2], inv[2]); noi_zz = pk2(noi[2], noi[2]); ainv_zz = pk2(ainv[2], ainv[2]); e_zz = pk2(E[2], E[2]);
    }
    bool any = false;
    for (uint32_t k = 0; k < P.n_mover_nodes; k++) {
        const ulonglong2 *np = reinterpret_cast<const ulonglong2 *>(P.mover_nodes + k);
        const ulonglong2 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
        const u64 T0 = fma2(q0.x, inv_xy, noi_xy), H0 = fma2(q0.y, ainv_xy, e_xy);
        const u64 T1 = fma2(q1.x, inv_xy, noi_xy), H1 = fma2(q1.y, ainv_xy, e_xy);
        const u64 Tz = fma2(q2.x, inv_zz, noi_zz), Hz = fma2(q2.y, ainv_zz, e_zz);
        float n0x, n0y, f0x, f0y, n1x, n1y, f1x, f1y, nz0, nz1, fz0, fz1;
        upk2(sub2(T0, H0), n0x, n0y); upk2(add2(T0, H0), f0x, f0y);
        upk2(sub2(T1, H1), n1x, n1y); upk2(add2(T1, H1), f1x, f1y);
        upk2(sub2(Tz, Hz), nz0, nz1); upk2(add2(Tz, Hz), fz0, fz1);
        const float tn0 = fmaxf(fmaxf(n0x, n0y), nz0), tf0 = fminf(fminf(f0x, f0y), fz0);
        const float tn1 = fmaxf(fmaxf(n1x, n1y), nz1), tf1 = fminf(fminf(f1x, f1y), fz1);
        any |= (fmaxf(tn0, 0.f) <= tf0) | (fmaxf(tn1, 0.f) <= tf1);
    }
    return any;
}

// First pulse: the closest hit among the triangles that never move, for every flagged ray of the queue.
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_WAVE_MIN_BLOCKS) k_wave1_fill(const __grid_constant__ WaveParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_front = (unsigned)min(*P.in_count, P.out_capacity);
    const unsigned n_in = n_front + (P.in_back ? (unsigned)min(*P.in_back, P.out_capacity - n_front) : 0u);
    unsigned *work = reinterpret_cast<unsigned *>(P.fill_counter);
    unsigned ovf = 0;
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(work, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_in) break;
        if (base + lane >= n_in) continue;
        const unsigned idx = queue_slot(P, base + lane, n_front);
        Ray r;
        load_ray_geom(P.in, idx, r);
        if (!(r.meta & M_COH)) continue;
        HitRec h;
        unsigned nn = 0, nt = 0;
        if (RTS_QNODES) traverse_q<false, true>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, ovf);
        else traverse<false, true>(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz), SCENE_EPS, h, nn, nt, ovf);
        const uint32_t ray = __ldcs(P.in.ray + idx);
        P.w1_static[w1_pixel(P, ray)] = h.pos >= 0 ? (((unsigned long long)__float_as_uint(h.t) << 32) | (unsigned long long)(uint32_t)h.pos) : ~0ull;
    }
    if (ovf) atomicAdd(&P.counters->overflow, (unsigned long long)ovf);
}

// Every pulse: the flagged rays, served from the kept hits unless a moving target is near.
template <bool RECORDS, bool TABLES = false>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_SHADE_MIN_BLOCKS) k_wave1_kept(const __grid_constant__ WaveParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_front = (unsigned)min(*P.in_count, P.out_capacity);
    const unsigned n_in = n_front + (P.in_back ? (unsigned)min(*P.in_back, P.out_capacity - n_front) : 0u);
    Local L = {0, 0, 0, 0, 0};
    unsigned served = 0;
    bins_smem_init(P);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned entry = blockIdx.x * blockDim.x + threadIdx.x; entry < n_in; entry += stride) {
        const unsigned idx = queue_slot(P, entry, n_front);
        Ray r;
        r.meta = __ldcs(P.in.meta + idx);
        bool mine = (r.meta & M_COH) != 0;
        unsigned long long kept = ~0ull;
        if (mine) {
            load_ray_geom(P.in, idx, r);
            mine = !near_movers(P, mk3(r.ox, r.oy, r.oz), mk3(r.dx, r.dy, r.dz));
            if (mine) {
                kept = P.w1_static[w1_pixel(P, __ldcs(P.in.ray + idx))];
                mine = kept != W1_UNKNOWN;         // the pixel was behind a moving target when the hits were kept
            }
        }
        if (!mine) {                               // for the ordinary wave kernel launched behind this one
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned long long at = 0;
            if (g.thread_rank() == 0) at = atomicAdd(P.todo_count, (unsigned long long)g.size());
            at = g.shfl(at, 0);
            P.todo_list[at + g.thread_rank()] = idx;
            continue;
        }
        load_ray_rest(P.in, idx, r, RECORDS || P.keep_first != 0, P.rMax != 0);
        r.meta &= ~M_COH;
        served++;
        if (kept != ~0ull) {
            HitRec h;
            h.pos = (int)(uint32_t)kept; h.t = __uint_as_float((unsigned)(kept >> 32)); h.id = 0;
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
    unsigned long long *c = reinterpret_cast<unsigned long long *>(P.counters);
    const unsigned f[7] = {(unsigned)(L.a & 0x1fffff), (unsigned)((L.a >> 21) & 0x1fffff), (unsigned)(L.a >> 42),
                           (unsigned)(L.b & 0x1fffff), (unsigned)((L.b >> 21) & 0x1fffff), (unsigned)(L.b >> 42), L.overflow};
    const int slot[7] = {1, 2, 3, 4, 5, 6, 9};
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const unsigned x = __reduce_add_sync(0xffffffffu, f[k]);
        if (lane == 0 && x) atomicAdd(c + slot[k], (unsigned long long)x);
    }
    {
        const unsigned x = __reduce_add_sync(0xffffffffu, served);
        if (lane == 0 && x) atomicAdd(&P.counters->kept, (unsigned long long)x);
    }
}
