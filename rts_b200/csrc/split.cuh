// split.cuh — the later bounce waves as two kernels (included inside trace.cu's anonymous namespace, after coherent.cuh).
//
// The fused k_wave keeps a ray's whole shading state in reach of the traversal loop: 80 registers, 6 CTAs of 128
// threads per SM, and a traversal that is latency-bound at that occupancy (profiles/r01_k_wave_ncu_full.json: 37 %
// warps active, 52 % issue slots, 24 of 32 lanes).  Here the closest-hit query (OptiX traversal + the intersect
// program, triangle_mesh.cu:121-200) is a kernel of its own:
//
//   k_traverse     persistent warps; a lane that has finished its ray takes the next one at once (the warp refills as
//                  soon as RTS_FETCH_THRESH lanes are idle) instead of waiting for the slowest of 32 rays; only origin
//                  and direction are read (48 of the 88 queued bytes).  Result: one 64-bit word per queue slot —
//                  (fp32 t bits << 32 | leaf position) for a hit, TRAV_MISS when nothing is hit but the ray crosses some
//                  receiver's sphere (the discriminant test of miss(), ray_tracer.cu:288-296, same arithmetic), TRAV_DEAD
//                  when nothing is hit and no sphere is crossed: such a ray changes no output outside records mode.
//   k_shade_wave   streams the hit words in queue order and runs closest_hit / miss / bin accumulation (the code of the
//                  fused kernel) for the survivors only.
//
// Which of the two forms serves a wave is decided on the device from the wave's size: waves below split_below rays
// are launch-latency-bound and stay with the fused kernel (which follows reflections in place).  Results are
// bit-identical: same traversal arithmetic, same closest-hit rule, same shading code.

#ifndef RTS_TRAV_MIN_BLOCKS
#define RTS_TRAV_MIN_BLOCKS 8
#endif
#ifndef RTS_FETCH_THRESH
#define RTS_FETCH_THRESH 32u           // idle lanes that make a warp fetch new rays (32: when all its rays are done — refilling
                                       // lane by lane mixes rays at different tree depths in one warp, and the node fetches,
                                       // which bound this kernel, stop coalescing: 1.44 ms at 32 against 1.74 ms at 8)
#endif
#ifndef RTS_TRAV_CHUNK
#define RTS_TRAV_CHUNK 64u             // rays a warp reserves from the global work counter at a time
#endif

constexpr unsigned long long TRAV_MISS = ~0ull;
constexpr unsigned long long TRAV_DEAD = ~0ull - 1ull;

// discriminant of the ray / receiver-sphere quadratic exactly as miss() forms it (ray_tracer.cu:288-296)
__device__ __forceinline__ double rx_discriminant(const d3 &po, const d3 &dir, const RxDev &rx)
{
    const double A = (dir.x) * (dir.x) + (dir.y) * (dir.y) + (dir.z) * (dir.z);
    const double B = 2 * (((po.x - rx.cx) * dir.x) + ((po.y - rx.cy) * dir.y) + ((po.z - rx.cz) * dir.z));
    const double C = po.x * po.x + po.y * po.y + po.z * po.z + (rx.cx * rx.cx) + (rx.cy * rx.cy) + (rx.cz * rx.cz) -
                     2 * ((rx.cx * po.x) + (rx.cy * po.y) + (rx.cz * po.z)) - rx.radius * rx.radius;
    return B * B - 4 * A * C;
}

template <bool COUNT>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_TRAV_MIN_BLOCKS) k_traverse(const __grid_constant__ WaveParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_front = (unsigned)min(*P.in_count, P.out_capacity);
    const unsigned n_all = n_front + (P.in_back ? (unsigned)min(*P.in_back, P.out_capacity - n_front) : 0u);
    const unsigned n_in = P.todo_list ? (unsigned)*P.todo_count : n_all;
    if (n_in < P.split_below || P.n_tris == 0) return;        // thin wave: the fused kernel behind this one takes it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)n_all);
        atomicAdd(&P.counters->segments, (unsigned long long)n_all);
    }
    unsigned *work = reinterpret_cast<unsigned *>(P.work_counter);
    constexpr int SENT = 0x7fffffff;
    constexpr unsigned FULL = 0xffffffffu;
    const double tmin_d = (double)SCENE_EPS, tmax_d = (double)RT_DEFAULT_MAX_F;
    // Register budget: the node loop needs the ray's slab constants (QRay: six floats, six PRMT selectors), the triangle
    // test needs the fp64 ray and ~50 registers of fp64 temporaries — never both.  So the constants are parked in shared
    // memory when a ray is fetched and re-read after every leaf visit (volatile: the old values are dead across the leaf
    // code), and the fp64 origin / direction are re-read from the queue (L1/L2-resident: this warp fetched them) when a
    // leaf is reached.
    __shared__ uint4 s_const[3][RTS_WAVE_BLOCK];
    int stack[RTS_STACK_DEPTH];
    int sp = 0, cur = SENT;
    bool have = false;
    unsigned idx = 0;
    QRay Q;
#pragma unroll
    for (int k = 0; k < 3; k++) { Q.a[k] = 0; Q.b[k] = 0; Q.sn[k] = 0; Q.sf[k] = 0; }
    float best_t = RT_DEFAULT_MAX_F, best_pad = CUDART_INF_F;
    int best_pos = -1;
    uint32_t best_id = 0xffffffffu;
    unsigned chunk_next = 0, chunk_end = 0;   // warp-uniform: the warp's reservation of queue entries
    bool exhausted = false;
    unsigned n_nodes = 0, n_tris = 0;

    for (;;) {
        const unsigned need = __ballot_sync(FULL, !have);
        if (need) {
            const unsigned n_need = __popc(need);
            if (n_need >= RTS_FETCH_THRESH && !(exhausted && chunk_next >= chunk_end)) {
                if (chunk_next >= chunk_end) {
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(work, RTS_TRAV_CHUNK);
                    base = __shfl_sync(FULL, base, 0);
                    chunk_next = min(base, n_in);
                    chunk_end = min(base + RTS_TRAV_CHUNK, n_in);
                    if (base + RTS_TRAV_CHUNK >= n_in) exhausted = true;
                }
                const unsigned avail = chunk_end - chunk_next;
                const unsigned rank = __popc(need & ((1u << lane) - 1u));
                if (!have && rank < avail) {
                    const unsigned entry = chunk_next + rank;
                    idx = P.todo_list ? P.todo_list[entry] : queue_slot(P, entry, n_front);
                    q_setup(P, mk3(__ldg(P.in.f[F_OX] + idx), __ldg(P.in.f[F_OY] + idx), __ldg(P.in.f[F_OZ] + idx)),
                            mk3(__ldg(P.in.f[F_DX] + idx), __ldg(P.in.f[F_DY] + idx), __ldg(P.in.f[F_DZ] + idx)), Q);
                    {
                        float a0, a1, a2, b0, b1, b2, dup;
                        upk2(Q.a[0], a0, dup); upk2(Q.a[1], a1, dup); upk2(Q.a[2], a2, dup);
                        upk2(Q.b[0], b0, dup); upk2(Q.b[1], b1, dup); upk2(Q.b[2], b2, dup);
                        s_const[0][threadIdx.x] = make_uint4(__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(b0));
                        s_const[1][threadIdx.x] = make_uint4(__float_as_uint(b1), __float_as_uint(b2), Q.sn[0], Q.sn[1]);
                        s_const[2][threadIdx.x] = make_uint4(Q.sn[2], Q.sf[0], Q.sf[1], Q.sf[2]);
                    }
                    best_t = RT_DEFAULT_MAX_F; best_pad = CUDART_INF_F; best_pos = -1; best_id = 0xffffffffu;
                    sp = 0; cur = P.root_ref;
                    have = true;
                }
                chunk_next += min(n_need, avail);
            }
            if (exhausted && chunk_next >= chunk_end && !__any_sync(FULL, have)) break;
        }
        // descend until this lane stands on a leaf or has nothing left (the "while-while" loop of traverse())
        while ((unsigned)cur < (unsigned)SENT) {
            if (COUNT) n_nodes++;
            bool h0, h1;
            float tn0, tn1;
            int2 refs;
            q_visit(P.qnodes, cur, Q, best_pad, h0, h1, tn0, tn1, refs);
            if (h0 & h1) {
                const bool swap = tn1 < tn0;
                if (sp < RTS_STACK_DEPTH) stack[sp++] = swap ? refs.x : refs.y;
                else atomicAdd(&P.counters->overflow, 1ull);    // never seen; not worth a register
                cur = swap ? refs.y : refs.x;
            } else if (h0) cur = refs.x;
            else if (h1) cur = refs.y;
            else cur = sp ? stack[--sp] : SENT;
        }
        __syncwarp();
        if (have) {
            unsigned il;   // idx, laundered: otherwise the six queue addresses formed at fetch time stay live across the node loop (12 registers)
            asm volatile("mov.u32 %0, %1;" : "=r"(il) : "r"(idx));
            const d3 o = mk3(__ldg(P.in.f[F_OX] + il), __ldg(P.in.f[F_OY] + il), __ldg(P.in.f[F_OZ] + il));
            const d3 dir = mk3(__ldg(P.in.f[F_DX] + il), __ldg(P.in.f[F_DY] + il), __ldg(P.in.f[F_DZ] + il));
            if (cur < 0) {
                const int code = ~cur;
                const int first = code >> 3, cnt = (code & 7) + 1;
                for (int k = 0; k < cnt; k++) {
                    if (COUNT) n_tris++;
                    const Tri T = load_tri(P.trirec, first + k);
                    double t;
                    if (tri_accept(T, o, dir, tmin_d, tmax_d, t)) {
                        const float tf = (float)t;
                        if (tf > SCENE_EPS && (tf < best_t || (tf == best_t && T.id < best_id))) {
                            best_t = tf; best_pos = first + k; best_id = T.id;
                            best_pad = tf * 1.000001f;
                        }
                    }
                }
                cur = sp ? stack[--sp] : SENT;
            }
            if (cur == SENT) {                         // this ray's closest hit is known
                unsigned long long word;
                if (best_pos >= 0) word = ((unsigned long long)__float_as_uint(best_t) << 32) | (unsigned long long)(uint32_t)best_pos;
                else {
                    bool keep = P.split_keep_all != 0;
                    for (unsigned j = 0; j < P.n_rx && !keep; j++) keep = rx_discriminant(o, dir, P.rx[j]) > 0.f;
                    word = keep ? TRAV_MISS : TRAV_DEAD;
                }
                __stcs(P.trav_hits + idx, word);
                have = false;
            }
            {
                // back to the node loop: the slab constants come back from shared memory (also on the path of a finished
                // ray, so that on no path the old values have to survive the triangle code above)
                const unsigned sa = (unsigned)__cvta_generic_to_shared(&s_const[0][threadIdx.x]);
                uint4 c0, c1, c2;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c0.x), "=r"(c0.y), "=r"(c0.z), "=r"(c0.w) : "r"(sa));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c1.x), "=r"(c1.y), "=r"(c1.z), "=r"(c1.w) : "r"(sa + (unsigned)sizeof(uint4) * RTS_WAVE_BLOCK));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c2.x), "=r"(c2.y), "=r"(c2.z), "=r"(c2.w) : "r"(sa + 2u * (unsigned)sizeof(uint4) * RTS_WAVE_BLOCK));
                Q.a[0] = pk2(__uint_as_float(c0.x), __uint_as_float(c0.x)); Q.a[1] = pk2(__uint_as_float(c0.y), __uint_as_float(c0.y));
                Q.a[2] = pk2(__uint_as_float(c0.z), __uint_as_float(c0.z)); Q.b[0] = pk2(__uint_as_float(c0.w), __uint_as_float(c0.w));
                Q.b[1] = pk2(__uint_as_float(c1.x), __uint_as_float(c1.x)); Q.b[2] = pk2(__uint_as_float(c1.y), __uint_as_float(c1.y));
                Q.sn[0] = c1.z; Q.sn[1] = c1.w; Q.sn[2] = c2.x; Q.sf[0] = c2.y; Q.sf[1] = c2.z; Q.sf[2] = c2.w;
            }
        }
    }
    if (COUNT) {
        unsigned long long *c = reinterpret_cast<unsigned long long *>(P.counters);
        const unsigned a = __reduce_add_sync(FULL, n_nodes), b = __reduce_add_sync(FULL, n_tris);
        if (lane == 0) { atomicAdd(c + 7, (unsigned long long)a); atomicAdd(c + 8, (unsigned long long)b); }
    }
}

// The shading half: queue order, survivors only.
template <bool RECORDS>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_SHADE_MIN_BLOCKS) k_shade_wave(const __grid_constant__ WaveParams P)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_front = (unsigned)min(*P.in_count, P.out_capacity);
    const unsigned n_all = n_front + (P.in_back ? (unsigned)min(*P.in_back, P.out_capacity - n_front) : 0u);
    const unsigned n_in = P.todo_list ? (unsigned)*P.todo_count : n_all;
    if (n_in < P.split_below || P.n_tris == 0) return;
    Local L = {0, 0, 0, 0, 0};
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned entry = blockIdx.x * blockDim.x + threadIdx.x; entry < n_in; entry += stride) {
        const unsigned idx = P.todo_list ? P.todo_list[entry] : queue_slot(P, entry, n_front);
        const unsigned long long word = __ldcs(P.trav_hits + idx);
        if (word == TRAV_DEAD) continue;
        Ray r;
        load_ray_geom(P.in, idx, r);
        r.meta &= ~M_COH;
        load_ray_rest(P.in, idx, r, RECORDS || P.keep_first != 0, P.rMax != 0);
        if (word != TRAV_MISS) {
            HitRec h;
            h.pos = (int)(uint32_t)word; h.t = __uint_as_float((unsigned)(word >> 32)); h.id = 0;
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
