// raster.cuh — primary visibility by projection (included inside trace.cu's anonymous namespace).
//
// Every primary ray leaves the transmitter origin, and before its two rotations the reference's ray direction is
// linear in the launch index (ray_tracer.cu:167-169): d0 = (bs.x, bs.y + sy*iy, bs.z + sz*iz) — a pinhole camera with
// image plane x = bs.x in the beam frame.  So instead of walking the BVH once per primary ray, each triangle is
// projected into the launch grid, and only the rays inside its (padded) footprint run the reference's fp64 test
// (triangle_mesh.cu:121-137) against it — with the very same ray direction the wave kernel would use (k_primary_dirs
// evaluates primary_direction once per ray) and the same operations in the same order, so t, beta and gamma are
// bit-identical.  The closest hit is an atomicMin over (fp32 t bits << 32 | triangle id): smallest fp32 t, ties to the
// lowest global triangle id — the rule of traverse().  The footprint only has to be conservative: projection is done
// in fp64 (errors ~1e-9 pixel); the pixel bounds are padded by 1e-3 pixel, the fp32 edge functions by 0.05 pixel plus
// their evaluation error.
//
// STATUS: the default primary wave for launches whose rays form one image (nx == 1; RTS_NO_RASTER=1 switches back to
// BVH traversal, which also serves cubic launches).  Bit-identical results on every parity test.  On the 1M-triangle
// benchmark the primary wave takes 1.11 ms against 1.60 ms by traversal (direction pass 0.18 ms, footprints 0.03 +
// 0.37 ms for 30 M candidates = 1.8 per ray, resolve + shading pass 0.5 ms).
//
// Work distribution: one thread per triangle walks small footprints (k_raster_small); large ones are cut into row
// chunks taken by warps (k_raster_big).  A device-side guard turns the whole path off for launches where the summed
// footprint exceeds RTS_RASTER_LIMIT candidates per ray (huge overlapping triangles); the BVH primary wave, launched
// right behind with the same control block, then does the work instead.  No host synchronisation either way.

#ifndef RTS_SHADE_MIN_BLOCKS
#define RTS_SHADE_MIN_BLOCKS 6
#endif
#ifndef RTS_RASTER_SMALL
#define RTS_RASTER_SMALL 160u          // footprints up to this many candidates are walked by their own thread (the slack of the edge functions assumes <= 160)
#endif
#define RTS_RASTER_CHUNK 2048u         // candidates per row chunk of a large footprint, at most (WaveParams::raster_chunk: fewer in small launches)
#define RTS_RASTER_LIMIT 16ull         // candidates per primary ray beyond which the BVH primary wave is used

__device__ __forceinline__ bool raster_on(const WaveParams &P)
{
    const unsigned long long area = P.raster_ctl->area + (P.raster_static ? P.raster_static->area : 0ull);
    return area <= RTS_RASTER_LIMIT * P.n_primary;
}
// The triangles a footprint pass covers: all leaf positions, or the triangles of the moving targets (by id).
__device__ __forceinline__ unsigned raster_count(const WaveParams &P) { return P.raster_list ? P.raster_list_count : P.n_tris; }
__device__ __forceinline__ unsigned raster_pos(const WaveParams &P, unsigned i) { return P.raster_list ? P.leaf_of_tri[P.raster_list[i]] : i; }

// shard-local index (relative to the batch) of launch-grid pixel (iy, iz), nx == 1.  When the shard's columns form a
// lattice (stride divides the row length: always for stride 1) the caller walks that lattice and passes the column
// counter k = (iy - c0) / stride, and the index is z * (ny / stride) + k - g0 without any division.
__device__ __forceinline__ bool pixel_local(const WaveParams &P, unsigned iy, unsigned iz, int k, unsigned &rel)
{
    if (P.lat_w) {
        const long long g = (long long)iz * P.lat_w + k - (long long)P.lat_g0 - (long long)P.batch_base;
        if (g < 0 || g >= (long long)P.n_primary) return false;
        rel = (unsigned)g;
        return true;
    }
    const unsigned long long rayIndex = (unsigned long long)iz * P.ny + iy;
    if (rayIndex < P.ray_begin) return false;
    unsigned long long off = rayIndex - P.ray_begin;
    if (P.ray_stride > 1) {
        if (off % P.ray_stride) return false;
        off /= P.ray_stride;
    }
    if (off < P.batch_base || off >= P.batch_base + P.n_primary) return false;
    rel = (unsigned)(off - P.batch_base);
    return true;
}

// ray directions of the batch, exactly as the wave kernel generates them, and the empty hit buffer
// 128 threads in at most 40 registers: small blocks that fit beside whatever else is resident (the pass runs on engine.h's side_dirs).
#define RTS_DIRS_BLOCK 128
__global__ void __launch_bounds__(RTS_DIRS_BLOCK, 12) k_primary_dirs(const __grid_constant__ WaveParams P)
{
    for (unsigned long long rel = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; rel < P.n_primary;
         rel += (unsigned long long)gridDim.x * blockDim.x) {
        // launch indices are < 2^32 (api.cu checks the grid): 32-bit division
        const uint32_t rayIndex = (uint32_t)(P.ray_begin + (P.batch_base + rel) * P.ray_stride);
        const uint32_t iz = rayIndex / P.ny, iy = rayIndex - iz * P.ny;
        const d3 d = primary_direction(P, 0, iy, iz);
        P.dirs[0][rel] = d.x; P.dirs[1][rel] = d.y; P.dirs[2][rel] = d.z;
        P.hits[rel] = ~0ull;
    }
}

struct TriFoot {
    d3 e0, e1, n, pmo;          // hoisted terms of the reference test: p1-p0, p0-p2, e1 x e0, p0-o
    uint32_t id;
    int y0, y1, z0, z1;         // inclusive pixel bounds, already clamped to the grid; y0 > y1: no coverage
    int ystep;                  // this shard's columns: y0 is on the shard's lattice, step = stride (or 1)
    int k0;                     // lattice column counter of y0: y = c0 + k * stride
    float ea[3], eb[3], ec[3];  // padded 2D edge functions relative to (y0, z0): inside iff all >= 0
    bool use2d;
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// Footprint of the triangle at leaf position `pos`.  Returns false when no ray of the grid can hit it.
__device__ __forceinline__ bool tri_footprint(const WaveParams &P, unsigned pos, TriFoot &F)
{
    const Tri T = load_tri(P.trirec, pos);
    if (P.raster_skip && P.raster_skip[T.target]) return false;   // static pass: the moving targets come separately
    const d3 o = mk3(P.origin[0], P.origin[1], P.origin[2]);
    F.e0 = T.p1 - T.p0;
    F.e1 = T.p0 - T.p2;
    F.n = cross3(F.e1, F.e0);
    F.pmo = T.p0 - o;
    F.id = T.id;
    // beam-frame coordinates q = A^T (p - o); pixel u = (q.y * bs.x / q.x - bs.y) / sy, w likewise
    const d3 v[3] = {T.p0 - o, T.p1 - o, T.p2 - o};
    d3 q[3];
    double scale = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        q[k].x = P.AT[0] * v[k].x + P.AT[1] * v[k].y + P.AT[2] * v[k].z;
        q[k].y = P.AT[3] * v[k].x + P.AT[4] * v[k].y + P.AT[5] * v[k].z;
        q[k].z = P.AT[6] * v[k].x + P.AT[7] * v[k].y + P.AT[8] * v[k].z;
        scale = fmax(scale, fabs(q[k].x) + fabs(q[k].y) + fabs(q[k].z));
    }
    const double eps = 1e-9 * scale + 1e-300;      // near plane: hits are at t >= SCENE_EPS in front of the origin
    const bool front[3] = {q[0].x > eps, q[1].x > eps, q[2].x > eps};
    const int n_front = (int)front[0] + (int)front[1] + (int)front[2];
    if (n_front == 0) return false;
    const double bx = P.beamStart[0];
    const double isy = P.ny > 1 ? 1.0 / P.slope[1] : 0.0, isz = P.nz > 1 ? 1.0 / P.slope[2] : 0.0;
    double umin = 1e300, umax = -1e300, wmin = 1e300, wmax = -1e300;
    double pu[3] = {0, 0, 0}, pw[3] = {0, 0, 0};
    auto project = [&](const d3 &c, double &u, double &w) {
        const double lam = bx / c.x;
        u = (c.y * lam - P.beamStart[1]) * isy;
        w = (c.z * lam - P.beamStart[2]) * isz;
        umin = fmin(umin, u); umax = fmax(umax, u); wmin = fmin(wmin, w); wmax = fmax(wmax, w);
    };
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (front[k]) project(q[k], pu[k], pw[k]);
    if (n_front < 3) {
        // clip against the near plane: the crossing points project far out, the grid clamp below takes care of them
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int a = k, b = (k + 1) % 3;
            if (front[a] != front[b]) {
                const double s = (eps - q[a].x) / (q[b].x - q[a].x);
                d3 c = mk3(eps, q[a].y + s * (q[b].y - q[a].y), q[a].z + s * (q[b].z - q[a].z));
                double u, w;
                project(c, u, w);
            }
        }
    }
    if (!(umin <= umax) || !(wmin <= wmax)) return false;   // NaN guard
    const double big = 2.0e9;
    // a ray passes exactly through its pixel centre (integer u, w); the projection is good to ~1e-9 pixel, so the
    // integer pixels inside [min - pad, max + pad] are all that can be hit
    const double pad = 1e-3;
    int y0 = P.ny > 1 ? (int)clampd(ceil(umin - pad), -big, big) : 0, y1 = P.ny > 1 ? (int)clampd(floor(umax + pad), -big, big) : 0;
    int z0 = P.nz > 1 ? (int)clampd(ceil(wmin - pad), -big, big) : 0, z1 = P.nz > 1 ? (int)clampd(floor(wmax + pad), -big, big) : 0;
    y0 = max(y0, 0); y1 = min(y1, (int)P.ny - 1); z0 = max(z0, 0); z1 = min(z1, (int)P.nz - 1);
    // rows of this batch only (a batch is a contiguous range of shard-local indices, i.e. of rows up to one partial row)
    {
        const unsigned long long first = P.ray_begin + P.batch_base * P.ray_stride;
        const unsigned long long last = P.ray_begin + (P.batch_base + P.n_primary - 1) * P.ray_stride;
        z0 = max(z0, (int)(first / P.ny)); z1 = min(z1, (int)(last / P.ny));
    }
    // columns of this shard only, when they form a lattice (stride divides the row length)
    F.ystep = 1;
    F.k0 = y0;
    if (P.lat_w && P.ray_stride > 1) {
        const int s = (int)P.ray_stride, c0 = (int)P.lat_c0;
        y0 += ((c0 - y0 % s) + s) % s;
        F.ystep = s;
        F.k0 = (y0 - c0) / s;
    }
    F.y0 = y0; F.y1 = y1; F.z0 = z0; F.z1 = z1;
    if (y0 > y1 || z0 > z1) return false;
    // padded edge functions (only for triangles entirely in front)
    F.use2d = false;
    if (n_front == 3 && P.ny > 1 && P.nz > 1) {
        const double ax = pu[0] - y0, ay = pw[0] - z0, bxx = pu[1] - y0, by = pw[1] - z0, cx = pu[2] - y0, cy = pw[2] - z0;
        const double area2 = (bxx - ax) * (cy - ay) - (by - ay) * (cx - ax);
        if (fabs(area2) > 1e-6 && fabs(ax) < 1e6 && fabs(ay) < 1e6 && fabs(bxx) < 1e6 && fabs(by) < 1e6 && fabs(cx) < 1e6 && fabs(cy) < 1e6) {
            const double sg = area2 > 0 ? 1.0 : -1.0;
            const double xs[3] = {ax, bxx, cx}, ys[3] = {ay, by, cy};
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int a = k, b = (k + 1) % 3;
                const double A = -(ys[b] - ys[a]) * sg, B = (xs[b] - xs[a]) * sg;
                // slack: 0.05 pixel along the edge normal plus the fp32 evaluation error of the edge function; the last term
                // covers k_raster_small's incremental evaluation along a row (at most 160 fp32 additions to a value below
                // 322 (|A| + |B|): 160 * 2^-24 * 322 = 3.1e-3)
                const double C = -(A * xs[a] + B * ys[a]) + 0.05 * (fabs(A) + fabs(B)) + 1e-5 * (fabs(A * xs[a]) + fabs(B * ys[a]) + (fabs(A) + fabs(B)) * 64.0) +
                                 4e-3 * (fabs(A) + fabs(B));
                F.ea[k] = (float)A; F.eb[k] = (float)B; F.ec[k] = (float)C;
            }
            F.use2d = true;
        }
    }
    return true;
}

__device__ __forceinline__ unsigned long long foot_area(const TriFoot &F)
{
    return (unsigned long long)((F.y1 - F.y0) / F.ystep + 1) * (unsigned long long)(F.z1 - F.z0 + 1);
}

__device__ __forceinline__ bool foot_inside2d(const TriFoot &F, int y, int z)
{
    if (!F.use2d) return true;
    const float fy = (float)(y - F.y0), fz = (float)(z - F.z0);
    const float e0 = fmaf(F.ea[0], fy, fmaf(F.eb[0], fz, F.ec[0]));
    const float e1 = fmaf(F.ea[1], fy, fmaf(F.eb[1], fz, F.ec[1]));
    const float e2 = fmaf(F.ea[2], fy, fmaf(F.eb[2], fz, F.ec[2]));
    return fminf(fminf(e0, e1), e2) >= 0.f;
}

// The reference test for one ray of the footprint; same operations and order as tri_accept().
__device__ __forceinline__ void foot_test(const WaveParams &P, const TriFoot &F, unsigned rel)
{
    const d3 dir = mk3(P.dirs[0][rel], P.dirs[1][rel], P.dirs[2][rel]);
    const d3 e2 = (1 / dot3(F.n, dir)) * F.pmo;
    const double t = dot3(F.n, e2);
    if (!((t < (double)RT_DEFAULT_MAX_F) & (t > (double)SCENE_EPS))) return;
    const d3 i = cross3(dir, e2);
    const double beta = dot3(i, F.e1);
    const double gamma = dot3(i, F.e0);
    if (!((beta >= 0.0f) & (gamma >= 0.0f) & (beta + gamma <= 1))) return;
    const float tf = (float)t;
    if (!(tf > SCENE_EPS)) return;
    atomicMin(P.hits + rel, ((unsigned long long)__float_as_uint(tf) << 32) | (unsigned long long)F.id);
}

// Pass 1 (part of k_raster_small): summed footprint (the guard) and the row chunks of the large footprints.  Called by every
// thread of the block with its triangle's footprint (ok: it has one).
__device__ __forceinline__ void raster_setup(const WaveParams &P, bool ok, unsigned pos, const TriFoot &F)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long area = 0;
    unsigned c_pos = 0, c_n = 0, c_at = 0, c_rows = 0;    // this lane's large footprint: leaf position, chunks, first item, rows per chunk
    int c_z0 = 0, c_z1 = 0;
    if (ok) {
        area = foot_area(F);
        if (area > RTS_RASTER_SMALL) {
            const unsigned per_row = (unsigned)((F.y1 - F.y0) / F.ystep + 1);
            const unsigned rows = max(1u, P.raster_chunk / per_row);
            // one reservation for all of the footprint's chunks (a triangle that fills the image has thousands:
            // an atomic per chunk would serialise them on one thread); the warp writes them together below
            const unsigned n_chunks = (unsigned)(F.z1 - F.z0) / rows + 1u;
            const unsigned at = atomicAdd(&P.raster_ctl->n_items, n_chunks);
            if (at < P.raster_item_cap && n_chunks <= P.raster_item_cap - at) {
                c_pos = pos; c_n = n_chunks; c_at = at; c_rows = rows; c_z0 = F.z0; c_z1 = F.z1;
            } else {
                area = 1ull << 56;   // too many chunks: turn the path off
            }
        }
    }
    for (unsigned todo = __ballot_sync(0xffffffffu, c_n != 0); todo; todo &= todo - 1u) {
        const int src = __ffs(todo) - 1;
        const unsigned spos = __shfl_sync(0xffffffffu, c_pos, src), n = __shfl_sync(0xffffffffu, c_n, src), at = __shfl_sync(0xffffffffu, c_at, src);
        const unsigned rows = __shfl_sync(0xffffffffu, c_rows, src);
        const int z0 = __shfl_sync(0xffffffffu, c_z0, src), z1 = __shfl_sync(0xffffffffu, c_z1, src);
        for (unsigned c = lane; c < n; c += 32u) {
            const int z = z0 + (int)(c * rows);
            RasterItem it; it.pos = spos; it.z0 = (unsigned)z; it.z1 = (unsigned)min(z1, z + (int)rows - 1); it.pad = 0;
            P.raster_items[at + c] = it;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if (lane == 0 && area) atomicAdd(&P.raster_ctl->area, area);
}

// Passes 1 + 2a: one thread per triangle — its footprint (summed for the guard; cut into row chunks for k_raster_big when
// large), then the small footprints are walked: each lane first scans forward to its next candidate that passes the
// cheap 2D test, then the warp runs the fp64 test together.  (One kernel since round 2: a separate set-up pass computed
// every footprint a second time, 0.03 ms per million triangles.)
#ifndef RTS_RASTER_MIN_BLOCKS
#define RTS_RASTER_MIN_BLOCKS 7      // 72 registers: the kernel is latency-bound (scattered direction loads, hit-word atomics); 5 / 6 / 7 / 8 CTAs: wave 0 2.199 / 2.170 / 2.145 / 2.156 ms
#endif
__global__ void __launch_bounds__(128, RTS_RASTER_MIN_BLOCKS) k_raster_small(const __grid_constant__ WaveParams P)
{
    // (no guard here: the guard is the sum this very kernel forms, and a small footprint is at most 160 candidates — what
    // can explode is the large footprints' work, and k_raster_big and the shading pass do look at the guard)
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    TriFoot F;
    const unsigned pos = idx < raster_count(P) ? raster_pos(P, idx) : 0u;
    bool have = idx < raster_count(P) && tri_footprint(P, pos, F);
    raster_setup(P, have, pos, F);
    have = have && foot_area(F) <= RTS_RASTER_SMALL;
    int y = have ? F.y0 : 0, z = have ? F.z0 : 1, k = have ? F.k0 : 0;
    const int z1 = have ? F.z1 : 0;
    // the three edge functions at the lane's current pixel: set at the start of a row, then one addition per step (the
    // slack of tri_footprint covers the rounding of up to 160 of them); without edge functions every pixel of the box passes
    const bool use2d = have && F.use2d;
    const float da0 = use2d ? F.ea[0] * (float)F.ystep : 0.f, da1 = use2d ? F.ea[1] * (float)F.ystep : 0.f, da2 = use2d ? F.ea[2] * (float)F.ystep : 0.f;
    float ex0 = use2d ? F.ec[0] : 0.f, ex1 = use2d ? F.ec[1] : 0.f, ex2 = use2d ? F.ec[2] : 0.f;      // (y0, z0): fy = fz = 0
    while (__any_sync(0xffffffffu, z <= z1)) {
        unsigned rel = 0;
        bool found = false;
        while (z <= z1) {
            const bool ok = fminf(fminf(ex0, ex1), ex2) >= 0.f && pixel_local(P, (unsigned)y, (unsigned)z, k, rel);
            y += F.ystep; k++;
            ex0 += da0; ex1 += da1; ex2 += da2;
            if (y > F.y1) {
                y = F.y0; k = F.k0; z++;
                if (use2d) {
                    const float fz = (float)(z - F.z0);
                    ex0 = fmaf(F.eb[0], fz, F.ec[0]); ex1 = fmaf(F.eb[1], fz, F.ec[1]); ex2 = fmaf(F.eb[2], fz, F.ec[2]);
                }
            }
            if (ok) { found = true; break; }
        }
        if (found) foot_test(P, F, rel);
    }
}

// Pass 2b: one warp per row chunk of a large footprint, lanes stride along the row.
__global__ void k_raster_big(const __grid_constant__ WaveParams P)
{
    if (!raster_on(P)) return;
    const unsigned n_items = min(P.raster_ctl->n_items, P.raster_item_cap);
    const unsigned lane = threadIdx.x & 31u;
    for (;;) {
        unsigned it = 0;
        if (lane == 0) it = atomicAdd(&P.raster_ctl->next_item, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const RasterItem I = P.raster_items[it];
        TriFoot F;
        if (!tri_footprint(P, I.pos, F)) continue;
        for (int z = (int)I.z0; z <= (int)I.z1; z++)
            for (int y = F.y0 + (int)lane * F.ystep, k = F.k0 + (int)lane; y <= F.y1; y += 32 * F.ystep, k += 32) {
                unsigned rel;
                if (foot_inside2d(F, y, z) && pixel_local(P, (unsigned)y, (unsigned)z, k, rel)) foot_test(P, F, rel);
            }
    }
}

// Pass 3a: winner's triangle id -> leaf position, in place, so that the shading pass below has one dependent
// (scattered) fetch per ray — the triangle record — instead of three.
__global__ void k_raster_resolve(const __grid_constant__ WaveParams P)
{
    if (!raster_on(P)) return;
    for (unsigned long long rel = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; rel < P.n_primary;
         rel += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long hit = P.hits[rel];
        if (hit != ~0ull) {
            // bit 31 of the position word: the winner is the kept static one, so this ray's first reflection is the
            // same ray as in the pulse the first-reflection hits were kept from (coherent.cuh)
            const unsigned long long same = (P.w1_static && P.hits_static[rel] == hit) ? 0x80000000ull : 0ull;
            P.hits[rel] = (hit & 0xffffffff00000000ull) | (unsigned long long)P.leaf_of_tri[(uint32_t)hit] | same;
        }
    }
}

// Pass 3b: the primary wave without traversal — the closest hit comes from the hit buffer.
template <bool RECORDS, bool TABLES = false>
__global__ void __launch_bounds__(RTS_WAVE_BLOCK, RTS_SHADE_MIN_BLOCKS) k_primary_shade(const __grid_constant__ WaveParams P)
{
    if (!raster_on(P)) return;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned n_in = (unsigned)P.n_primary;
    Local L = {0, 0, 0, 0, 0};
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(P.wave_segs + P.wave_index, (unsigned long long)n_in);
        atomicAdd(&P.counters->segments, (unsigned long long)n_in);
    }
    // Every ray costs about the same here (no traversal), so the rays are dealt out statically; the next ray's
    // hit word and direction are fetched, and its triangle record prefetched, while the current one is shaded.
    bins_smem_init(P);
    const unsigned stride = gridDim.x * blockDim.x;
    unsigned rel = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long hit_n = ~0ull;
    double dnx = 0, dny = 0, dnz = 0;
    if (rel < n_in) { hit_n = __ldcs(P.hits + rel); dnx = __ldcs(P.dirs[0] + rel); dny = __ldcs(P.dirs[1] + rel); dnz = __ldcs(P.dirs[2] + rel); }
    while (rel < n_in) {
        const unsigned long long hit = hit_n;
        Ray r;
        r.ox = P.origin[0]; r.oy = P.origin[1]; r.oz = P.origin[2];
        r.dx = dnx; r.dy = dny; r.dz = dnz;
        const unsigned nrel = rel + stride;
        if (nrel < n_in) {
            hit_n = __ldcs(P.hits + nrel); dnx = __ldcs(P.dirs[0] + nrel); dny = __ldcs(P.dirs[1] + nrel); dnz = __ldcs(P.dirs[2] + nrel);
        }
        r.meta = m_make(0, 0, 0, false, true, 0);
        r.len = 0; r.pw = 0; r.dop = 0; r.fx = 0; r.fy = 0; r.fz = 0; r.n0 = 1; r.n1 = 1;
        r.key = 0; r.ray = (uint32_t)(P.ray_begin + (P.batch_base + rel) * P.ray_stride);
        if (hit != ~0ull) {
            HitRec h;
            // low word: leaf position (+ kept flag) after k_raster_resolve, or still the triangle id when that pass was skipped
            h.pos = P.hits_resolved ? (int)((uint32_t)hit & 0x7fffffffu) : (int)P.leaf_of_tri[(uint32_t)hit];
            h.t = __uint_as_float((unsigned)(hit >> 32)); h.id = 0;   // shade() takes the id from the record
            L.a += C_HIT;
            shade<RECORDS, TABLES>(P, r, h, L, false, (P.hits_resolved && ((uint32_t)hit & 0x80000000u)) ? M_COH : 0u);
        } else {
            const int received = miss<RECORDS>(P, r, L);
            if (received >= 0) {
                L.a += C_CAPTURED;
                if (P.flags & RTS_OUT_BINS) accumulate_bin<TABLES>(P, r, received);
            }
        }
        rel = nrel;
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
}
