// api.cu — the C-ABI of librts_b200.so (include/rts_b200.h) and the per-pulse orchestration.
//
// Host-side counterpart of the reference's ray_tracer.cpp pulse loop (ray_tracer.cpp:843-1333),
// minus everything that talks to SOARS or OptiX: scene upload, per-pulse pose update + BVH refit,
// the wavefront launch sequence, result hand-off.  Also exports the C++ symbol
// rs::kernel_wrapper with the reference's exact signature (aggregation.cuh:18-23).
#include "engine.h"
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <vector>

static thread_local char g_err[512] = "";

int rts_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *rts_last_error(void) { return g_err; }
extern "C" const char *rts_version(void) { return "rts_b200 0.1 (sm_100a; wavefront LBVH tracer; no CPU fallback)"; }

extern "C" int rts_abi_sizes(uint32_t s[10])
{
    if (!s) return rts_fail(RTS_ERR_ARG, "sizes is NULL");
    s[0] = sizeof(rts_ray_record); s[1] = sizeof(rts_target_mesh); s[2] = sizeof(rts_rx_sphere); s[3] = sizeof(rts_rx_desc);
    s[4] = sizeof(rts_pulse); s[5] = sizeof(rts_bin); s[6] = sizeof(rts_stats); s[7] = sizeof(rts_pose);
    s[8] = sizeof(rts_response); s[9] = sizeof(rts_sizes);
    return RTS_OK;
}

static_assert(sizeof(rts_ray_record) == 144, "PerRayData contract: 144 bytes");
static_assert(offsetof(rts_ray_record, refrIndex) == 16 && offsetof(rts_ray_record, reflDepth) == 32 &&
              offsetof(rts_ray_record, rayDirection) == 48 && offsetof(rts_ray_record, firstHitPoint) == 72 &&
              offsetof(rts_ray_record, prevHitPoint) == 96 && offsetof(rts_ray_record, power) == 120 &&
              offsetof(rts_ray_record, doppler) == 128 && offsetof(rts_ray_record, received) == 136 &&
              offsetof(rts_ray_record, end) == 140, "PerRayData contract: offsets");

extern "C" int rts_result_sizes(const rts_pulse *p, rts_sizes *out)
{
    if (!p || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    const uint32_t rMax = p->max_refr > 0 ? 2u : 0u;               // ray_tracer.cpp:604-605
    out->rays = (uint64_t)p->nx * p->ny * p->nz;
    out->slots = rMax == 2 ? 1 + (p->max_refl + 1) + 1 : 1;         // ray_tracer.cpp:608-613
    out->ray_total = out->rays * out->slots;                        // ray_tracer.cpp:626
    out->depth_total = p->max_refl + rMax;                          // ray_tracer.cpp:655
    out->tri_cols = p->max_refl + 3;
    out->_pad = 0;
    return RTS_OK;
}

// ---------------------------------------------------------------------------------------------
static int set_option(rts_engine *e, const char *name, long long v)
{
    Knobs &k = e->knobs;
    if (!strcmp(name, "bvh")) e->builder_forced = (v == 1 || v == 2) ? (int)v : 0;           // 1 = Morton radix tree, 2 = PLOC, 0 = by SAH cost
    else if (!strcmp(name, "leaf_max")) { if (v < 1 || v > 8) return rts_fail(RTS_ERR_ARG, "leaf_max must be 1..8"); e->leaf_max = (int)v; }
    else if (!strcmp(name, "no_chain")) k.no_chain = v != 0;
    else if (!strcmp(name, "no_raster")) k.no_raster = v != 0;
    else if (!strcmp(name, "no_tiles")) k.no_tiles = v != 0;
    else if (!strcmp(name, "one_ended_queue")) k.one_ended_queue = v != 0;
    else if (!strcmp(name, "debug_raster")) k.debug_raster = v != 0;
    else if (!strcmp(name, "no_static_hits")) k.no_static_hits = v != 0;
    else if (!strcmp(name, "no_kept_reflections")) k.no_kept_reflections = v != 0;
    else if (!strcmp(name, "no_split")) {
        const bool was = k.no_split != 0;
        k.no_split = v != 0;
        // the split form walks the quantised nodes, which are only maintained while it is enabled (bvh.cu: want_qnodes)
        if (was && !k.no_split && e->scene_ready) { int rc = bvh_build(e); if (rc) return rc; }
    }
    else if (!strcmp(name, "split_below")) { if (v < 0 || v > (1ll << 30)) return rts_fail(RTS_ERR_ARG, "split_below out of range"); k.split_below = (uint32_t)v; }
    else if (!strcmp(name, "no_follow")) k.no_follow = v != 0;
    else if (!strcmp(name, "no_smem_bins")) k.no_smem_bins = v != 0;
    else if (!strcmp(name, "no_split_raster")) k.no_split_raster = v < 0 ? -1 : (v != 0);   // -1: split even when nothing is in flight (tests)
    else if (!strcmp(name, "debug_timeline")) { k.debug_timeline = v != 0; e->tl_n = 0; }
    else if (!strcmp(name, "no_overlap")) { bvh_join(e); k.no_overlap = v != 0; }
    else if (!strcmp(name, "hash_bins")) k.hash_bins = v != 0;
    else if (!strcmp(name, "hash_log2")) { if (v < 4 || v > 28) return rts_fail(RTS_ERR_ARG, "hash_log2 must be 4..28"); k.hash_log2 = (uint32_t)v; e->hash_ready = false; }
    else if (!strcmp(name, "batch")) { if (v != 0 && (v < 32 || v > (1ll << 24))) return rts_fail(RTS_ERR_ARG, "batch must be 0 or 32..2^24"); k.batch = v; }
    else return rts_fail(RTS_ERR_ARG, "unknown option '%s'", name);
    return RTS_OK;
}

extern "C" int rts_set_option(rts_engine *e, const char *name, int64_t value)
{
    if (!e || !name) return rts_fail(RTS_ERR_ARG, "NULL argument");
    return set_option(e, name, (long long)value);
}

extern "C" int rts_create(int device, rts_engine **out)
{
    if (!out) return rts_fail(RTS_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return rts_fail(RTS_ERR_NO_DEVICE, "no CUDA device available (%s); librts_b200 has no CPU fallback",
                        err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
    if (device < 0 || device >= count) return rts_fail(RTS_ERR_ARG, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    RTS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return rts_fail(RTS_ERR_NO_DEVICE, "device %d is sm_%d%d; librts_b200 is built for sm_100a only", device, prop.major,
                        prop.minor);
    RTS_CUDA(cudaSetDevice(device));
    rts_engine *e = new rts_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    // the engine's own stream outranks side_dirs (created below at the lowest priority): the direction pass should take what
    // the thin last waves of the previous pulse leave, not the other way round (a caller's stream: see INTEGRATION.md)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&e->own_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        delete e;
        return rts_fail(RTS_ERR_CUDA, "cudaStreamCreate failed");
    }
    e->stream = e->own_stream;
    // tuning / test switches: the environment is read here, once; afterwards only rts_set_option changes them
    for (const char *name : {"bvh", "leaf_max", "no_chain", "no_raster", "no_tiles", "one_ended_queue", "debug_raster", "no_static_hits",
                             "no_kept_reflections", "no_split", "split_below", "no_follow", "no_smem_bins", "no_overlap", "debug_timeline", "no_split_raster", "batch", "hash_bins", "hash_log2"}) {
        std::string env = "RTS_";
        for (const char *c = name; *c; c++) env += (char)toupper(*c);
        if (const char *v = getenv(env.c_str())) {
            if (!strcmp(name, "bvh")) e->builder_forced = !strcmp(v, "ploc") ? 2 : (!strcmp(v, "lbvh") ? 1 : 0);
            else set_option(e, name, atoll(v));
        }
    }
    for (auto &ev : e->ev) cudaEventCreate(&ev);
    for (auto &ev : e->wave_ev) cudaEventCreate(&ev);
    for (auto &ev : e->split_ev) cudaEventCreate(&ev);
    for (auto &ev : e->follow_ev) cudaEventCreate(&ev);
    for (auto &s : e->stage) cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->sah_ev, cudaEventDisableTiming);
    {   // side streams (engine.h); the refit's is the more urgent one: the kernel that walks the tree waits for it
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&e->side_bvh, cudaStreamNonBlocking, hi) != cudaSuccess) e->side_bvh = nullptr;
        if (cudaStreamCreateWithPriority(&e->side_dirs, cudaStreamNonBlocking, lo) != cudaSuccess) e->side_dirs = nullptr;
        for (cudaEvent_t *ev : {&e->ev_dirs_free, &e->ev_dirs_done, &e->ev_bvh_fork, &e->ev_bvh_done}) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        cudaGetLastError();
    }
    if (cudaMallocHost((void **)&e->h_rb, sizeof(Readback)) != cudaSuccess) {
        rts_destroy(e);
        return rts_fail(RTS_ERR_CUDA, "cudaMallocHost failed");
    }
    memset(e->h_rb, 0, sizeof(Readback));
    cudaMalloc(&e->d_wave_segs, sizeof(unsigned long long) * 32);
    cudaMalloc(&e->d_counts, sizeof(unsigned long long) * 96);
    cudaMalloc(&e->d_counters, sizeof(Counters));
    cudaMalloc(&e->d_rx, sizeof(RxDev) * RTS_MAX_RX);
    *out = e;
    return RTS_OK;
}

static void free_scene(rts_engine *e)
{
    void **ptrs[] = {(void **)&e->d_base_verts, (void **)&e->d_base_normals, (void **)&e->d_world_verts,
                     (void **)&e->d_world_normals, (void **)&e->d_tris, (void **)&e->d_tri_target, (void **)&e->d_vert_target,
                     (void **)&e->d_norm_target, (void **)&e->d_t_vert_off, (void **)&e->d_t_norm_off, (void **)&e->d_t_tri_off,
                     (void **)&e->d_t_per_face, (void **)&e->d_t_refl, (void **)&e->d_t_refr, (void **)&e->d_t_vel, (void **)&e->d_t_rcs,
                     (void **)&e->d_poses};
    for (void **p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    bvh_free(e);
    e->scene_ready = false;
}

extern "C" void rts_destroy(rts_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    rts_comm_destroy(e);
    for (void *p : {(void *)e->d_rcs_tab, (void *)e->d_rcs_values, (void *)e->d_ant_tx, (void *)e->d_ant_rx, (void *)e->d_ant_values}) if (p) cudaFree(p);
    free_scene(e);
    for (int k = 0; k < 2; k++) if (e->q_slab[k]) cudaFree(e->q_slab[k]);
    void *ptrs[] = {e->d_ckeys, e->d_csums, e->d_cmins, e->d_hash_keys, e->d_hash_used, e->d_hash_count, e->d_trav_hits, e->d_todo, e->d_coop_stacks, e->d_w1_static, e->d_target_box, e->d_mover_nodes, e->d_dirs, e->d_hits, e->d_hits_static, e->d_raster_ctl, e->d_raster_ctl_static, e->d_raster_items, e->d_counts, e->d_counters, e->d_rx, e->d_bin_sums, e->d_bin_mins, e->d_bins_out, e->d_bins_out_count, e->d_rx_sums, e->d_rx_mins,
                    e->d_results, e->d_targ_intersect, e->d_tri_path, e->d_rcs_angle};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->wave_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->split_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->follow_ev) if (ev) cudaEventDestroy(ev);
    for (auto &s : e->stage) { if (s.done) cudaEventDestroy(s.done); if (s.host) cudaFreeHost(s.host); }
    if (e->sah_ev) cudaEventDestroy(e->sah_ev);
    for (auto &row : e->tl_ev) for (auto &ev : row) if (ev) cudaEventDestroy(ev);
    for (cudaStream_t sd : {e->side_dirs, e->side_bvh}) if (sd) { cudaStreamSynchronize(sd); cudaStreamDestroy(sd); }
    for (cudaEvent_t ev : {e->ev_dirs_free, e->ev_dirs_done, e->ev_bvh_fork, e->ev_bvh_done}) if (ev) cudaEventDestroy(ev);
    if (e->h_rb) cudaFreeHost(e->h_rb);
    for (int k = 0; k < 2; k++) { if (e->h_bins_buf[k]) cudaFreeHost(e->h_bins_buf[k]); if (e->bins_ev[k]) cudaEventDestroy(e->bins_ev[k]); }
    if (e->h_bins_count) cudaFreeHost(e->h_bins_count);
    if (e->d_wave_segs) cudaFree(e->d_wave_segs);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

// ---- pinned staging ring: the caller's small per-pulse arrays are copied here, then to the device
// asynchronously, so that no entry point has to wait for a pageable-memory copy ----
char *stage_acquire(rts_engine *e, size_t bytes)
{
    StageSlot &s = e->stage[e->stage_next];
    if (s.in_flight) { cudaEventSynchronize(s.done); s.in_flight = false; }
    if (s.cap < bytes) {
        if (s.host) cudaFreeHost(s.host);
        s.host = nullptr; s.cap = 0;
        const size_t cap = std::max<size_t>(bytes, 16384);
        if (cudaMallocHost((void **)&s.host, cap) != cudaSuccess) return nullptr;
        s.cap = cap;
    }
    return s.host;
}
void stage_release(rts_engine *e)
{
    StageSlot &s = e->stage[e->stage_next];
    cudaEventRecord(s.done, e->stream);
    s.in_flight = true;
    e->stage_next = (e->stage_next + 1) % 8;
}

extern "C" int rts_set_stream(rts_engine *e, void *cuda_stream)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    bvh_join(e);
    cudaStreamSynchronize(e->stream);
    if (e->side_dirs) cudaStreamSynchronize(e->side_dirs);
    e->dirs_free_valid = false;      // ev_dirs_free belongs to the stream that is being left
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return RTS_OK;
}

// ---------------------------------------------------------------------------------------------
template <class T> static int upload(T **dst, const std::vector<T> &src)
{
    if (*dst) { cudaFree(*dst); *dst = nullptr; }
    RTS_CUDA(cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(1, src.size())));
    if (!src.empty()) RTS_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return RTS_OK;
}

extern "C" int rts_scene_set_targets(rts_engine *e, const rts_target_mesh *targets, uint32_t n_targets)
{
    NvtxRange nvtx_range("rts:scene_set_targets (upload + BVH build)");
    if (!e || (!targets && n_targets)) return rts_fail(RTS_ERR_ARG, "NULL argument");
    RTS_CUDA(cudaSetDevice(e->device));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    free_scene(e);
    e->n_targets = n_targets;
    e->tri_off.assign(n_targets, 0); e->vert_off.assign(n_targets, 0); e->norm_off.assign(n_targets, 0);
    e->t_nverts.assign(n_targets, 0); e->t_ntris.assign(n_targets, 0); e->t_nnormals.assign(n_targets, 0);
    uint64_t T = 0, V = 0, N = 0;
    for (uint32_t k = 0; k < n_targets; k++) {
        const rts_target_mesh &m = targets[k];
        if ((m.n_verts && !m.verts) || (m.n_tris && !m.tris) || (m.n_normals && !m.normals))
            return rts_fail(RTS_ERR_ARG, "target %u has NULL arrays", k);
        e->tri_off[k] = (uint32_t)T; e->vert_off[k] = (uint32_t)V; e->norm_off[k] = (uint32_t)N;
        e->t_nverts[k] = m.n_verts; e->t_ntris[k] = m.n_tris; e->t_nnormals[k] = m.n_normals;
        T += m.n_tris; V += m.n_verts; N += m.n_normals;
    }
    if (T >= (1ull << 28)) return rts_fail(RTS_ERR_CAPACITY, "%llu triangles exceed the 2^28 leaf-reference limit", (unsigned long long)T);
    e->n_tris = (uint32_t)T; e->n_verts = (uint32_t)V; e->n_normals = (uint32_t)N;

    std::vector<double> verts(3 * V), normals(3 * N), refl(n_targets), refr(n_targets), vel(3 * (size_t)n_targets, 0.0);
    std::vector<uint32_t> tris(3 * T), tri_target(T), vert_target(V), norm_target(N), per_face(n_targets);
    std::vector<rts_pose> poses(n_targets);
    for (uint32_t k = 0; k < n_targets; k++) {
        const rts_target_mesh &m = targets[k];
        for (uint32_t t = 0; t < m.n_tris; t++)
            for (int c = 0; c < 3; c++) {
                const uint32_t vi = m.tris[3 * (size_t)t + c];
                if (vi >= m.n_verts) return rts_fail(RTS_ERR_ARG, "target %u triangle %u references vertex %u of %u", k, t, vi, m.n_verts);
                tris[3 * ((size_t)e->tri_off[k] + t) + c] = vi;
            }
        if (m.n_verts) memcpy(&verts[3 * (size_t)e->vert_off[k]], m.verts, sizeof(double) * 3 * m.n_verts);
        if (m.n_normals) memcpy(&normals[3 * (size_t)e->norm_off[k]], m.normals, sizeof(double) * 3 * m.n_normals);
        std::fill(tri_target.begin() + e->tri_off[k], tri_target.begin() + e->tri_off[k] + m.n_tris, k);
        std::fill(vert_target.begin() + e->vert_off[k], vert_target.begin() + e->vert_off[k] + m.n_verts, k);
        std::fill(norm_target.begin() + e->norm_off[k], norm_target.begin() + e->norm_off[k] + m.n_normals, k);
        per_face[k] = m.n_normals > m.n_verts ? 1u : 0u;   // triangle_mesh.cu:180
        refl[k] = m.refl_coeff; refr[k] = m.refr_index;
        memset(&poses[k], 0, sizeof(rts_pose));
        poses[k].R[0] = poses[k].R[4] = poses[k].R[8] = 1.0;
    }
    int rc;
    if ((rc = upload(&e->d_base_verts, verts))) return rc;
    if ((rc = upload(&e->d_base_normals, normals))) return rc;
    if ((rc = upload(&e->d_world_verts, verts))) return rc;
    if ((rc = upload(&e->d_world_normals, normals))) return rc;
    if ((rc = upload(&e->d_tris, tris))) return rc;
    if ((rc = upload(&e->d_tri_target, tri_target))) return rc;
    if ((rc = upload(&e->d_vert_target, vert_target))) return rc;
    if ((rc = upload(&e->d_norm_target, norm_target))) return rc;
    if ((rc = upload(&e->d_t_vert_off, e->vert_off))) return rc;
    if ((rc = upload(&e->d_t_norm_off, e->norm_off))) return rc;
    if ((rc = upload(&e->d_t_tri_off, e->tri_off))) return rc;
    if ((rc = upload(&e->d_t_per_face, per_face))) return rc;
    if ((rc = upload(&e->d_t_refl, refl))) return rc;
    if ((rc = upload(&e->d_t_refr, refr))) return rc;
    if ((rc = upload(&e->d_t_vel, vel))) return rc;
    if ((rc = upload(&e->d_t_rcs, refl))) return rc;   // sized [n_targets]; filled per pulse when targ_rcs is given
    if ((rc = upload(&e->d_poses, poses))) return rc;
    e->h_poses = poses;
    e->moving.assign(n_targets, 0);
    e->scene_version++;
    e->moving_version++;
    e->builder = e->builder_forced;
    if ((rc = bvh_alloc(e))) return rc;
    if ((rc = bvh_build(e))) return rc;
    e->scene_ready = true;
    return RTS_OK;
}

static bool same_pose(const rts_pose &a, const rts_pose &b)
{
    if ((a.has_rotation != 0) != (b.has_rotation != 0)) return false;
    if (memcmp(a.t, b.t, sizeof(a.t)) != 0) return false;
    return !a.has_rotation || memcmp(a.R, b.R, sizeof(a.R)) == 0;
}

extern "C" int rts_scene_set_poses(rts_engine *e, const rts_pose *poses, uint32_t n_targets)
{
    NvtxRange nvtx_range("rts:scene_set_poses (transform + refit)");
    if (!e || !poses) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->scene_ready) return rts_fail(RTS_ERR_STATE, "no scene committed");
    if (n_targets != e->n_targets) return rts_fail(RTS_ERR_ARG, "%u poses for %u targets", n_targets, e->n_targets);
    RTS_CUDA(cudaSetDevice(e->device));
    // a target counts as moving from the first pose that differs from the one it had before
    bool changed = false, grew = false;
    for (uint32_t k = 0; k < n_targets; k++) {
        if (!same_pose(poses[k], e->h_poses[k])) {
            changed = true;
            if (!e->moving[k]) { e->moving[k] = 1; grew = true; }
        }
        e->h_poses[k] = poses[k];
    }
    if (grew) { e->partial_ready = false; e->moving_version++; }
    if (!changed) { e->refit_timed = false; e->bvh_info.ms_refit = 0.f; return RTS_OK; }
    cudaEventRecord(e->ev[4], e->stream);
    char *st = stage_acquire(e, sizeof(rts_pose) * n_targets);
    if (!st) return rts_fail(RTS_ERR_CUDA, "pinned staging allocation failed");
    memcpy(st, poses, sizeof(rts_pose) * n_targets);
    RTS_CUDA(cudaMemcpyAsync(e->d_poses, st, sizeof(rts_pose) * n_targets, cudaMemcpyHostToDevice, e->stream));
    stage_release(e);
    int rc = bvh_refit(e);
    if (rc) return rc;
    cudaEventRecord(e->ev[5], e->bvh_join_pending ? e->side_bvh : e->stream);   // where the refit's last kernel went
    e->refit_timed = true;
    return RTS_OK;   // nothing waited for: the transform and refit run on the engine's stream
}

static void collect_refit_time(rts_engine *e)
{
    if (!e->refit_timed) return;
    cudaEventSynchronize(e->ev[5]);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev[4], e->ev[5]) == cudaSuccess) e->bvh_info.ms_refit = ms;
    e->refit_timed = false;
}

extern "C" int rts_scene_rebuild(rts_engine *e)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->scene_ready) return rts_fail(RTS_ERR_STATE, "no scene committed");
    RTS_CUDA(cudaSetDevice(e->device));
    return bvh_build(e);
}

extern "C" int rts_scene_bvh_info(rts_engine *e, rts_bvh_info *out)
{
    if (!e || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (e->scene_ready) {
        RTS_CUDA(cudaSetDevice(e->device));
        collect_refit_time(e);
        bvh_sync_info(e);
        int rc = bvh_read_scene_box(e);
        if (rc) return rc;
    }
    *out = e->bvh_info;
    out->sah_at_build = e->sah_at_build;
    out->builds = e->builds;
    out->builder = (uint32_t)e->builder;
    return RTS_OK;
}

extern "C" int rts_scene_get_world_vertices(rts_engine *e, uint32_t target, double *out)
{
    if (!e || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->scene_ready || target >= e->n_targets) return rts_fail(RTS_ERR_ARG, "bad target %u", target);
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    RTS_CUDA(cudaMemcpy(out, e->d_world_verts + 3 * (size_t)e->vert_off[target], sizeof(double) * 3 * e->t_nverts[target],
                        cudaMemcpyDeviceToHost));
    return RTS_OK;
}

extern "C" int rts_scene_get_tri_bounds(rts_engine *e, float *out6)
{
    if (!e || !out6) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->scene_ready) return rts_fail(RTS_ERR_STATE, "no scene committed");
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    RTS_CUDA(cudaMemcpy(out6, e->d_tri_box, sizeof(float) * 6 * (size_t)e->n_tris, cudaMemcpyDeviceToHost));
    return RTS_OK;
}

extern "C" int rts_scene_check_bvh(rts_engine *e, uint64_t *violations)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->scene_ready) return rts_fail(RTS_ERR_STATE, "no scene committed");
    return bvh_check(e, violations);
}

// ---------------------------------------------------------------------------------------------
static void sph_to_cart(double azi, double ele, double out[3]) // ray_tracer.cu:132-139
{
    out[0] = cos(azi) * cos(ele);
    out[1] = sin(azi) * cos(ele);
    out[2] = sin(ele);
}

// Launch-invariant part of ray_generation (ray_tracer.cu:155-196), evaluated once per pulse on the
// host in the source's operation order; the kernels then use only +,-,*,/ and sqrt per ray.
static void fill_launch_constants(WaveParams &P, const rts_pulse *p)
{
    const double az = p->tx_dir[0], el = p->tx_dir[1];
    double beamEnd[3];
    sph_to_cart(-p->tx_span[0] / 2, -p->tx_span[1] / 2, P.beamStart);
    sph_to_cart(p->tx_span[0] / 2, p->tx_span[1] / 2, beamEnd);
    P.slope[0] = p->nx > 1 ? (((beamEnd[0] * (1 + p->tx_span[2])) - P.beamStart[0]) / (p->nx - 1)) : 0.0;
    P.slope[1] = p->ny > 1 ? ((beamEnd[1] - P.beamStart[1]) / (p->ny - 1)) : 0.0;
    P.slope[2] = p->nz > 1 ? ((beamEnd[2] - P.beamStart[2]) / (p->nz - 1)) : 0.0;
    const double Rot[9] = {cos(az), -sin(az), 0, sin(az), cos(az), 0, 0, 0, 1};
    memcpy(P.Rot, Rot, sizeof(Rot));
    double rx = 0, ry = 0, rz = 0;
    rx += Rot[1]; ry += Rot[4]; rz += Rot[7];
    const double norm = sqrt(rx * rx + ry * ry + rz * rz);
    const double ox = rx / norm, oy = ry / norm, oz = rz / norm;
    const double c = cos(el), s = sin(el);
    const double Rot1[9] = {c + ox * ox * (1 - c), ox * oy * (1 - c) + oz * s, ox * oz * (1 - c) - oy * s,
                            oy * ox * (1 - c) - oz * s, c + oy * oy * (1 - c), oy * oz * (1 - c) + ox * s,
                            oz * ox * (1 - c) + oy * s, oz * oy * (1 - c) - ox * s, c + oz * oz * (1 - c)};
    memcpy(P.Rot1, Rot1, sizeof(Rot1));
    // A = Rot1 * Rot takes the beam-frame direction to the world; its transpose takes world offsets back (raster.cuh)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double a = 0;
            for (int k = 0; k < 3; k++) a += Rot1[3 * i + k] * Rot[3 * k + j];
            P.AT[3 * j + i] = a;
        }
    sph_to_cart(az, el, P.boresight);
    P.single_ray = (p->nx == 1 && p->ny == 1 && p->nz == 1) ? 1 : 0;
}

static int ensure_records(rts_engine *e, const rts_sizes &sz)
{
    if (e->rec_alloc_rays >= sz.ray_total && e->rec_alloc_D >= sz.depth_total && e->rec_alloc_W >= sz.tri_cols) return RTS_OK;
    void **ptrs[] = {(void **)&e->d_results, (void **)&e->d_targ_intersect, (void **)&e->d_rcs_angle, (void **)&e->d_tri_path};
    for (void **p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
    e->rec_alloc_rays = 0;
    const size_t n = sz.ray_total, D = std::max<uint32_t>(1, sz.depth_total), W = sz.tri_cols;
    RTS_CUDA(cudaMalloc(&e->d_results, sizeof(rts_ray_record) * n));
    RTS_CUDA(cudaMalloc(&e->d_targ_intersect, sizeof(int32_t) * n * D));
    RTS_CUDA(cudaMalloc(&e->d_rcs_angle, sizeof(double) * 2 * n * D));
    RTS_CUDA(cudaMalloc(&e->d_tri_path, sizeof(int32_t) * n * W));
    e->rec_alloc_rays = n; e->rec_alloc_D = sz.depth_total; e->rec_alloc_W = W;
    return RTS_OK;
}

extern "C" int rts_trace_pulse(rts_engine *e, const rts_pulse *p, uint32_t flags)
{
    NvtxRange nvtx_range("rts:trace_pulse");
    if (!e || !p) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->scene_ready) return rts_fail(RTS_ERR_STATE, "no scene committed (rts_scene_set_targets)");
    if (!(flags & (RTS_OUT_BINS | RTS_OUT_RECORDS))) return rts_fail(RTS_ERR_ARG, "flags must request RTS_OUT_BINS and/or RTS_OUT_RECORDS");
    if (p->n_targets != e->n_targets) return rts_fail(RTS_ERR_ARG, "pulse carries %u target velocities, scene has %u targets", p->n_targets, e->n_targets);
    if (p->n_rx > RTS_MAX_RX) return rts_fail(RTS_ERR_CAPACITY, "%u receivers > %u", p->n_rx, RTS_MAX_RX);
    if (p->n_rx && !p->rx) return rts_fail(RTS_ERR_ARG, "rx is NULL");
    if (p->n_targets && !p->targ_vel) return rts_fail(RTS_ERR_ARG, "targ_vel is NULL");
    if (!p->nx || !p->ny || !p->nz) return rts_fail(RTS_ERR_ARG, "empty launch grid");
    if (p->max_refl > 12) return rts_fail(RTS_ERR_CAPACITY, "max_refl %u too deep for the ray-state encoding (12)", p->max_refl);   // before any size arithmetic
    rts_sizes sz;
    rts_result_sizes(p, &sz);
    if (sz.depth_total > RTS_MAX_DEPTH) return rts_fail(RTS_ERR_CAPACITY, "depth_total %u > %u", sz.depth_total, RTS_MAX_DEPTH);
    if (sz.rays >= (1ull << 32)) return rts_fail(RTS_ERR_CAPACITY, "%llu primary rays per launch exceed 2^32", (unsigned long long)sz.rays);
    RTS_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;

    const uint32_t rMax = p->max_refr > 0 ? 2u : 0u;
    const uint64_t stride = p->ray_stride ? p->ray_stride : 1;
    uint64_t begin = std::min<uint64_t>(p->ray_begin, sz.rays);
    uint64_t end = p->ray_count ? std::min<uint64_t>(sz.rays, p->ray_begin + p->ray_count) : sz.rays;
    if (end < begin) end = begin;
    const uint64_t n_primary_total = (end - begin + stride - 1) / stride;

    WaveParams P;
    memset(&P, 0, sizeof(P));
    P.nodes = e->d_nodes; P.qnodes = e->d_qnodes; P.qframe = e->d_qframe; P.trirec = e->d_trirec; P.root_ref = e->root_ref; P.n_tris = e->n_tris;
    P.scene_abs = e->d_scene_abs;
    P.world_normals = e->d_world_normals; P.tris = e->d_tris;
    P.t_norm_off = e->d_t_norm_off; P.t_tri_off = e->d_t_tri_off; P.t_per_face = e->d_t_per_face;
    P.t_refl = e->d_t_refl; P.t_refr = e->d_t_refr; P.t_vel = e->d_t_vel;
    P.nx = p->nx; P.ny = p->ny; P.nz = p->nz; P.R3 = sz.rays;
    P.dMax = p->max_refl + 1;                                     // ray_tracer.cpp:776
    P.rMax = rMax; P.D = sz.depth_total; P.M = sz.slots; P.W = sz.tri_cols;
    P.interpolate = p->interpolate_smooth ? 1 : 0;
    memcpy(P.origin, p->tx_origin, sizeof(P.origin));
    fill_launch_constants(P, p);
    P.n_rx = p->n_rx; P.rx = e->d_rx;
    P.cspeed = p->cspeed; P.carrier = p->carrier;
    {
        const double Wl = p->cspeed / p->carrier;                 // ray_tracer.cpp:815
        const double Gt = p->gain_tx != 0 ? p->gain_tx : 1.0;     // Transmitter::GetGain stand-in (:1233)
        const double Gr = p->gain_rx != 0 ? p->gain_rx : 1.0;     // Receiver::GetGain stand-in (:1234-1235)
        P.wl2gain = (Wl * Wl * Gt * Gr);                          // ray_tracer.cpp:1247
    }
    P.t_rcs = p->targ_rcs ? e->d_t_rcs : nullptr;
    if (flags & RTS_TABLES) {   // tabulated Target::GetRCS / GetGain on the device (rts_set_rcs_tables / rts_set_antennas)
        if (!(flags & RTS_OUT_BINS) || (flags & RTS_OUT_RECORDS))
            return rts_fail(RTS_ERR_ARG, "RTS_TABLES applies to the fused bins: trace with RTS_OUT_BINS and without RTS_OUT_RECORDS");
        if (!e->d_rcs_tab && !e->d_ant_rx) return rts_fail(RTS_ERR_STATE, "RTS_TABLES without tables: rts_set_rcs_tables / rts_set_antennas first");
        if (e->d_rcs_tab) {
            if (rMax) return rts_fail(RTS_ERR_ARG, "tabulated RCS is not available with refraction (max_refr > 0): use the two-phase path (rts_get_received + rts_aggregate)");
            if (e->n_rcs_tab != e->n_targets) return rts_fail(RTS_ERR_ARG, "%u RCS tables for %u targets", e->n_rcs_tab, e->n_targets);
            P.rcs_tab = e->d_rcs_tab;
        }
        if (e->d_ant_rx) {
            if (e->n_ant_rx != p->n_rx) return rts_fail(RTS_ERR_ARG, "%u receiver antennas for %u receivers", e->n_ant_rx, p->n_rx);
            P.ant_tx = e->d_ant_tx; P.ant_rx = e->d_ant_rx;
            P.keep_first = 1;
            const double Wl = p->cspeed / p->carrier;
            P.wl2 = Wl * Wl;
            P.gain_tx_scalar = p->gain_tx != 0 ? p->gain_tx : 1.0;
            P.gain_rx_scalar = p->gain_rx != 0 ? p->gain_rx : 1.0;
        }
    }
    P.B = (uint64_t)e->n_targets + 1;
    // path key: digits base B = n_targets + 1 (digit 0 <=> -1)
    const uint64_t B = (uint64_t)e->n_targets + 1;
    P.powB[0] = 1;
    bool key_overflow = false;
    for (uint32_t c = 1; c <= RTS_MAX_DEPTH; c++) {
        if (c <= sz.depth_total && P.powB[c - 1] > (~0ull) / B / 2) key_overflow = true;
        P.powB[c] = P.powB[c - 1] * B;
    }
    if (key_overflow) return rts_fail(RTS_ERR_CAPACITY, "path key (targets+1)^depth does not fit 63 bits");
    P.key_all = 0;
    for (uint32_t c = 0; c < sz.depth_total; c++) P.key_all += P.powB[c];
    P.flags = flags;
    P.chain_below = e->knobs.no_chain ? 0u : (1u << 18);

    // receivers, target velocities, per-target RCS: through the pinned staging ring, no waiting
    {
        const size_t rx_bytes = sizeof(RxDev) * p->n_rx, vel_bytes = sizeof(double) * 3 * p->n_targets;
        const size_t rcs_bytes = p->targ_rcs ? sizeof(double) * p->n_targets : 0;
        char *stg = stage_acquire(e, rx_bytes + vel_bytes + rcs_bytes + 64);
        if (!stg) return rts_fail(RTS_ERR_CUDA, "pinned staging allocation failed");
        RxDev *rx = reinterpret_cast<RxDev *>(stg);
        for (uint32_t j = 0; j < p->n_rx; j++) {
            rx[j].cx = p->rx[j].centre[0]; rx[j].cy = p->rx[j].centre[1]; rx[j].cz = p->rx[j].centre[2];
            rx[j].radius = p->rx[j].radius;
            rx[j].min_theta = p->rx[j].min_theta; rx[j].max_theta = p->rx[j].max_theta;
            rx[j].min_phi = p->rx[j].min_phi; rx[j].max_phi = p->rx[j].max_phi;
        }
        if (p->n_rx) RTS_CUDA(cudaMemcpyAsync(e->d_rx, rx, rx_bytes, cudaMemcpyHostToDevice, st));
        if (p->n_targets) {
            memcpy(stg + rx_bytes, p->targ_vel, vel_bytes);
            RTS_CUDA(cudaMemcpyAsync(e->d_t_vel, stg + rx_bytes, vel_bytes, cudaMemcpyHostToDevice, st));
        }
        if (rcs_bytes) {
            memcpy(stg + rx_bytes + vel_bytes, p->targ_rcs, rcs_bytes);
            RTS_CUDA(cudaMemcpyAsync(e->d_t_rcs, stg + rx_bytes + vel_bytes, rcs_bytes, cudaMemcpyHostToDevice, st));
        }
        stage_release(e);
    }

    // bins
    if (flags & RTS_OUT_BINS) {
        const uint64_t per_rx = P.powB[sz.depth_total];
        if (per_rx > (~0ull >> 1) / std::max<uint32_t>(1, p->n_rx))
            return rts_fail(RTS_ERR_CAPACITY, "bin key (%u targets+1)^%u x %u receivers does not fit 63 bits", e->n_targets, sz.depth_total, p->n_rx);
        // dense table while it is small; beyond 2^20 bins (the reference groups arbitrary path rows, aggregation.cu:43-57) a
        // hash table whose per-pulse cost follows the number of occupied bins
        e->bins_hashed = e->knobs.hash_bins || per_rx * std::max<uint32_t>(1, p->n_rx) > (1ull << 20);
        e->bins_per_rx = per_rx;
        const uint64_t nb = e->bins_hashed ? (1ull << e->knobs.hash_log2) : per_rx * std::max<uint32_t>(1, p->n_rx);
        if (e->bins_alloc < nb) {
            if (e->d_bin_sums) cudaFree(e->d_bin_sums);
            if (e->d_bin_mins) cudaFree(e->d_bin_mins);
            e->d_bin_sums = nullptr; e->d_bin_mins = nullptr; e->bins_alloc = 0; e->hash_ready = false;
            RTS_CUDA(cudaMalloc(&e->d_bin_sums, sizeof(double) * 5 * nb));
            RTS_CUDA(cudaMalloc(&e->d_bin_mins, sizeof(unsigned long long) * nb));
            e->bins_alloc = nb;
        }
        e->n_bins_dense = p->n_rx ? nb : 0;
        if (e->bins_hashed) {
            int rc = agg_hash_prepare(e, nb);
            if (rc) return rc;
            P.hash_keys = e->d_hash_keys; P.hash_used = e->d_hash_used; P.hash_count = e->d_hash_count;
        } else {
            e->hash_ready = false;   // the dense table shares the arrays; cleared below by agg_pulse_clear (sums 0, slots 0x7f7f…7f:
                                     // positive as int64, so a signed MIN all-reduce keeps it last)
        }
        P.bin_sums = e->d_bin_sums; P.bin_mins = e->d_bin_mins; P.n_bins = e->n_bins_dense;
        // small dense tables are pre-reduced per CTA in shared memory (trace.cu: bins_smem_*); not in the split form of the waves
        P.smem_bins = (!e->bins_hashed && e->n_bins_dense <= RTS_SMEM_BINS && e->knobs.no_split && !e->knobs.no_smem_bins) ? (uint32_t)e->n_bins_dense : 0u;
    } else {
        e->n_bins_dense = 0;
    }

    // records
    const bool records = (flags & RTS_OUT_RECORDS) != 0;
    if (records) {
        int rc = ensure_records(e, sz);
        if (rc) return rc;
        P.results = e->d_results; P.targ_intersect = e->d_targ_intersect; P.rcs_angle = e->d_rcs_angle; P.tri_path = e->d_tri_path;
    }

    // queues: a batch of primaries and up to three live chains per primary with refraction
    const uint64_t batch_max = e->knobs.batch ? (uint64_t)e->knobs.batch : 1ull << 24;   // knob: batching at test sizes
    const uint64_t batch = std::min<uint64_t>(n_primary_total ? n_primary_total : 1, batch_max);
    const uint64_t cap = batch * (rMax ? 3 : 1);
    {
        int rc = trace_alloc_queues(e, cap);
        if (rc) return rc;
    }
    // primary visibility by projection (raster.cuh; RTS_NO_RASTER=1 turns it off): launches whose rays form one image
    // (nx == 1) with a forward image plane; node/triangle counting needs the traversal
    const bool use_raster = p->nx == 1 && !P.single_ray && e->n_tris > 0 && !(flags & RTS_COUNT_NODES) && !e->knobs.no_raster &&
                            P.beamStart[0] > 1e-6 && (p->ny == 1 || P.slope[1] != 0.0) && (p->nz == 1 || P.slope[2] != 0.0) &&
                            n_primary_total > 0;
    P.lat_w = 0; P.lat_c0 = 0; P.lat_g0 = 0;
    if (use_raster) {
        int rc = trace_raster_alloc(e, batch);
        if (rc) return rc;
        if (p->ny % stride == 0) {
            P.lat_w = (uint32_t)(p->ny / stride);
            P.lat_c0 = (uint32_t)(begin % stride);
            P.lat_g0 = (begin - P.lat_c0) / stride;
        }
    }
    P.out_capacity = e->q_capacity;
    P.counters = e->d_counters;
    {   // counters, per-wave counts, the first batch's queue counters, the dense bin table, the emission's accumulators: one launch
        const bool dense = (flags & RTS_OUT_BINS) && !e->bins_hashed && e->n_bins_dense;
        int rc = agg_pulse_clear(e, dense, e->n_bins_dense, p->n_rx);
        if (rc) return rc;
    }

    e->split_timed = false;
    e->follow_timed = false;
    e->followed = false;
    cudaEventRecord(e->ev[0], st);
    if (records) {
        int rc = agg_fill_records(e, sz.ray_total, sz.depth_total, sz.tri_cols);
        if (rc) return rc;
    }
    const uint32_t max_waves = p->max_refl + 1 + (rMax ? 2 : 0); // longest chain: see DESIGN.md (wave count)
    uint64_t waves = 0;
    P.wave_segs = e->d_wave_segs;
    for (int w = 0; w < 32; w++) e->wave_ms[w] = 0.f;
    e->n_waves = std::min<uint32_t>(max_waves, 31);
    const bool single_batch = n_primary_total <= batch;
    // primary-ray tiling (trace.cu: k_wave): with nx == 1 the shard-local index space is a (y,z) plane whose rows
    // hold ny/stride of this shard's rays; whole bands of four rows are walked in 8x4 tiles
    P.swz_w = 0; P.swz_limit = 0;
    // Not when the projected primary wave is on: its batches take their rays in index order, and a batch whose guard
    // trips falls back to k_wave<PRIMARY> — which must then cover exactly that batch's indices, not a tile permutation
    // of them that reaches into the neighbouring batches (found by tests/test_gpu_fuzz.py with RTS_BATCH).
    if (!use_raster && p->nx == 1 && !e->knobs.no_tiles && p->ny % stride == 0 && (p->ny / stride) % 8 == 0 && (p->ny / stride) >= 8) {
        P.swz_w = (uint32_t)(p->ny / stride);
        P.swz_limit = n_primary_total / (4ull * P.swz_w) * (4ull * P.swz_w);
    }
    bool kept = false;
    WaveParams kept_params;
    memset(&kept_params, 0, sizeof(kept_params));
    for (uint64_t done = 0; done < n_primary_total; done += batch) {
        const uint64_t nb = std::min<uint64_t>(batch, n_primary_total - done);
        // d_counts: [0..31] queue counts per wave, [32..63] work counters per wave, [64..95] queue counts from the far end
        if (done) RTS_CUDA(cudaMemsetAsync(e->d_counts, 0, sizeof(unsigned long long) * 96, st));   // (the first batch's: agg_pulse_clear)
        for (uint32_t w = 0; w < max_waves && w < 31; w++) {
            WaveParams Q = P;
            Q.ray_begin = begin; Q.ray_stride = stride; Q.n_primary = nb; Q.batch_base = done;
            Q.in = e->q[w & 1]; Q.out = e->q[(w + 1) & 1];
            Q.in_count = e->d_counts + w;
            Q.out_count = e->d_counts + w + 1;
            if (rMax && !e->knobs.one_ended_queue) { Q.in_back = e->d_counts + 64 + w; Q.out_back = e->d_counts + 64 + w + 1; }
            Q.work_counter = e->d_counts + 32 + w;
            Q.wave_index = w;
            if (single_batch) cudaEventRecord(e->wave_ev[w], st);
            char wave_name[24];
            snprintf(wave_name, sizeof(wave_name), "rts:wave %u", w);
            NvtxRange wave_range(wave_name);
            if (w == 0 && use_raster) {   // projected primary wave; the BVH one behind it runs only if the guard trips
                int rr = trace_launch_raster(e, Q, records, single_batch);
                if (rr) return rr;
                kept = e->coh_on;
                kept_params = Q;
            }
            if (w == 1 && e->knobs.debug_timeline && e->tl_n && e->tl_ev[e->tl_n - 1][5]) cudaEventRecord(e->tl_ev[e->tl_n - 1][5], st);
            if (w == 1 && kept) {         // first reflections served from the kept hits (coherent.cuh)
                Q.w1_static = kept_params.w1_static; Q.hits_static = kept_params.hits_static; Q.moving_flags = kept_params.moving_flags;
                trace_wave_grid(e);
                int rr = trace_launch_kept(e, Q, records);
                if (rr) return rr;
            }
            if (w >= 1 && !e->knobs.no_split && !(Q.rcs_tab || Q.ant_rx)) {   // (the split form has no RTS_TABLES instantiation)   // the two-kernel form for waves that are large (split.cuh); decided on the device
                Q.split_below = e->knobs.split_below;
                int rr = trace_launch_split(e, Q, records);
                if (rr) return rr;
            }
            int rc = trace_launch_wave(e, Q, w == 0, records);
            if (rc) return rc;
            // the shading pass is the last reader of the direction / hit-word buffers, the BVH primary wave behind it (which
            // returns at once unless the guard tripped) the last reader of the footprint passes' control blocks: from here on
            // the next batch's or pulse's direction pass and static footprints may run on side_dirs
            if (w == 0 && use_raster && e->side_dirs) { cudaEventRecord(e->ev_dirs_free, st); e->dirs_free_valid = true; }
            waves++;
        }
        if (single_batch) cudaEventRecord(e->wave_ev[e->n_waves], st);
        if (e->knobs.debug_timeline && e->tl_n && e->tl_ev[e->tl_n - 1][6]) cudaEventRecord(e->tl_ev[e->tl_n - 1][6], st);
    }
    cudaEventRecord(e->ev[1], st);
    // read-back of the counters into pinned memory; folded into the stats by pulse_collect()
    RTS_CUDA(cudaMemcpyAsync(&e->h_rb->counters, e->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    RTS_CUDA(cudaMemcpyAsync(e->h_rb->wave_segs, e->d_wave_segs, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost, st));
    e->pulse_raster = use_raster;
    if (use_raster) {
        RTS_CUDA(cudaMemcpyAsync(&e->h_rb->raster, e->d_raster_ctl, sizeof(RasterCtl), cudaMemcpyDeviceToHost, st));
        e->h_rb->raster_static.area = 0;
        if (e->static_valid || e->split_static)
            RTS_CUDA(cudaMemcpyAsync(&e->h_rb->raster_static, e->d_raster_ctl_static, sizeof(RasterCtl), cudaMemcpyDeviceToHost, st));
    }
    e->pulse_pending = true;
    e->pulse_single_batch = single_batch && n_primary_total;
    e->pulse_primary = n_primary_total; e->pulse_waves = waves;
    e->have_pulse = true; e->last_flags = flags; e->last_sizes = sz;
    e->last_B = (uint32_t)B; e->last_D = sz.depth_total; e->last_nrx = p->n_rx;
    e->last_begin = begin; e->last_stride = stride; e->last_n_primary = n_primary_total;
    e->bins_finalised = !(flags & RTS_NO_FINALISE);
    e->prev_bins_eager = e->bins_eager; e->prev_bins_slot = e->bins_slot;   // what rts_get_bins_previous will read
    e->bins_eager = false;
    if ((flags & RTS_OUT_BINS) && e->bins_finalised) {
        int rc = agg_emit_bins_async(e);
        if (rc) return rc;
    }
    if (flags & RTS_ASYNC) return RTS_OK;
    return pulse_collect(e);
}

// Wait for the pulse in flight and fold its read-back into e->stats / the wave profile.
int pulse_collect(rts_engine *e)
{
    NvtxRange nvtx_range("rts:pulse_collect (wait)");
    if (!e->pulse_pending) return RTS_OK;
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    e->pulse_pending = false;
    const Counters c = e->h_rb->counters;
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
    memcpy(e->wave_segs, e->h_rb->wave_segs, sizeof(e->wave_segs));
    for (int w = 0; w < 32; w++) e->wave_ms[w] = 0.f;
    if (e->pulse_single_batch)
        for (uint32_t w = 0; w < e->n_waves; w++) cudaEventElapsedTime(&e->wave_ms[w], e->wave_ev[w], e->wave_ev[w + 1]);
    e->split_ms[0] = e->split_ms[1] = 0.f;
    if (e->split_timed) {
        cudaEventElapsedTime(&e->split_ms[0], e->split_ev[0], e->split_ev[1]);
        cudaEventElapsedTime(&e->split_ms[1], e->split_ev[1], e->split_ev[2]);
    }
    e->follow_ms = 0.f;
    if (e->follow_timed) cudaEventElapsedTime(&e->follow_ms, e->follow_ev[0], e->follow_ev[1]);
    collect_refit_time(e);
    rts_stats &s = e->stats;
    memset(&s, 0, sizeof(s));
    s.primary_rays = e->pulse_primary; s.segments = c.segments; s.hits = c.hits; s.shaded_hits = c.shaded;
    s.captured = c.captured; s.multi_captured = c.multi; s.edge_rays = c.edge; s.refracted = c.refracted;
    s.nodes_visited = c.nodes; s.tris_tested = c.tris; s.waves = e->pulse_waves; s.kept_reflections = c.kept;
    s.ms_trace = ms; s.ms_update = e->bvh_info.ms_refit; s.ms_total = ms;
    // the guard of raster.cuh, evaluated on the last batch's control block (16 candidates per ray)
    s.primary_projected = (e->pulse_raster && e->h_rb->raster.area + e->h_rb->raster_static.area <= 16ull * std::min<uint64_t>(e->pulse_primary, 1ull << 24)) ? 1u : 0u;
    if (e->knobs.debug_raster && e->pulse_raster)
        fprintf(stderr, "[raster] candidates %llu (+ %llu kept from the static pass), %u row chunks, projected %u\n", e->h_rb->raster.area, e->h_rb->raster_static.area, e->h_rb->raster.n_items, s.primary_projected);
    if (c.overflow) return rts_fail(RTS_ERR_CAPACITY, "%llu ray states dropped (queue/stack overflow)", (unsigned long long)c.overflow);
    if (e->comm.check_timeout) {
        e->comm.check_timeout = false;
        if (e->h_rb->comm_timed_out) return rts_fail(RTS_ERR_STATE, "peer-memory bin exchange: a rank did not publish its bins within 2 s (every rank must call rts_comm_allreduce_bins once per pulse)");
    }
    return RTS_OK;
}

// ---- tabulated callbacks ------------------------------------------------------------------------
static int check_table(const rts_table2d &t, const char *what, uint32_t k)
{
    if (t.n_az == 0) return RTS_OK;
    if (!t.values || t.n_el == 0 || !(t.az_step > 0) || !(t.el_step > 0) || (uint64_t)t.n_az * t.n_el > (1ull << 26))
        return rts_fail(RTS_ERR_ARG, "%s table %u: needs values, n_el > 0, positive steps and at most 2^26 samples", what, k);
    return RTS_OK;
}
static DevTable dev_table(const rts_table2d &t, const double *values)
{
    DevTable d;
    memset(&d, 0, sizeof(d));
    d.n_az = t.n_az; d.n_el = t.n_az ? t.n_el : 0;
    d.az0 = t.az0; d.az_step = t.az_step; d.el0 = t.el0; d.el_step = t.el_step;
    d.values = t.n_az ? values : nullptr;
    return d;
}

extern "C" int rts_set_rcs_tables(rts_engine *e, const rts_table2d *tables, uint32_t n_targets)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    RTS_CUDA(cudaSetDevice(e->device));
    RTS_CUDA(cudaStreamSynchronize(e->stream));      // a pulse in flight may still read the old tables
    if (e->d_rcs_tab) { cudaFree(e->d_rcs_tab); e->d_rcs_tab = nullptr; }
    if (e->d_rcs_values) { cudaFree(e->d_rcs_values); e->d_rcs_values = nullptr; }
    e->n_rcs_tab = 0;
    if (!tables || !n_targets) return RTS_OK;
    size_t total = 0;
    for (uint32_t k = 0; k < n_targets; k++) {
        int rc = check_table(tables[k], "RCS", k);
        if (rc) return rc;
        total += (size_t)tables[k].n_az * tables[k].n_el;
    }
    RTS_CUDA(cudaMalloc(&e->d_rcs_values, sizeof(double) * std::max<size_t>(total, 1)));
    RTS_CUDA(cudaMalloc(&e->d_rcs_tab, sizeof(DevTable) * n_targets));
    std::vector<DevTable> dev(n_targets);
    size_t at = 0;
    for (uint32_t k = 0; k < n_targets; k++) {
        const size_t n = (size_t)tables[k].n_az * tables[k].n_el;
        dev[k] = dev_table(tables[k], e->d_rcs_values + at);
        if (n) RTS_CUDA(cudaMemcpy(e->d_rcs_values + at, tables[k].values, sizeof(double) * n, cudaMemcpyHostToDevice));
        at += n;
    }
    RTS_CUDA(cudaMemcpy(e->d_rcs_tab, dev.data(), sizeof(DevTable) * n_targets, cudaMemcpyHostToDevice));
    e->n_rcs_tab = n_targets;
    return RTS_OK;
}

extern "C" int rts_set_antennas(rts_engine *e, const rts_antenna *tx, const rts_antenna *rx, uint32_t n_rx)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    RTS_CUDA(cudaSetDevice(e->device));
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    for (void *p : {(void *)e->d_ant_tx, (void *)e->d_ant_rx, (void *)e->d_ant_values}) if (p) cudaFree(p);
    e->d_ant_tx = nullptr; e->d_ant_rx = nullptr; e->d_ant_values = nullptr; e->n_ant_rx = 0;
    if (!rx || !n_rx) {
        if (tx) return rts_fail(RTS_ERR_ARG, "a transmitter antenna needs the receivers' antennas too (their positions orient the direct ray, ray_tracer.cpp:1206)");
        return RTS_OK;
    }
    if (n_rx > RTS_MAX_RX) return rts_fail(RTS_ERR_CAPACITY, "%u receivers > %u", n_rx, RTS_MAX_RX);
    size_t total = 0;
    if (tx) { int rc = check_table(tx->gain, "transmitter gain", 0); if (rc) return rc; total += (size_t)tx->gain.n_az * tx->gain.n_el; }
    for (uint32_t j = 0; j < n_rx; j++) {
        int rc = check_table(rx[j].gain, "receiver gain", j);
        if (rc) return rc;
        total += (size_t)rx[j].gain.n_az * rx[j].gain.n_el;
    }
    RTS_CUDA(cudaMalloc(&e->d_ant_values, sizeof(double) * std::max<size_t>(total, 1)));
    size_t at = 0;
    auto make = [&](const rts_antenna &a, DevAntenna &d) -> int {
        const size_t n = (size_t)a.gain.n_az * a.gain.n_el;
        memset(&d, 0, sizeof(d));
        d.gain = dev_table(a.gain, e->d_ant_values + at);
        d.bore_az = a.bore_az; d.bore_el = a.bore_el; d.rate_az = a.rate_az; d.rate_el = a.rate_el;
        memcpy(d.pos, a.position, sizeof(d.pos));
        if (n) RTS_CUDA(cudaMemcpy(e->d_ant_values + at, a.gain.values, sizeof(double) * n, cudaMemcpyHostToDevice));
        at += n;
        return RTS_OK;
    };
    if (tx) {
        DevAntenna d;
        int rc = make(*tx, d);
        if (rc) return rc;
        RTS_CUDA(cudaMalloc(&e->d_ant_tx, sizeof(DevAntenna)));
        RTS_CUDA(cudaMemcpy(e->d_ant_tx, &d, sizeof(d), cudaMemcpyHostToDevice));
    }
    std::vector<DevAntenna> dev(n_rx);
    for (uint32_t j = 0; j < n_rx; j++) { int rc = make(rx[j], dev[j]); if (rc) return rc; }
    RTS_CUDA(cudaMalloc(&e->d_ant_rx, sizeof(DevAntenna) * n_rx));
    RTS_CUDA(cudaMemcpy(e->d_ant_rx, dev.data(), sizeof(DevAntenna) * n_rx, cudaMemcpyHostToDevice));
    e->n_ant_rx = n_rx;
    return RTS_OK;
}

extern "C" int rts_sync(rts_engine *e)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    RTS_CUDA(cudaSetDevice(e->device));
    int rc = RTS_OK;
    if (e->pulse_pending) rc = pulse_collect(e);
    else RTS_CUDA(cudaStreamSynchronize(e->stream));
    if (e->knobs.debug_timeline && e->tl_n) {   // when each of the last pulses' direction pass / footprints / shading pass ran
        if (e->side_dirs) cudaStreamSynchronize(e->side_dirs);
        for (int i = 0; i < e->tl_n; i++) {
            float t[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 7; k++) cudaEventElapsedTime(&t[k], e->tl_ev[0][0], e->tl_ev[i][k]);
            fprintf(stderr, "[timeline] pulse %2d: dirs %8.3f .. %8.3f   footprints from %8.3f   shading pass %8.3f .. %8.3f   later waves %8.3f .. %8.3f ms\n", i, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
        }
        e->tl_n = 0;
    }
    return rc;
}

extern "C" int rts_get_stats(rts_engine *e, rts_stats *out)
{
    if (!e || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    int rc = pulse_collect(e);
    *out = e->stats;
    return rc;
}

extern "C" int rts_get_wave_profile(rts_engine *e, uint32_t cap, float *ms, uint64_t *segments, uint32_t *n)
{
    if (!e || !n) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->have_pulse) return rts_fail(RTS_ERR_STATE, "no pulse traced yet");
    pulse_collect(e);
    *n = e->n_waves;
    for (uint32_t w = 0; w < e->n_waves && w < cap; w++) {
        if (ms) ms[w] = e->wave_ms[w];
        if (segments) segments[w] = e->wave_segs[w];
    }
    return RTS_OK;
}

extern "C" int rts_get_split_profile(rts_engine *e, float ms[2])
{
    if (!e || !ms) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->have_pulse) return rts_fail(RTS_ERR_STATE, "no pulse traced yet");
    pulse_collect(e);
    ms[0] = e->split_ms[0]; ms[1] = e->split_ms[1];
    return RTS_OK;
}

extern "C" int rts_get_follow_profile(rts_engine *e, float *ms)
{
    if (!e || !ms) return rts_fail(RTS_ERR_ARG, "NULL argument");
    if (!e->have_pulse) return rts_fail(RTS_ERR_STATE, "no pulse traced yet");
    pulse_collect(e);
    *ms = e->follow_ms;
    return RTS_OK;
}

extern "C" int rts_kernel_launches(rts_engine *e, uint64_t *out)
{
    if (!e || !out) return rts_fail(RTS_ERR_ARG, "NULL argument");
    *out = e->launches;
    return RTS_OK;
}

// Read bandwidth of a buffer that stays resident in L2: every CTA streams the whole buffer `reps` times with 128-bit
// loads (the access width of the node and triangle fetches), so after the first pass all sectors are L2 hits.
__global__ void __launch_bounds__(512) k_probe_read(const uint4 *__restrict__ buf, size_t n16, uint32_t reps, uint32_t *sink)
{
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (uint32_t r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n16; i += 4 * stride) {
            uint4 a, b, c, d;    // ld.global.cg: cached in L2 only, so the figure is L2 -> SM bandwidth, not L1 hits
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(buf + i));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(buf + i + stride));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(buf + i + 2 * stride));
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(buf + i + 3 * stride));
            acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
        }
    if (acc == 0x9e3779b9u) *sink = acc;     // keeps the loads alive
}

extern "C" int rts_probe_read_bandwidth(rts_engine *e, uint64_t bytes, uint32_t reps, double *gbs)
{
    if (!e || !gbs || bytes < (1u << 20) || reps == 0) return rts_fail(RTS_ERR_ARG, "bad argument");
    RTS_CUDA(cudaSetDevice(e->device));
    int sms = 0;
    RTS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
    const unsigned grid = (unsigned)sms * 4u, block = 512u;
    const size_t quantum = (size_t)grid * block * 4u * 16u;        // whole unrolled iterations only
    const size_t n = std::max<size_t>(1, bytes / quantum) * quantum;
    void *buf = nullptr;
    uint32_t *sink = nullptr;
    RTS_CUDA(cudaMalloc(&buf, n + quantum));
    RTS_CUDA(cudaMalloc(&sink, 4));
    RTS_CUDA(cudaMemsetAsync(buf, 1, n + quantum, e->stream));
    cudaEvent_t a, b;
    RTS_CUDA(cudaEventCreate(&a));
    RTS_CUDA(cudaEventCreate(&b));
    k_probe_read<<<grid, block, 0, e->stream>>>((const uint4 *)buf, n / 16 + 1, 2, sink);      // warm L2
    RTS_CUDA(cudaEventRecord(a, e->stream));
    k_probe_read<<<grid, block, 0, e->stream>>>((const uint4 *)buf, n / 16 + 1, reps, sink);
    RTS_CUDA(cudaEventRecord(b, e->stream));
    e->launches += 2;
    RTS_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    RTS_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    *gbs = (double)n * reps / (ms * 1e-3) / 1e9;
    return RTS_OK;
}

extern "C" int rts_get_bins(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n)
{
    NvtxRange nvtx_range("rts:get_bins");
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    RTS_CUDA(cudaSetDevice(e->device));
    int rc = pulse_collect(e);
    if (rc) return rc;
    return agg_collect_bins(e, out, cap, n);
}

extern "C" int rts_get_bins_previous(rts_engine *e, rts_bin *out, uint32_t cap, uint32_t *n)
{
    NvtxRange nvtx_range("rts:get_bins_previous");
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    RTS_CUDA(cudaSetDevice(e->device));
    return agg_collect_bins_previous(e, out, cap, n);
}

// ray_tracer.cpp:1289-1320 on the fused bins.  d_pathMatch of a ray = smallest received-list index among the
// rays it summed (aggregation.cu:68-69); the received list is in result-slot order (ray_tracer.cpp:1190), so
// slot order == list order.  Non-direct bin: its members share min_slot, which is one of them -> one response
// with the bin's values.  Direct bin: its members' d_pathMatch is the receiver-wide minimum; that ray is a
// member of the direct bin only when own_min_slot == min_slot, otherwise the value coincides with another
// bin's representative and unique() drops it.
extern "C" int rts_get_responses(rts_engine *e, rts_response *out, uint32_t cap, uint32_t *n)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    if (!e->bins_finalised) return rts_fail(RTS_ERR_STATE, "bins are not finalised (RTS_NO_FINALISE without rts_finalise_bins)");
    RTS_CUDA(cudaSetDevice(e->device));
    uint32_t nb = 0;
    int rc = pulse_collect(e);
    if (rc) return rc;
    rc = agg_collect_bins(e, nullptr, 0, &nb);
    if (rc) return rc;
    std::vector<rts_bin> bins(nb ? nb : 1);
    if (nb && (rc = agg_collect_bins(e, bins.data(), nb, &nb))) return rc;
    std::vector<rts_response> r;
    r.reserve(nb);
    for (uint32_t i = 0; i < nb; i++) {
        const rts_bin &b = bins[i];
        if (b.direct && b.own_min_slot != b.min_slot) continue;
        rts_response x;
        x.rx = b.rx; x._pad = 0; x.slot = b.min_slot;
        x.power = b.power; x.delay = b.delay; x.doppler = b.doppler; x.phase = b.phase;
        r.push_back(x);
    }
    std::sort(r.begin(), r.end(), [](const rts_response &a, const rts_response &b) { return a.slot < b.slot; });
    if (n) *n = (uint32_t)r.size();
    if (out) for (size_t i = 0; i < r.size() && i < cap; i++) out[i] = r[i];
    return RTS_OK;
}

extern "C" int rts_bins_device(rts_engine *e, void **sums, uint64_t *n_sum, void **mins, uint64_t *n_mins)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    if (e->bins_hashed) return rts_fail(RTS_ERR_STATE, "the last pulse used the sparse bin table: exchange it with rts_bins_compact_device / rts_bins_load_compact");
    if (sums) *sums = e->d_bin_sums;
    if (n_sum) *n_sum = e->n_bins_dense * 5;
    if (mins) *mins = e->d_bin_mins;
    if (n_mins) *n_mins = e->n_bins_dense;
    return RTS_OK;
}

extern "C" int rts_bins_compact_device(rts_engine *e, void **keys_device, void **sums_device, void **mins_device, uint32_t *n)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    if (!e->bins_hashed) return rts_fail(RTS_ERR_STATE, "the last pulse used the dense bin table: rts_bins_device");
    RTS_CUDA(cudaSetDevice(e->device));
    return agg_hash_compact(e, keys_device, sums_device, mins_device, n);
}

extern "C" int rts_bins_load_compact(rts_engine *e, const void *keys_device, const void *sums_device, const void *mins_device, uint32_t n)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS) || !e->bins_hashed) return rts_fail(RTS_ERR_STATE, "last pulse did not use the sparse bin table");
    if (n && (!keys_device || !sums_device || !mins_device)) return rts_fail(RTS_ERR_ARG, "NULL array");
    RTS_CUDA(cudaSetDevice(e->device));
    return agg_hash_load(e, keys_device, sums_device, mins_device, n);
}

extern "C" int rts_finalise_bins(rts_engine *e)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_BINS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce bins");
    e->bins_finalised = true; // myKernel2 is applied when bins are emitted
    RTS_CUDA(cudaSetDevice(e->device));
    return agg_emit_bins_async(e);   // behind the caller's reduction, which runs on the engine's stream
}

extern "C" int rts_get_records(rts_engine *e, rts_ray_record *results, int32_t *targ_intersect, double *rcs_angle,
                               int32_t *tri_path)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_RECORDS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce records");
    RTS_CUDA(cudaSetDevice(e->device));
    {
        int rc = pulse_collect(e);
        if (rc) return rc;
    }
    RTS_CUDA(cudaStreamSynchronize(e->stream));
    const rts_sizes &sz = e->last_sizes;
    const size_t n = sz.ray_total, D = sz.depth_total, W = sz.tri_cols;
    if (results) RTS_CUDA(cudaMemcpy(results, e->d_results, sizeof(rts_ray_record) * n, cudaMemcpyDeviceToHost));
    if (targ_intersect && D) RTS_CUDA(cudaMemcpy(targ_intersect, e->d_targ_intersect, sizeof(int32_t) * n * D, cudaMemcpyDeviceToHost));
    if (rcs_angle && D) RTS_CUDA(cudaMemcpy(rcs_angle, e->d_rcs_angle, sizeof(double) * 2 * n * D, cudaMemcpyDeviceToHost));
    if (tri_path) RTS_CUDA(cudaMemcpy(tri_path, e->d_tri_path, sizeof(int32_t) * n * W, cudaMemcpyDeviceToHost));
    return RTS_OK;
}

extern "C" int rts_get_records_shard(rts_engine *e, uint64_t *n_shard, rts_ray_record *results, int32_t *targ_intersect,
                                     double *rcs_angle, int32_t *tri_path)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_RECORDS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce records");
    RTS_CUDA(cudaSetDevice(e->device));
    int rc = pulse_collect(e);
    if (rc) return rc;
    return agg_get_records_shard(e, n_shard, results, targ_intersect, rcs_angle, tri_path);
}

extern "C" int rts_get_received(rts_engine *e, uint64_t cap, uint64_t *n, uint64_t *slots, rts_ray_record *results,
                                int32_t *targ_intersect, double *rcs_angle)
{
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (!e->have_pulse || !(e->last_flags & RTS_OUT_RECORDS)) return rts_fail(RTS_ERR_STATE, "last pulse did not produce records");
    RTS_CUDA(cudaSetDevice(e->device));
    int rc = pulse_collect(e);
    if (rc) return rc;
    return agg_get_received(e, cap, n, slots, results, targ_intersect, rcs_angle);
}

extern "C" int rts_aggregate(rts_engine *e, rts_ray_record *rx_results, const int32_t *rx_intersects, uint32_t received,
                             uint32_t depth_total, double cspeed, double carrier, double *npath, double *power,
                             double *doppler, double *delay, double *phase, int32_t *path_match)
{
    NvtxRange nvtx_range("rts:aggregate (kernel_wrapper)");
    if (!e) return rts_fail(RTS_ERR_ARG, "engine is NULL");
    if (received && (!rx_results || !npath || !power || !doppler || !delay || !phase || !path_match || (depth_total && !rx_intersects)))
        return rts_fail(RTS_ERR_ARG, "NULL array");
    RTS_CUDA(cudaSetDevice(e->device));
    return agg_kernel_wrapper(e, rx_results, rx_intersects, received, depth_total, cspeed, carrier, npath, power, doppler,
                              delay, phase, path_match);
}

// ---------------------------------------------------------------------------------------------
// The reference's C++ entry point, same symbol and signature (aggregation.cuh:18-23), so a host
// simulator that links rs::kernel_wrapper keeps linking.  PerRayData here is the layout twin
// declared in include/rts_prd.h.  MaxThreads / MaxBlocks were launch-shape hints and are ignored.
// Errors abort like the reference's cudaCheckErrors (aggregation.cu:17-27) because the signature
// has no way to report them.
#include "../../include/rts_prd.h"
namespace rs {
void kernel_wrapper(PerRayData *h_rx_results_arr, int *h_rx_intersects_arr, unsigned int receivedRays,
                    unsigned int depthTotal, unsigned int MaxThreads, unsigned int MaxBlocks, double cspeed, double carrier,
                    double *h_npath_arr, double *h_power_arr, double *h_doppler_arr, double *h_delay_arr,
                    double *h_phase_arr, int *h_pathMatch)
{
    (void)MaxThreads; (void)MaxBlocks;
    // one engine per calling thread and device, created on first use and released when the thread ends
    struct Holder {
        rts_engine *eng = nullptr;
        int dev = -1;
        ~Holder() { if (eng) rts_destroy(eng); }
    };
    static thread_local Holder h;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!h.eng || h.dev != dev) {
        if (h.eng) { rts_destroy(h.eng); h.eng = nullptr; }
        if (rts_create(dev, &h.eng) != RTS_OK) {
            fprintf(stderr, "Fatal error: rs::kernel_wrapper: %s\n*** FAILED - ABORTING\n\n", rts_last_error());
            exit(1);
        }
        h.dev = dev;
    }
    rts_engine *eng = h.eng;
    if (rts_aggregate(eng, reinterpret_cast<rts_ray_record *>(h_rx_results_arr), h_rx_intersects_arr, receivedRays, depthTotal,
                      cspeed, carrier, h_npath_arr, h_power_arr, h_doppler_arr, h_delay_arr, h_phase_arr, h_pathMatch) != RTS_OK) {
        fprintf(stderr, "Fatal error: rs::kernel_wrapper: %s\n*** FAILED - ABORTING\n\n", rts_last_error());
        exit(1);
    }
}
} // namespace rs
