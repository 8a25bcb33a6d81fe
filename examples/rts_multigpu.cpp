// rts_multigpu.cpp — a C++ host driving librts_b200 on N GPUs of one node through the C-ABI only
// (include/rts_b200.h), with the receiver bins reduced by NCCL: the shape of the multi-GPU path that
// BASELINE.json's north star describes (scene and BVH replicated on every GPU, primary rays sharded,
// bins all-reduced over NVLink), without Python in the loop.  rts_b200/dist.py + bench.py are the
// torch.distributed form of the same exchange.
//
//   one host thread per GPU  ->  rts_create(g) · rts_scene_set_targets · own cudaStream (rts_set_stream)
//   per pulse                ->  rts_scene_set_poses · rts_trace_pulse(BINS | NO_FINALISE | ASYNC) on rays g, g+N, ...
//                                rts_comm_allreduce_bins: the library's own exchange over peer memory (default), or
//                                (--exchange nccl) ncclAllReduce(SUM, double) over bins[n][5] + ncclAllReduce(MIN, uint64)
//                                over the slots, then rts_finalise_bins; rts_get_bins
//   check                    ->  a second engine on GPU 0 traces the whole launch un-sharded; the reduced bins must
//                                have the same keys, counts and representative slots, and sums within 1e-9
//
// The reference has no multi-GPU path; what is sharded is the launch index space of
// rtContextLaunch3D (ray_tracer.cpp:1165), what is reduced are the accumulators of myKernel1 (aggregation.cu:56-69).
//
// Build: make -C examples        Run: examples/rts_multigpu [--gpus N] [--grid 2048] [--pulses 8] [--cells 256]
#include <cuda_runtime_api.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/rts_b200.h"

namespace {

struct Mesh {
    std::vector<double> verts, normals;
    std::vector<uint32_t> tris;
    double refl = 0.9, refr = 1.0;
    rts_target_mesh view() const
    {
        rts_target_mesh m;
        std::memset(&m, 0, sizeof(m));
        m.n_verts = (uint32_t)(verts.size() / 3); m.n_tris = (uint32_t)(tris.size() / 3); m.n_normals = (uint32_t)(normals.size() / 3);
        m.verts = verts.data(); m.tris = tris.data(); m.normals = normals.data();
        m.refl_coeff = refl; m.refr_index = refr;
        return m;
    }
};

// Height field over [0, lx] x [-ly/2, ly/2]: two triangles per cell, per-face normals (n_normals = n_tris > n_verts
// is the reference's "file mesh" convention, triangle_mesh.cu:180).
Mesh terrain(int cx, int cy, double lx, double ly)
{
    Mesh m;
    auto h = [](double x, double y) { return 12.0 * std::sin(x * 0.011) * std::cos(y * 0.017) + 5.0 * std::sin(x * 0.047 + 1.3) + 3.0 * std::cos(y * 0.061); };
    for (int j = 0; j <= cy; j++)
        for (int i = 0; i <= cx; i++) {
            const double x = lx * i / cx, y = -ly / 2 + ly * j / cy;
            m.verts.insert(m.verts.end(), {x, y, h(x, y)});
        }
    auto vid = [&](int i, int j) { return (uint32_t)(j * (cx + 1) + i); };
    for (int j = 0; j < cy; j++)
        for (int i = 0; i < cx; i++) {
            const uint32_t a = vid(i, j), b = vid(i + 1, j), c = vid(i + 1, j + 1), d = vid(i, j + 1);
            m.tris.insert(m.tris.end(), {a, b, c, a, c, d});
        }
    for (size_t t = 0; t < m.tris.size() / 3; t++) {
        const double *p0 = &m.verts[3 * m.tris[3 * t]], *p1 = &m.verts[3 * m.tris[3 * t + 1]], *p2 = &m.verts[3 * m.tris[3 * t + 2]];
        const double u[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, v[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const double l = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        m.normals.insert(m.normals.end(), {n[0] / l, n[1] / l, n[2] / l});
    }
    m.refl = 0.7;
    return m;
}

template <class F> Mesh generated(F call)
{
    Mesh m;
    uint32_t nv = 0, nt = 0, nn = 0;
    call(nullptr, &nv, nullptr, &nt, nullptr, &nn);
    m.verts.resize(3 * (size_t)nv); m.tris.resize(3 * (size_t)nt); m.normals.resize(3 * (size_t)nn);
    call(m.verts.data(), &nv, m.tris.data(), &nt, m.normals.data(), &nn);
    return m;
}

struct Barrier {       // C++17 stand-in for std::barrier
    std::mutex mu; std::condition_variable cv; int n, waiting = 0; unsigned long gen = 0;
    explicit Barrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> l(mu);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};

#define CHECK_RTS(x) do { int rc_ = (x); if (rc_ != RTS_OK) { std::fprintf(stderr, "%s:%d %s -> %d: %s\n", __FILE__, __LINE__, #x, rc_, rts_last_error()); std::exit(2); } } while (0)
#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); std::exit(2); } } while (0)
#define CHECK_NCCL(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { std::fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, ncclGetErrorString(r_)); std::exit(2); } } while (0)

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<rts_target_mesh> views;
    std::vector<double> pos0, vel;          // per target, 3 each
    std::vector<rts_rx_sphere> rx;
    rts_pulse pulse;
    std::vector<rts_pose> poses(int p, double dt) const
    {
        std::vector<rts_pose> out(meshes.size());
        for (size_t k = 0; k < meshes.size(); k++) {
            std::memset(&out[k], 0, sizeof(rts_pose));
            out[k].R[0] = out[k].R[4] = out[k].R[8] = 1.0;
            for (int a = 0; a < 3; a++) out[k].t[a] = pos0[3 * k + a] + vel[3 * k + a] * dt * p;
        }
        return out;
    }
};

} // namespace

int main(int argc, char **argv)
{
    int n_gpus = 0, grid = 2048, pulses = 8, cells = 256, n_rx = 4, movers = 8;
    bool peer = true;      // bins reduced by the library's peer-memory kernels (rts_comm_*); --exchange nccl: two NCCL all-reduces
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        const int v = std::atoi(argv[i + 1]);
        if (a == "--exchange") { peer = std::string(argv[i + 1]) != "nccl"; continue; }
        if (a == "--gpus") n_gpus = v; else if (a == "--grid") grid = v; else if (a == "--pulses") pulses = v;
        else if (a == "--cells") cells = v; else if (a == "--rx") n_rx = v; else if (a == "--movers") movers = v;
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    int visible = 0;
    CHECK_CUDA(cudaGetDeviceCount(&visible));
    if (n_gpus <= 0 || n_gpus > visible) n_gpus = visible;
    if (n_gpus < 1) { std::fprintf(stderr, "no CUDA device (there is no CPU fallback)\n"); return 2; }

    // ---- scene: terrain + moving spheres and boxes, Tx looking down at a grazing angle, receivers behind the patch
    Scene S;
    const double lx = 2000.0, ly = 1000.0;
    S.meshes.push_back(terrain(cells, cells / 2, lx, ly));
    for (int k = 0; k < movers; k++) {
        if (k % 2) S.meshes.push_back(generated([&](double *v, uint32_t *nv, uint32_t *t, uint32_t *nt, double *n, uint32_t *nn) { rts_rect_mesh(8.f, 3.f, 3.f, 0.3f * k, 0.f, 0.f, v, nv, t, nt, n, nn); }));
        else S.meshes.push_back(generated([&](double *v, uint32_t *nv, uint32_t *t, uint32_t *nt, double *n, uint32_t *nn) { rts_sphere_mesh(3, 2.5f, 0.f, 0.f, 0.f, v, nv, t, nt, n, nn); }));
    }
    const size_t K = S.meshes.size();
    S.pos0.assign(3 * K, 0.0); S.vel.assign(3 * K, 0.0);
    for (size_t k = 1; k < K; k++) {
        S.pos0[3 * k] = lx * (0.25 + 0.5 * ((k * 37) % 100) / 100.0); S.pos0[3 * k + 1] = ly * (-0.3 + 0.6 * ((k * 61) % 100) / 100.0); S.pos0[3 * k + 2] = 40.0 + 5.0 * k;
        S.vel[3 * k] = 30.0 - 7.0 * k; S.vel[3 * k + 1] = 11.0 * ((k % 3) - 1.0); S.vel[3 * k + 2] = 0.5 * k;
    }
    for (const Mesh &m : S.meshes) S.views.push_back(m.view());
    const double tx[3] = {-1500.0, 0.0, 500.0};
    for (int j = 0; j < n_rx; j++) {
        rts_rx_desc d;
        const double ang = (j - (n_rx - 1) / 2.0) * 0.15;
        d.position[0] = lx + 3000.0 * std::cos(ang); d.position[1] = 3000.0 * std::sin(ang); d.position[2] = 900.0 + 150.0 * j;
        d.azimuth = M_PI + ang; d.elevation = 0.0; d.radius = 200.0; d.theta_span = 2.0; d.phi_span = 2.0;
        rts_rx_sphere s;
        rts_rx_sphere_from_desc(&d, &s);
        S.rx.push_back(s);
    }
    rts_pulse &P = S.pulse;
    std::memset(&P, 0, sizeof(P));
    P.nx = 1; P.ny = (uint32_t)grid; P.nz = (uint32_t)grid; P.max_refl = 3; P.max_refr = 0; P.interpolate_smooth = 0;
    std::memcpy(P.tx_origin, tx, sizeof(tx));
    P.tx_dir[0] = 0.0; P.tx_dir[1] = std::atan2(-tx[2], lx / 2 - tx[0]);
    P.tx_span[0] = 0.40; P.tx_span[1] = 0.16; P.tx_span[2] = 0.0;
    P.cspeed = 299792458.0; P.carrier = 10e9;
    P.n_rx = (uint32_t)S.rx.size(); P.rx = S.rx.data();
    P.n_targets = (uint32_t)K; P.targ_vel = S.vel.data();
    P.gain_tx = 1.0; P.gain_rx = 1.0;
    const double dt = 1e-3;

    // ---- one engine, stream and communicator per GPU
    std::vector<ncclComm_t> comms(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int g = 0; g < n_gpus; g++) devs[g] = g;
    CHECK_NCCL(ncclCommInitAll(comms.data(), n_gpus, devs.data()));
    Barrier bar(n_gpus);
    std::vector<void *> xchg(n_gpus, nullptr);             // every rank's exchange block (peer-memory exchange)
    std::vector<std::vector<rts_bin>> reduced(pulses);     // rank 0's view of every pulse
    std::vector<float> ms_pulse(pulses, 0.f);
    uint64_t launches = 0;

    auto worker = [&](int g) {
        CHECK_CUDA(cudaSetDevice(g));
        cudaStream_t st;
        CHECK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        rts_engine *e = nullptr;
        CHECK_RTS(rts_create(g, &e));
        CHECK_RTS(rts_set_stream(e, st));
        CHECK_RTS(rts_scene_set_targets(e, S.views.data(), (uint32_t)K));
        cudaEvent_t e0, e1;
        CHECK_CUDA(cudaEventCreate(&e0)); CHECK_CUDA(cudaEventCreate(&e1));
        if (peer) {
            // one exchange block per GPU, mapped by every other GPU of this process: peer access + raw pointers
            CHECK_RTS(rts_comm_create(e, (uint32_t)g, (uint32_t)n_gpus, 1u << 16));
            CHECK_RTS(rts_comm_local_ptr(e, &xchg[g]));
            for (int q = 0; q < n_gpus; q++)
                if (q != g) {
                    cudaError_t pe = cudaDeviceEnablePeerAccess(q, 0);
                    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { std::fprintf(stderr, "GPU %d cannot map GPU %d: %s\n", g, q, cudaGetErrorString(pe)); std::exit(2); }
                    cudaGetLastError();
                }
            bar.wait();
            CHECK_RTS(rts_comm_connect_ptrs(e, xchg.data()));
            bar.wait();
        }
        rts_pulse mine = S.pulse;
        mine.ray_begin = (uint64_t)g; mine.ray_count = 0; mine.ray_stride = (uint64_t)n_gpus;   // rays g, g+N, g+2N, ...
        for (int p = -1; p < pulses; p++) {                 // p = -1: warm-up
            const int pp = p < 0 ? 0 : p;
            const std::vector<rts_pose> poses = S.poses(pp, dt);
            bar.wait();
            CHECK_CUDA(cudaEventRecord(e0, st));
            CHECK_RTS(rts_scene_set_poses(e, poses.data(), (uint32_t)K));
            CHECK_RTS(rts_trace_pulse(e, &mine, RTS_OUT_BINS | RTS_NO_FINALISE | RTS_ASYNC | RTS_NO_REUSE));
            if (peer) {
                CHECK_RTS(rts_comm_allreduce_bins(e));         // publish, wait for the peers, reduce in rank order, finalise
            } else {
                void *sums = nullptr, *mins = nullptr;
                uint64_t n_sums = 0, n_mins = 0;
                CHECK_RTS(rts_bins_device(e, &sums, &n_sums, &mins, &n_mins));
                CHECK_NCCL(ncclGroupStart());
                CHECK_NCCL(ncclAllReduce(sums, sums, n_sums, ncclDouble, ncclSum, comms[g], st));
                CHECK_NCCL(ncclAllReduce(mins, mins, n_mins, ncclUint64, ncclMin, comms[g], st));
                CHECK_NCCL(ncclGroupEnd());
                CHECK_RTS(rts_finalise_bins(e));
            }
            CHECK_CUDA(cudaEventRecord(e1, st));
            uint32_t n = 0;
            std::vector<rts_bin> bins(4096);
            CHECK_RTS(rts_get_bins(e, bins.data(), (uint32_t)bins.size(), &n));      // waits for the pulse
            if (n > bins.size()) { bins.resize(n); CHECK_RTS(rts_get_bins(e, bins.data(), n, &n)); }
            bins.resize(n);
            if (g == 0 && p >= 0) {
                CHECK_CUDA(cudaEventElapsedTime(&ms_pulse[p], e0, e1));
                reduced[p] = bins;
            }
        }
        if (g == 0) CHECK_RTS(rts_kernel_launches(e, &launches));
        bar.wait();                                         // nobody frees a block a peer may still be reading
        rts_destroy(e);
        cudaStreamDestroy(st);
    };
    std::vector<std::thread> threads;
    for (int g = 0; g < n_gpus; g++) threads.emplace_back(worker, g);
    for (auto &t : threads) t.join();
    for (auto &c : comms) ncclCommDestroy(c);

    // ---- check: the whole launch on one engine
    CHECK_CUDA(cudaSetDevice(0));
    rts_engine *ref = nullptr;
    CHECK_RTS(rts_create(0, &ref));
    CHECK_RTS(rts_scene_set_targets(ref, S.views.data(), (uint32_t)K));
    double max_rel = 0.0;
    bool ok = true;
    size_t n_bins = 0;
    for (int p = 0; p < pulses; p++) {
        const std::vector<rts_pose> poses = S.poses(p, dt);
        CHECK_RTS(rts_scene_set_poses(ref, poses.data(), (uint32_t)K));
        CHECK_RTS(rts_trace_pulse(ref, &S.pulse, RTS_OUT_BINS | RTS_NO_REUSE));
        uint32_t n = 0;
        std::vector<rts_bin> bins(4096);
        CHECK_RTS(rts_get_bins(ref, bins.data(), (uint32_t)bins.size(), &n));
        if (n > bins.size()) { bins.resize(n); CHECK_RTS(rts_get_bins(ref, bins.data(), n, &n)); }
        bins.resize(n);
        const std::vector<rts_bin> &got = reduced[p];
        n_bins = bins.size();
        if (got.size() != bins.size()) { ok = false; std::fprintf(stderr, "pulse %d: %zu bins reduced, %zu expected\n", p, got.size(), bins.size()); continue; }
        for (size_t i = 0; i < bins.size(); i++) {
            const rts_bin &a = got[i], &b = bins[i];
            if (a.rx != b.rx || std::memcmp(a.path, b.path, sizeof(a.path)) || a.npath != b.npath || a.min_slot != b.min_slot || a.direct != b.direct) {
                ok = false;
                std::fprintf(stderr, "pulse %d bin %zu: key/count/slot mismatch (npath %.0f vs %.0f)\n", p, i, a.npath, b.npath);
                continue;
            }
            const double pa[8] = {a.sum_sqrt_power, a.sum_delay, a.sum_phase, a.sum_doppler, a.power, a.delay, a.phase, a.doppler};
            const double pb[8] = {b.sum_sqrt_power, b.sum_delay, b.sum_phase, b.sum_doppler, b.power, b.delay, b.phase, b.doppler};
            for (int k = 0; k < 8; k++)
                if (pa[k] != pb[k]) max_rel = std::fmax(max_rel, std::fabs(pa[k] - pb[k]) / std::fmax(std::fabs(pb[k]), 1e-300));
        }
    }
    rts_destroy(ref);
    if (max_rel > 1e-9) ok = false;
    double ms = 0;
    for (float v : ms_pulse) ms += v;
    ms /= pulses;
    const double rays = (double)grid * grid;
    size_t tris = 0;
    for (const Mesh &m : S.meshes) tris += m.tris.size() / 3;
    std::printf("{\"gpus\": %d, \"exchange\": \"%s\", \"pulses\": %d, \"rays_per_pulse\": %.0f, \"triangles\": %zu, \"receivers\": %d, \"bins\": %zu, "
                "\"ms_per_pulse\": %.4f, \"Mrays_per_s\": %.1f, \"kernel_launches_rank0\": %llu, \"max_rel_vs_single_gpu\": %.3e, \"ok\": %s}\n",
                n_gpus, peer ? "peer" : "nccl", pulses, rays, tris, n_rx, n_bins, ms, rays / (ms * 1e-3) / 1e6, (unsigned long long)launches, max_rel, ok ? "true" : "false");
    return ok ? 0 : 1;
}
