// rts_soars_adapter.hpp — the body of rs::RTS over the rts_b200 C-ABI (header-only, C++17).
//
// The reference's single public entry is `void rs::RTS(World*, unsigned MaxThreads, unsigned MaxBlocks)`
// (/root/reference/ray_tracer.cpp:512).  It cannot be compiled without the host simulator (SOARS), so this
// header restates its host flow as a template over the simulator's types: inside SOARS a maintainer writes
//
//     struct Soars { using World = rs::World; using Transmitter = rs::Transmitter; using Receiver = rs::Receiver;
//                    using Target = rs::Target; using TransmitterPulse = rs::TransmitterPulse;
//                    using Parameters = rs::rsParameters; using Vec3 = rs::Vec3; using SVec3 = rs::SVec3;
//                    using InterpPoint = rs::InterpPoint; using Response = rs::Response; };
//     namespace rs { void RTS(World *w, unsigned mt, unsigned mb) { rts_b200::RTS<Soars>(w, mt, mb); } }
//
// and the simulator links librts_b200.so instead of the OptiX objects.  tests/mock_soars/ instantiates the same
// template with a stand-in World to check it end to end.  Every step cites the reference lines it replaces.
//
// Two paths:
//   exact (default)  records on the device -> received rays compacted on the device (rts_get_received) -> the
//                    reference's own per-ray host loop with the Target::GetRCS / GetGain callbacks
//                    (ray_tracer.cpp:1190-1258), only over received rays -> rs::kernel_wrapper -> unique paths ->
//                    responses (ray_tracer.cpp:1263-1321).
//   fused            Options::fused: RCS and gains are sampled once per pulse as scalars and folded on the device
//                    (rts_pulse.targ_rcs / gain_tx / gain_rx); responses come from rts_get_responses.  Exact only
//                    when the callbacks do not depend on the angles.
//   tabulated        Options::tabulated: Target::GetRCS and the GetGain patterns are sampled on regular grids
//                    (Options::table_az x table_el) — once per pulse, a few hundred thousand callback calls instead of one
//                    set per received ray — and evaluated on the device by bilinear interpolation per hop / per captured
//                    ray (rts_set_rcs_tables / rts_set_antennas, RTS_TABLES): no per-ray data crosses to the host.  Exact
//                    for callbacks that are themselves tables on that grid; otherwise within the interpolation error
//                    (h^2/8 * |f''|: 1e-5 for smooth patterns at the default resolution).  Reflection-only pulses.
#ifndef RTS_SOARS_ADAPTER_HPP
#define RTS_SOARS_ADAPTER_HPP

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "rts_b200.h"

namespace rts_b200 {

struct Options {
    int device = 0;
    bool fused = false;     // fold scalar RCS / gains on the device instead of calling the callbacks per ray
    bool tabulated = false; // sample the callbacks on grids and evaluate them on the device (angle-dependent, no per-ray D2H)
    unsigned table_az = 1441, table_el = 721;   // samples over [-2pi, 2pi] x [-pi, pi]
    bool verbose = false;   // the reference's progress prints (ray_tracer.cpp:521, 1260, 1362)
};

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace detail {
inline void check(int rc, const char *what)
{
    if (rc != RTS_OK) throw Error(std::string(what) + ": " + rts_last_error());
}
struct EngineGuard {
    rts_engine *e = nullptr;
    ~EngineGuard() { if (e) rts_destroy(e); }
};
// One target's mesh at its t = 0 orientation: rect_mesh / sphere_mesh / file_mesh (ray_tracer.cpp:956-987).
struct Mesh {
    std::vector<double> verts, normals;
    std::vector<uint32_t> tris;
};
template <class Target> Mesh generate(Target *tg)
{
    const float yaw = (float)tg->GetTargetRotation(0).yaw, pitch = (float)tg->GetTargetRotation(0).pitch,
                roll = (float)tg->GetTargetRotation(0).roll;                          // :956-958
    const std::string shape = tg->GetShape();                                          // :963
    float w = 0, h = 0, d = 0, radius = 0;
    unsigned subdivs = 0;
    std::string v_file, n_file;
    if (shape == "rect") tg->GetRect(w, h, d);                                         // :966-969
    else if (shape == "sphere") tg->GetSphere(subdivs, radius);                        // :973-979
    else if (shape == "file") tg->GetFile(v_file, n_file);                             // :983-987
    else throw Error("unknown target shape '" + shape + "'");
    Mesh m;
    uint32_t nv = 0, nt = 0, nn = 0;
    auto call = [&](double *pv, uint32_t *pt, double *pn) {
        if (shape == "rect") return rts_rect_mesh(w, h, d, yaw, pitch, roll, pv, &nv, pt, &nt, pn, &nn);
        if (shape == "sphere") return rts_sphere_mesh(subdivs, radius, yaw, pitch, roll, pv, &nv, pt, &nt, pn, &nn);
        return rts_file_mesh(v_file.c_str(), n_file.c_str(), yaw, pitch, roll, pv, &nv, pt, &nt, pn, &nn);
    };
    check(call(nullptr, nullptr, nullptr), "mesh generator (counts)");
    m.verts.resize(3 * (size_t)nv); m.tris.resize(3 * (size_t)nt); m.normals.resize(3 * (size_t)nn);
    check(call(m.verts.data(), m.tris.data(), m.normals.data()), "mesh generator");
    return m;
}
} // namespace detail

template <class S>
void RTS(typename S::World *world, unsigned int MaxThreads, unsigned int MaxBlocks, const Options &opt = Options())
{
    using Transmitter = typename S::Transmitter;
    using Receiver = typename S::Receiver;
    using Target = typename S::Target;
    using Params = typename S::Parameters;
    using Vec3 = typename S::Vec3;
    using SVec3 = typename S::SVec3;
    if (opt.verbose) printf("Setting up RTS...\n");                                   // :521

    detail::EngineGuard g;
    detail::check(rts_create(opt.device, &g.e), "rts_create");
    rts_engine *eng = g.e;

    // ray counts and depths (:600-605), simulation constants (:645-648)
    const auto rts_vars = Params::GetRTSVariables();
    const unsigned h_numRays = rts_vars.x, h_maxReflDepth = rts_vars.y;
    const unsigned h_maxRefrDepth = rts_vars.z > 0 ? 2u : 0u;
    const double cspeed = Params::c();
    const double sim_starttime = Params::start_time();
    const double sample_time = 1.0 / Params::cw_sample_rate();
    const bool interpolate_smooth = Params::interpolate_smooth();
    const unsigned depthTotal = h_maxRefrDepth + h_maxReflDepth;                       // :655

    Transmitter **trans_arr = world->transmitters.data();
    Receiver **recv_arr = world->receivers.data();
    Target **targ_arr = world->targets.data();
    const unsigned txsize = (unsigned)world->transmitters.size(), rxsize = (unsigned)world->receivers.size(),
                   targsize = (unsigned)world->targets.size();

    // Scene: every target's mesh once, at its t = 0 orientation.  The reference regenerates (and, for file meshes,
    // re-reads) them every pulse (:936-987); here the per-pulse rotation/translation runs on the device.
    std::vector<detail::Mesh> meshes(targsize);
    std::vector<rts_target_mesh> tm(targsize);
    for (unsigned k = 0; k < targsize; k++) {
        meshes[k] = detail::generate<Target>(targ_arr[k]);
        tm[k].n_verts = (uint32_t)(meshes[k].verts.size() / 3);
        tm[k].n_tris = (uint32_t)(meshes[k].tris.size() / 3);
        tm[k].n_normals = (uint32_t)(meshes[k].normals.size() / 3);
        tm[k]._pad = 0;
        tm[k].verts = meshes[k].verts.data(); tm[k].tris = meshes[k].tris.data(); tm[k].normals = meshes[k].normals.data();
        tm[k].refl_coeff = targ_arr[k]->GetReflCoeff();                                // :1043
        tm[k].refr_index = targ_arr[k]->GetRefrIndex();                                // :1044
    }
    detail::check(rts_scene_set_targets(eng, tm.data(), targsize), "rts_scene_set_targets");

    for (unsigned tx_i = 0; tx_i < txsize; tx_i++) {                                   // :806
        Transmitter *trans = trans_arr[tx_i];
        const unsigned pulseCount = (unsigned)trans->GetPulseCount();                  // :810
        typename S::TransmitterPulse signal;
        trans->GetPulse(&signal, 0);                                                   // :812
        auto *wave = signal.wave;
        const double carrier = wave->GetCarrier();                                     // :814
        const double Wl = cspeed / carrier;                                            // :815
        const auto h_txSpan = trans->GetTxSpan();                                      // :818
        for (unsigned j = 0; j < rxsize; j++)                                          // :829
            recv_arr[j]->SetNoiseTemperature(wave->GetTemp() + recv_arr[j]->GetNoiseTemperature());

        for (unsigned k = 0; k < pulseCount; k++) {                                    // :843
            trans->GetPulse(&signal, (int)k);
            const double time_t = signal.time;                                         // :846-850

            // receivers: sphere centre and angular window (:894-918)
            std::vector<rts_rx_sphere> rx(rxsize);
            for (unsigned j = 0; j < rxsize; j++) {
                const auto rxsphere = recv_arr[j]->GetRxSphere();
                const Vec3 repos = recv_arr[j]->GetPosition(0);
                rts_rx_desc d;
                d.position[0] = repos.x; d.position[1] = repos.y; d.position[2] = repos.z;
                d.azimuth = recv_arr[j]->GetRotation(time_t).azimuth;
                d.elevation = recv_arr[j]->GetRotation(time_t).elevation;
                d.radius = rxsphere.x; d.theta_span = rxsphere.y; d.phi_span = rxsphere.z;
                rts_rx_sphere_from_desc(&d, &rx[j]);
            }

            // targets: position now and one sample later (:941-948), time-varying rotation on top of the t = 0
            // one only when rotating and past the start time (:993-1005), velocity over one sample (:1144-1145)
            std::vector<rts_pose> pose(targsize);
            std::vector<double> vel(3 * (size_t)targsize), rcs_scalar(targsize, 1.0);
            for (unsigned i = 0; i < targsize; i++) {
                Target *tg = targ_arr[i];
                const Vec3 p0 = tg->GetPosition(time_t), p1 = tg->GetPosition(time_t + sample_time);
                rts_pose &P = pose[i];
                for (int a = 0; a < 9; a++) P.R[a] = (a % 4 == 0) ? 1.0 : 0.0;
                P.t[0] = p0.x; P.t[1] = p0.y; P.t[2] = p0.z;
                P.has_rotation = (tg->GetRotating() == true) && (time_t > sim_starttime);
                P._pad = 0;
                if (P.has_rotation)   // doubles narrowed to vertex_rotation's float parameters (:156, :998-1002)
                    rts_rotation_matrix((float)tg->GetTargetRotation(time_t).yaw, (float)tg->GetTargetRotation(time_t).pitch,
                                        (float)tg->GetTargetRotation(time_t).roll, P.R);
                vel[3 * i] = (p1.x - p0.x) / sample_time;
                vel[3 * i + 1] = (p1.y - p0.y) / sample_time;
                vel[3 * i + 2] = (p1.z - p0.z) / sample_time;
                if (opt.fused) rcs_scalar[i] = tg->GetRCS(0.0, 0.0, Wl);
            }
            detail::check(rts_scene_set_poses(eng, pose.data(), targsize), "rts_scene_set_poses");

            rts_pulse pl = {};
            pl.nx = pl.ny = pl.nz = h_numRays;                                         // rtContextLaunch3D(N,N,N), :1165
            pl.max_refl = h_maxReflDepth; pl.max_refr = h_maxRefrDepth;
            pl.interpolate_smooth = interpolate_smooth ? 1 : 0;
            const Vec3 trpos = trans->GetPosition(0);                                  // :881
            pl.tx_origin[0] = trpos.x; pl.tx_origin[1] = trpos.y; pl.tx_origin[2] = trpos.z;
            pl.tx_dir[0] = trans->GetRotation(time_t).azimuth;                         // :888
            pl.tx_dir[1] = trans->GetRotation(time_t).elevation;                       // :889
            pl.tx_span[0] = h_txSpan.x; pl.tx_span[1] = h_txSpan.y; pl.tx_span[2] = h_txSpan.z;
            pl.cspeed = cspeed; pl.carrier = carrier;
            pl.n_rx = rxsize; pl.rx = rx.data();
            pl.n_targets = targsize; pl.targ_vel = vel.data();
            const double h_rayOrigin[3] = {trpos.x, trpos.y, trpos.z};

            if (opt.tabulated) {
                if (h_maxRefrDepth) throw Error("Options::tabulated: pulses with refraction need the exact path");
                // the callbacks as tables: GetRCS over the summed in/out angles of dbuf_rcs_angle, GetGain over the look
                // direction minus the boresight; the receiver's boresight drift from two GetRotation samples
                const double PI = 3.14159265358979323846;
                const unsigned na = opt.table_az, ne = opt.table_el;
                const double az0 = -2 * PI, azs = 4 * PI / (na - 1), el0 = -PI, els = 2 * PI / (ne - 1);
                std::vector<std::vector<double>> rcs_vals(targsize, std::vector<double>((size_t)na * ne));
                std::vector<rts_table2d> rcs_tab(targsize);
                for (unsigned i = 0; i < targsize; i++) {
                    for (unsigned a = 0; a < na; a++)
                        for (unsigned b = 0; b < ne; b++) rcs_vals[i][(size_t)a * ne + b] = targ_arr[i]->GetRCS(az0 + a * azs, el0 + b * els, Wl);
                    rcs_tab[i] = rts_table2d{na, ne, az0, azs, el0, els, rcs_vals[i].data()};
                }
                detail::check(rts_set_rcs_tables(eng, rcs_tab.data(), targsize), "rts_set_rcs_tables");
                auto sample_gain = [&](auto *ant, const SVec3 &bore, std::vector<double> &vals) {
                    vals.resize((size_t)na * ne);
                    for (unsigned a = 0; a < na; a++)
                        for (unsigned b = 0; b < ne; b++)
                            vals[(size_t)a * ne + b] = ant->GetGain(SVec3(1, bore.azimuth + (az0 + a * azs), bore.elevation + (el0 + b * els)), bore, Wl);
                    return rts_table2d{na, ne, az0, azs, el0, els, vals.data()};
                };
                std::vector<double> tx_vals;
                std::vector<std::vector<double>> rx_vals(rxsize);
                rts_antenna txa = {};
                const SVec3 tx_bore = trans->GetRotation(time_t);
                txa.gain = sample_gain(trans, tx_bore, tx_vals);
                txa.bore_az = tx_bore.azimuth; txa.bore_el = tx_bore.elevation;
                txa.position[0] = trpos.x; txa.position[1] = trpos.y; txa.position[2] = trpos.z;
                std::vector<rts_antenna> rxa(rxsize);
                for (unsigned j = 0; j < rxsize; j++) {
                    const SVec3 b0 = recv_arr[j]->GetRotation(time_t), b1 = recv_arr[j]->GetRotation(time_t + sample_time);
                    rxa[j] = rts_antenna{};
                    rxa[j].gain = sample_gain(recv_arr[j], b0, rx_vals[j]);
                    rxa[j].bore_az = b0.azimuth; rxa[j].bore_el = b0.elevation;
                    rxa[j].rate_az = (b1.azimuth - b0.azimuth) / sample_time; rxa[j].rate_el = (b1.elevation - b0.elevation) / sample_time;
                    const Vec3 rp = recv_arr[j]->GetPosition(0);
                    rxa[j].position[0] = rp.x; rxa[j].position[1] = rp.y; rxa[j].position[2] = rp.z;
                }
                detail::check(rts_set_antennas(eng, &txa, rxa.data(), rxsize), "rts_set_antennas");
                detail::check(rts_trace_pulse(eng, &pl, RTS_OUT_BINS | RTS_TABLES), "rts_trace_pulse");
                uint32_t n = 0;
                detail::check(rts_get_responses(eng, nullptr, 0, &n), "rts_get_responses");
                std::vector<rts_response> resp(n ? n : 1);
                detail::check(rts_get_responses(eng, resp.data(), n, &n), "rts_get_responses");
                for (uint32_t u = 0; u < n; u++) {                                     // :1301-1320
                    const rts_response &a = resp[u];
                    typename S::InterpPoint point(a.power, time_t + a.delay, a.delay, a.doppler, a.phase,
                                                  recv_arr[a.rx]->GetNoiseTemperature());
                    auto *response = new typename S::Response(wave, trans);
                    response->AddInterpPoint(point);
                    recv_arr[a.rx]->AddResponse(response);
                }
                continue;
            }

            if (opt.fused) {
                // scalar callbacks folded on the device, responses straight from the bins
                pl.targ_rcs = rcs_scalar.data();
                pl.gain_tx = trans->GetGain(trans->GetRotation(time_t), trans->GetRotation(time_t), Wl);
                pl.gain_rx = rxsize ? recv_arr[0]->GetGain(recv_arr[0]->GetRotation(time_t), recv_arr[0]->GetRotation(time_t), Wl) : 1.0;
                detail::check(rts_trace_pulse(eng, &pl, RTS_OUT_BINS), "rts_trace_pulse");
                uint32_t n = 0;
                detail::check(rts_get_responses(eng, nullptr, 0, &n), "rts_get_responses");
                std::vector<rts_response> resp(n ? n : 1);
                detail::check(rts_get_responses(eng, resp.data(), n, &n), "rts_get_responses");
                for (uint32_t u = 0; u < n; u++) {                                     // :1301-1320
                    const rts_response &a = resp[u];
                    typename S::InterpPoint point(a.power, time_t + a.delay, a.delay, a.doppler, a.phase,
                                                  recv_arr[a.rx]->GetNoiseTemperature());
                    auto *response = new typename S::Response(wave, trans);
                    response->AddInterpPoint(point);
                    recv_arr[a.rx]->AddResponse(response);
                }
                continue;
            }

            // ---- exact path ----
            detail::check(rts_trace_pulse(eng, &pl, RTS_OUT_RECORDS), "rts_trace_pulse");
            rts_sizes sz;
            rts_result_sizes(&pl, &sz);
            uint64_t receivedRays64 = 0;
            detail::check(rts_get_received(eng, 0, &receivedRays64, nullptr, nullptr, nullptr, nullptr), "rts_get_received");
            const unsigned receivedRays = (unsigned)receivedRays64;
            if (opt.verbose) printf("Rays: %d,\n", receivedRays);                      // :1260
            if (receivedRays == 0) continue;                                           // :1263
            std::vector<rts_ray_record> h_rx_results(receivedRays);
            std::vector<int32_t> h_rx_intersects((size_t)receivedRays * std::max(1u, depthTotal));
            std::vector<double> rcs_angle((size_t)receivedRays * std::max(1u, depthTotal) * 2);
            detail::check(rts_get_received(eng, receivedRays, &receivedRays64, nullptr, h_rx_results.data(), h_rx_intersects.data(),
                                           rcs_angle.data()), "rts_get_received");

            // the reference's per-ray host loop (:1190-1258), over received rays only
            for (unsigned i = 0; i < receivedRays; i++) {
                rts_ray_record &r = h_rx_results[i];
                const Vec3 repos = recv_arr[r.received]->GetPosition(0);               // :1201
                Vec3 tv, rv;
                if ((r.reflDepth == 0) && (r.refrDepth == 0)) {                        // :1205-1208
                    tv = Vec3(h_rayOrigin[0] - repos.x, h_rayOrigin[1] - repos.y, h_rayOrigin[2] - repos.z);
                    rv = Vec3(repos.x - h_rayOrigin[0], repos.y - h_rayOrigin[1], repos.z - h_rayOrigin[2]);
                } else {                                                               // :1210-1212
                    tv = Vec3(r.firstHitPoint[0] - h_rayOrigin[0], r.firstHitPoint[1] - h_rayOrigin[1], r.firstHitPoint[2] - h_rayOrigin[2]);
                    rv = Vec3(r.prevHitPoint[0] - repos.x, r.prevHitPoint[1] - repos.y, r.prevHitPoint[2] - repos.z);
                }
                SVec3 transvec(tv), recvvec(rv);
                transvec.length = 1;                                                   // :1216
                recvvec.length = 1;                                                    // :1217
                const double delay = (r.rayLength) / cspeed;                           // :1218
                for (unsigned kk = 0; kk < depthTotal; kk++) {                         // :1221-1231
                    const size_t at = kk + (size_t)i * depthTotal;
                    const int targ_k = h_rx_intersects[at];
                    if (targ_k >= 0) {
                        const double targRCS = targ_arr[targ_k]->GetRCS(rcs_angle[2 * at], rcs_angle[2 * at + 1], Wl);
                        r.power *= targRCS;
                    }
                }
                const double Gt = trans->GetGain(transvec, trans->GetRotation(time_t), Wl);                        // :1233
                const double Gr = recv_arr[r.received]->GetGain(recvvec, recv_arr[r.received]->GetRotation(delay + time_t), Wl); // :1234-1235
                r.power *= (Wl * Wl * Gt * Gr);                                        // :1247
                const double Vr = r.doppler / 2;                                       // :1252
                r.doppler = carrier * (((1 + Vr / cspeed) / (1 - Vr / cspeed)) - 1);   // :1253
            }

            // aggregation (:1266-1285) through the reference's own entry point, exported by librts_b200.so
            std::vector<double> h_npath(receivedRays, 0), h_power(receivedRays, 0), h_doppler(receivedRays, 0),
                h_delay(receivedRays, 0), h_phase(receivedRays, 0);
            std::vector<int32_t> h_pathMatch(receivedRays, (int32_t)(sz.ray_total + 1));   // :1271
            detail::check(rts_aggregate(eng, h_rx_results.data(), h_rx_intersects.data(), receivedRays, depthTotal, cspeed, carrier,
                                        h_npath.data(), h_power.data(), h_doppler.data(), h_delay.data(), h_phase.data(),
                                        h_pathMatch.data()), "rts_aggregate");
            (void)MaxThreads; (void)MaxBlocks;   // launch-shape hints of the reference's kernels (:1283)

            // unique paths -> one response each (:1289-1321)
            std::vector<int32_t> unique_path_rays(h_pathMatch.begin(), h_pathMatch.end());
            std::sort(unique_path_rays.begin(), unique_path_rays.end());
            unique_path_rays.erase(std::unique(unique_path_rays.begin(), unique_path_rays.end()), unique_path_rays.end());
            for (size_t u = 0; u < unique_path_rays.size(); u++) {
                const unsigned i = (unsigned)unique_path_rays[u];
                if (i >= receivedRays) continue;
                const unsigned rxi = (unsigned)h_rx_results[i].received;
                typename S::InterpPoint point(h_rx_results[i].power, time_t + h_delay[i], h_delay[i], h_rx_results[i].doppler,
                                              h_phase[i], recv_arr[rxi]->GetNoiseTemperature());
                auto *response = new typename S::Response(wave, trans);
                response->AddInterpPoint(point);
                recv_arr[rxi]->AddResponse(response);
            }
        }
    }
    if (opt.verbose) printf("Exiting RTS...\n");                                      // :1362
}

} // namespace rts_b200
#endif // RTS_SOARS_ADAPTER_HPP
