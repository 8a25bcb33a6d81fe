// oracle_aggregate.cpp — host post-process and ray aggregation of the reference, restated.
// TEST INFRASTRUCTURE (see rts_oracle.h).
//   post-process        /root/reference/ray_tracer.cpp:1190-1258
//   accumulator init    /root/reference/ray_tracer.cpp:1266-1271 (done by the caller, as in the reference)
//   myKernel1/myKernel2 /root/reference/aggregation.cu:32-97
//   unique paths        /root/reference/ray_tracer.cpp:1289-1294
#include "oracle_common.h"
#include <algorithm>
#include <map>
#include <vector>

extern "C" int orc_postprocess(const rts_ray_record *results, const int32_t *targ_intersect, uint64_t ray_total,
                               uint32_t depth_total, double cspeed, double carrier, const double *rcs_per_target,
                               double gain, rts_ray_record *rx_results, int32_t *rx_intersects, uint64_t *rx_slots,
                               uint64_t cap, uint64_t *n_received)
{
    const double Wl = cspeed / carrier; // ray_tracer.cpp:815
    uint64_t receivedRays = 0;
    for (uint64_t i = 0; i < ray_total; i++) {
        if (results[i].received >= 0) {
            rts_ray_record r = results[i];
            if (receivedRays < cap) {
                for (uint32_t k = 0; k < depth_total; k++) {
                    uint64_t depth_ray_index = k + i * depth_total;
                    int targ_k = targ_intersect ? targ_intersect[depth_ray_index] : -1;
                    if (rx_intersects) rx_intersects[receivedRays * depth_total + k] = targ_k;
                    if (targ_k >= 0) {
                        double targRCS = rcs_per_target ? rcs_per_target[targ_k] : 1.0;
                        r.power *= targRCS;
                    }
                }
                // Gt, Gr are SOARS callbacks; here a single caller-supplied product with Gr = 1
                const double Gt = gain, Gr = 1.0;
                r.power *= (Wl * Wl * Gt * Gr);
                double Vr = r.doppler / 2;
                r.doppler = carrier * (((1 + Vr / cspeed) / (1 - Vr / cspeed)) - 1);
                if (rx_results) rx_results[receivedRays] = r;
                if (rx_slots) rx_slots[receivedRays] = i;
            }
            receivedRays++;
        }
    }
    if (n_received) *n_received = receivedRays;
    return 0;
}

extern "C" int orc_aggregate_literal(rts_ray_record *res, const int32_t *rows, uint32_t receivedRays,
                                     uint32_t depthTotal, double cspeed, double carrier, double *npath, double *power,
                                     double *doppler, double *delay_arr, double *phase_arr, int32_t *pathMatch)
{
    // myKernel1, aggregation.cu:38-73 (serial over i)
    for (uint32_t i = 0; i < receivedRays; i++) {
        for (uint32_t r = 0; r < receivedRays; r++) {
            if (res[i].received == res[r].received) {
                bool row_equal = true;
                for (uint32_t k = 0; k < depthTotal; k++) {
                    if (rows[k + (size_t)i * depthTotal] != rows[k + (size_t)r * depthTotal]) {
                        row_equal = false;
                        break;
                    }
                }
                if ((row_equal == true) || ((res[i].reflDepth == 0) && (res[i].refrDepth == 0))) {
                    double delay = (res[r].rayLength) / cspeed;
                    double phase = -fmod(delay * 2 * M_PI * carrier, 2 * M_PI);
                    npath[i] += 1;
                    power[i] += sqrt(res[r].power);
                    delay_arr[i] += delay;
                    phase_arr[i] += phase;
                    doppler[i] += res[r].doppler;
                    if ((int32_t)r < pathMatch[i]) pathMatch[i] = (int32_t)r;
                }
            }
        }
    }
    // myKernel2, aggregation.cu:83-95
    for (uint32_t i = 0; i < receivedRays; i++) {
        if (npath[i] > 0) {
            res[i].power = pow(power[i] / npath[i], 2);
            delay_arr[i] /= npath[i];
            phase_arr[i] /= npath[i];
            res[i].doppler = doppler[i] / npath[i];
        }
    }
    return 0;
}

extern "C" int orc_aggregate_binned(rts_ray_record *res, const int32_t *rows, uint32_t receivedRays,
                                    uint32_t depthTotal, double cspeed, double carrier, double *npath, double *power,
                                    double *doppler, double *delay_arr, double *phase_arr, int32_t *pathMatch)
{
    struct Acc { double n = 0, sp = 0, sd = 0, sph = 0, sdop = 0; int32_t mn = INT32_MAX; };
    typedef std::vector<int32_t> Key;
    std::map<Key, Acc> groups;     // key = (rx, row)
    std::map<int32_t, Acc> per_rx; // totals for the direct-ray rule (aggregation.cu:56)
    std::vector<Key> keys(receivedRays);
    for (uint32_t r = 0; r < receivedRays; r++) {
        Key k(depthTotal + 1);
        k[0] = res[r].received;
        for (uint32_t c = 0; c < depthTotal; c++) k[c + 1] = rows[c + (size_t)r * depthTotal];
        keys[r] = k;
        double delay = (res[r].rayLength) / cspeed;
        double phase = -fmod(delay * 2 * M_PI * carrier, 2 * M_PI);
        Acc *targets[2] = {&groups[k], &per_rx[res[r].received]};
        for (Acc *a : targets) {
            a->n += 1; a->sp += sqrt(res[r].power); a->sd += delay; a->sph += phase; a->sdop += res[r].doppler;
            a->mn = std::min(a->mn, (int32_t)r);
        }
    }
    // inputs are read before any res[].power/doppler is overwritten, as the two-kernel split guarantees
    std::vector<const Acc *> pick(receivedRays);
    for (uint32_t i = 0; i < receivedRays; i++) {
        const bool direct = (res[i].reflDepth == 0) && (res[i].refrDepth == 0);
        pick[i] = direct ? &per_rx[res[i].received] : &groups[keys[i]];
    }
    for (uint32_t i = 0; i < receivedRays; i++) {
        const Acc &a = *pick[i];
        npath[i] += a.n; power[i] += a.sp; delay_arr[i] += a.sd; phase_arr[i] += a.sph; doppler[i] += a.sdop;
        if (a.mn < pathMatch[i]) pathMatch[i] = a.mn;
        if (npath[i] > 0) {
            res[i].power = pow(power[i] / npath[i], 2);
            delay_arr[i] /= npath[i];
            phase_arr[i] /= npath[i];
            res[i].doppler = doppler[i] / npath[i];
        }
    }
    return 0;
}

extern "C" uint32_t orc_unique_paths(const int32_t *path_match, uint32_t received, int32_t *out)
{
    std::vector<int32_t> u(path_match, path_match + received);
    std::sort(u.begin(), u.end());
    u.erase(std::unique(u.begin(), u.end()), u.end());
    if (out) std::copy(u.begin(), u.end(), out);
    return (uint32_t)u.size();
}

// ray_tracer.cpp:1289-1320
extern "C" uint32_t orc_responses(const rts_ray_record *rx_results, const double *delay, const double *phase,
                                  const int32_t *path_match, const uint64_t *rx_slots, uint32_t received,
                                  rts_response *out, uint32_t cap)
{
    std::vector<int32_t> unique_path_rays(path_match, path_match + received);   // :1290
    std::sort(unique_path_rays.begin(), unique_path_rays.end());               // :1291
    unique_path_rays.erase(std::unique(unique_path_rays.begin(), unique_path_rays.end()), unique_path_rays.end()); // :1292
    uint32_t n = 0;
    for (size_t k = 0; k < unique_path_rays.size(); k++) {                      // :1301
        const uint32_t i = (uint32_t)unique_path_rays[k];                       // :1304
        if (i >= received) continue; // untouched pre-fill value (never happens for rays that went through aggregation)
        if (out && n < cap) {
            rts_response &r = out[n];
            r.rx = rx_results[i].received; r._pad = 0;                          // :1305
            r.slot = rx_slots ? rx_slots[i] : i;
            r.power = rx_results[i].power;                                      // :1312
            r.delay = delay[i];                                                 // :1308
            r.doppler = rx_results[i].doppler;                                  // :1314
            r.phase = phase[i];                                                 // :1309
        }
        n++;
    }
    return n;
}
