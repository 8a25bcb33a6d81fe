/* rts_prd.h — the per-ray record with CUDA vector member types, for C++/CUDA hosts.
 *
 * Layout twin of the reference contract `struct PerRayData` (/root/reference/ray_tracer.h:13-28):
 * same member names, order, types, size (144) and alignment (16), so a host simulator that was
 * compiled against the reference header can pass its arrays to this library unchanged.
 * A translation unit that already includes the reference's own ray_tracer.h must define
 * RTS_HAVE_REFERENCE_PRD before including this file (the struct has no include guard there).
 */
#ifndef RTS_PRD_H
#define RTS_PRD_H

#include <cstddef>
#include <vector_types.h>
#include "rts_types.h"

#ifndef RTS_HAVE_REFERENCE_PRD
struct PerRayData {
    double       rayLength;     /* total path length                                   */
    double2      refrIndex;     /* previous (x) and current (y) refractive index       */
    unsigned int reflDepth;     /* reflections so far                                  */
    unsigned int refrDepth;     /* refractions so far                                  */
    unsigned int maxRayIndex;   /* result-slot offset of this chain (k * rays)         */
    double3      rayDirection;  /* fp64 direction                                      */
    double3      firstHitPoint;
    double3      prevHitPoint;  /* ray origin until the first bounce                   */
    double       power;
    double       doppler;
    int          received;      /* -1, or the receiver index                           */
    bool         end;
};
#endif

static_assert(sizeof(PerRayData) == 144 && alignof(PerRayData) == 16, "PerRayData: size 144, align 16");
static_assert(offsetof(PerRayData, rayLength) == 0 && offsetof(PerRayData, refrIndex) == 16 &&
              offsetof(PerRayData, reflDepth) == 32 && offsetof(PerRayData, refrDepth) == 36 &&
              offsetof(PerRayData, maxRayIndex) == 40 && offsetof(PerRayData, rayDirection) == 48 &&
              offsetof(PerRayData, firstHitPoint) == 72 && offsetof(PerRayData, prevHitPoint) == 96 &&
              offsetof(PerRayData, power) == 120 && offsetof(PerRayData, doppler) == 128 &&
              offsetof(PerRayData, received) == 136 && offsetof(PerRayData, end) == 140,
              "PerRayData: member offsets of the reference contract");
static_assert(sizeof(rts_ray_record) == sizeof(PerRayData), "rts_ray_record mirrors PerRayData");

/* The reference's aggregation entry point (aggregation.cuh:18-23), exported by librts_b200.so. */
namespace rs {
void kernel_wrapper(PerRayData *h_rx_results_arr, int *h_rx_intersects_arr, unsigned int receivedRays,
                    unsigned int depthTotal, unsigned int MaxThreads, unsigned int MaxBlocks, double cspeed,
                    double carrier, double *h_npath_arr, double *h_power_arr, double *h_doppler_arr,
                    double *h_delay_arr, double *h_phase_arr, int *h_pathMatch);
}
#endif /* RTS_PRD_H */
