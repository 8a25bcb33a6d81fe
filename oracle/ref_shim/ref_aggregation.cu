// Compiles the reference's aggregation.cu UNMODIFIED (from /root/reference, via -I) and exposes
// rs::kernel_wrapper through a C symbol for ctypes.  TEST INFRASTRUCTURE (GPU box only).
#include "aggregation.cu"
extern "C" void ref_kernel_wrapper(void *rx_results, int *rx_intersects, unsigned receivedRays, unsigned depthTotal,
                                   unsigned MaxThreads, unsigned MaxBlocks, double cspeed, double carrier,
                                   double *npath, double *power, double *doppler, double *delay, double *phase,
                                   int *pathMatch)
{
    rs::kernel_wrapper((PerRayData *)rx_results, rx_intersects, receivedRays, depthTotal, MaxThreads, MaxBlocks, cspeed,
                       carrier, npath, power, doppler, delay, phase, pathMatch);
}
