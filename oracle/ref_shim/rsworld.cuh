// Stub of the SOARS header the reference's aggregation.cuh includes (aggregation.cuh:11);
// supplies only what aggregation.cu needs: the CUDA vector types and the PerRayData contract.
#pragma once
#include <vector_types.h>
#include "ray_tracer.h"
