// OptiX emulation shim (test infrastructure) — see rts_optix_shim.h
#include "../rts_optix_shim.h"
