// Internal helpers shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/hs_b200.h"

namespace hs {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int device_sm_count();
int compute_sm_count();      // SMs the MVAR kernels size their persistent grids for (hs_set_compute_sm_limit; default: all)

// Experiment switches (environment variables selecting alternative kernels, chunk sizes, phase skips ...) exist only in
// builds made with `make HS_EXPERIMENT=1`; the product library takes the default and never reads the environment.
#ifdef HS_EXPERIMENT
int exp_env_int(const char* name, int dflt);
#else
inline int exp_env_int(const char*, int dflt) { return dflt; }
#endif


}  // namespace hs
