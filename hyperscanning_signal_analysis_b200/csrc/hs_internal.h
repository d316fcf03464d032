// Internal helpers shared by the .cu translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/hs_b200.h"

namespace hs {

int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int device_sm_count();


}  // namespace hs
