// Kernel-level C ABI of libarapgs: plain device pointers + sizes + an explicit
// cudaStream_t.  Declared in include/arapgs_kernels.h (with the reference
// interface each group replaces); the session layer (session.cu) is built on them.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <cmath>
#include <vector>

// declarations live in the public header
#include "../../include/arapgs_kernels.h"
