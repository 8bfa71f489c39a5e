// Setup-side kernels (preconditioner construction).  Filled in incrementally; see precond_setup.c.
#include <cuda_runtime.h>
#include "b200_dev.h"
#include "b200_cuda_util.cuh"
