// Definitions for tests/emu/cuda_emu.h (test infrastructure only).
#include "cuda_emu.h"
namespace emu {
thread_local ThreadCtx tctx;
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
}  // namespace emu
