// Host-side helpers shared by the d2dx translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/d2dx.h"

struct d2dx_handle {
  int device;
  int sm_count;
};

namespace d2dx {
int set_error(int code, const char* fmt, ...);
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace d2dx

// NVTX range around every launching ABI entry: shows up under nsys / ncu --nvtx as the C-ABI call that enqueued the
// kernels; header-only NVTX v3 is a no-op unless a profiler injects itself
namespace d2dx {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace d2dx
#define D2DX_NVTX(name) d2dx::NvtxRange d2dx_nvtx_range_(name)

#define D2DX_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return d2dx::set_error(D2DX_EINVAL, __VA_ARGS__); } while (0)
#define D2DX_CUDA(call) \
  do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
         return d2dx::set_error(D2DX_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
#define D2DX_LAUNCH_CHECK(name) \
  do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
         return d2dx::set_error(D2DX_ECUDA, "launch of %s: %s", name, cudaGetErrorString(e_)); } while (0)
