// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dit4sr_b200.h"

namespace host {

void set_error(const char* fmt, ...);

#define DIT4SR_REQUIRE(cond, code, ...)      \
  do {                                       \
    if (!(cond)) {                           \
      host::set_error(__VA_ARGS__);          \
      return (code);                         \
    }                                        \
  } while (0)

#define DIT4SR_CUDA(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      host::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                 \
      return (int)_e;                                                                  \
    }                                                                                  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Encode a bf16 tensor map with 128-byte swizzle. dims/strides innermost first; strides in BYTES for dims 1..rank-1.
// Out-of-bounds box elements read as zero. Returns 0 or a DIT4SR_E_* code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

// Same, without swizzle: the shared-memory image is the plain row-major box.
int make_tmap_bf16_noswizzle(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box);

// fp32 tensor map with 128-byte swizzle (inner box extent must be 32 floats = 128 bytes).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

int sm_count();                                      // of the CURRENT device (cached per device)
// true the first time it is called with `mask` while a given device is current: function attributes such as the opt-in
// to > 48 KB of dynamic shared memory are per (function, device), so a process-wide "configured" flag is not enough
bool first_use_on_device(unsigned long long* mask);

// Programmatic dependent launch (griddepcontrol): a kernel launched through launch_pdl may start while the previous kernel
// of the stream is still draining; it must execute ptx-level `griddepcontrol.wait` (host::pdl_wait() on the device side
// is in ptx.cuh) before it touches anything the previous kernel wrote or read.  OFF by default (DIT4SR_PDL=1 turns it
// on): at the C2 shape it measured 43.3 ms/step against 43.0 without (the GEMM / attention CTAs fill an SM, so a
// successor cannot become resident before its predecessor drains) and one model-level parity number moved with it,
// which has not been explained; with a plain launch griddepcontrol.wait / launch_dependents are no-ops.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace host
