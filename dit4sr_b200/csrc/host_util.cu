#include "host_util.h"
#include <stdlib.h>

#include <mutex>
#include <string.h>

namespace host {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

static int make_tmap_impl(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                          CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap_impl(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap_impl(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_bf16_noswizzle(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_impl(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

static int make_tmap_impl(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                          CUtensorMapDataType dt) {
  EncodeTiledFn enc = get_encode();
  DIT4SR_REQUIRE(enc != nullptr, DIT4SR_E_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  DIT4SR_REQUIRE(aligned16(base), DIT4SR_E_ALIGN, "TMA base pointer %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      DIT4SR_REQUIRE((strides_bytes[i - 1] & 15u) == 0, DIT4SR_E_ALIGN, "TMA stride %llu not a multiple of 16 bytes",
                     (unsigned long long)strides_bytes[i - 1]);
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIT4SR_REQUIRE(r == CUDA_SUCCESS, DIT4SR_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r,
                 rank);
  return 0;
}

int sm_count() {
  static int cache[64] = {};   // per device: a process may drive several GPUs
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cache[dev] != 0) return cache[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DIT4SR_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;   // opt-in: measured slightly slower at C2 (43.3 vs 43.0 ms), see DESIGN.md
  }
  return on != 0;
}

bool first_use_on_device(unsigned long long* mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;   // beyond the mask: configure every time (cheap, idempotent)
  if (*mask >> dev & 1ull) return false;
  *mask |= 1ull << dev;                     // benign race: the attribute set is idempotent
  return true;
}

}  // namespace host

extern "C" {

int dit4sr_abi_version(void) { return DIT4SR_ABI_VERSION; }

const char* dit4sr_last_error(void) { return host::g_err; }

int dit4sr_check_device(void) {
  int dev = 0;
  DIT4SR_CUDA(cudaGetDevice(&dev));
  int major = 0;
  DIT4SR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DIT4SR_REQUIRE(major == 10, DIT4SR_E_ARCH, "device compute capability %d.x is not 10.x (sm_100a kernels only)", major);
  return 0;
}
}
