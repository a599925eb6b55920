// Library-level entry points of libelastic_b200 + host helpers shared by the kernels' launchers.
#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace ed {
thread_local int g_last_cuda_error = 0;

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs, from several threads).
int current_sm_count(int* sms) {
  static std::atomic<int> cache[64];
  int dev = 0;
  ED_CUDA_CHECK(cudaGetDevice(&dev));
  int v = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (!v) {
    ED_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) cache[dev].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return ED_OK;
}

int encode_tmap_3d_f32(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                       uint32_t b1, uint32_t b2) {
  return encode_tmap_3d(map, base, ED_F32, d0, d1, d2, b0, b1, b2);
}

int encode_tmap_3d(CUtensorMap* map, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                   uint32_t b1, uint32_t b2) {
  const uint64_t es = dtype == ED_F32 ? 4 : 2;
  const CUtensorMapDataType dt = dtype == ED_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == ED_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  // TMA constraints: 16-byte aligned base and strides, inner box a multiple of 16 bytes, box dims <= 256.
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((d0 * es) & 15) || ((b0 * es) & 15) || b0 > 256 || b1 > 256 ||
      b2 > 256 || b0 == 0 || b1 == 0 || b2 == 0)
    return ED_ERR_UNSUPPORTED;
  auto fn = get_encode_fn();
  if (!fn) return ED_ERR_NO_DEVICE;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * es, d0 * d1 * es};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_last_cuda_error = (int)r;
    return ED_ERR_CUDA;
  }
  return ED_OK;
}
}  // namespace ed

extern "C" {

int ed_abi_version(void) { return ED_ABI_VERSION; }

const char* ed_strerror(int status) {
  switch (status) {
    case ED_OK: return "ok";
    case ED_ERR_INVALID: return "invalid argument";
    case ED_ERR_UNSUPPORTED: return "unsupported shape / alignment for this kernel";
    case ED_ERR_CUDA: return "CUDA error (see ed_last_cuda_error)";
    case ED_ERR_NO_DEVICE: return "no sm_100 CUDA device / driver available";
    default: return "unknown status";
  }
}

int ed_last_cuda_error(void) { return ed::g_last_cuda_error; }

int ed_device_check(int dev, int* sm_count) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || dev < 0 || dev >= n) {
    cudaGetLastError();
    return ED_ERR_NO_DEVICE;
  }
  int major = 0, sms = 0;
  ED_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ED_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = sms;
  return major == 10 ? ED_OK : ED_ERR_NO_DEVICE;
}

int ed_upload_step_params(void* d_params, const ed_step_params_t* h_params, void* stream) {
  if (!d_params || !h_params) return ED_ERR_INVALID;
  ED_CUDA_CHECK(cudaMemcpyAsync(d_params, h_params, sizeof(ed_step_params_t), cudaMemcpyHostToDevice,
                                static_cast<cudaStream_t>(stream)));
  return ED_OK;
}

}  // extern "C"
