// Error slot, device queries and TMA tensor-map encoders for the C-ABI library.
#include "common.cuh"

namespace me {

static thread_local char g_err[512] = {0};

char* error_slot() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};   // per device ordinal
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev < 0 || dev >= 64 || cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev < 0 || dev >= 64) return n;
    cached[dev] = n;
  }
  return cached[dev];
}

bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*encode_im2col_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void* driver_symbol(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
  if (q != cudaDriverEntryPointSuccess) return nullptr;
  return fn;
}

// Drivers up to 13.1 mis-encode maps over tensors smaller than 128 KiB (CUTLASS carries the
// same fix-up in cute/atom/copy_traits_sm90_tma.hpp): clear bit 21 of the second qword.
static void small_tensor_fixup(CUtensorMap* map, uint64_t span_bytes) {
  static int drv = -1;
  if (drv < 0) {
    int v = 0;
    drv = (cudaDriverGetVersion(&v) == cudaSuccess) ? v : 0;
  }
  if (drv <= 13010 && span_bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
}

int encode_tiled_2d(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, uint64_t cols,
                    uint64_t rows, uint64_t pitch_elems, uint32_t box_cols, uint32_t box_rows,
                    CUtensorMapSwizzle swz) {
  static encode_tiled_fn fn = reinterpret_cast<encode_tiled_fn>(driver_symbol("cuTensorMapEncodeTiled"));
  if (!fn) return fail(ME_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(ME_ERR_ARG, "tensor base not 16B aligned");
  if ((pitch_elems * esize) % 16 != 0) return fail(ME_ERR_ARG, "row pitch %llu B not a multiple of 16",
                                                   (unsigned long long)(pitch_elems * esize));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * static_cast<uint64_t>(esize)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ME_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu pitch=%llu box=%ux%u", (int)r,
                (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems, box_cols,
                box_rows);
  small_tensor_fixup(map, rows * pitch_elems * esize);
  return ME_OK;
}

int encode_im2col_nhwc(CUtensorMap* map, const void* base, int n, int h, int w, int c, int pitch_elems, int ksize,
                       int pad, int stride, uint32_t channels_per_pixel, uint32_t pixels_per_col,
                       CUtensorMapSwizzle swz) {
  static encode_im2col_fn fn = reinterpret_cast<encode_im2col_fn>(driver_symbol("cuTensorMapEncodeIm2col"));
  if (!fn) return fail(ME_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(ME_ERR_ARG, "tensor base not 16B aligned");
  if ((pitch_elems * 2) % 16 != 0) return fail(ME_ERR_ARG, "pixel pitch %d B not a multiple of 16", pitch_elems * 2);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)pitch_elems * 2, (cuuint64_t)w * pitch_elems * 2,
                           (cuuint64_t)h * w * pitch_elems * 2};
  // Base pixels run over [lower, dim + upper): lower = -pad, upper = pad - (ksize-1)
  // (fprop convention of cutlass/conv/collective/detail.hpp compute_{lower,upper}_corner_whd).
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper,
                  channels_per_pixel, pixels_per_col, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ME_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d): n=%d h=%d w=%d c=%d pitch=%d k=%d s=%d", (int)r, n,
                h, w, c, pitch_elems, ksize, stride);
  small_tensor_fixup(map, (uint64_t)n * h * w * pitch_elems * 2);
  return ME_OK;
}

}  // namespace me

extern "C" {

int me_version(void) { return 100; }

int me_last_error(char* buf, size_t n) {
  const char* e = me::error_slot();
  size_t len = strlen(e);
  if (buf && n > 0) {
    size_t k = len < n - 1 ? len : n - 1;
    memcpy(buf, e, k);
    buf[k] = 0;
  }
  return static_cast<int>(len);
}

int me_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  ME_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  ME_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ME_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  ME_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) return me::fail(ME_ERR_UNSUPPORTED, "millieye_b200 needs sm_100 (found sm_%d%d)", maj, min);
  return ME_OK;
}

}  // extern "C"

// ---- stream memory operations + peer copies (millieye_b200/dist.py::PeerGather) ----------------------------------------
extern "C" {

// Blocks `stream` (no SM is used) until *(volatile uint32*)addr >= value.  addr: 4-byte aligned device memory of the current
// device.  The flag is typically written by another GPU (me_peer_copy / me_stream_write_value32 on the peer's stream).
int me_stream_wait_value32(const void* addr, unsigned int value, me_stream_t stream) {
  using namespace me;
  typedef CUresult (*wait_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static wait_fn fn = reinterpret_cast<wait_fn>(driver_symbol("cuStreamWaitValue32"));
  if (!fn) return fail(ME_ERR_CUDA, "cuStreamWaitValue32 entry point not available");
  ME_REQUIRE(addr && (reinterpret_cast<uintptr_t>(addr) & 3) == 0, "stream_wait_value32: bad address");
  const CUresult r = fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) return fail(ME_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", static_cast<int>(r));
  return ME_OK;
}

// Stream-ordered 4-byte store without a kernel.
int me_stream_write_value32(void* addr, unsigned int value, me_stream_t stream) {
  using namespace me;
  typedef CUresult (*write_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  static write_fn fn = reinterpret_cast<write_fn>(driver_symbol("cuStreamWriteValue32"));
  if (!fn) return fail(ME_ERR_CUDA, "cuStreamWriteValue32 entry point not available");
  ME_REQUIRE(addr && (reinterpret_cast<uintptr_t>(addr) & 3) == 0, "stream_write_value32: bad address");
  const CUresult r = fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(addr), value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) return fail(ME_ERR_CUDA, "cuStreamWriteValue32 failed (%d)", static_cast<int>(r));
  return ME_OK;
}

// Lets the CURRENT device read and write allocations of `peer_device` directly (NVLink / PCIe peer-to-peer).  Without it a
// copy into a peer's IPC-mapped buffer is staged through the host.  Already enabled / same device: ME_OK.
int me_peer_enable(int peer_device) {
  using namespace me;
  int cur = -1;
  ME_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return ME_OK;
  int can = 0;
  ME_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) return fail(ME_ERR_UNSUPPORTED, "device %d cannot access device %d directly", cur, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return ME_OK;
  }
  ME_CUDA(e);
  return ME_OK;
}

// cudaMemcpyAsync between two device allocations that may live on different GPUs (unified addressing + peer access, e.g. a
// buffer of another process opened through CUDA IPC): a copy-engine transfer over NVLink, ordered on `stream` only.
int me_peer_copy(void* dst, const void* src, size_t bytes, me_stream_t stream) {
  using namespace me;
  ME_REQUIRE(dst && src, "peer_copy: null argument");
  if (bytes == 0) return ME_OK;
  ME_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return ME_OK;
}

}  // extern "C"
