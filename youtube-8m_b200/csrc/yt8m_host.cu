// yt8m_b200 -- error reporting, version, TMA tensor-map construction (driver entry point resolved at
// run time so the library links without libcuda and loads on a GPU-less build box).
#include "yt8m_host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <atomic>
#include <mutex>

namespace yt8m {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void* lib_scratch(int slot, size_t bytes, size_t ticket_bytes, cudaStream_t stream) {
  constexpr int kMaxDev = 16;
  static void* buf[kMaxDev][kScratchSlots] = {};
  static size_t cap[kMaxDev][kScratchSlots] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kScratchSlots) {
    set_error("lib_scratch: device ordinal %d / slot %d out of range", dev, slot);
    return nullptr;
  }
  if (cap[dev][slot] < bytes) {
    cudaStreamSynchronize(stream);
    if (buf[dev][slot]) cudaFree(buf[dev][slot]);
    buf[dev][slot] = nullptr;
    cap[dev][slot] = 0;
    const size_t want = bytes + bytes / 2;                 // head-room: shapes grow a few times, not every call
    if (cudaMalloc(&buf[dev][slot], want) != cudaSuccess || cudaMemset(buf[dev][slot], 0, ticket_bytes) != cudaSuccess) {
      set_error("lib_scratch: cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(cudaGetLastError()));
      buf[dev][slot] = nullptr;
      return nullptr;
    }
    cap[dev][slot] = want;
  }
  return buf[dev][slot];
}

int& host_debug_flags() {
  static int flags = 0;
  return flags;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return YT8M_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return YT8M_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  auto enc = get_encode();
  YT8M_REQUIRE(enc != nullptr, YT8M_E_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  YT8M_REQUIRE(aligned16(ptr), YT8M_E_BADPTR, "TMA operand %p is not 16-byte aligned", ptr);
  YT8M_REQUIRE((ld_elems * 2) % 16 == 0, YT8M_E_BADSHAPE, "TMA row stride %llu elements is not a multiple of 8",
               (unsigned long long)ld_elems);
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YT8M_REQUIRE(r == CUDA_SUCCESS, YT8M_E_CUDA, "cuTensorMapEncodeTiled(2D %llux%llu ld %llu) failed: %d",
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, (int)r);
  return YT8M_OK;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1_elems,
                      uint64_t ld2_elems, uint32_t box_d1, uint32_t box_d0) {
  auto enc = get_encode();
  YT8M_REQUIRE(enc != nullptr, YT8M_E_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  YT8M_REQUIRE(aligned16(ptr), YT8M_E_BADPTR, "TMA operand %p is not 16-byte aligned", ptr);
  YT8M_REQUIRE((ld1_elems * 2) % 16 == 0 && (ld2_elems * 2) % 16 == 0, YT8M_E_BADSHAPE,
               "TMA strides must be multiples of 8 elements");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {ld1_elems * 2, ld2_elems * 2};
  cuuint32_t box[3] = {box_d0, box_d1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YT8M_REQUIRE(r == CUDA_SUCCESS, YT8M_E_CUDA, "cuTensorMapEncodeTiled(3D) failed: %d", (int)r);
  return YT8M_OK;
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  auto enc = get_encode();
  YT8M_REQUIRE(enc != nullptr, YT8M_E_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  YT8M_REQUIRE(rank >= 2 && rank <= 5, YT8M_E_BADSHAPE, "tensor map rank %d", rank);
  YT8M_REQUIRE(aligned16(ptr), YT8M_E_BADPTR, "TMA operand %p is not 16-byte aligned", ptr);
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) {
    YT8M_REQUIRE(strides_bytes[i] % 16 == 0, YT8M_E_BADSHAPE, "TMA stride %llu bytes is not a multiple of 16",
                 (unsigned long long)strides_bytes[i]);
    gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  YT8M_REQUIRE(r == CUDA_SUCCESS, YT8M_E_CUDA, "cuTensorMapEncodeTiled(rank %d) failed: %d", rank, (int)r);
  return YT8M_OK;
}

}  // namespace yt8m

extern "C" {
int yt8m_version(void) { return 100; }  // 0.1.0
const char* yt8m_last_error(void) { return yt8m::g_err; }
long long yt8m_launch_count(void) { return yt8m::g_launches.load(std::memory_order_relaxed); }
}
