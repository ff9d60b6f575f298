// yt8m_b200 -- host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/yt8m_b200.h"

namespace yt8m {

void set_error(const char* fmt, ...);

// bf16 row-major [rows, cols] (row stride ld_elems) -> 2D tensor map, SWIZZLE_128B, box = box_rows x 64 cols
int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols = 64);
// bf16 [d2, d1, d0] (d0 contiguous; strides in elements) -> 3D tensor map, box = 1 x box_d1 x 64
int make_tmap_bf16_3d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1_elems,
                      uint64_t ld2_elems, uint32_t box_d1, uint32_t box_d0 = 64);

// generic bf16 tensor map, SWIZZLE_128B: dims/box innermost first, strides (bytes) for dims 1..rank-1
int make_tmap_bf16_nd(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);

int check_launch(const char* what);

// acc += A . W^T with split-K atomics straight into acc (yt8m_gemm.cu).  pdl: launch with programmatic stream serialization (the
// kernel's prologue overlaps the previous kernel on the stream; it waits for that kernel's completion before touching memory)
int linear_accumulate(const yt8m_bf16* a_hi, const yt8m_bf16* a_lo, long long lda, const yt8m_bf16* w, long long ldw, int M, int N,
                      int K, float* acc, long long ld_acc, cudaStream_t stream, bool pdl = false);

// Library-owned device scratch, one buffer per (device, slot), grown on demand (the growth synchronises `stream` and frees the
// old buffer).  The first `ticket_bytes` are zero after allocation and are kept zero by their users (self-cleaning tickets of
// last-block reductions).  Users of one slot must be ordered on one stream per device.  Returns nullptr on failure
// (yt8m_last_error is set).
enum ScratchSlot : int { kScratchDcw2 = 0, kScratchColsum = 1, kScratchSlots = 2 };
void* lib_scratch(int slot, size_t bytes, size_t ticket_bytes, cudaStream_t stream);

// debug / experiment switches set through yt8m_debug_set_flags (host copy; 0 in normal operation)
int& host_debug_flags();

// device buffer registered through yt8m_debug_set_timeline (nullptr in normal operation)
unsigned long long*& host_debug_timeline();

// persistent LSTM recurrence (yt8m_lstm_rec.cu): one launch per layer and per lstm_rec_batch_chunk() videos
bool lstm_rec_supported(int H);                 // shape rule only (used to size the workspace)
bool lstm_rec_available(int H);                 // + the launch's clusters can be co-resident on this GPU
int lstm_rec_batch_chunk();
int launch_lstm_rec(const float* xw, const int* num_frames, int B, int T, int H, const yt8m_bf16* w_rec, long long ldw,
                    float forget_bias, yt8m_bf16* h_hi, yt8m_bf16* h_lo, float* out_seq, float* c_out, float* h_out,
                    long long ld_state, unsigned int* counters, cudaStream_t stream);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace yt8m

#define YT8M_REQUIRE(cond, code, ...)     \
  do {                                    \
    if (!(cond)) {                        \
      ::yt8m::set_error(__VA_ARGS__);     \
      return (code);                      \
    }                                     \
  } while (0)

#define YT8M_CUDA(expr)                                                                \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      ::yt8m::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));              \
      return YT8M_E_CUDA;                                                              \
    }                                                                                  \
  } while (0)
