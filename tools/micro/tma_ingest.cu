// Microbenchmark: per-SM TMA ingest rate vs box shape / ring depth / producer count / source residency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I youtube-8m_b200/csrc -o /tmp/tma_ingest tools/micro/tma_ingest.cu
#include "yt8m_common.cuh"
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace yt8m;

template <int BOX_ROWS>
__global__ void __launch_bounds__(128, 1)
ingest_kernel(const __grid_constant__ CUtensorMap tm, int slots, int iters, int rows_total, int nprod, unsigned long long* out_ns) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  constexpr int kSlot = BOX_ROWS * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint64_t* empty = full + 32;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const unsigned long long t0 = global_timer_ns();
  if (warp < nprod) {
    // producer warp p handles slots p, p+nprod, ...
    uint32_t phase = 0;
    int slot = warp;
    const int row_tiles = rows_total / BOX_ROWS;
    for (int it = warp; it < iters; it += nprod) {
      mbar_wait(&empty[slot], phase ^ 1u);
      const int tile = (blockIdx.x * 977 + it) % row_tiles;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[slot], kSlot);
        tma_load_2d(smem + slot * kSlot, &tm, &full[slot], 0, tile * BOX_ROWS, kEvictNormal);
      }
      __syncwarp();
      slot += nprod;
      if (slot >= slots) { slot -= slots; phase ^= 1u; }
    }
  } else if (warp == 3) {
    uint32_t phase = 0;
    int slot = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[slot], phase);
      if (lane == 0) mbar_arrive(&empty[slot]);
      __syncwarp();
      if (++slot == slots) { slot = 0; phase ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_ns[blockIdx.x] = global_timer_ns() - t0;
}

static PFN_cuTensorMapEncodeTiled_v12000 enc() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (PFN_cuTensorMapEncodeTiled_v12000)p;
}

template <int BOX_ROWS>
void run(const void* buf, long long rows_total, int grid, int slots, int nprod, const char* label) {
  CUtensorMap tm;
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows_total}; cuuint64_t gstr[1] = {128}; cuuint32_t box[2] = {64, BOX_ROWS}; cuuint32_t es[2] = {1, 1};
  enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  unsigned long long* d_ns; cudaMalloc(&d_ns, grid * 8);
  const int iters = 2048 * 128 / BOX_ROWS;
  auto k = ingest_kernel<BOX_ROWS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<<<grid, 128, 202 * 1024>>>(tm, slots, iters, (int)rows_total, nprod, d_ns);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<unsigned long long> ns(grid); cudaMemcpy(ns.data(), d_ns, grid * 8, cudaMemcpyDeviceToHost);
  double mx = 0; for (auto v : ns) mx = v > mx ? v : mx;
  const double bytes = (double)iters * BOX_ROWS * 128;
  printf("%-28s box=%3dx128B slots=%2d (%3d KB in flight) prod=%d grid=%3d : %7.1f GB/s/SM  %6.0f ns/op  aggregate %6.2f TB/s  %s\n", label, BOX_ROWS,
         slots, slots * BOX_ROWS * 128 / 1024, nprod, grid, bytes / mx, mx / (iters / (double)nprod) , bytes * grid / mx / 1e3, cudaGetErrorString(e));
  cudaFree(d_ns);
}

int main() {
  void *big, *small;
  const long long big_rows = (1LL << 30) / 128, small_rows = (32LL << 20) / 128;   // 1 GiB (DRAM) and 32 MiB (L2 resident)
  cudaMalloc(&big, big_rows * 128); cudaMalloc(&small, small_rows * 128);
  cudaMemset(big, 1, big_rows * 128); cudaMemset(small, 1, small_rows * 128);
  for (int grid : {1, 148}) {
    run<128>(small, small_rows, grid, 8, 1, "L2-resident");
    run<128>(small, small_rows, grid, 12, 1, "L2-resident");
    run<128>(small, small_rows, grid, 12, 2, "L2-resident");
    run<128>(small, small_rows, grid, 12, 3, "L2-resident");
    run<256>(small, small_rows, grid, 6, 1, "L2-resident");
    run<64>(small, small_rows, grid, 24, 1, "L2-resident");
    run<64>(small, small_rows, grid, 24, 3, "L2-resident");
    run<128>(big, big_rows, grid, 8, 1, "DRAM");
    run<128>(big, big_rows, grid, 12, 1, "DRAM");
    run<128>(big, big_rows, grid, 12, 3, "DRAM");
    run<256>(big, big_rows, grid, 6, 1, "DRAM");
    run<256>(big, big_rows, grid, 6, 3, "DRAM");
  }
  return 0;
}
