// Microbenchmark: back-to-back tcgen05.mma issue rate for K-major vs MN-major operands (operands resident in smem).
#include "yt8m_common.cuh"
#include <cstdio>
using namespace yt8m;

template <int N, int AMN, int BMN>
__global__ void __launch_bounds__(128, 1) umma_rate(int iters, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, AMN, BMN);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    const uint64_t ad = AMN ? make_sdesc_sw128(a, 16384, 1024) : make_sdesc_sw128(a, 16, 1024);
    const uint64_t bd = BMN ? make_sdesc_sw128(b, 16384, 1024) : make_sdesc_sw128(b, 16, 1024);
    t0 = global_timer_ns();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm, sdesc_advance(ad, AMN ? k * 2048 : k * 32), sdesc_advance(bd, BMN ? k * 2048 : k * 32), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = global_timer_ns();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int AMN, int BMN>
void run(const char* label) {
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  auto k = umma_rate<N, AMN, BMN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int grid : {1, 148}) {
    k<<<grid, 128, 100 * 1024>>>(iters, d);
    k<<<grid, 128, 100 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long ns[148]; cudaMemcpy(ns, d, grid * 8, cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int i = 0; i < grid; ++i) mx = ns[i] > mx ? ns[i] : mx;
    const double per = (double)mx / (iters * 4);
    printf("%-22s M=128 N=%3d grid=%3d : %6.1f ns/UMMA  -> %6.1f TFLOP/s/SM-equivalent chip %7.1f TFLOP/s  %s\n", label, N, grid, per,
           2.0 * 128 * N * 16 / per / 1e3, 2.0 * 128 * N * 16 / per / 1e3 * 148, cudaGetErrorString(e));
  }
  cudaFree(d);
}

int main() {
  run<64, 0, 0>("K-major A, K-major B");
  run<64, 1, 1>("MN-major A, MN-major B");
  run<64, 1, 0>("MN-major A, K-major B");
  run<64, 0, 1>("K-major A, MN-major B");
  run<128, 0, 0>("K-major A, K-major B");
  run<128, 1, 1>("MN-major A, MN-major B");
  run<256, 0, 0>("K-major A, K-major B");
  run<256, 1, 1>("MN-major A, MN-major B");
  return 0;
}
