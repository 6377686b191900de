// Micro-benchmark (not part of the library): sustained cycles per tcgen05.mma for the operand
// configurations the attention kernel can choose between.  Build+run:  scripts/gpu_probe.sh
//   SS  : A and B from shared memory (K-major, SW128)     TS : A from TMEM
#include <cstdio>
#include <cstdlib>
#include "../jenga_b200/csrc/sm100_ptx.cuh"
using namespace jenga;

template <int N, bool kTS>
__global__ void __launch_bounds__(128) probe(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  uint8_t* smem = raw + (((a0 + 1023u) & ~1023u) - a0);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(true, false, 128, N);
    const uint64_t ad = umma_smem_desc(smem_u32(smem), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t bd = umma_smem_desc(smem_u32(smem + 32768), 16, 1024, UMMA_LAYOUT_SW128);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t off = static_cast<uint64_t>(((kk & 3) * 32 + (kk >> 2) * 16384) >> 4);
        if (kTS) umma_ts(tm, tm + 256 + kk * 8, bd + off, idesc, 1u);
        else umma_ss(tm, ad + off, bd + off, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, nullptr);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, bool kTS>
static void run(const char* name, int ctas) {
  long long* d; cudaMalloc(&d, 8);
  auto k = probe<N, kTS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  k<<<ctas, 128, 100 * 1024>>>(d, iters);
  k<<<ctas, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / (iters * 8.0);
  printf("%-28s ctas=%3d  %7.2f cycles/MMA  ideal %3d  eff %.2f  (%s)\n", name, ctas, per, N / 2, (N / 2) / per,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, false>("SS M128 N64  K16", 148);
  run<128, false>("SS M128 N128 K16", 148);
  run<256, false>("SS M128 N256 K16", 148);
  run<64, true>("TS M128 N64  K16", 148);
  run<128, true>("TS M128 N128 K16", 148);
  run<256, true>("TS M128 N256 K16", 148);
  return 0;
}
