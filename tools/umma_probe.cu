// Micro-benchmark (not part of the library): sustained cycles per tcgen05.mma for the operand
// configurations the attention kernel can choose between.  Build+run:  scripts/gpu_probe.sh
//   SS  : A and B from shared memory (K-major, SW128)     TS : A from TMEM
#include <cstdio>
#include <cstdlib>
#include <cstring>
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


// Interleaved patterns: does the shared-memory operand fetch of consecutive MMAs overlap, so that an
// N=64 SS MMA (6 KB of operands for 32 cycles of math) issued next to an N=128 TS MMA (4 KB for 64
// cycles) no longer pays the 128 B/clk port limit?   'P' = TS M128 N128 MN-major B (a P.V step),
// 'Q' = SS M128 N64 (a Q.K^T step).  The pattern is a compile-time string so that the issue loop is
// fully unrolled (a data-dependent issue loop is itself ~140 cycles per MMA).
template <int kPat>
struct Pat;
template <> struct Pat<0> { static constexpr const char* s = "PQQPQQPQQPQQ"; };
template <> struct Pat<1> { static constexpr const char* s = "PPPPQQQQQQQQ"; };
template <> struct Pat<2> { static constexpr const char* s = "QQQQQQQQ"; };
template <> struct Pat<3> { static constexpr const char* s = "PPPP"; };
template <> struct Pat<4> { static constexpr const char* s = "PQPQPQPQ"; };
template <> struct Pat<5> { static constexpr const char* s = "QPQQPQQPQQPQ"; };
template <> struct Pat<6> { static constexpr const char* s = "PPQQQQPPQQQQ"; };
constexpr int cstrlen(const char* s) { int n = 0; while (s[n]) ++n; return n; }

template <int kPat>
__global__ void __launch_bounds__(128) probe_pattern(long long* out, int iters) {
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
  if (threadIdx.x < 32 && elect_one()) {
    constexpr uint32_t idesc_q = umma_idesc_f16(true, false, 128, 64);
    constexpr uint32_t idesc_p = umma_idesc_f16(true, true, 128, 128);
    const uint64_t ad = umma_smem_desc(smem_u32(smem), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t bd = umma_smem_desc(smem_u32(smem + 32768), 16, 1024, UMMA_LAYOUT_SW128);
    const uint64_t vd = umma_smem_desc(smem_u32(smem + 65536), 8192, 1024, UMMA_LAYOUT_SW128);
    constexpr int plen = cstrlen(Pat<kPat>::s);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      int kq = 0, kp = 0;
#pragma unroll
      for (int i = 0; i < plen; ++i) {
        if (Pat<kPat>::s[i] == 'Q') {
          const uint64_t off = static_cast<uint64_t>((((kq & 3) * 32) + ((kq >> 2) & 1) * 16384) >> 4);
          umma_ss(tm + ((kq >> 3) & 1) * 64, ad + off, bd + off, idesc_q, 1u);
          ++kq;
        } else {
          const uint64_t off = static_cast<uint64_t>(((kp & 3) * 16 * 128) >> 4);
          umma_ts(tm + 128, tm + 256 + (kp & 3) * 8, vd + off, idesc_p, 1u);
          ++kp;
        }
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

template <int kPat>
static void run_pattern(int ctas) {
  long long* d; cudaMalloc(&d, 8);
  const char* pat = Pat<kPat>::s;
  const int plen = (int)strlen(pat);
  cudaFuncSetAttribute(probe_pattern<kPat>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 1000;
  probe_pattern<kPat><<<ctas, 128, 100 * 1024>>>(d, iters);
  probe_pattern<kPat><<<ctas, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  int nq = 0, np = 0;
  for (int i = 0; i < plen; ++i) (pat[i] == 'Q' ? nq : np)++;
  const double per = double(h) / iters;
  printf("pattern %-26s ctas=%3d  %8.1f cycles/rep  ideal %4d  burst-serial %4d  (%s)\n", pat, ctas, per,
         nq * 32 + np * 64, nq * 48 + np * 64, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, false>("SS M128 N64  K16", 148);
  run<128, false>("SS M128 N128 K16", 148);
  run<256, false>("SS M128 N256 K16", 148);
  run<64, true>("TS M128 N64  K16", 148);
  run<128, true>("TS M128 N128 K16", 148);
  run<256, true>("TS M128 N256 K16", 148);
  run_pattern<0>(148); run_pattern<1>(148); run_pattern<2>(148); run_pattern<3>(148);
  run_pattern<4>(148); run_pattern<5>(148); run_pattern<6>(148);
  return 0;
}
