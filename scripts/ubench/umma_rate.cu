// How many clocks does one tcgen05.mma (kind::f16, K = 16) take as a function of N, when both operands come from shared
// memory (SS) or A comes from tensor memory (TS)?  The floor is M*N*16*2 flop / 8192 flop per clock per SM (M = 128 per
// CTA); small N cannot reach it when the operand reads (A: 4 KB per instruction whatever N is) become the limit.  Decides
// the chunk width of the fused feed-forward kernel (drag_mlp.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ai-dial-rag_b200/csrc -o scripts/ubench/umma_rate scripts/ubench/umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "drag_tc.cuh"

using namespace drag;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int KBLOCKS = 6;          // A: 6 k-blocks of [128 rows][64] (96 KB), walked round robin like the X tile of the MLP kernel
constexpr int A_KB = 128 * 128;
constexpr int B_KBLOCKS = 2;        // B: two k-blocks of [<= 256 rows][64] (64 KB), alternating
constexpr int ITERS = 64;           // groups of 24 instructions (6 k-blocks x 4 k-steps)

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_pair(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// MODE 0: SS, MODE 1: TS (A = 8 columns of tensor memory per k-step).  PAIR: cta_group::2 (M = 256 over two CTAs, each
// holds N/2 rows of B).  ACCS accumulators are used round robin (1: one dependent chain).
template <int N, int MODE, bool PAIR, int ACCS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + KBLOCKS * A_KB;                 // [k-block][N (or N/2) rows][64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + B_KBLOCKS * 256 * 128);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (KBLOCKS * A_KB + B_KBLOCKS * 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  if (PAIR) tc::cluster_sync_all();
  if (warp == 0) {
    if (PAIR) { tc::tmem_alloc_pair(tmem_ptr, 512); tc::tmem_relinquish_pair(); }
    else { tc::tmem_alloc(tmem_ptr, 512); tc::tmem_relinquish(); }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
  if (warp == 1 && rank == 0 && tc::elect_one()) {
    const uint32_t idesc = tc::umma_idesc_bf16(PAIR ? 256 : 128, N);
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
      const uint32_t d = tmem + (uint32_t)((it % ACCS) * N);
#pragma unroll
      for (int kb = 0; kb < KBLOCKS; ++kb) {
        const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(a_s + kb * A_KB));
        const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(b_s + (kb % B_KBLOCKS) * 256 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 0) {
            if (PAIR) tc::umma_bf16_pair(d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            else tc::umma_bf16(d, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          } else {
            const uint32_t a_t = tmem + 448 + (uint32_t)((kb & 1) * 32 + k * 8);
            if (PAIR) umma_ts_pair(d, a_t, b_desc + 2 * k, idesc, (kb | k) != 0);
            else umma_ts(d, a_t, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
        }
      }
    }
    if (PAIR) tc::umma_commit_pair(bar, 1); else tc::umma_commit(bar);
    tc::mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync_all();
  if (warp == 0) {
    tc::tc_fence_after();
    if (PAIR) tc::tmem_dealloc_pair(tmem, 512); else tc::tmem_dealloc(tmem, 512);
  }
}

template <int N, int MODE, bool PAIR, int ACCS>
int run(const char* what, int ctas) {
  long long* d_out;
  CHECK(cudaMalloc(&d_out, 8));
  const size_t smem = KBLOCKS * A_KB + B_KBLOCKS * 256 * 128 + 64 + 1024;
  auto kern = rate_kernel<N, MODE, PAIR, ACCS>;
  CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  long long clocks = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CHECK(cudaLaunchKernelEx(&cfg, kern, d_out));
    CHECK(cudaDeviceSynchronize());
  }
  CHECK(cudaMemcpy(&clocks, d_out, 8, cudaMemcpyDeviceToHost));
  const double per = (double)clocks / (ITERS * KBLOCKS * 4);
  const double floor_clk = 128.0 * N * 16 * 2 / 8192.0;
  printf("%-44s N=%3d %s accs=%d ctas=%3d: %7.1f clocks per instruction (floor %5.1f) -> %5.1f%% of the tensor peak\n", what, N,
         PAIR ? "pair" : "cta ", ACCS, ctas, per, floor_clk, 100.0 * floor_clk / per);
  cudaFree(d_out);
  return 0;
}

int main() {
  // one CTA / one pair alone, and all SMs busy (shared-memory bandwidth is per SM: should not matter)
  run<64, 0, false, 1>("SS", 1);
  run<64, 0, false, 2>("SS", 1);
  run<128, 0, false, 1>("SS", 1);
  run<128, 0, false, 2>("SS", 1);
  run<256, 0, false, 1>("SS", 1);
  run<64, 0, true, 1>("SS", 2);
  run<64, 0, true, 2>("SS", 2);
  run<64, 0, true, 2>("SS", 148);
  run<96, 0, true, 2>("SS", 2);
  run<128, 0, true, 1>("SS", 2);
  run<128, 0, true, 2>("SS", 2);
  run<128, 0, true, 2>("SS", 148);
  run<192, 0, true, 2>("SS", 2);
  run<256, 0, true, 1>("SS", 2);
  run<256, 0, true, 1>("SS", 148);
  run<192, 1, true, 2>("TS (A from tensor memory)", 2);
  run<192, 1, true, 2>("TS (A from tensor memory)", 148);
  run<192, 1, false, 2>("TS (A from tensor memory)", 1);
  run<64, 1, true, 2>("TS (A from tensor memory)", 2);
  run<32, 1, false, 2>("TS (A from tensor memory)", 1);
  return 0;
}
