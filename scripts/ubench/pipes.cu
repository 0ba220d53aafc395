// Pipe-throughput microbenchmarks for the attention/GELU design decisions (B200, sm_100a):
// MUFU ex2 f32 / f16x2, tanh f32 / f16x2, HFMA2, mma.sync m16n8k16 bf16, tcgen05.ld 32x32b.x32.
// Prints per-SM-per-clock rates.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024) alu_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  uint32_t r[ILP];
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { r[i] = seed + threadIdx.x * 7 + i; f[i] = -0.001f * (float)(threadIdx.x + i + seed); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 4) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(r[i]));
      if (OP == 5) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
      if (OP == 6) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "+r"(r[i]) : "f"(f[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= r[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(1024) mma_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  float d[4][4];
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b0 = seed + 5, b1 = seed + 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// tcgen05.ld throughput: `warps` warps each read 32 lanes x 32 columns per instruction, back to back
template <int X>
__global__ void __launch_bounds__(512) ldtm_kernel(uint32_t* out, long long* cycles, int wait_every) {
  __shared__ uint32_t tmem_addr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_addr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_addr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    uint32_t r[32];
    const uint32_t col = (uint32_t)((it * X) & 255) + (uint32_t)((warp >> 2) & 1) * 256;
    if (X == 32) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(base + col)
          : "memory");
    } else {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(base + col)
          : "memory");
#pragma unroll
      for (int i = 16; i < 32; ++i) r[i] = 0;
    }
    if ((it % wait_every) == wait_every - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // the loaded registers must be consumed or the load is dead; xor two of them after the wait only
    if ((it % wait_every) == wait_every - 1) acc ^= r[0] ^ r[X - 1];
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_addr) : "memory");
  }
}


// Do the legacy HMMA pipe and the MUFU pipe run concurrently?  Warps with (warp >> 2) & 1 == 0 issue
// mma.sync, the others ex2 (every SM sub-partition gets both kinds).  mode 1: mma only, 2: ex2 only, 3: both.
__global__ void __launch_bounds__(1024) mix_kernel(uint32_t* out, long long* cycles, uint32_t seed, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool is_mma = ((warp >> 2) & 1) == 0;
  float d[4][4];
  float f[ILP];
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b0 = seed + 5, b1 = seed + 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) f[i] = -0.001f * (float)(threadIdx.x + i + seed);
  __syncthreads();
  const long long t0 = clock64();
  if (is_mma && (mode & 1)) {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
  } else if (!is_mma && (mode & 2)) {
#pragma unroll 1
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc += d[i][0] + d[i][1] + d[i][2] + d[i][3];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc);
  __shared__ long long tmax[2];
  if (threadIdx.x < 2) tmax[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&tmax[is_mma ? 0 : 1], (unsigned long long)(t1 - t0));
  __syncthreads();
  if (threadIdx.x == 0) { cycles[2 * blockIdx.x] = tmax[0]; cycles[2 * blockIdx.x + 1] = tmax[1]; }
}

static double avg_cycles(long long* d_cycles, int blocks) {
  long long h[1024];
  cudaMemcpy(h, d_cycles, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < blocks; ++i) s += (double)h[i];
  return s / blocks;
}

int main() {
  int sms = 0;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  uint32_t* out;
  long long* cyc;
  CHECK(cudaMalloc(&out, (size_t)sms * 1024 * 4));
  CHECK(cudaMalloc(&cyc, (size_t)sms * 8));
  const char* names[] = {"ex2.f32", "ex2.f16x2", "tanh.f32", "tanh.f16x2", "fma.f16x2", "fma.f32", "ex2.bf16x2", "cvt.f16x2.f32"};
#define RUN_ALU(OP)                                                                                   \
  for (int threads = 256; threads <= 1024; threads *= 2) {                                            \
    alu_kernel<OP><<<sms, threads>>>(out, cyc, 1);                                                    \
    CHECK(cudaDeviceSynchronize());                                                                   \
    alu_kernel<OP><<<sms, threads>>>(out, cyc, 2);                                                    \
    CHECK(cudaDeviceSynchronize());                                                                   \
    const double c = avg_cycles(cyc, sms);                                                            \
    printf("%-14s threads/SM %4d: %.2f lane-instr/clk/SM\n", names[OP], threads, (double)threads * ITERS * ILP / c); \
  }
  RUN_ALU(0) RUN_ALU(1) RUN_ALU(2) RUN_ALU(3) RUN_ALU(4) RUN_ALU(5) RUN_ALU(6) RUN_ALU(7)
  for (int threads = 128; threads <= 1024; threads *= 2) {
    mma_kernel<<<sms, threads>>>(out, cyc, 1);
    CHECK(cudaDeviceSynchronize());
    mma_kernel<<<sms, threads>>>(out, cyc, 2);
    CHECK(cudaDeviceSynchronize());
    const double c = avg_cycles(cyc, sms);
    const double mmas = (double)(threads / 32) * ITERS * 4;
    printf("mma.m16n8k16   threads/SM %4d: %.3f mma/clk/SM = %.0f flop/clk/SM\n", threads, mmas / c, mmas * 4096 / c);
  }
  for (int wait_every = 1; wait_every <= 4; wait_every *= 4)
    for (int threads = 128; threads <= 512; threads *= 2) {
      ldtm_kernel<32><<<sms, threads>>>(out, cyc, wait_every);
      CHECK(cudaDeviceSynchronize());
      double c = avg_cycles(cyc, sms);
      printf("tcgen05.ld x32 warps/SM %2d wait_every %d: %.1f B/clk/SM\n", threads / 32, wait_every, (double)(threads / 32) * ITERS * 32 * 32 * 4 / c);
      ldtm_kernel<16><<<sms, threads>>>(out, cyc, wait_every);
      CHECK(cudaDeviceSynchronize());
      c = avg_cycles(cyc, sms);
      printf("tcgen05.ld x16 warps/SM %2d wait_every %d: %.1f B/clk/SM\n", threads / 32, wait_every, (double)(threads / 32) * ITERS * 32 * 16 * 4 / c);
    }
  {
    long long* cyc2;
    CHECK(cudaMalloc(&cyc2, (size_t)sms * 16));
    for (int mode = 1; mode <= 3; ++mode) {
      mix_kernel<<<sms, 512>>>(out, cyc2, 1, mode);
      CHECK(cudaDeviceSynchronize());
      mix_kernel<<<sms, 512>>>(out, cyc2, 2, mode);
      CHECK(cudaDeviceSynchronize());
      long long h[2048];
      cudaMemcpy(h, cyc2, sms * 16, cudaMemcpyDeviceToHost);
      double cm = 0, cx = 0;
      for (int i = 0; i < sms; ++i) { cm += (double)h[2 * i]; cx += (double)h[2 * i + 1]; }
      cm /= sms; cx /= sms;
      // 8 mma warps x ITERS x 4 mma; 8 ex2 warps x ITERS/2 x ILP x 32 lanes
      printf("mix mode %d (1 mma, 2 ex2, 3 both): mma warps %.0f cycles (%.3f mma/clk/SM), ex2 warps %.0f cycles (%.2f ex2/clk/SM)\n", mode,
             cm, 8.0 * ITERS * 4 / cm, cx, 8.0 * 32 * (ITERS / 2) * ILP / cx);
    }
  }
  printf("done\n");
  return 0;
}
