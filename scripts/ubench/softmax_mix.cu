// Pipe rates behind the tcgen05 attention kernel's softmax loop (B200, sm_100a):
//   - F2FP (cvt.rn.bf16x2.f32), FMNMX3 (max.f32 a,b,c), FFMA2 / FADD2 (fma/add.rn.f32x2), LEA-style integer ops
//   - the loop body itself, from registers:  fma2 -> 4 x ex2 -> add2 x2 -> pack x2   (optionally every POLY-th pair by
//     the degree-3 polynomial), with 4, 8 or 16 warps per SM -- the ceiling of weights per clock per SM for that mix
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_mix softmax_mix.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ uint64_t pack2f(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

// OP: 0 F2FP, 1 FMNMX3, 2 FFMA2, 3 FADD2, 4 shl+add (LEA)
template <int OP>
__global__ void __launch_bounds__(1024) op_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  float f[8];
  uint32_t r[8];
  uint64_t d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = -0.001f * (float)(threadIdx.x + i + seed); r[i] = seed + i; d[i] = pack2f(f[i], f[i] * 0.5f); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) r[i] ^= packbf(f[i], __uint_as_float(r[i] | 0x3f000000u));
      if (OP == 1) f[i] = max3(f[i], f[(i + 1) & 7], -1.0f);
      if (OP == 2) d[i] = fma2(d[i], d[i], d[i]);
      if (OP == 3) d[i] = add2(d[i], d[i]);
      if (OP == 4) r[i] = (r[i] << 23) + r[(i + 1) & 7];
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; unpack2f(d[i], a, b); acc ^= r[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(a + b); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the loop kernels use NON-volatile forms (the compiler may schedule them like it does in the attention kernel)
__device__ __forceinline__ uint64_t fma2n(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2n(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2n(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3n(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ uint32_t round_bf16x2(float lo, float hi) { return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632); }

// the attention kernel's loop body as shipped: 8 max chains, fma2 -> ex2 -> add2 (row sum) -> add-half-ulp + PRMT
// MODE 0: all of it; 1: without the row-sum additions; 2: without the packing; 3: exponentials only (fma2 + ex2)
template <int MODE>
__global__ void __launch_bounds__(512) loop2_kernel(uint32_t* out, long long* cycles, uint32_t seed, int blocks) {
  float r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13 + seed) & 1023);
  const uint64_t scale2 = pack2f(0.255f, 0.255f);
  uint32_t acc = 0;
  float l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int b = 0; b < blocks; ++b) {
    float mx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) mx[c] = max3n(r[3 * c], r[3 * c + 1], r[3 * c + 2]);
#pragma unroll
    for (int i = 24; i + 15 < 128; i += 16)
#pragma unroll
      for (int c = 0; c < 8; ++c) mx[c] = max3n(mx[c], r[i + 2 * c], r[i + 2 * c + 1]);
#pragma unroll
    for (int c = 0; c < 4; ++c) mx[c] = max3n(mx[c], r[120 + 2 * c], r[121 + 2 * c]);
    const float mb = max3n(max3n(mx[0], mx[1], mx[2]), max3n(mx[3], mx[4], mx[5]), fmaxf(mx[6], mx[7]));
    const float off = mb * 0.255f;
    const uint64_t noff2 = pack2f(-off, -off);
    uint64_t l2a = pack2f(0.f, 0.f), l2b = l2a;
    uint32_t pk[64];
    float keep = 0.f;
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      float p0, p1, p2, p3;
      unpack2f(fma2n(pack2f(r[i], r[i + 1]), scale2, noff2), p0, p1);
      unpack2f(fma2n(pack2f(r[i + 2], r[i + 3]), scale2, noff2), p2, p3);
      p0 = ex2n(p0); p1 = ex2n(p1); p2 = ex2n(p2); p3 = ex2n(p3);
      if (MODE == 0 || MODE == 2) { l2a = add2n(l2a, pack2f(p0, p1)); l2b = add2n(l2b, pack2f(p2, p3)); }
      if (MODE == 0 || MODE == 1) { pk[i >> 1] = round_bf16x2(p0, p1); pk[(i >> 1) + 1] = round_bf16x2(p2, p3); }
      else { pk[i >> 1] = __float_as_uint(p0) ^ __float_as_uint(p1); pk[(i >> 1) + 1] = __float_as_uint(p2) ^ __float_as_uint(p3); }
    }
    float la, lb, lc, ld;
    unpack2f(l2a, la, lb);
    unpack2f(l2b, lc, ld);
    l += (la + lb) + (lc + ld) + keep;
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= pk[i];
#pragma unroll
    for (int i = 0; i < 128; i += 16) r[i] += __uint_as_float((acc & 0x7fu) | 0x3a000000u) * 1e-3f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(l);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void exp2_poly_pair(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2f(x2, x0, x1);
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t xc = pack2f(x0, x1);
  const uint64_t t2 = add2(xc, pack2f(12582912.f, 12582912.f));
  const uint64_t j2 = add2(t2, pack2f(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(j2, pack2f(-1.f, -1.f), xc);
  uint64_t q2 = fma2(f2, pack2f(0.05500893f, 0.05500893f), pack2f(0.24221095f, 0.24221095f));
  q2 = fma2(q2, f2, pack2f(0.6932829f, 0.6932829f));
  q2 = fma2(q2, f2, pack2f(1.f, 1.f));
  float t0, t1, q0, q1;
  unpack2f(t2, t0, t1);
  unpack2f(q2, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// the softmax loop body on 128 scores held in registers: max, then p = 2^(s c - m c), row sum, bf16 pairs
template <int POLY>
__global__ void __launch_bounds__(512) loop_kernel(uint32_t* out, long long* cycles, uint32_t seed, int blocks) {
  float r[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) r[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13 + seed) & 1023);
  const uint64_t scale2 = pack2f(0.255f, 0.255f);
  uint32_t acc = 0;
  float l = 0.f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int b = 0; b < blocks; ++b) {
    float mb = max3(r[0], r[1], r[2]), mb2 = max3(r[3], r[4], r[5]);
#pragma unroll
    for (int i = 6; i + 3 < 128; i += 4) { mb = max3(mb, r[i], r[i + 1]); mb2 = max3(mb2, r[i + 2], r[i + 3]); }
    mb = max3(mb, mb2, fmaxf(r[126], r[127]));
    const float off = mb * 0.255f;
    const uint64_t noff2 = pack2f(-off, -off);
    uint64_t l2a = pack2f(0.f, 0.f), l2b = l2a;
    uint32_t pk[64];
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
      float p0, p1, p2, p3;
      const uint64_t xa = fma2(pack2f(r[i], r[i + 1]), scale2, noff2);
      const uint64_t xb = fma2(pack2f(r[i + 2], r[i + 3]), scale2, noff2);
      if (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == 0) exp2_poly_pair(xa, p0, p1);
      else { unpack2f(xa, p0, p1); p0 = ex2(p0); p1 = ex2(p1); }
      if (POLY > 0 && (((i >> 1) + 1) % (POLY > 0 ? POLY : 1)) == 0) exp2_poly_pair(xb, p2, p3);
      else { unpack2f(xb, p2, p3); p2 = ex2(p2); p3 = ex2(p3); }
      l2a = add2(l2a, pack2f(p0, p1));
      l2b = add2(l2b, pack2f(p2, p3));
      pk[i >> 1] = packbf(p0, p1);
      pk[(i >> 1) + 1] = packbf(p2, p3);
    }
    float la, lb, lc, ld;
    unpack2f(l2a, la, lb);
    unpack2f(l2b, lc, ld);
    l += (la + lb) + (lc + ld);
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= pk[i];
    // the next block's scores depend on this one's (keeps the loop from being hoisted)
#pragma unroll
    for (int i = 0; i < 128; i += 16) r[i] += __uint_as_float((acc & 0x7fu) | 0x3a000000u) * 1e-3f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ __float_as_uint(l);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
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
  const char* names[] = {"F2FP.bf16x2", "FMNMX3", "FFMA2", "FADD2", "shl+add"};
#define RUN_OP(OP)                                                                                              \
  for (int threads = 256; threads <= 1024; threads *= 2) {                                                      \
    op_kernel<OP><<<sms, threads>>>(out, cyc, 1);                                                               \
    CHECK(cudaDeviceSynchronize());                                                                             \
    op_kernel<OP><<<sms, threads>>>(out, cyc, 2);                                                               \
    CHECK(cudaDeviceSynchronize());                                                                             \
    printf("%-12s threads/SM %4d: %.2f lane-instr/clk/SM\n", names[OP], threads, (double)threads * ITERS * 8 / avg_cycles(cyc, sms)); \
  }
  RUN_OP(0) RUN_OP(1) RUN_OP(2) RUN_OP(3) RUN_OP(4)
#define RUN_LOOP(POLY)                                                                                          \
  for (int threads = 128; threads <= 512; threads *= 2) {                                                       \
    loop_kernel<POLY><<<sms, threads>>>(out, cyc, 1, 64);                                                       \
    CHECK(cudaDeviceSynchronize());                                                                             \
    loop_kernel<POLY><<<sms, threads>>>(out, cyc, 2, 64);                                                       \
    CHECK(cudaDeviceSynchronize());                                                                             \
    const double c = avg_cycles(cyc, sms);                                                                      \
    printf("softmax loop poly 1/%d warps/SM %2d: %.2f weights/clk/SM (%.0f clk per 128-score block and warp)\n", POLY, threads / 32, \
           (double)threads * 64 * 128 / c, c / 64);                                                             \
  }
  RUN_LOOP(0) RUN_LOOP(4) RUN_LOOP(2)
  const char* modes[] = {"as shipped (max, fma2, ex2, add2, +half ulp, PRMT)", "without the row-sum additions", "without the bf16 packing", "fma2 + ex2 only"};
#define RUN_LOOP2(MODE)                                                                                         \
  for (int threads = 128; threads <= 512; threads *= 2) {                                                       \
    loop2_kernel<MODE><<<sms, threads>>>(out, cyc, 1, 64);                                                      \
    CHECK(cudaDeviceSynchronize());                                                                             \
    loop2_kernel<MODE><<<sms, threads>>>(out, cyc, 2, 64);                                                      \
    CHECK(cudaDeviceSynchronize());                                                                             \
    const double c = avg_cycles(cyc, sms);                                                                      \
    printf("scheduled loop, %s, warps/SM %2d: %.2f weights/clk/SM (%.0f clk per 128-score block and warp)\n", modes[MODE], threads / 32, \
           (double)threads * 64 * 128 / c, c / 64);                                                             \
  }
  RUN_LOOP2(0) RUN_LOOP2(1) RUN_LOOP2(2) RUN_LOOP2(3)
  printf("done\n");
  return 0;
}
