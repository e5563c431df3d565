// Instruction-throughput probe for the integer pipes of sm_100a (SASS-verified:
// tools/bin/inst_bench.sass.txt lists the loop mix of every kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/inst_bench tools/inst_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef uint64_t u64;
constexpr int CH = 12;
__device__ __forceinline__ unsigned lo32(u64 x) { return (unsigned)x; }
__device__ __forceinline__ unsigned hi32(u64 x) { return (unsigned)(x >> 32); }

struct OpImad { static constexpr const char* name = "IMAD narrow";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = lo32(x), h = hi32(x); l = l * a + h; h = h * b + l; x = ((u64)h << 32) | l; } };
struct OpHi { static constexpr const char* name = "IMAD.HI";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = lo32(x), h = hi32(x); l = __umulhi(l | 0x80000000u, a) + h; h = __umulhi(h | 0x80000000u, b) + l; x = ((u64)h << 32) | l; } };
struct OpWide { static constexpr const char* name = "IMAD.WIDE (acc)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { x = (u64)lo32(x) * a + x; x = (u64)hi32(x) * b + x; } };
struct OpWideNoAcc { static constexpr const char* name = "IMAD.WIDE (no acc)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 p = (u64)lo32(x) * a; u64 r = (u64)hi32(x) * b; x = p ^ r; } };
struct OpIadd3 { static constexpr const char* name = "IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = lo32(x), h = hi32(x); l = l + a + h; h = h + b + l; x = ((u64)h << 32) | l; } };
struct OpAdd64 { static constexpr const char* name = "add.u64 pair";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)a << 32) | b; x = x + c; x = x + (c ^ 0x55); } };
struct OpLop3 { static constexpr const char* name = "LOP3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = lo32(x), h = hi32(x); l = (l & a) ^ h; h = (h | b) ^ l; x = ((u64)h << 32) | l; } };
struct OpCsub { static constexpr const char* name = "csub64";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)a << 32) | b; u64 d = x - c; x = x >= c ? d : x + 12345; } };
struct OpMulhi64 { static constexpr const char* name = "__umul64hi + mul.lo.u64";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)a << 40) | b; x = __umul64hi(x, c) ^ (x * c); } };
struct OpMixWI { static constexpr const char* name = "1 WIDE + 2 IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 p = (u64)lo32(x) * a + x; unsigned l = lo32(p) + a + hi32(p); unsigned h = hi32(p) + b + l; x = ((u64)h << 32) | l; } };
struct OpMixNI { static constexpr const char* name = "2 IMAD + 2 IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = lo32(x), h = hi32(x); l = l * a + h; h = h * b + l; l = l + a + h; h = h + b + l; x = ((u64)h << 32) | l; } };

template <class Op>
__global__ void __launch_bounds__(256) kern(u64* out, unsigned a, unsigned b, int iters) {
  u64 x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = (u64)(threadIdx.x * 7 + c) * 0x9E3779B97F4A7C15ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) Op::step(x[c], a, b);
  }
  u64 acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class Op> void run(u64* out) {
  const int iters = 20000, grid = 148 * 4;
  kern<Op><<<grid, 256>>>(out, 3, 5, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<Op><<<grid, 256>>>(out, 3, 5, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = ms * 1e-3 * clk * 1e3;                                  // SM cycles at max clock
  double steps_per_smsp = (double)grid * 8 * iters * CH / (148 * 4);   // warp-steps per SMSP
  printf("%-28s %8.3f ms  %7.3f SMSP-cycles per warp-step (at %d MHz)  %s\n", Op::name, ms, cyc / steps_per_smsp,
         clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  u64* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  run<OpImad>(out); run<OpHi>(out); run<OpWide>(out); run<OpWideNoAcc>(out); run<OpIadd3>(out); run<OpAdd64>(out);
  run<OpLop3>(out); run<OpCsub>(out); run<OpMulhi64>(out); run<OpMixWI>(out); run<OpMixNI>(out);
  return 0;
}
