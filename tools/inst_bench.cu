// Instruction-throughput probe for the integer pipes of sm_100a.
// Each kernel runs ITER iterations of an unrolled body of CH independent chains
// of one instruction type (or a mix); prints warp-instructions / cycle / SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/inst_bench tools/inst_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef uint64_t u64;
constexpr int CH = 16;

struct OpImad   { static constexpr int N = 1; static constexpr const char* name = "IMAD (mad.lo.u32)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(l) : "r"(a), "r"(b)); x = l; } };
struct OpImadHi { static constexpr int N = 1; static constexpr const char* name = "IMAD.HI (mad.hi.u32)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(l) : "r"(a), "r"(b)); x = l; } };
struct OpWide   { static constexpr int N = 1; static constexpr const char* name = "IMAD.WIDE (mad.wide.u32)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)b << 32) | a; asm volatile("{.reg .u32 l, h; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %2;}" : "+l"(x) : "r"(a), "l"(c)); } };
struct OpIadd3  { static constexpr int N = 1; static constexpr const char* name = "IADD3 (a+b+c)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(l) : "r"(a), "r"(b)); x = l; } };
struct OpLop3   { static constexpr int N = 1; static constexpr const char* name = "LOP3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l) : "r"(a), "r"(b)); x = l; } };
struct OpShf    { static constexpr int N = 1; static constexpr const char* name = "SHF";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(l) : "r"(a), "r"(b)); x = l; } };
struct OpAdd64  { static constexpr int N = 2; static constexpr const char* name = "add.u64 (2 instr)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)a << 32) | b; asm volatile("add.u64 %0, %0, %1;" : "+l"(x) : "l"(c)); } };
struct OpCsub64 { static constexpr int N = 6; static constexpr const char* name = "csub u64 (sub,setp,selp ~6)";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)a << 32) | b; asm volatile("{.reg .pred p; .reg .u64 d; sub.u64 d, %0, %1; setp.ge.u64 p, %0, %1; selp.u64 %0, d, %0, p;}" : "+l"(x) : "l"(c)); } };
struct OpMixWideIadd { static constexpr int N = 2; static constexpr const char* name = "mix WIDE + IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)b << 32) | a; asm volatile("{.reg .u32 l, h, t; .reg .u64 w; mov.b64 {l,h}, %0; add.u32 t, h, %2; add.u32 h, t, %1; mad.wide.u32 w, l, %1, %3; mov.b64 {l,t}, w; xor.b32 h, h, t; mov.b64 %0, {l,h};}" : "+l"(x) : "r"(a), "r"(b), "l"(c)); } };
struct OpMixImadIadd { static constexpr int N = 2; static constexpr const char* name = "mix IMAD + IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { asm volatile("{.reg .u32 l, h, t; mov.b64 {l,h}, %0; add.u32 t, h, %2; add.u32 h, t, %1; mad.lo.u32 l, l, %1, %2; mov.b64 %0, {l,h};}" : "+l"(x) : "r"(a), "r"(b)); } };
struct OpMixWideImad { static constexpr int N = 2; static constexpr const char* name = "mix WIDE + IMAD";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)b << 32) | a; asm volatile("{.reg .u32 l, h; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %3; mov.b64 {l,h}, %0; mad.lo.u32 h, h, %1, %2; mov.b64 %0, {l,h};}" : "+l"(x) : "r"(a), "r"(b), "l"(c)); } };
struct OpMix1W2I2A { static constexpr int N = 5; static constexpr const char* name = "mix 1 WIDE + 2 IMAD + 2 IADD3";
  __device__ static void step(u64& x, unsigned a, unsigned b) { u64 c = ((u64)b << 32) | a; asm volatile("{.reg .u32 l, h, t; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %3; mov.b64 {l,h}, %0; mad.lo.u32 h, h, %1, %2; mad.lo.u32 l, l, %2, %1; add.u32 t, h, %2; add.u32 h, t, l; add.u32 t, l, %1; add.u32 l, t, h; mov.b64 %0, {l,h};}" : "+l"(x) : "r"(a), "r"(b), "l"(c)); } };
struct OpDfma   { static constexpr int N = 1; static constexpr const char* name = "DFMA";
  __device__ static void step(u64& x, unsigned a, unsigned b) { double d = __longlong_as_double(x); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0 + a * 1e-9), "d"(b * 1e-9)); x = __double_as_longlong(d); } };
struct OpFfma   { static constexpr int N = 1; static constexpr const char* name = "FFMA";
  __device__ static void step(u64& x, unsigned a, unsigned b) { float f = __uint_as_float((unsigned)x); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0f + a * 1e-9f), "f"(b * 1e-9f)); x = __float_as_uint(f); } };
struct OpMixDfmaImad { static constexpr int N = 2; static constexpr const char* name = "mix DFMA + IMAD";
  __device__ static void step(u64& x, unsigned a, unsigned b) { unsigned l = (unsigned)x; double d = __longlong_as_double(x | 0x3ff0000000000000ull); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0 + a * 1e-9), "d"(b * 1e-9)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(l) : "r"(a), "r"(b)); x = __double_as_longlong(d) ^ l; } };

template <class Op>
__global__ void __launch_bounds__(256) kern(u64* out, unsigned a, unsigned b, int iters) {
  u64 x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 7 + c;
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
  const int iters = 4000, grid = 148 * 4;
  kern<Op><<<grid, 256>>>(out, 3, 5, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<Op><<<grid, 256>>>(out, 3, 5, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double winst = (double)grid * 8 * iters * CH * Op::N;
  double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-34s %8.3f ms  %6.3f nominal warp-inst/clk/SMSP = %5.2f clk per group  (%d MHz) %s\n", Op::name, ms,
         winst / cyc / (148 * 4), cyc * 148 * 4 / ((double)grid * 8 * iters * CH), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  u64* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  run<OpImad>(out); run<OpImadHi>(out); run<OpWide>(out); run<OpIadd3>(out); run<OpLop3>(out); run<OpShf>(out);
  run<OpAdd64>(out); run<OpCsub64>(out); run<OpMixWideIadd>(out); run<OpMixImadIadd>(out); run<OpMixWideImad>(out);
  run<OpMix1W2I2A>(out); run<OpDfma>(out); run<OpFfma>(out); run<OpMixDfmaImad>(out);
  return 0;
}
