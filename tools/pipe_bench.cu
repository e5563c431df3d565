// Instruction-throughput probe, round 2: which pipes of an sm_100a SM can be kept busy at
// the same time.  The NTT butterflies are bound by the FMA-heavy pipe (IMAD / IMAD.WIDE);
// this measures what the FP64 pipe, the FMA-lite pipe (FFMA) and the ALU pipe can absorb
// next to it.  Every kernel is a register-only loop of independent dependency chains; the
// SASS mix of each loop is checked with `cuobjdump -sass tools/bin/pipe_bench`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/pipe_bench tools/pipe_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef uint64_t u64;
constexpr int CH = 8;

__device__ __forceinline__ unsigned lo32(u64 x) { return (unsigned)x; }
__device__ __forceinline__ unsigned hi32(u64 x) { return (unsigned)(x >> 32); }
__device__ __forceinline__ u64 madwide(unsigned a, unsigned b, u64 c) {
  u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ u64 mulwide(unsigned a, unsigned b) {
  u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ unsigned madlo(unsigned a, unsigned b, unsigned c) {
  unsigned r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ unsigned madhi(unsigned a, unsigned b, unsigned c) {
  unsigned r; asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ unsigned add3(unsigned a, unsigned b, unsigned c) {
  unsigned r; asm("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ double xdfma(double a, double b, double c) {
  double r; asm("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r;
}
__device__ __forceinline__ double xdadd(double a, double b) {
  double r; asm("add.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r;
}
__device__ __forceinline__ double xdmul(double a, double b) {
  double r; asm("mul.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b)); return r;
}
__device__ __forceinline__ float xffma(float a, float b, float c) {
  float r; asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

struct State {
  u64 x[CH];       // integer chains
  unsigned n[CH];  // narrow chains
  double d[CH];    // fp64 chains
  float f[CH];     // fp32 chains
  unsigned s[CH];  // alu chains
};
struct Args { unsigned a, b; double da, db; float fa, fb; };

#define OPDEF(NAME, LABEL, NINST, BODY)                                                   \
  struct NAME {                                                                            \
    static constexpr const char* name = LABEL;                                             \
    static constexpr int ninst = NINST;                                                    \
    __device__ static __forceinline__ void step(State& S, const Args& A, int c) { BODY }   \
  };

// building blocks: every multiplicand changes from step to step (ptxas strength-reduces loop-invariant products)
#define W_ S.x[c] = madwide(S.n[c], S.s[c], S.x[c]);
#define N_ S.n[c] = madlo(S.n[c], A.a, A.b);
#define A_ S.s[c] = add3(S.s[c], A.a, A.b);
#define D_ S.d[c] = xdfma(S.d[c], A.da, A.db);
#define F_ S.f[c] = xffma(S.f[c], A.fa, A.fb);
// ---- single-instruction loops
OPDEF(OpImad, "IMAD (narrow)", 1, N_)
OPDEF(OpImadHi, "IMAD.HI", 1, S.n[c] = madhi(S.n[c], A.a, A.b);)
OPDEF(OpWideSelf, "IMAD.WIDE d=lo(d)*a+d", 1, S.x[c] = madwide(lo32(S.x[c]), A.a, S.x[c]);)
OPDEF(OpWideOther, "IMAD.WIDE d=lo(d)*a+e", 1, S.x[c] = madwide(lo32(S.x[c]), A.a, (u64)__double_as_longlong(S.d[c]));)
OPDEF(OpWideHiUse, "IMAD.WIDE d=hi(d)*a (no acc)", 1, S.x[c] = mulwide(hi32(S.x[c]), A.a);)
OPDEF(OpIadd3, "IADD3", 1, A_)
OPDEF(OpDfma, "DFMA", 1, D_)
OPDEF(OpDadd, "DADD", 1, S.d[c] = xdadd(S.d[c], A.db);)
OPDEF(OpDmul, "DMUL", 1, S.d[c] = xdmul(S.d[c], A.da);)
OPDEF(OpFfma, "FFMA", 1, F_)
// I2F.F64.U32 (+ one IADD3 to keep the chain integer): is the conversion unit a usable extra pipe?
OPDEF(OpI2F, "I2F.F64.U32 + IADD3", 2, S.n[c] = (unsigned)__double2loint(__uint2double_rn(S.n[c])) + A.a;)
OPDEF(OpI2FW, "I2F.F64.U32 + IADD3 + W + N", 4, S.s[c] = (unsigned)__double2loint(__uint2double_rn(S.s[c])) + A.a; W_ N_)
// ---- pairs: does the second instruction hide under the first?
OPDEF(OpWN, "W N (narrow feeds wide)", 2, W_ N_)
OPDEF(OpWNA, "W N A", 3, W_ N_ A_)
OPDEF(OpWND, "W N A D", 4, W_ N_ A_ D_)
OPDEF(OpWNDD, "W N A D D", 5, W_ N_ A_ D_ D_)
OPDEF(OpND, "N D", 2, N_ D_)
OPDEF(OpNDD, "N D D", 3, N_ D_ D_)
OPDEF(OpNF, "N F", 2, N_ F_)
OPDEF(OpNFF, "N F F", 3, N_ F_ F_)
OPDEF(OpNA, "N A", 2, N_ A_)
OPDEF(OpNAA, "N A A", 3, N_ A_ A_)
OPDEF(OpDA, "D A", 2, D_ A_)
OPDEF(OpDF, "D F", 2, D_ F_)
// ---- the candidate butterfly mixes (instruction counts only, no meaning)
OPDEF(OpMixNow, "mix 5W+4N+7A (today)", 16, W_ N_ A_ W_ N_ A_ W_ N_ A_ W_ N_ A_ W_ A_ A_ A_)
OPDEF(OpMixHyb2, "mix 3W+4N+2D+6A (hybrid)", 15, W_ N_ A_ D_ W_ N_ A_ D_ W_ N_ A_ N_ A_ A_ A_)
OPDEF(OpMixHyb3, "mix 3W+4N+3D+6A", 16, W_ N_ A_ D_ W_ N_ A_ D_ W_ N_ A_ D_ N_ A_ A_ A_)
OPDEF(OpMixHyb4, "mix 3W+4N+4D+7A", 18, W_ N_ A_ D_ W_ N_ A_ D_ W_ N_ A_ D_ N_ A_ D_ A_ A_ A_)
OPDEF(OpMixHyb6, "mix 3W+4N+6D+7A", 20, W_ N_ A_ D_ D_ W_ N_ A_ D_ D_ W_ N_ A_ D_ N_ A_ D_ A_ A_ A_)

template <class Op>
__global__ void __launch_bounds__(256) kern(u64* out, Args A, int iters) {
  State S;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    S.x[c] = (u64)(threadIdx.x * 7 + c) * 0x9E3779B97F4A7C15ull;
    S.n[c] = threadIdx.x * 13 + c;
    S.s[c] = threadIdx.x * 17 + c;
    S.d[c] = 1.0 + 1e-9 * (threadIdx.x + c);
    S.f[c] = 1.0f + 1e-6f * (threadIdx.x + c);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) Op::step(S, A, c);
  }
  u64 acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += S.x[c] + S.n[c] + S.s[c] + (u64)__double_as_longlong(S.d[c]) + __float_as_uint(S.f[c]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class Op> void run(u64* out, int blocks_per_sm) {
  const int iters = 8000, grid = 148 * blocks_per_sm;
  Args A{3u, 5u, 1.0000001, 1e-7, 1.0001f, 1e-5f};
  kern<Op><<<grid, 256>>>(out, A, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<Op><<<grid, 256>>>(out, A, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * clk * 1e3;                                    // SM cycles at max clock
  const double steps_per_smsp = (double)grid * 8 * iters * CH / (148 * 4);     // warp-steps per SMSP
  printf("%-36s warps/SMSP %d  %8.3f ms  %7.3f SMSP-cycles per warp-step (%2d inst: %5.2f each)  %s\n", Op::name,
         blocks_per_sm * 2, ms, cyc / steps_per_smsp, Op::ninst, cyc / steps_per_smsp / Op::ninst,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  u64* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("# pipe_bench: max SM clock %d MHz (cycle counts assume the kernel ran at it)\n", clk / 1000);
  for (int bps : {4, 2}) {
    run<OpImad>(out, bps); run<OpImadHi>(out, bps); run<OpWideSelf>(out, bps); run<OpWideOther>(out, bps);
    run<OpWideHiUse>(out, bps); run<OpIadd3>(out, bps);
    run<OpDfma>(out, bps); run<OpDadd>(out, bps); run<OpDmul>(out, bps); run<OpFfma>(out, bps); run<OpI2F>(out, bps); run<OpI2FW>(out, bps);
    run<OpWN>(out, bps); run<OpWNA>(out, bps); run<OpWND>(out, bps); run<OpWNDD>(out, bps);
    run<OpND>(out, bps); run<OpNDD>(out, bps); run<OpNF>(out, bps); run<OpNFF>(out, bps);
    run<OpNA>(out, bps); run<OpNAA>(out, bps); run<OpDA>(out, bps); run<OpDF>(out, bps);
    run<OpMixNow>(out, bps); run<OpMixHyb2>(out, bps); run<OpMixHyb3>(out, bps); run<OpMixHyb4>(out, bps); run<OpMixHyb6>(out, bps);
  }
  return 0;
}
