// Micro-benchmark: register-resident radix-16 butterfly networks, no memory
// traffic in the loop.  Finds the issue-bound ceiling of each arithmetic
// formulation so kernel changes can be judged against it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/bfly_bench tools/bfly_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
typedef uint64_t u64;
struct Tw { u64 w, wp; };
struct Mod { u64 q, two_q, four_q, mu; unsigned n0, n1; };
__device__ __forceinline__ unsigned lo32(u64 x) { return (unsigned)x; }
__device__ __forceinline__ unsigned hi32(u64 x) { return (unsigned)(x >> 32); }
__device__ __forceinline__ u64 join(unsigned lo, unsigned hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 csub(u64 x, u64 b) { u64 d = x - b; return x >= b ? d : x; }

// ---- variant 0: GENERIC (Harvey) as first written
__device__ __forceinline__ void bf0(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 tx = csub(X, m.two_q);
  u64 Q = __umul64hi(Y, w.wp);
  u64 T = Y * w.w - Q * m.q;
  X = tx + T; Y = tx + m.two_q - T;
}
// ---- variant 1: FAST as in the library now
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {
  const unsigned a0 = lo32(a), a1 = hi32(a), b0 = lo32(b), b1 = hi32(b);
  return (u64)a1 * b1 + (u64)__umulhi(a1, b0) + (u64)__umulhi(a0, b1);
}
__device__ __forceinline__ u64 mad_chain(u64 x, u64 w, u64 Q, const Mod& m) {
  const unsigned x0 = lo32(x), x1 = hi32(x), w0 = lo32(w), w1 = hi32(w), q0 = lo32(Q), q1 = hi32(Q);
  u64 t = (u64)x0 * w0; t = (u64)q0 * m.n0 + t;
  unsigned h = hi32(t);
  h = x0 * w1 + h; h = x1 * w0 + h; h = q0 * m.n1 + h; h = q1 * m.n0 + h;
  return join(lo32(t), h);
}
__device__ __forceinline__ void bf1(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 T = mad_chain(Y, w.w, mulhi_approx(Y, w.wp), m);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 2: FAST, PTX-pinned instruction selection
__device__ __forceinline__ u64 madwide(unsigned a, unsigned b, u64 c) {
  u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
__device__ __forceinline__ u64 mulwide(unsigned a, unsigned b) {
  u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ unsigned madlo(unsigned a, unsigned b, unsigned c) {
  unsigned r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ void bf2(u64& X, u64& Y, Tw w, const Mod& m) {
  unsigned y0, y1, p0, p1, w0, w1;
  asm("mov.b64 {%0,%1}, %2;" : "=r"(y0), "=r"(y1) : "l"(Y));
  asm("mov.b64 {%0,%1}, %2;" : "=r"(p0), "=r"(p1) : "l"(w.wp));
  asm("mov.b64 {%0,%1}, %2;" : "=r"(w0), "=r"(w1) : "l"(w.w));
  // Q ~ y1*p1 + hi(y1*p0) + hi(y0*p1), all as wide mads (FMA pipe only)
  u64 Q = mulwide(y1, p1);
  Q = madwide(__umulhi(y1, p0), 1u, Q);
  Q = madwide(__umulhi(y0, p1), 1u, Q);
  unsigned q0, q1;
  asm("mov.b64 {%0,%1}, %2;" : "=r"(q0), "=r"(q1) : "l"(Q));
  u64 t = mulwide(y0, w0);
  t = madwide(q0, m.n0, t);
  unsigned tl, th;
  asm("mov.b64 {%0,%1}, %2;" : "=r"(tl), "=r"(th) : "l"(t));
  th = madlo(y0, w1, th); th = madlo(y1, w0, th); th = madlo(q0, m.n1, th); th = madlo(q1, m.n0, th);
  u64 T; asm("mov.b64 %0, {%1,%2};" : "=l"(T) : "r"(tl), "r"(th));
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 3: FAST, X' folded into the mad chain:  X' = X + y*w + Q*negq ; Y' = 2X + 4q - X'
__device__ __forceinline__ void bf3(u64& X, u64& Y, Tw w, const Mod& m) {
  const unsigned y0 = lo32(Y), y1 = hi32(Y), w0 = lo32(w.w), w1 = hi32(w.w);
  u64 Q = mulhi_approx(Y, w.wp);
  const unsigned q0 = lo32(Q), q1 = hi32(Q);
  u64 t = (u64)y0 * w0 + X; t = (u64)q0 * m.n0 + t;
  unsigned h = hi32(t);
  h = y0 * w1 + h; h = y1 * w0 + h; h = q0 * m.n1 + h; h = q1 * m.n0 + h;
  u64 Xn = join(lo32(t), h);
  Y = X + X + m.four_q - Xn; X = Xn;
}
// ---- variant 4: GENERIC with mad chain and carry-based csub
__device__ __forceinline__ void bf4(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 tx = csub(X, m.two_q);
  u64 T = mad_chain(Y, w.w, __umul64hi(Y, w.wp), m);
  X = tx + T; Y = tx + m.two_q - T;
}

// ---- variant 5: FAST as in the library now (PTX-pinned wide/narrow mads, no IMAD.HI)
__device__ __forceinline__ void split(u64 x, unsigned& lo, unsigned& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
__device__ __forceinline__ u64 join2(unsigned lo, unsigned hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ u64 mulhi_approx5(u64 a, u64 b) {
  unsigned a0, a1, b0, b1; split(a, a0, a1); split(b, b0, b1);
  u64 Q = mulwide(a1, b1);
  Q = madwide(hi32(mulwide(a1, b0)), 1u, Q);
  Q = madwide(hi32(mulwide(a0, b1)), 1u, Q);
  return Q;
}
__device__ __forceinline__ u64 mad_chain5(u64 x, u64 w, u64 Q, const Mod& m) {
  unsigned x0, x1, w0, w1, q0, q1, t0, t1;
  split(x, x0, x1); split(w, w0, w1); split(Q, q0, q1);
  split(madwide(q0, m.n0, mulwide(x0, w0)), t0, t1);
  t1 = madlo(x0, w1, t1); t1 = madlo(x1, w0, t1); t1 = madlo(q0, m.n1, t1); t1 = madlo(q1, m.n0, t1);
  return join2(t0, t1);
}
__device__ __forceinline__ void bf5(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 T = mad_chain5(Y, w.w, mulhi_approx5(Y, w.wp), m);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 12/13: the quotient sum as plain 64-bit adds of zero-extended high halves (IADD3 with two carry-outs
//   + IADD3.X) instead of multiply-by-one wide mads (which ptxas turns into register-pair moves plus adds)
__device__ __forceinline__ void bf12(u64& X, u64& Y, Tw w, const Mod& m) {
  unsigned a0, a1, b0, b1; split(Y, a0, a1); split(w.wp, b0, b1);
  const u64 Q = mulwide(a1, b1) + (u64)hi32(mulwide(a1, b0)) + (u64)hi32(mulwide(a0, b1));
  u64 T = mad_chain5(Y, w.w, Q, m);
  Y = X + m.four_q - T; X = X + T;
}
__device__ __forceinline__ void bf13(u64& X, u64& Y, Tw w, const Mod& m) {
  unsigned a0, a1, b0, b1; split(Y, a0, a1); split(w.wp, b0, b1);
  const u64 Q = mulwide(a1, b1) + (u64)__umulhi(a1, b0) + (u64)__umulhi(a0, b1);
  u64 T = mad_chain5(Y, w.w, Q, m);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 14: the two high halves are summed first (one 64-bit add) and ride in as the accumulator of a1*b1
__device__ __forceinline__ void bf14(u64& X, u64& Y, Tw w, const Mod& m) {
  unsigned a0, a1, b0, b1; split(Y, a0, a1); split(w.wp, b0, b1);
  const u64 hs = (u64)hi32(mulwide(a1, b0)) + (u64)hi32(mulwide(a0, b1));
  const u64 Q = madwide(a1, b1, hs);
  u64 T = mad_chain5(Y, w.w, Q, m);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 15: same with explicit carry instructions
__device__ __forceinline__ void bf15(u64& X, u64& Y, Tw w, const Mod& m) {
  unsigned a0, a1, b0, b1; split(Y, a0, a1); split(w.wp, b0, b1);
  unsigned h1 = hi32(mulwide(a1, b0)), h2 = hi32(mulwide(a0, b1)), s0, s1;
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(s0), "=r"(s1) : "r"(h1), "r"(h2));
  const u64 Q = madwide(a1, b1, join2(s0, s1));
  u64 T = mad_chain5(Y, w.w, Q, m);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 16: Montgomery butterfly (R = 2^64, twiddle in Montgomery form, lazy [0,2q) output) -- SURVEY 8(f)-4:
//   P = Y*W (128 bits), m = lo(P) * (-q^-1) mod 2^64, T = hi(P) + hi(m*q) + (lo(P) != 0)  in [0, 2q)
__device__ __forceinline__ void bf16(u64& X, u64& Y, Tw w, const Mod& m) {
  const u64 plo = Y * w.w, phi = __umul64hi(Y, w.w);
  const u64 mm = plo * m.mu;                      // m.mu stands in for -q^-1 mod 2^64
  const u64 T = phi + __umul64hi(mm, m.q) + (plo != 0);
  Y = X + m.two_q - T; X = X + T;
}
// ---- variant 6: exact mulhi (compiler's __umul64hi) + pinned mad chain, no csub (FAST-exact)
__device__ __forceinline__ void bf6(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 T = mad_chain5(Y, w.w, __umul64hi(Y, w.wp), m);
  Y = X + m.two_q - T; X = X + T;
}
// ---- variant 7: multiplies only (no adds for X',Y'): how fast can the 9-IMAD core go
__device__ __forceinline__ void bf7(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 T = mad_chain5(Y, w.w, mulhi_approx5(Y, w.wp), m);
  Y = T ^ X; X = T;
}
// ---- variant 8: adds only (no multiplies)
__device__ __forceinline__ void bf8(u64& X, u64& Y, Tw w, const Mod& m) {
  u64 T = Y ^ w.w;
  Y = X + m.four_q - T; X = X + T;
}

// ---- variant 9: quotient cross terms on the FP64 pipe.
//   Q ~ a1*p1 + K,  K = round(a1*(p0/2^32) + a0*(p1/2^32) - 1)  in {floor-1, floor}
//   (w.wp reinterpreted: here b0s/b1s are derived on the fly from wp only to keep the
//   benchmark self-contained; a real table would store them)
struct TwD { u64 w; double b0s, b1s; unsigned p1; };
__device__ __forceinline__ double u32_to_double(unsigned a) {
  // bit pattern of 2^52 + a, then remove 2^52 exactly
  return __hiloint2double(0x43300000, (int)a) - 4503599627370496.0;
}
__device__ __forceinline__ void bf9(u64& X, u64& Y, const TwD& w, const Mod& m) {
  unsigned y0, y1; split(Y, y0, y1);
  const double d = fma(u32_to_double(y0), w.b1s, u32_to_double(y1) * w.b0s);   // (a1*p0 + a0*p1) / 2^32
  const double mg = d + 4503599627370495.0;                                      // 2^52 - 1 + d -> integer in the mantissa
  u64 K = (u64)__double_as_longlong(mg) - 0x4330000000000000ull;
  u64 Q = madwide(y1, w.p1, K);
  u64 T = mad_chain5(Y, w.w, Q, m);
  Y = X + m.four_q - T; X = X + T;
}

// ---- variant 10: quotient cross terms as two round-down DFMAs on biased operands (round 2).
//   A0 = 2^52 + y0, A1 = 2^52 + y1 are formed by pairing a word with the constant 0x43300000 (no arithmetic);
//   u = fma.rm(A0, beta1, K), R = fma.rm(A1, beta0, u) = 2^52 + 2 + cross - [0,2), K = 2^52 + 2 - 2^20 (b0 + b1),
//   beta_i = b_i / 2^32;  Q + C = y1*b1 + bits(R) with C = 0x4330000000000002, and C*q is folded into the
//   accumulator of the first product of the mad chain (m.bias).  Q is low by at most 2, as in v5.
struct TwH { u64 w; double beta0, beta1, K; unsigned b1; };
struct ModH { u64 q, four_q; unsigned n0, n1; u64 bias; };
__device__ __forceinline__ double fma_rm(double a, double b, double c) {
  double r; asm("fma.rm.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r;
}
__device__ __forceinline__ void bf10(u64& X, u64& Y, const TwH& w, const ModH& m) {
  unsigned y0, y1; split(Y, y0, y1);
  const double A0 = __hiloint2double(0x43300000, (int)y0), A1 = __hiloint2double(0x43300000, (int)y1);
  const double R = fma_rm(A1, w.beta0, fma_rm(A0, w.beta1, w.K));
  const u64 Qc = madwide(y1, w.b1, (u64)__double_as_longlong(R));
  unsigned w0, w1, q0, q1, t0, t1;
  split(w.w, w0, w1); split(Qc, q0, q1);
  split(madwide(q0, m.n0, madwide(y0, w0, m.bias)), t0, t1);
  t1 = madlo(y0, w1, t1); t1 = madlo(y1, w0, t1); t1 = madlo(q0, m.n1, t1); t1 = madlo(q1, m.n0, t1);
  const u64 T = join2(t0, t1);
  Y = X + m.four_q - T; X = X + T;
}
// ---- variant 11: as v10 but the operand words are converted with I2F.F64.U32 (conversion unit) instead of
//   being paired with a constant register (which costs a MOV per word): K is then the constant 2^52 + 2.
__device__ __forceinline__ void bf11(u64& X, u64& Y, const TwH& w, const ModH& m) {
  unsigned y0, y1; split(Y, y0, y1);
  const double R = fma_rm(__uint2double_rn(y1), w.beta0, fma_rm(__uint2double_rn(y0), w.beta1, 4503599627370498.0));
  const u64 Qc = madwide(y1, w.b1, (u64)__double_as_longlong(R));
  unsigned w0, w1, q0, q1, t0, t1;
  split(w.w, w0, w1); split(Qc, q0, q1);
  split(madwide(q0, m.n0, madwide(y0, w0, m.bias)), t0, t1);
  t1 = madlo(y0, w1, t1); t1 = madlo(y1, w0, t1); t1 = madlo(q0, m.n1, t1); t1 = madlo(q1, m.n0, t1);
  const u64 T = join2(t0, t1);
  Y = X + m.four_q - T; X = X + T;
}
template <int MINB, int VAR = 10>
__global__ void __launch_bounds__(256, MINB) kern10(u64* out, const Tw* tw, Mod m0, int iters) {
  auto bfx = [](u64& X, u64& Y, const TwH& w, const ModH& m) { if (VAR == 10) bf10(X, Y, w, m); else bf11(X, Y, w, m); };
  u64 v[16];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (u64)tid * 0x9E3779B97F4A7C15ull + e * 0x1234567ull;
  ModH m; m.q = m0.q; m.four_q = m0.four_q; m.n0 = m0.n0; m.n1 = m0.n1; m.bias = 0x4330000000000002ull * m0.q;
  TwH w[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    Tw t = tw[(tid + g) & 1023];
    w[g].w = t.w; w[g].b1 = hi32(t.wp);
    w[g].beta0 = (double)lo32(t.wp) * (1.0 / 4294967296.0); w[g].beta1 = (double)hi32(t.wp) * (1.0 / 4294967296.0);
    w[g].K = 4503599627370498.0 - 1048576.0 * ((double)lo32(t.wp) + (double)hi32(t.wp));
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int l = 0; l < 8; ++l) bfx(v[l], v[l | 8], w[0], m);
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int l = 0; l < 4; ++l) bfx(v[(g << 3) | l], v[((g << 3) | l) | 4], w[g], m);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int l = 0; l < 2; ++l) bfx(v[(g << 2) | l], v[((g << 2) | l) | 2], w[g], m);
#pragma unroll
    for (int g = 0; g < 8; ++g) bfx(v[g << 1], v[(g << 1) | 1], w[g], m);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] &= (1ull << 60) - 1;
  }
  u64 acc = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) acc ^= v[e];
  out[tid] = acc;
}
template <int MINB, int VAR = 10> void run10(const char* name, u64* out, const Tw* tw, Mod m, int blocks_per_sm) {
  const int iters = 2000, grid = 148 * blocks_per_sm;
  kern10<MINB, VAR><<<grid, 256>>>(out, tw, m, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern10<MINB, VAR><<<grid, 256>>>(out, tw, m, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bfl = (double)grid * 256 * iters * 32;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern10<MINB, VAR>);
  printf("%-28s blocks/SM %d regs %3d : %8.1f G bfly/s = %6.3f M NTT(2^16)/s = %5.1f SMSP-cycles per warp-bfly  %s\n", name,
         blocks_per_sm, fa.numRegs, bfl / ms / 1e6, bfl / ms / 1e6 / 524288.0 * 1e3, 592.0 * 1.965e9 / (bfl / ms * 1e3 / 32),
         cudaGetErrorString(cudaGetLastError()));
}

template <int V> __device__ __forceinline__ void bf(u64& X, u64& Y, Tw w, const Mod& m) {
  if (V == 5) { bf5(X, Y, w, m); return; } if (V == 6) { bf6(X, Y, w, m); return; }
  if (V == 16) { bf16(X, Y, w, m); return; }
  if (V == 14) { bf14(X, Y, w, m); return; } if (V == 15) { bf15(X, Y, w, m); return; }
  if (V == 12) { bf12(X, Y, w, m); return; } if (V == 13) { bf13(X, Y, w, m); return; }
  if (V == 7) { bf7(X, Y, w, m); return; } if (V == 8) { bf8(X, Y, w, m); return; }
  if (V == 0) bf0(X, Y, w, m); else if (V == 1) bf1(X, Y, w, m); else if (V == 2) bf2(X, Y, w, m);
  else if (V == 3) bf3(X, Y, w, m); else bf4(X, Y, w, m);
}

template <int V, int EB>
__device__ __forceinline__ void stage(u64 (&v)[16], const Tw (&w)[8], const Mod& m) {
#pragma unroll
  for (int g = 0; g < (8 >> EB); ++g)
#pragma unroll
    for (int l = 0; l < (1 << EB); ++l) bf<V>(v[(g << (EB + 1)) | l], v[((g << (EB + 1)) | l) | (1 << EB)], w[g], m);
}

template <int V, int MINB>
__global__ void __launch_bounds__(256, MINB) kern(u64* out, const Tw* tw, Mod m, int iters) {
  u64 v[16];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (u64)tid * 0x9E3779B97F4A7C15ull + e * 0x1234567ull;
  Tw w[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) w[g] = tw[(tid + g) & 1023];
  for (int it = 0; it < iters; ++it) {
    stage<V, 3>(v, w, m); stage<V, 2>(v, w, m); stage<V, 1>(v, w, m); stage<V, 0>(v, w, m);
    if (V != 0 && V != 4) {  // keep FAST values bounded: mask to 60 bits (cheap, not part of the count)
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] &= (1ull << 60) - 1;
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) acc ^= v[e];
  out[tid] = acc;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) kern9(u64* out, const Tw* tw, Mod m, int iters) {
  u64 v[16];
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = (u64)tid * 0x9E3779B97F4A7C15ull + e * 0x1234567ull;
  TwD w[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    Tw t = tw[(tid + g) & 1023];
    w[g].w = t.w; w[g].p1 = hi32(t.wp);
    w[g].b0s = (double)lo32(t.wp) * (1.0 / 4294967296.0); w[g].b1s = (double)hi32(t.wp) * (1.0 / 4294967296.0);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 1; ++g)
#pragma unroll
      for (int l = 0; l < 8; ++l) bf9(v[l], v[l | 8], w[0], m);
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int l = 0; l < 4; ++l) bf9(v[(g << 3) | l], v[((g << 3) | l) | 4], w[g], m);
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int l = 0; l < 2; ++l) bf9(v[(g << 2) | l], v[((g << 2) | l) | 2], w[g], m);
#pragma unroll
    for (int g = 0; g < 8; ++g) bf9(v[g << 1], v[(g << 1) | 1], w[g], m);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] &= (1ull << 60) - 1;
  }
  u64 acc = 0;
#pragma unroll
  for (int e = 0; e < 16; ++e) acc ^= v[e];
  out[tid] = acc;
}
template <int MINB> void run9(const char* name, u64* out, const Tw* tw, Mod m, int blocks_per_sm) {
  const int iters = 2000, grid = 148 * blocks_per_sm;
  kern9<MINB><<<grid, 256>>>(out, tw, m, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern9<MINB><<<grid, 256>>>(out, tw, m, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bfl = (double)grid * 256 * iters * 32;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern9<MINB>);
  printf("%-28s blocks/SM %d regs %3d : %8.1f G bfly/s = %6.3f M NTT(2^16)/s = %5.1f SMSP-cycles per warp-bfly  %s\n", name,
         blocks_per_sm, fa.numRegs, bfl / ms / 1e6, bfl / ms / 1e6 / 524288.0 * 1e3, 592.0 * 1.965e9 / (bfl / ms * 1e3 / 32),
         cudaGetErrorString(cudaGetLastError()));
}

template <int V, int MINB> void run(const char* name, u64* out, const Tw* tw, Mod m, int blocks_per_sm) {
  const int iters = 2000, grid = 148 * blocks_per_sm;
  kern<V, MINB><<<grid, 256>>>(out, tw, m, 10);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  kern<V, MINB><<<grid, 256>>>(out, tw, m, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double bfl = (double)grid * 256 * iters * 32;
  int regs = 0; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern<V, MINB>); regs = fa.numRegs;
  printf("%-28s blocks/SM %d regs %3d : %8.1f G bfly/s = %6.3f M NTT(2^16)/s = %5.1f SMSP-cycles per warp-bfly  %s\n", name,
         blocks_per_sm, regs, bfl / ms / 1e6, bfl / ms / 1e6 / 524288.0 * 1e3, 592.0 * 1.965e9 / (bfl / ms * 1e3 / 32),
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  u64* out; Tw* tw; cudaMalloc(&out, 148 * 8 * 256 * 8); cudaMalloc(&tw, 1024 * sizeof(Tw));
  Tw h[1024]; u64 q = 36028797019488257ull;
  for (int i = 0; i < 1024; ++i) { h[i].w = (0x9E3779B97F4A7C15ull * (i + 1)) % q; h[i].wp = (u64)(((unsigned __int128)h[i].w << 64) / q); }
  cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
  Mod m; m.q = q; m.two_q = 2 * q; m.four_q = 4 * q; m.mu = (u64)(((unsigned __int128)1 << 64) / q);
  u64 nq = 0 - q; m.n0 = (unsigned)nq; m.n1 = (unsigned)(nq >> 32);
  run<0, 2>("v0 generic first", out, tw, m, 2);
  run<4, 2>("v4 generic madchain", out, tw, m, 2);
  run<1, 2>("v1 fast IMAD.HI", out, tw, m, 2);
  for (int bps = 2; bps <= 3; ++bps) {
    run<5, 2>("v5 fast pinned (lib)", out, tw, m, bps);
    run<6, 2>("v6 fast exact-mulhi", out, tw, m, bps);
    run<7, 2>("v7 multiplies only", out, tw, m, bps);
    run<8, 2>("v8 adds only", out, tw, m, bps);
  }
  run9<2>("v9 fp64 cross terms", out, tw, m, 2);
  run9<2>("v9 fp64 cross terms", out, tw, m, 3);
  run9<3>("v9 mb3", out, tw, m, 3);
  run10<2>("v10 fp64 biased 2xDFMA", out, tw, m, 2);
  run10<2>("v10 fp64 biased 2xDFMA", out, tw, m, 3);
  run10<3>("v10 mb3", out, tw, m, 3);
  run10<3>("v10 mb3 x4", out, tw, m, 4);
  run10<2, 11>("v11 fp64 I2F 2xDFMA", out, tw, m, 2);
  run10<2, 11>("v11 fp64 I2F 2xDFMA", out, tw, m, 3);
  run10<3, 11>("v11 mb3", out, tw, m, 3);
  run10<3, 11>("v11 mb3 x4", out, tw, m, 4);
  run<12, 2>("v12 plain 64-bit quotient adds", out, tw, m, 2);
  run<12, 3>("v12 mb3", out, tw, m, 3);
  run<14, 2>("v14 hs as accumulator", out, tw, m, 2);
  run<14, 3>("v14 mb3", out, tw, m, 3);
  run<15, 2>("v15 hs via add.cc", out, tw, m, 2);
  run<15, 3>("v15 mb3", out, tw, m, 3);
  run<16, 2>("v16 Montgomery R=2^64", out, tw, m, 2);
  run<16, 3>("v16 mb3", out, tw, m, 3);
  run<13, 2>("v13 IMAD.HI + plain adds", out, tw, m, 2);
  run<13, 3>("v13 mb3", out, tw, m, 3);
  run<5, 3>("v5 mb3", out, tw, m, 3);
  run<6, 3>("v6 mb3", out, tw, m, 3);
  run<7, 3>("v7 mb3", out, tw, m, 3);
  return 0;
}
