// Device-side 64-bit modular arithmetic for the sm_100a kernels.
//
// B200 has no 64-bit integer multiplier: mul.lo.u64 lowers to 3 IMAD and
// mul.hi.u64 to 4 IMAD.WIDE.U32 plus carry adds, all on the FMA pipe
// (64 lanes/clk/SM).  Everything here is written to minimise those.
//
// The arithmetic restates (not copies) the scalar definitions of the reference:
//   Shoup/Harvey lazy multiply   hexl/include/hexl/number-theory/number-theory.hpp:127-141
//   conditional reductions       number-theory.hpp:214-258
//   Barrett-64                   number-theory.hpp:195-205
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hexl_b200 {

typedef uint64_t u64;  // same type as the C ABI's uint64_t (unsigned long on LP64)

// {w, w_shoup}: a twiddle and floor(w * 2^64 / q), fetched with one 128-bit load
struct __align__(16) Twiddle {
  u64 w;
  u64 wp;
};
// the same for moduli below 2^30, where every lazy value fits one 32-bit word:
// {w, floor(w * 2^32 / q)}, one 64-bit load
struct __align__(8) Twiddle32 {
  uint32_t w;
  uint32_t wp;
};

__device__ __forceinline__ u64 mulhi(u64 a, u64 b) {
  return __umul64hi((unsigned long long)a, (unsigned long long)b);
}

// x*w - floor(x*wp / 2^64)*q  in [0, 2q) for ANY 64-bit x (wp = floor(w*2^64/q), w < q)
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 wp, u64 q) {
  return x * w - mulhi(x, wp) * q;
}

// x in [0, 2b)  ->  [0, b)
__device__ __forceinline__ u64 csub(u64 x, u64 b) {
  u64 d = x - b;
  return x >= b ? d : x;
}

// the same for b < 2^63 (every modulus multiple the NTT butterflies subtract): x - b as a signed number is negative
// exactly when x < b, so the test is one compare of the high word instead of a 64-bit compare (ISETP + ISETP.EX)
__device__ __forceinline__ u64 csub_s(u64 x, u64 b) {
  const u64 d = x - b;
  return (long long)d < 0 ? x : d;
}

// [0, k*q) -> [0, q) for k in {1,2,4,8} by conditional subtractions
template <int K>
__device__ __forceinline__ u64 reduce_from(u64 x, u64 q) {
  if (K >= 8) x = csub(x, q << 2);
  if (K >= 4) x = csub(x, q << 1);
  if (K >= 2) x = csub(x, q);
  return x;
}
__device__ __forceinline__ u64 reduce_from_rt(u64 x, u64 q, int k) {
  if (k >= 8) x = csub(x, q << 2);
  if (k >= 4) x = csub(x, q << 1);
  if (k >= 2) x = csub(x, q);
  return x;
}

// x - floor(x*mu/2^64)*q in [0, 2q) for any 64-bit x, mu = floor(2^64/q)
__device__ __forceinline__ u64 barrett64_lazy(u64 x, u64 q, u64 mu) {
  return x - mulhi(x, mu) * q;
}

// 128-bit global accesses with streaming (evict-first) policy: every coefficient
// is touched once per kernel, keep L2 for the twiddle tables.
__device__ __forceinline__ ulonglong2 ld_stream2(const u64* p) {
  return __ldcs(reinterpret_cast<const ulonglong2*>(p));
}
__device__ __forceinline__ void st_stream2(u64* p, ulonglong2 v) {
  __stcs(reinterpret_cast<ulonglong2*>(p), v);
}

}  // namespace hexl_b200
