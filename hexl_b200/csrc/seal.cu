// Device kernels for the SEAL-shaped composites that sit directly on top of the
// hot path: DyadicMultiply and the element-wise glue of
// CKKS KeySwitch.  The NTTs inside KeySwitch are the kernels of ntt.cu; what is
// here is memory-bound streaming work.
//   DyadicMultiply   hexl/experimental/seal/dyadic-multiply-internal.cpp:17-73
//   KeySwitch        hexl/experimental/seal/key-switch-internal.cpp:25-201
#include "internal.h"

namespace hexl_b200 {
namespace {

constexpr int kThreads = 256;

// generalised Barrett x*y mod q for x, y < q (same arithmetic as EltwiseMultMod,
// eltwise-mult-mod-internal.hpp:71-99)
struct MulCtx {
  u64 q, mu;
  int shift;
};
__device__ __forceinline__ u64 mulmod(u64 x, u64 y, const MulCtx& c) {
  const u64 lo = x * y, hi = mulhi(x, y);
  const u64 c1 = c.shift ? ((lo >> c.shift) | (hi << (64 - c.shift))) : lo;
  return csub(lo - mulhi(c1, c.mu) * c.q, c.q);
}

// ---- DyadicMultiply: (x0*y0, x0*y1 + x1*y0, x1*y1) for every RNS modulus.
// One thread per coefficient slot: 4 loads (32 B), 3 stores (24 B), all
// coalesced; inputs are read before any output is written, so result may alias
// either operand (test-dyadic-multiply.cpp:38-112).
__global__ void __launch_bounds__(kThreads)
    dyadic_kernel(u64* result, const u64* op1, const u64* op2, u64 n, u64 num_moduli, u64 first, u64 count,
                  const __grid_constant__ DyadicModuli mods) {
  const u64 total = n * count, poly = n * num_moduli, base = first * n;
  const u64 stride = (u64)gridDim.x * kThreads;
  for (u64 i = (u64)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const DyadicModulus& dm = mods.m[i / n];
    const MulCtx c{dm.q, dm.mu, dm.shift};
    const u64 o = base + i;
    const u64 x0 = __ldcs(op1 + o), x1 = __ldcs(op1 + o + poly);
    const u64 y0 = __ldcs(op2 + o), y1 = __ldcs(op2 + o + poly);
    const u64 r0 = mulmod(x0, y0, c);
    const u64 r1 = csub(mulmod(x0, y1, c) + mulmod(x1, y0, c), c.q);
    const u64 r2 = mulmod(x1, y1, c);
    __stcs(result + o, r0);
    __stcs(result + o + poly, r1);
    __stcs(result + o + 2 * poly, r2);
  }
}

// ---- KeySwitch: lazy 128-bit multiply-accumulate with the switching keys and
// the final reduction (key-switch-internal.cpp:93-130).  For RNS modulus i
// (key_index = its slot in the key), thread (k, l) sums over the decomposition
// digits j:  acc += operand[j][l] * key_j[n*key_index + k*key_modulus_size*n + l]
// and stores acc mod q.  acc = hi*2^64 + lo is reduced as
// Shoup(hi, 2^64 mod q) + Barrett(lo), both lazy, then two conditional subtractions.
__global__ void __launch_bounds__(kThreads)
    ks_mac_kernel(u64* prod_i /* [kcc][rns*n] slice base + i*n */, const u64* operands /* [count][n] */,
                  const __grid_constant__ KeyPointers keys, u64 n, u64 count, u64 kcc, u64 key_index,
                  u64 key_modulus_size, u64 prod_stride_k, u64 q, u64 mu, Twiddle r64, int accumulate) {
  const u64 total = kcc * n;
  const u64 g = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (g >= total) return;
  const u64 k = g / n, l = g - k * n;
  const u64 key_off = n * key_index + k * key_modulus_size * n + l;
  u64 lo = 0, hi = 0;
  for (u64 j = 0; j < count; ++j) {
    const u64 a = operands[j * n + l];
    const u64 b = __ldcs(keys.p[j] + key_off);
    const u64 plo = a * b, phi = mulhi(a, b);
    lo += plo;
    hi += phi + (lo < plo);
  }
  u64 r = shoup_lazy(hi, r64.w, r64.wp, q) + barrett64_lazy(lo, q, mu);  // < 4q
  r = csub(csub(r, q << 1), q);
  u64* dst = prod_i + k * prod_stride_k + l;
  if (accumulate) r = csub(r + *dst, q);
  *dst = r;
}

// ---- KeySwitch tail, first half (key-switch-internal.cpp:148-178): the special
// prime's part, already in coefficient form in [0, 2*q_last):
//   t = (x + q_last/2) mod q_last;   out_i = (t mod q_i) + (q_i - (q_last/2 mod q_i))
// written for every target modulus i at out + i*kcc_stride (lazy, < 2 q_i).
__global__ void __launch_bounds__(kThreads)
    ks_round_kernel(u64* out, const u64* t_last, u64 n, u64 q_last, u64 mu_last, u64 q_i, u64 mu_i, u64 fix) {
  const u64 l = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (l >= n) return;
  u64 x = t_last[l] + (q_last >> 1);
  x = csub(barrett64_lazy(x, q_last, mu_last), q_last);
  if (q_last > q_i) x = csub(barrett64_lazy(x, q_i, mu_i), q_i);
  out[l] = x + fix;
}

// ---- KeySwitch tail, second half (:183-197):
//   ith = prod + 4 q_i - t_ntt;  r = ith * modswitch mod q_i (inputs < 8 q_i);  result = (result + r) mod q_i
__global__ void __launch_bounds__(kThreads)
    ks_finish_kernel(u64* result, const u64* prod, const u64* t_ntt, u64 n, u64 q, u64 ms, u64 ms_p) {
  const u64 l = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (l >= n) return;
  u64 x = prod[l] + (q << 2) - t_ntt[l];
  x = reduce_from<8>(x, q);
  const u64 r = csub(shoup_lazy(x, ms, ms_p, q), q);
  result[l] = csub(result[l] + r, q);
}

unsigned blocks_for(u64 items) { return (unsigned)((items + kThreads - 1) / kThreads); }

}  // namespace

cudaError_t launch_dyadic_multiply(u64* result, const u64* op1, const u64* op2, u64 n, u64 num_moduli, u64 first,
                                   u64 count, const DyadicModuli& mods, cudaStream_t stream) {
  const u64 total = n * count;
  if (total == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  u64 blocks = blocks_for(total);
  if (blocks > (u64)sms * 8) blocks = (u64)sms * 8;
  dyadic_kernel<<<(unsigned)blocks, kThreads, 0, stream>>>(result, op1, op2, n, num_moduli, first, count, mods);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_mac(u64* prod_i, const u64* operands, const KeyPointers& keys, u64 n, u64 count, u64 kcc,
                          u64 key_index, u64 key_modulus_size, u64 prod_stride_k, u64 q, u64 mu, Twiddle r64,
                          int accumulate, cudaStream_t stream) {
  ks_mac_kernel<<<blocks_for(kcc * n), kThreads, 0, stream>>>(prod_i, operands, keys, n, count, kcc, key_index,
                                                             key_modulus_size, prod_stride_k, q, mu, r64, accumulate);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_round(u64* out, const u64* t_last, u64 n, u64 q_last, u64 mu_last, u64 q_i, u64 mu_i, u64 fix,
                            cudaStream_t stream) {
  ks_round_kernel<<<blocks_for(n), kThreads, 0, stream>>>(out, t_last, n, q_last, mu_last, q_i, mu_i, fix);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_finish(u64* result, const u64* prod, const u64* t_ntt, u64 n, u64 q, u64 ms, u64 ms_p,
                             cudaStream_t stream) {
  ks_finish_kernel<<<blocks_for(n), kThreads, 0, stream>>>(result, prod, t_ntt, n, q, ms, ms_p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hexl_b200
