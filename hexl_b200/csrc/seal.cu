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
// A thread owns VEC consecutive coefficient slots of one modulus: 4 loads and 3 stores of
// VEC*8 bytes, all coalesced and streaming (56 B of traffic per slot); inputs are read
// before any output is written, so result may alias either operand
// (test-dyadic-multiply.cpp:38-112).  VEC = 2 (128-bit accesses) when n is even and every
// pointer is 16-byte aligned, else 1.
template <int VEC>
struct Slots {
  u64 v[VEC];
};
template <int VEC>
__device__ __forceinline__ Slots<VEC> ld_slots(const u64* p) {
  Slots<VEC> r;
  if constexpr (VEC == 2) {
    const ulonglong2 t = ld_stream2(p);
    r.v[0] = t.x;
    r.v[1] = t.y;
  } else {
    r.v[0] = __ldcs(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void st_slots(u64* p, const Slots<VEC>& r) {
  if constexpr (VEC == 2) {
    st_stream2(p, make_ulonglong2(r.v[0], r.v[1]));
  } else {
    __stcs(p, r.v[0]);
  }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
    dyadic_kernel(u64* result, const u64* op1, const u64* op2, u64 n, u64 num_moduli, u64 first, u64 count,
                  const __grid_constant__ DyadicModuli mods) {
  const u64 total = n * count / VEC, poly = n * num_moduli, base = first * n;
  const u64 stride = (u64)gridDim.x * kThreads;
  for (u64 i = (u64)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const DyadicModulus& dm = mods.m[i * VEC / n];
    const MulCtx c{dm.q, dm.mu, dm.shift};
    const u64 o = base + i * VEC;
    const Slots<VEC> x0 = ld_slots<VEC>(op1 + o), x1 = ld_slots<VEC>(op1 + o + poly);
    const Slots<VEC> y0 = ld_slots<VEC>(op2 + o), y1 = ld_slots<VEC>(op2 + o + poly);
    Slots<VEC> r0, r1, r2;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      r0.v[k] = mulmod(x0.v[k], y0.v[k], c);
      r1.v[k] = csub(mulmod(x0.v[k], y1.v[k], c) + mulmod(x1.v[k], y0.v[k], c), c.q);
      r2.v[k] = mulmod(x1.v[k], y1.v[k], c);
    }
    st_slots<VEC>(result + o, r0);
    st_slots<VEC>(result + o + poly, r1);
    st_slots<VEC>(result + o + 2 * poly, r2);
  }
}

// ---- EltwiseMultMod / AddMod / SubMod over an RNS batch: block e of per_mod elements under
// modulus e (eltwise-mult-mod-internal.hpp:33-101, eltwise-add-mod.cpp:16-40, eltwise-sub-mod.cpp:16-40
// per element; MultMod inputs < in_mf * q_e with in_mf in {1,2,4}, Add/Sub inputs < q_e).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
    rns_eltwise_kernel(u64* result, const u64* a, const u64* b, u64 per_mod, u64 count, int op, int in_mf,
                       const __grid_constant__ DyadicModuli mods) {
  const u64 total = per_mod * count / VEC;
  const u64 stride = (u64)gridDim.x * kThreads;
  for (u64 i = (u64)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const DyadicModulus& dm = mods.m[i * VEC / per_mod];
    const MulCtx c{dm.q, dm.mu, dm.shift};
    const Slots<VEC> x = ld_slots<VEC>(a + i * VEC), y = ld_slots<VEC>(b + i * VEC);
    Slots<VEC> r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      u64 xv = x.v[k], yv = y.v[k];
      if (op == kRnsAdd) {
        r.v[k] = csub(xv + yv, c.q);
        continue;
      }
      if (op == kRnsSub) {
        r.v[k] = xv >= yv ? xv - yv : xv + c.q - yv;
        continue;
      }
      if (in_mf >= 4) {
        xv = csub(xv, c.q << 1);
        yv = csub(yv, c.q << 1);
      }
      if (in_mf >= 2) {
        xv = csub(xv, c.q);
        yv = csub(yv, c.q);
      }
      r.v[k] = mulmod(xv, yv, c);
    }
    st_slots<VEC>(result + i * VEC, r);
  }
}

// ---- KeySwitch glue (key-switch-internal.cpp:60-198), every kernel batched over the RNS
// moduli of one parameter block; layouts are [modulus][component or digit][n].

// (:77-85, every digit reduced into every modulus, is folded into the forward transform: NttMulti::gather)

// :93-130: lazy 128-bit multiply-accumulate of the digits with the switching keys, one
// Shoup(hi, 2^64 mod q) + Barrett(lo) at the end, two conditional subtractions.
__global__ void __launch_bounds__(kThreads)
    ks_mac_kernel(u64* prod, const u64* ops, u64 ops_stride, const __grid_constant__ KeyPointers keys, u64 n,
                  u64 jcount, u64 kcc, u64 key_modulus_size, u64 count, const __grid_constant__ KsModuli mods,
                  int accumulate) {
  const u64 per_mod = kcc * n;
  const u64 g = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (g >= per_mod * count) return;
  const u64 e = g / per_mod, r = g - e * per_mod;
  const u64 k = r / n, l = r - k * n;
  const KsModulus& md = mods.m[e];
  const u64 key_off = n * md.c + k * key_modulus_size * n + l;
  const u64* op = ops + e * ops_stride + l;
  u64 lo = 0, hi = 0;
  for (u64 j = 0; j < jcount; ++j) {
    const u64 a = op[j * n];
    const u64 b = __ldcs(keys.p[j] + key_off);
    const u64 plo = a * b, phi = mulhi(a, b);
    lo += plo;
    hi += phi + (lo < plo);
  }
  u64 v = shoup_lazy(hi, md.a, md.b, md.q) + barrett64_lazy(lo, md.q, md.mu);  // < 4q
  v = csub(csub(v, md.q << 1), md.q);
  if (accumulate) v = csub(v + prod[g], md.q);
  prod[g] = v;
}

// :148-178: the special prime's part (coefficient form, [0, 2 q_last)), rounded and moved
// into modulus e:  t = (x + q_last/2) mod q_last;  out = (t mod q_e) + (q_e - (q_last/2 mod q_e))
__global__ void __launch_bounds__(kThreads)
    ks_round_kernel(u64* tmp, const u64* t_last, u64 per_mod /* kcc*n */, u64 q_last, u64 mu_last, u64 count,
                    const __grid_constant__ KsModuli mods) {
  const u64 g = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (g >= per_mod * count) return;
  const u64 e = g / per_mod, r = g - e * per_mod;
  const KsModulus& md = mods.m[e];
  u64 x = t_last[r] + (q_last >> 1);
  x = csub(barrett64_lazy(x, q_last, mu_last), q_last);
  if (q_last > md.q) x = csub(barrett64_lazy(x, md.q, md.mu), md.q);
  tmp[g] = x + md.a;
}

// :183-197:  r = (prod + 4 q_e - t_ntt) * modswitch mod q_e (operand < 8 q_e);  result += r mod q_e
__global__ void __launch_bounds__(kThreads)
    ks_finish_kernel(u64* result, const u64* prod, const u64* tmp, u64 n, u64 kcc, u64 decomp, u64 i0, u64 count,
                     const __grid_constant__ KsModuli mods) {
  const u64 per_mod = kcc * n;
  const u64 g = (u64)blockIdx.x * kThreads + threadIdx.x;
  if (g >= per_mod * count) return;
  const u64 e = g / per_mod, r = g - e * per_mod;
  const u64 k = r / n, l = r - k * n;
  const KsModulus& md = mods.m[e];
  u64 x = prod[g] + (md.q << 2) - tmp[g];
  x = reduce_from<8>(x, md.q);
  const u64 v = csub(shoup_lazy(x, md.a, md.b, md.q), md.q);
  u64* dst = result + n * (decomp * k + i0 + e) + l;
  *dst = csub(*dst + v, md.q);
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
  const bool vec = n % 2 == 0 &&
                   ((reinterpret_cast<uintptr_t>(result) | reinterpret_cast<uintptr_t>(op1) |
                     reinterpret_cast<uintptr_t>(op2)) & 15) == 0;
  u64 blocks = blocks_for(vec ? total / 2 : total);
  if (blocks > (u64)sms * 16) blocks = (u64)sms * 16;
  if (vec)
    dyadic_kernel<2><<<(unsigned)blocks, kThreads, 0, stream>>>(result, op1, op2, n, num_moduli, first, count, mods);
  else
    dyadic_kernel<1><<<(unsigned)blocks, kThreads, 0, stream>>>(result, op1, op2, n, num_moduli, first, count, mods);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_rns_eltwise(int op, u64* result, const u64* a, const u64* b, u64 per_mod, u64 count, int in_mf,
                               const DyadicModuli& mods, cudaStream_t stream) {
  const u64 total = per_mod * count;
  if (total == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool vec = per_mod % 2 == 0 &&
                   ((reinterpret_cast<uintptr_t>(result) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  u64 blocks = blocks_for(vec ? total / 2 : total);
  if (blocks > (u64)sms * 16) blocks = (u64)sms * 16;
  if (vec)
    rns_eltwise_kernel<2><<<(unsigned)blocks, kThreads, 0, stream>>>(result, a, b, per_mod, count, op, in_mf, mods);
  else
    rns_eltwise_kernel<1><<<(unsigned)blocks, kThreads, 0, stream>>>(result, a, b, per_mod, count, op, in_mf, mods);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_mac(u64* prod, const u64* ops, u64 ops_stride, const KeyPointers& keys, u64 n, u64 jcount,
                          u64 kcc, u64 key_modulus_size, u64 count, const KsModuli& mods, int accumulate,
                          cudaStream_t stream) {
  ks_mac_kernel<<<blocks_for(kcc * n * count), kThreads, 0, stream>>>(prod, ops, ops_stride, keys, n, jcount, kcc,
                                                                     key_modulus_size, count, mods, accumulate);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_round(u64* tmp, const u64* t_last, u64 n, u64 kcc, u64 q_last, u64 mu_last, u64 count,
                            const KsModuli& mods, cudaStream_t stream) {
  ks_round_kernel<<<blocks_for(kcc * n * count), kThreads, 0, stream>>>(tmp, t_last, kcc * n, q_last, mu_last, count,
                                                                       mods);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_ks_finish(u64* result, const u64* prod, const u64* tmp, u64 n, u64 kcc, u64 decomp, u64 i0,
                             u64 count, const KsModuli& mods, cudaStream_t stream) {
  ks_finish_kernel<<<blocks_for(kcc * n * count), kThreads, 0, stream>>>(result, prod, tmp, n, kcc, decomp, i0, count,
                                                                        mods);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hexl_b200
