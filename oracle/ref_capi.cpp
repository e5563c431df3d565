// TEST INFRASTRUCTURE ONLY -- C-callable veneer over the UNMODIFIED reference
// (intel/hexl v1.2.5 compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libhexl_ref.so).  Nothing under hexl_b200/ may link or load this.
//
// Every entry point forwards to the reference's own public API
// (hexl/include/hexl/ntt/ntt.hpp:99,109; hexl/include/hexl/eltwise/*.hpp) or, for
// the "*_native" variants, to the reference's scalar C++ path
// (hexl/ntt/ntt-internal.hpp:34,98; hexl/eltwise/*-internal.hpp), which is the
// tier BASELINE.json's north_star names as the bit-exact oracle.
#include <stdint.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "eltwise/eltwise-add-mod-internal.hpp"
#include "eltwise/eltwise-cmp-add-internal.hpp"
#include "eltwise/eltwise-cmp-sub-mod-internal.hpp"
#include "eltwise/eltwise-fma-mod-internal.hpp"
#include "eltwise/eltwise-mult-mod-internal.hpp"
#include "eltwise/eltwise-reduce-mod-internal.hpp"
#include "eltwise/eltwise-sub-mod-internal.hpp"
#include "hexl/eltwise/eltwise-add-mod.hpp"
#include "hexl/eltwise/eltwise-cmp-add.hpp"
#include "hexl/eltwise/eltwise-cmp-sub-mod.hpp"
#include "hexl/eltwise/eltwise-fma-mod.hpp"
#include "hexl/eltwise/eltwise-mult-mod.hpp"
#include "hexl/eltwise/eltwise-reduce-mod.hpp"
#include "hexl/eltwise/eltwise-sub-mod.hpp"
#include "hexl/experimental/seal/dyadic-multiply.hpp"
#include "hexl/experimental/seal/key-switch.hpp"
#include "hexl/ntt/ntt.hpp"
#include "hexl/number-theory/number-theory.hpp"
#include "hexl/util/util.hpp"
#include "ntt/ntt-internal.hpp"
#include "util/cpu-features.hpp"
#ifdef HEXL_HAS_AVX512DQ
#include "eltwise/eltwise-reduce-mod-avx512.hpp"
#endif

using namespace intel::hexl;

namespace {
template <class F>
void parallel_units(uint64_t units, int threads, F&& body) {
  if (threads <= 1 || units <= 1) {
    for (uint64_t u = 0; u < units; ++u) body(u);
    return;
  }
  int nt = static_cast<int>(std::min<uint64_t>(threads, units));
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (int t = 0; t < nt; ++t) {
    pool.emplace_back([=, &body]() {
      // contiguous block split, one polynomial per call (README.md:264-265:
      // the library is single-threaded and thread-safe)
      uint64_t lo = units * t / nt, hi = units * (t + 1) / nt;
      for (uint64_t u = lo; u < hi; ++u) body(u);
    });
  }
  for (auto& th : pool) th.join();
}
}  // namespace

extern "C" {

// ---- which tier the reference will dispatch on this host ---------------------
int ref_has_avx512dq() { return has_avx512dq ? 1 : 0; }
int ref_has_avx512ifma() { return has_avx512ifma ? 1 : 0; }

// ---- number theory ---------------------------------------------------------
uint64_t ref_minimal_primitive_root(uint64_t degree, uint64_t q) {
  return MinimalPrimitiveRoot(degree, q);
}
int ref_is_prime(uint64_t n) { return IsPrime(n) ? 1 : 0; }
int ref_generate_primes(uint64_t* out, uint64_t num, uint64_t bits,
                        int prefer_small, uint64_t ntt_size) {
  std::vector<uint64_t> p = GeneratePrimes(num, bits, prefer_small != 0, ntt_size);
  for (size_t i = 0; i < p.size() && i < num; ++i) out[i] = p[i];
  return static_cast<int>(p.size());
}
uint64_t ref_inverse_mod(uint64_t x, uint64_t q) { return InverseMod(x, q); }
uint64_t ref_pow_mod(uint64_t b, uint64_t e, uint64_t q) { return PowMod(b, e, q); }
uint64_t ref_multiply_mod(uint64_t x, uint64_t y, uint64_t q) {
  return MultiplyMod(x, y, q);
}
uint64_t ref_reverse_bits(uint64_t x, uint64_t w) { return ReverseBits(x, w); }

// ---- NTT object --------------------------------------------------------------
void* ref_ntt_create(uint64_t n, uint64_t q) { return new NTT(n, q); }
void* ref_ntt_create_root(uint64_t n, uint64_t q, uint64_t root) {
  return new NTT(n, q, root);
}
void ref_ntt_destroy(void* h) { delete static_cast<NTT*>(h); }
uint64_t ref_ntt_root(void* h) {
  return static_cast<NTT*>(h)->GetMinimalRootOfUnity();
}
void ref_ntt_tables(void* h, uint64_t* w, uint64_t* w_precon64, uint64_t* inv_w,
                    uint64_t* inv_w_precon64) {
  NTT* t = static_cast<NTT*>(h);
  uint64_t n = t->GetDegree();
  if (w) std::copy_n(t->GetRootOfUnityPowers().data(), n, w);
  if (w_precon64) std::copy_n(t->GetPrecon64RootOfUnityPowers().data(), n, w_precon64);
  if (inv_w) std::copy_n(t->GetInvRootOfUnityPowers().data(), n, inv_w);
  if (inv_w_precon64)
    std::copy_n(t->GetPrecon64InvRootOfUnityPowers().data(), n, inv_w_precon64);
}

// public-API dispatch (AVX-512 when the host has it), `batch` polynomials of
// degree n laid out back to back, `threads` host threads.
void ref_ntt_forward(void* h, uint64_t* result, const uint64_t* operand,
                     uint64_t in_mf, uint64_t out_mf, uint64_t batch, int threads) {
  NTT* t = static_cast<NTT*>(h);
  uint64_t n = t->GetDegree();
  parallel_units(batch, threads, [&](uint64_t u) {
    t->ComputeForward(result + u * n, operand + u * n, in_mf, out_mf);
  });
}
void ref_ntt_inverse(void* h, uint64_t* result, const uint64_t* operand,
                     uint64_t in_mf, uint64_t out_mf, uint64_t batch, int threads) {
  NTT* t = static_cast<NTT*>(h);
  uint64_t n = t->GetDegree();
  parallel_units(batch, threads, [&](uint64_t u) {
    t->ComputeInverse(result + u * n, operand + u * n, in_mf, out_mf);
  });
}

// scalar ("native C++") tier, reached directly
void ref_ntt_forward_native(void* h, uint64_t* result, const uint64_t* operand,
                            uint64_t in_mf, uint64_t out_mf, uint64_t batch,
                            int threads) {
  NTT* t = static_cast<NTT*>(h);
  uint64_t n = t->GetDegree();
  parallel_units(batch, threads, [&](uint64_t u) {
    ForwardTransformToBitReverseRadix2(
        result + u * n, operand + u * n, n, t->GetModulus(),
        t->GetRootOfUnityPowers().data(), t->GetPrecon64RootOfUnityPowers().data(),
        in_mf, out_mf);
  });
}
void ref_ntt_inverse_native(void* h, uint64_t* result, const uint64_t* operand,
                            uint64_t in_mf, uint64_t out_mf, uint64_t batch,
                            int threads) {
  NTT* t = static_cast<NTT*>(h);
  uint64_t n = t->GetDegree();
  parallel_units(batch, threads, [&](uint64_t u) {
    InverseTransformFromBitReverseRadix2(
        result + u * n, operand + u * n, n, t->GetModulus(),
        t->GetInvRootOfUnityPowers().data(),
        t->GetPrecon64InvRootOfUnityPowers().data(), in_mf, out_mf);
  });
}
// textbook O(N log N) transforms the reference's own tests use as ground truth
// (hexl/ntt/ntt-radix-2.cpp:263-328); in place.
void ref_ntt_forward_textbook(void* h, uint64_t* operand) {
  NTT* t = static_cast<NTT*>(h);
  ReferenceForwardTransformToBitReverse(operand, t->GetDegree(), t->GetModulus(),
                                        t->GetRootOfUnityPowers().data());
}
void ref_ntt_inverse_textbook(void* h, uint64_t* operand) {
  NTT* t = static_cast<NTT*>(h);
  ReferenceInverseTransformFromBitReverse(operand, t->GetDegree(), t->GetModulus(),
                                          t->GetInvRootOfUnityPowers().data());
}
void ref_ntt_forward_radix4(void* h, uint64_t* result, const uint64_t* operand,
                            uint64_t in_mf, uint64_t out_mf) {
  NTT* t = static_cast<NTT*>(h);
  ForwardTransformToBitReverseRadix4(
      result, operand, t->GetDegree(), t->GetModulus(),
      t->GetRootOfUnityPowers().data(), t->GetPrecon64RootOfUnityPowers().data(),
      in_mf, out_mf);
}
void ref_ntt_inverse_radix4(void* h, uint64_t* result, const uint64_t* operand,
                            uint64_t in_mf, uint64_t out_mf) {
  NTT* t = static_cast<NTT*>(h);
  InverseTransformFromBitReverseRadix4(
      result, operand, t->GetDegree(), t->GetModulus(),
      t->GetInvRootOfUnityPowers().data(),
      t->GetPrecon64InvRootOfUnityPowers().data(), in_mf, out_mf);
}

// ---- eltwise, public dispatch; `batch` rows of n elements, `threads` threads ---
#define ROWS(expr)                                        \
  parallel_units(batch, threads, [&](uint64_t u) {        \
    const uint64_t o = u * n;                             \
    (void)o;                                              \
    expr;                                                 \
  })

void ref_eltwise_add_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                         uint64_t q, uint64_t batch, int threads) {
  ROWS(EltwiseAddMod(r + o, a + o, b + o, n, q));
}
void ref_eltwise_add_mod_scalar(uint64_t* r, const uint64_t* a, uint64_t b, uint64_t n,
                                uint64_t q, uint64_t batch, int threads) {
  ROWS(EltwiseAddMod(r + o, a + o, b, n, q));
}
void ref_eltwise_sub_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                         uint64_t q, uint64_t batch, int threads) {
  ROWS(EltwiseSubMod(r + o, a + o, b + o, n, q));
}
void ref_eltwise_sub_mod_scalar(uint64_t* r, const uint64_t* a, uint64_t b, uint64_t n,
                                uint64_t q, uint64_t batch, int threads) {
  ROWS(EltwiseSubMod(r + o, a + o, b, n, q));
}
void ref_eltwise_mult_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                          uint64_t q, uint64_t in_mf, uint64_t batch, int threads) {
  ROWS(EltwiseMultMod(r + o, a + o, b + o, n, q, in_mf));
}
void ref_eltwise_fma_mod(uint64_t* r, const uint64_t* a, uint64_t b, const uint64_t* c,
                         uint64_t n, uint64_t q, uint64_t in_mf, uint64_t batch,
                         int threads) {
  ROWS(EltwiseFMAMod(r + o, a + o, b, c ? c + o : nullptr, n, q, in_mf));
}
void ref_eltwise_reduce_mod(uint64_t* r, const uint64_t* a, uint64_t n, uint64_t q,
                            uint64_t in_mf, uint64_t out_mf, uint64_t batch,
                            int threads) {
  ROWS(EltwiseReduceMod(r + o, a + o, n, q, in_mf, out_mf));
}
void ref_eltwise_cmp_add(uint64_t* r, const uint64_t* a, uint64_t n, int cmp,
                         uint64_t bound, uint64_t diff, uint64_t batch, int threads) {
  ROWS(EltwiseCmpAdd(r + o, a + o, n, static_cast<CMPINT>(cmp), bound, diff));
}
void ref_eltwise_cmp_sub_mod(uint64_t* r, const uint64_t* a, uint64_t n, uint64_t q,
                             int cmp, uint64_t bound, uint64_t diff, uint64_t batch,
                             int threads) {
  ROWS(EltwiseCmpSubMod(r + o, a + o, n, q, static_cast<CMPINT>(cmp), bound, diff));
}

// ---- eltwise, scalar ("native") tier ------------------------------------------
void ref_eltwise_add_mod_native(uint64_t* r, const uint64_t* a, const uint64_t* b,
                                uint64_t n, uint64_t q) {
  EltwiseAddModNative(r, a, b, n, q);
}
void ref_eltwise_add_mod_scalar_native(uint64_t* r, const uint64_t* a, uint64_t b,
                                       uint64_t n, uint64_t q) {
  EltwiseAddModNative(r, a, b, n, q);
}
void ref_eltwise_sub_mod_native(uint64_t* r, const uint64_t* a, const uint64_t* b,
                                uint64_t n, uint64_t q) {
  EltwiseSubModNative(r, a, b, n, q);
}
void ref_eltwise_sub_mod_scalar_native(uint64_t* r, const uint64_t* a, uint64_t b,
                                       uint64_t n, uint64_t q) {
  EltwiseSubModNative(r, a, b, n, q);
}
void ref_eltwise_mult_mod_native(uint64_t* r, const uint64_t* a, const uint64_t* b,
                                 uint64_t n, uint64_t q, uint64_t in_mf) {
  switch (in_mf) {
    case 1: EltwiseMultModNative<1>(r, a, b, n, q); break;
    case 2: EltwiseMultModNative<2>(r, a, b, n, q); break;
    default: EltwiseMultModNative<4>(r, a, b, n, q); break;
  }
}
void ref_eltwise_fma_mod_native(uint64_t* r, const uint64_t* a, uint64_t b,
                                const uint64_t* c, uint64_t n, uint64_t q,
                                uint64_t in_mf) {
  switch (in_mf) {
    case 1: EltwiseFMAModNative<1>(r, a, b, c, n, q); break;
    case 2: EltwiseFMAModNative<2>(r, a, b, c, n, q); break;
    case 4: EltwiseFMAModNative<4>(r, a, b, c, n, q); break;
    default: EltwiseFMAModNative<8>(r, a, b, c, n, q); break;
  }
}
void ref_eltwise_reduce_mod_native(uint64_t* r, const uint64_t* a, uint64_t n,
                                   uint64_t q, uint64_t in_mf, uint64_t out_mf) {
  EltwiseReduceModNative(r, a, n, q, in_mf, out_mf);
}
void ref_eltwise_cmp_add_native(uint64_t* r, const uint64_t* a, uint64_t n, int cmp,
                                uint64_t bound, uint64_t diff) {
  EltwiseCmpAddNative(r, a, n, static_cast<CMPINT>(cmp), bound, diff);
}
void ref_eltwise_cmp_sub_mod_native(uint64_t* r, const uint64_t* a, uint64_t n,
                                    uint64_t q, int cmp, uint64_t bound,
                                    uint64_t diff) {
  EltwiseCmpSubModNative(r, a, n, q, static_cast<CMPINT>(cmp), bound, diff);
}

// ---- Montgomery-form helpers: the reference's scalar definitions (number-theory.hpp:269-336) for any r, and its
// AVX-512 element-wise helpers (eltwise/eltwise-reduce-mod-avx512.hpp:156-352) for the two r its own tests use.
uint64_t ref_hensel_lemma_2adic_root(uint32_t r, uint64_t q) { return HenselLemma2adicRoot(r, q); }
uint64_t ref_montgomery_reduce(uint64_t T_hi, uint64_t T_lo, uint64_t q, int r, uint64_t inv_mod) {
  return MontgomeryReduce<64>(T_hi, T_lo, q, r, (1ULL << r) - 1, inv_mod);
}
// kind: 0 = a*b*R^-1, 1 = a*R (b[0] = R^2 mod q), 2 = a*R^-1.  Returns 1 when the AVX-512 helper ran, 0 for the scalar loop.
int ref_eltwise_montgomery(int kind, uint64_t* res, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t q, int r,
                           uint64_t inv_mod) {
#ifdef HEXL_HAS_AVX512DQ
  if (has_avx512dq && (r == 46 || r == 61)) {
    if (kind == 0) {
      if (r == 46) EltwiseMontReduceModAVX512<64, 46>(res, a, b, n, q, inv_mod);
      else EltwiseMontReduceModAVX512<64, 61>(res, a, b, n, q, inv_mod);
    } else if (kind == 1) {
      if (r == 46) EltwiseMontgomeryFormInAVX512<64, 46>(res, a, b[0], n, q, inv_mod);
      else EltwiseMontgomeryFormInAVX512<64, 61>(res, a, b[0], n, q, inv_mod);
    } else {
      if (r == 46) EltwiseMontgomeryFormOutAVX512<64, 46>(res, a, n, q, inv_mod);
      else EltwiseMontgomeryFormOutAVX512<64, 61>(res, a, n, q, inv_mod);
    }
    return 1;
  }
#endif
  const uint64_t mask = (1ULL << r) - 1;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t hi = 0, lo = a[i];
    if (kind == 0) MultiplyUInt64(a[i], b[i], &hi, &lo);
    if (kind == 1) MultiplyUInt64(a[i], b[0], &hi, &lo);
    res[i] = MontgomeryReduce<64>(hi, lo, q, r, mask, inv_mod);
  }
  return 0;
}

// ---- SEAL-shaped composites (hexl/include/hexl/experimental/seal/*.hpp) --------
void ref_dyadic_multiply(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                         uint64_t n, const uint64_t* moduli, uint64_t num_moduli) {
  DyadicMultiply(result, operand1, operand2, n, moduli, num_moduli);
}
void ref_key_switch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
                    uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                    uint64_t rns_modulus_size, uint64_t key_component_count,
                    const uint64_t* moduli, const uint64_t** k_switch_keys,
                    const uint64_t* modswitch_factors) {
  KeySwitch(result, t_target_iter_ptr, n, decomp_modulus_size, key_modulus_size, rns_modulus_size,
            key_component_count, moduli, k_switch_keys, modswitch_factors);
}

}  // extern "C"
