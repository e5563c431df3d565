// Internal interfaces between the C ABI (capi.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "modarith.cuh"

namespace hexl_b200 {

// ------------------------------------------------------------------ eltwise
enum class EltOp : int {
  AddVV, AddVS, SubVV, SubVS, MultVV, Fma, FmaNoAdd, Reduce, Copy, CmpAdd, CmpSubMod,
  MontMult, MontIn, MontOut  // Montgomery form, R = 2^shift: a*b/R, a*scalar/R (scalar = R^2 mod q), a/R; mu = -q^-1 mod R
};

struct EltParams {
  u64* result;
  const u64* a;
  const u64* b;  // second vector operand (AddVV/SubVV/MultVV: op2; Fma: arg3)
  u64 n;
  u64 q;
  u64 scalar;    // AddVS/SubVS: operand2; Fma: reduced arg2; Cmp*: bound
  u64 scalar_p;  // Fma: floor(arg2*2^64/q); Cmp*: diff
  u64 mu;        // MultVV: generalised-Barrett mu; Reduce/CmpSubMod: floor(2^64/q)
  int in_mf;     // MultVV/Fma: 1,2,4,8; Reduce: 0 means "== q", else 2 or 4
  int out_mf;    // Reduce: 1 or 2
  int shift;     // MultVV: ceil_log2(q) - 2
  int cmp;       // CMPINT 0..7
};

cudaError_t launch_eltwise(EltOp op, const EltParams& p, cudaStream_t stream);

// ---------------------------------------------------------------------- NTT
// Device-resident tables of one (N, q, root) on one device.
//   fwd[k], k in [1, N): forward twiddle of tree node k  ( = psi^bitrev(k), the
//           reference's root_of_unity_powers[k], ntt-internal.cpp:60-72 )
//   inv[k]: its modular inverse (the reference stores these re-ordered,
//           ntt-internal.cpp:144-154; here they keep the tree indexing)
// The children of node k are 2k and 2k+1; a sub-transform rooted at node b uses
// node b*2^s + i for its stage s, group i.
//
// NttDeviceParams: what a kernel needs to know about one (N, q), resident in device
// memory next to the tables, so that ONE launch can transform polynomials of several
// moduli (RNS batches): the launch carries a short list of pointers to these records.
struct NttDeviceParams {
  const Twiddle* fwd;
  const Twiddle* inv;
  u64 q, mu;
  Twiddle inv_n, inv_n_w;
  // generalised-Barrett constants of the point-wise product (eltwise-mult-mod-internal.hpp:52-99, alpha = 62,
  // beta = -2): prod_shift = bits(q) - 2, prod_mu = floor(2^(prod_shift + 64) / q).  Used by the inverse transform
  // that multiplies on load (NttMulti::mul).
  u64 prod_mu;
  int prod_shift;
};
struct NttDeviceTables {
  const Twiddle* fwd;
  const Twiddle* inv;
  const Twiddle32* fwd32;  // 32-bit copies of both tables, only for q < 2^30 (else nullptr)
  const Twiddle32* inv32;
  u64 n;
  int log_n;
  u64 q;
  u64 mu;           // floor(2^64 / q)
  Twiddle inv_n;    // N^-1 and its Shoup factor
  Twiddle inv_n_w;  // N^-1 * inv[1] and its Shoup factor
  Twiddle32 inv_n32, inv_n_w32;
  const NttDeviceParams* dparams;  // the same facts as a device-resident record
};
constexpr u64 kSmallModulusLimit = 1ull << 30;  // below: 4q < 2^32, the 32-bit kernels apply

// result/operand: `batch` polynomials back to back on the current device.
cudaError_t launch_ntt_forward(const NttDeviceTables& t, u64* result, const u64* operand,
                               int in_mf, int out_mf, u64 batch, cudaStream_t stream);
cudaError_t launch_ntt_inverse(const NttDeviceTables& t, u64* result, const u64* operand,
                               int in_mf, int out_mf, u64 batch, cudaStream_t stream);

// Multi-modulus launch: polynomial u (of `units` back to back) belongs to entry u / group.
// At most kParamBlock entries per call; all moduli share the degree 2^log_n.
constexpr int kParamBlock = 64;
constexpr int kMaxMirrors = 15;
struct NttMulti {
  const NttDeviceParams* p[kParamBlock];
  unsigned group;
  // Inverse transforms only: the kernel that writes the final values also writes them to `mirrors` more buffers at the
  // same offsets -- peer-mapped memory of other GPUs (P2P stores over NVLink).  This is how the sharded key switch
  // all-gathers its digits inside the transform that produces them instead of copying afterwards.
  unsigned mirrors;
  u64* mirror[kMaxMirrors];
  // Forward transforms only: when non-zero, unit u READS polynomial (u % gather) of `operand` (result still goes to
  // unit u) and every value is first reduced into its own modulus (any 64-bit value -> [0, q)).  This is KeySwitch's
  // "every digit into every modulus" step (key-switch-internal.cpp:77-85) folded into the transform that consumes it:
  // the digits are read from L2 instead of a decomp x rns x n intermediate being written to and read back from HBM.
  unsigned gather;
  // Inverse transforms only: when non-null, the kernel that reads `operand` multiplies every value by the value at the
  // same offset of `mul` (both canonical, i.e. forward outputs with output_mod_factor 1) before its first butterfly:
  // InvNTT(a (.) b) in one pass over the data -- the FwdNTT -> MultMod -> InvNTT chain of
  // dyadic-multiply-internal.cpp:17-73 / the product pipelines of the callers without the MultMod kernel and without
  // the product's round trip through HBM (24 B per coefficient less).
  const u64* mul;
};
// max_q = the largest modulus of the call: it selects the butterflies every entry can run
cudaError_t launch_ntt_multi(bool forward, const NttMulti& multi, int log_n, u64 min_q, u64 max_q, u64* result,
                             const u64* operand, int out_mf, u64 units, cudaStream_t stream);

// ----------------------------------------------------- SEAL-shaped composites
struct DyadicModulus {  // per RNS modulus: q and its generalised-Barrett constants
  u64 q, mu;
  int shift;
};
// Small per-call tables travel as kernel parameters (no upload, no synchronisation, capturable
// in a CUDA graph); longer lists are processed in blocks of kParamBlock entries.
struct DyadicModuli {
  DyadicModulus m[kParamBlock];
};
struct KeyPointers {
  const u64* p[kParamBlock];
};
// moduli [first, first + count) of a DyadicMultiply over `num_moduli` moduli
cudaError_t launch_dyadic_multiply(u64* result, const u64* op1, const u64* op2, u64 n, u64 num_moduli, u64 first,
                                   u64 count, const DyadicModuli& mods, cudaStream_t stream);
// EltwiseMultMod / AddMod / SubMod of `count` blocks of per_mod elements, block e under mods.m[e]
enum : int { kRnsMult = 0, kRnsAdd = 1, kRnsSub = 2 };
cudaError_t launch_rns_eltwise(int op, u64* result, const u64* a, const u64* b, u64 per_mod, u64 count, int in_mf,
                               const DyadicModuli& mods, cudaStream_t stream);

// KeySwitch glue, batched over the RNS moduli of one parameter block (entry e of `mods`
// describes modulus i0 + e).  KsModulus.a/b/c mean, per kernel:
//   reduce: -            mac: a,b = 2^64 mod q and its Shoup factor, c = slot of q in the key
//   round:  a = q - (q_last/2 mod q)        finish: a,b = mod-switch factor and its Shoup factor
struct KsModulus {
  u64 q, mu, a, b, c;
};
struct KsModuli {
  KsModulus m[kParamBlock];
};
// prod[e][k][l] (+)= sum_{j < jcount} ops[e][j][l] * keys[j][k][c_e][l]  mod q_e ; ops_stride = elements between e's
cudaError_t launch_ks_mac(u64* prod, const u64* ops, u64 ops_stride, const KeyPointers& keys, u64 n, u64 jcount,
                          u64 kcc, u64 key_modulus_size, u64 count, const KsModuli& mods, int accumulate,
                          cudaStream_t stream);
// tmp[e][k][l] = ((t_last[k][l] + q_last/2) mod q_last) mod q_e + a_e
cudaError_t launch_ks_round(u64* tmp, const u64* t_last, u64 n, u64 kcc, u64 q_last, u64 mu_last, u64 count,
                            const KsModuli& mods, cudaStream_t stream);
// result[k][i0+e][l] = (result + (prod[e][k][l] + 4 q_e - tmp[e][k][l]) * a_e) mod q_e ; result has `decomp` moduli per k
cudaError_t launch_ks_finish(u64* result, const u64* prod, const u64* tmp, u64 n, u64 kcc, u64 decomp, u64 i0,
                             u64 count, const KsModuli& mods, cudaStream_t stream);

// Stream-ordered scratch from the library's own memory pool (capi.cu): kept warm between calls, capturable.
cudaError_t scratch_alloc_async(void** p, size_t bytes, cudaStream_t stream);
void scratch_free_async(void* p, cudaStream_t stream);

// launches issued so far (all kernels of this library)
void count_launch(unsigned n = 1);
uint64_t launches_so_far();

}  // namespace hexl_b200
