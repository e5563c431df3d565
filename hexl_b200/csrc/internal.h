// Internal interfaces between the C ABI (capi.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "modarith.cuh"

namespace hexl_b200 {

// ------------------------------------------------------------------ eltwise
enum class EltOp : int {
  AddVV, AddVS, SubVV, SubVS, MultVV, Fma, FmaNoAdd, Reduce, Copy, CmpAdd, CmpSubMod
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
};
constexpr u64 kSmallModulusLimit = 1ull << 30;  // below: 4q < 2^32, the 32-bit kernels apply

// result/operand: `batch` polynomials back to back on the current device.
cudaError_t launch_ntt_forward(const NttDeviceTables& t, u64* result, const u64* operand,
                               int in_mf, int out_mf, u64 batch, cudaStream_t stream);
cudaError_t launch_ntt_inverse(const NttDeviceTables& t, u64* result, const u64* operand,
                               int in_mf, int out_mf, u64 batch, cudaStream_t stream);

// ----------------------------------------------------- SEAL-shaped composites
struct DyadicModulus {  // per RNS modulus: q and its generalised-Barrett constants
  u64 q, mu;
  int shift;
};
// Small per-call tables travel as kernel parameters (no upload, no synchronisation, capturable
// in a CUDA graph); longer lists are processed in blocks of this many entries.
constexpr int kParamBlock = 64;
struct DyadicModuli {
  DyadicModulus m[kParamBlock];
};
struct KeyPointers {
  const u64* p[kParamBlock];
};
// moduli [first, first + count) of a DyadicMultiply over `num_moduli` moduli
cudaError_t launch_dyadic_multiply(u64* result, const u64* op1, const u64* op2, u64 n, u64 num_moduli, u64 first,
                                   u64 count, const DyadicModuli& mods, cudaStream_t stream);
// digits [0, count) of `keys`/`operands`; accumulate != 0 adds to the value already in prod_i (mod q)
cudaError_t launch_ks_mac(u64* prod_i, const u64* operands, const KeyPointers& keys, u64 n, u64 count, u64 kcc,
                          u64 key_index, u64 key_modulus_size, u64 prod_stride_k, u64 q, u64 mu, Twiddle r64,
                          int accumulate, cudaStream_t stream);
cudaError_t launch_ks_round(u64* out, const u64* t_last, u64 n, u64 q_last, u64 mu_last, u64 q_i, u64 mu_i, u64 fix,
                            cudaStream_t stream);
cudaError_t launch_ks_finish(u64* result, const u64* prod, const u64* t_ntt, u64 n, u64 q, u64 ms, u64 ms_p,
                             cudaStream_t stream);

// launches issued so far (all kernels of this library)
void count_launch(unsigned n = 1);
uint64_t launches_so_far();

}  // namespace hexl_b200
