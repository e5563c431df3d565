/* hexl_b200.h -- C ABI of libhexl_b200.so, the Blackwell (sm_100a) drop-in for the
 * intel/hexl hot path: NTT::ComputeForward / ComputeInverse, the seven Eltwise*Mod
 * operations, and the SEAL-shaped callers built on them (DyadicMultiply, KeySwitch,
 * the NTT cache).
 *
 * Every entry point names the reference interface it replaces (file:line relative
 * to the intel/hexl v1.2.5 tree).  The C++ headers under include/hexl/ re-create
 * the reference's `intel::hexl` API (same names, overloads, defaults) as inline
 * forwarders to these symbols, so SEAL/OpenFHE-style callers re-link unchanged;
 * INTEGRATION.md shows the binding a maintainer would add on the reference side.
 *
 * Conventions
 *  - Plain pointers and sizes only; no C++ or torch types.
 *  - Every data pointer may be a DEVICE pointer (cudaMalloc / torch tensor
 *    storage; the call is enqueued on `stream` and returns without synchronising)
 *    or a HOST pointer (pageable or pinned; the call stages the buffers through
 *    the GPU -- H2D, kernel, D2H, chunked so copies and kernels overlap -- and
 *    returns when `result` is complete).  Unified-memory pointers
 *    (hexl_b200_managed_alloc) are worked on in place like device pointers, and
 *    with stream == NULL the call returns with the result complete.  All data
 *    pointers of one call must be of the same kind.  `result` may alias an input (in place), as in the
 *    reference (test/test-ntt.cpp:240-243).
 *  - `stream` is a cudaStream_t passed as void*; NULL = the legacy default stream.
 *  - Batched calls take `batch` independent units laid out back to back
 *    (unit u at offset u*n elements); batch = 1 is the reference's one-call shape.
 *  - Return value: 0 on success, negative hexl_b200_status otherwise;
 *    hexl_b200_last_error() returns a thread-local message.  There is NO CPU
 *    fallback: without a usable CUDA device every compute entry point fails
 *    with HEXL_B200_ERR_NO_DEVICE.
 *  - Argument validation mirrors the reference's HEXL_CHECKs
 *    (e.g. hexl/ntt/ntt-internal.cpp:191-200, hexl/eltwise/eltwise-fma-mod.cpp:20-40).
 *    Cheap checks (null, n == 0, mod factors, modulus range) are always on;
 *    the O(n) input-range checks only when hexl_b200_set_debug(1) was called
 *    (the reference does them only in HEXL_DEBUG builds, check.hpp:12-44).
 */
#ifndef HEXL_B200_H
#define HEXL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hexl_b200_status {
  HEXL_B200_OK = 0,
  HEXL_B200_ERR_INVALID_ARG = -1, /* a HEXL_CHECK of the reference would fire */
  HEXL_B200_ERR_NO_DEVICE = -2,   /* no CUDA device / driver */
  HEXL_B200_ERR_CUDA = -3,        /* a CUDA runtime call failed */
  HEXL_B200_ERR_ALLOC = -4,
  HEXL_B200_ERR_MIXED_POINTERS = -5 /* host and device pointers in one call */
} hexl_b200_status;

typedef struct hexl_b200_ntt hexl_b200_ntt; /* opaque, reference-counted */

/* ---- library / device management (no counterpart in the reference: it is CPU-only) */
const char* hexl_b200_version(void);
const char* hexl_b200_last_error(void);
int hexl_b200_device_count(void);
/* Devices used for HOST-pointer calls with batch > 1: units are split into
 * contiguous blocks, one block per listed device (no inter-GPU traffic).
 * Default: the calling thread's current device only. */
int hexl_b200_set_host_devices(const int* devices, int count);
void hexl_b200_set_debug(int on); /* O(n) input-range checks, like HEXL_DEBUG */
int hexl_b200_sync(void* stream);
/* Pinned host memory so host-pointer calls DMA at full PCIe rate (the
 * reference's AllocatorBase hook, hexl/include/hexl/util/allocator.hpp:12-24,
 * is the place a caller would plug these in). */
void* hexl_b200_host_alloc(size_t bytes);
void hexl_b200_host_free(void* p);
/* Unified (managed) memory: one pointer valid on the host and on every GPU, so a
 * caller's buffers are device-visible with no staging copy.  Calls on managed
 * buffers with stream == NULL return after the result is complete (the
 * reference's synchronous semantics); with a stream they are asynchronous. */
void* hexl_b200_managed_alloc(size_t bytes);
void hexl_b200_managed_free(void* p);
/* number of kernel launches this library has issued in this process */
uint64_t hexl_b200_launch_count(void);

/* ---- number theory (host side; hexl/include/hexl/number-theory/number-theory.hpp) */
uint64_t hexl_b200_multiply_mod(uint64_t x, uint64_t y, uint64_t q);      /* :83  */
uint64_t hexl_b200_add_uint_mod(uint64_t x, uint64_t y, uint64_t q);      /* :95  */
uint64_t hexl_b200_sub_uint_mod(uint64_t x, uint64_t y, uint64_t q);      /* :99  */
uint64_t hexl_b200_pow_mod(uint64_t base, uint64_t exp, uint64_t q);      /* :102 */
uint64_t hexl_b200_inverse_mod(uint64_t x, uint64_t q);                   /* :79  */
uint64_t hexl_b200_reverse_bits(uint64_t x, uint64_t bit_width);          /* :75  */
int hexl_b200_is_prime(uint64_t n);                                       /* :166 */
int hexl_b200_is_primitive_root(uint64_t root, uint64_t degree, uint64_t q); /* :108 */
uint64_t hexl_b200_generate_primitive_root(uint64_t degree, uint64_t q);  /* :112 */
uint64_t hexl_b200_minimal_primitive_root(uint64_t degree, uint64_t q);   /* :117 */
/* floor(operand * 2^bit_shift / q), bit_shift in {32, 52, 64} (MultiplyFactor, :19-51) */
uint64_t hexl_b200_multiply_factor(uint64_t operand, uint64_t bit_shift, uint64_t q);
/* GeneratePrimes (:181): writes up to num primes, returns how many were found */
int hexl_b200_generate_primes(uint64_t* out, size_t num, size_t bit_size, int prefer_small,
                              size_t ntt_size);

/* ---- NTT object (class NTT, hexl/include/hexl/ntt/ntt.hpp:22-293) ---------------- */
/* NTT(degree, q): ntt.hpp:54 -- uses the minimal primitive 2N-th root */
int hexl_b200_ntt_create(hexl_b200_ntt** out, uint64_t degree, uint64_t q);
/* NTT(degree, q, root_of_unity): ntt.hpp:75 */
int hexl_b200_ntt_create_with_root(hexl_b200_ntt** out, uint64_t degree, uint64_t q,
                                   uint64_t root_of_unity);
void hexl_b200_ntt_retain(hexl_b200_ntt* h);  /* NTT is copyable in the reference */
void hexl_b200_ntt_release(hexl_b200_ntt* h); /* ~NTT */
/* NTT::CheckArguments (ntt.hpp:90, ntt-internal.cpp:171-186): 1 if valid */
int hexl_b200_ntt_check_arguments(uint64_t degree, uint64_t q);
uint64_t hexl_b200_ntt_degree(const hexl_b200_ntt* h);           /* GetDegree  :119 */
uint64_t hexl_b200_ntt_modulus(const hexl_b200_ntt* h);          /* GetModulus :122 */
uint64_t hexl_b200_ntt_minimal_root(const hexl_b200_ntt* h);     /* GetMinimalRootOfUnity :116 */
/* Host copies of the tables in the reference's layouts (getters ntt.hpp:125-194).
 * `which`: 0 root powers (bit-reversed slots), 1 their 64-bit Shoup factors,
 * 2 inverse root powers (stage-sequential order of ntt-internal.cpp:144-154),
 * 3 their 64-bit Shoup factors.  Returns a pointer valid for the handle's life. */
const uint64_t* hexl_b200_ntt_table(const hexl_b200_ntt* h, int which);

/* Uploads the handle's tables to `device` (-1: the calling thread's current device) now instead of on
 * the first transform there.  The first use of a handle on a device allocates and copies synchronously,
 * which is not allowed while a stream is being captured into a CUDA graph: warm the handles (this call,
 * or one ordinary call) before capturing.  No counterpart in the reference (its tables live in host
 * memory, ntt-internal.cpp:54-169). */
int hexl_b200_ntt_prepare(hexl_b200_ntt* h, int device);

/* NTT::ComputeForward (ntt.hpp:99; ntt-internal.cpp:188-250): natural-order input,
 * bit-reversed output.  in_mf in {1,2,4}: inputs < in_mf*q; out_mf in {1,4}:
 * outputs in [0, out_mf*q).  `batch` polynomials back to back. */
int hexl_b200_ntt_forward(hexl_b200_ntt* h, uint64_t* result, const uint64_t* operand,
                          uint64_t input_mod_factor, uint64_t output_mod_factor,
                          uint64_t batch, void* stream);
/* NTT::ComputeInverse (ntt.hpp:109; ntt-internal.cpp:252-310): bit-reversed input,
 * natural-order output, includes the 1/N scale.  in_mf in {1,2}, out_mf in {1,2}. */
int hexl_b200_ntt_inverse(hexl_b200_ntt* h, uint64_t* result, const uint64_t* operand,
                          uint64_t input_mod_factor, uint64_t output_mod_factor,
                          uint64_t batch, void* stream);

/* RNS batches in ONE launch (the shape of the reference's callers: every ciphertext
 * polynomial exists once per modulus, key-switch-internal.cpp:49-55,82-88): `count`
 * handles of the same degree; of the count * batch_per_modulus polynomials laid out back
 * to back, polynomial u is transformed under handles[u / batch_per_modulus], with the
 * semantics of hexl_b200_ntt_forward / _inverse.  Device (or unified) pointers run as one
 * launch per kernel stage regardless of `count`; host pointers fall back to one staged
 * call per handle. */
int hexl_b200_ntt_forward_multi(hexl_b200_ntt* const* handles, uint64_t count, uint64_t* result,
                                const uint64_t* operand, uint64_t input_mod_factor,
                                uint64_t output_mod_factor, uint64_t batch_per_modulus, void* stream);
int hexl_b200_ntt_inverse_multi(hexl_b200_ntt* const* handles, uint64_t count, uint64_t* result,
                                const uint64_t* operand, uint64_t input_mod_factor,
                                uint64_t output_mod_factor, uint64_t batch_per_modulus, void* stream);

/* The other two steps of an RNS polynomial product, with the same batching (not in the
 * reference, whose callers loop over the moduli: dyadic-multiply-internal.cpp:50-72):
 * EltwiseMultMod (eltwise-mult-mod.hpp:23) over `num_moduli` blocks of n_per_modulus
 * elements, block e under moduli[e], in one launch;  and the whole negacyclic product
 * result = InvNTT(FwdNTT(a) .* FwdNTT(b)) of count * batch_per_modulus polynomials,
 * polynomial u under handles[u / batch_per_modulus] (BASELINE configs[3]: FwdNTT ->
 * EltwiseMultMod -> InvNTT; the point-wise product is folded into the inverse transform,
 * which multiplies on load), 4 to 6 launches whatever `count`.  Inputs < q, outputs
 * in [0, q); result may be a, b or a separate buffer. */
int hexl_b200_eltwise_mult_mod_multi(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                                     uint64_t n_per_modulus, const uint64_t* moduli, uint64_t num_moduli,
                                     uint64_t input_mod_factor, void* stream);
/* EltwiseAddMod / EltwiseSubMod (eltwise-add-mod.hpp:22, eltwise-sub-mod.hpp:22) over an RNS batch */
int hexl_b200_eltwise_add_mod_multi(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                                    uint64_t n_per_modulus, const uint64_t* moduli, uint64_t num_moduli, void* stream);
int hexl_b200_eltwise_sub_mod_multi(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                                    uint64_t n_per_modulus, const uint64_t* moduli, uint64_t num_moduli, void* stream);
int hexl_b200_poly_multiply_multi(hexl_b200_ntt* const* handles, uint64_t count, uint64_t* result,
                                  const uint64_t* a, const uint64_t* b, uint64_t batch_per_modulus, void* stream);

/* ---- element-wise operations (hexl/include/hexl/eltwise/ *.hpp) -------------------
 * n = number of elements (for batched use pass n = batch * N: the ops are
 * position-independent). */
/* EltwiseAddMod vector-vector, eltwise-add-mod.hpp:22 */
int hexl_b200_eltwise_add_mod(uint64_t* result, const uint64_t* operand1,
                              const uint64_t* operand2, uint64_t n, uint64_t modulus,
                              void* stream);
/* EltwiseAddMod vector-scalar, eltwise-add-mod.hpp:36 */
int hexl_b200_eltwise_add_mod_scalar(uint64_t* result, const uint64_t* operand1,
                                     uint64_t operand2, uint64_t n, uint64_t modulus,
                                     void* stream);
/* EltwiseSubMod vector-vector, eltwise-sub-mod.hpp:22 */
int hexl_b200_eltwise_sub_mod(uint64_t* result, const uint64_t* operand1,
                              const uint64_t* operand2, uint64_t n, uint64_t modulus,
                              void* stream);
/* EltwiseSubMod vector-scalar, eltwise-sub-mod.hpp:36 */
int hexl_b200_eltwise_sub_mod_scalar(uint64_t* result, const uint64_t* operand1,
                                     uint64_t operand2, uint64_t n, uint64_t modulus,
                                     void* stream);
/* EltwiseMultMod, eltwise-mult-mod.hpp:23; in_mf in {1,2,4} */
int hexl_b200_eltwise_mult_mod(uint64_t* result, const uint64_t* operand1,
                               const uint64_t* operand2, uint64_t n, uint64_t modulus,
                               uint64_t input_mod_factor, void* stream);
/* EltwiseFMAMod, eltwise-fma-mod.hpp:22; arg3 may be NULL; in_mf in {1,2,4,8} */
int hexl_b200_eltwise_fma_mod(uint64_t* result, const uint64_t* arg1, uint64_t arg2,
                              const uint64_t* arg3, uint64_t n, uint64_t modulus,
                              uint64_t input_mod_factor, void* stream);
/* EltwiseReduceMod, eltwise-reduce-mod.hpp:24; in_mf in {modulus,2,4}, out_mf in {1,2} */
int hexl_b200_eltwise_reduce_mod(uint64_t* result, const uint64_t* operand, uint64_t n,
                                 uint64_t modulus, uint64_t input_mod_factor,
                                 uint64_t output_mod_factor, void* stream);
/* EltwiseCmpAdd, eltwise-cmp-add.hpp:22; cmp = CMPINT value 0..7 (util.hpp:16-25) */
int hexl_b200_eltwise_cmp_add(uint64_t* result, const uint64_t* operand1, uint64_t n,
                              int cmp, uint64_t bound, uint64_t diff, void* stream);
/* EltwiseCmpSubMod, eltwise-cmp-sub-mod.hpp:24 */
int hexl_b200_eltwise_cmp_sub_mod(uint64_t* result, const uint64_t* operand1, uint64_t n,
                                  uint64_t modulus, int cmp, uint64_t bound, uint64_t diff,
                                  void* stream);

/* ---- Montgomery-form helpers (R = 2^r > q, q odd, r <= 62: the BitShift = 64 forms of the reference)
 * HenselLemma2adicRoot (number-theory.hpp:303): x in [0, 2^r) with q*x = -1 mod 2^r; 0 for invalid arguments */
uint64_t hexl_b200_hensel_lemma_2adic_root(uint32_t r, uint64_t q);
/* MontgomeryReduce<64> (number-theory.hpp:269-301): T * 2^-r mod q for T = T_hi*2^64 + T_lo < q * 2^r */
uint64_t hexl_b200_montgomery_reduce(uint64_t T_hi, uint64_t T_lo, uint64_t q, int r, uint64_t inv_mod);
/* EltwiseMontReduceModAVX512<64, r> (hexl/eltwise/eltwise-reduce-mod-avx512.hpp:156): result = a*b*R^-1 mod q;
 * EltwiseMontgomeryFormInAVX512 (:227): result = a*R mod q, given R^2 mod q;
 * EltwiseMontgomeryFormOutAVX512 (:298): result = a*R^-1 mod q.  Inputs < q, outputs in [0, q). */
int hexl_b200_eltwise_mont_reduce_mod(uint64_t* result, const uint64_t* a, const uint64_t* b, uint64_t n,
                                      uint64_t modulus, int r, uint64_t neg_inv_mod, void* stream);
int hexl_b200_eltwise_montgomery_form_in(uint64_t* result, const uint64_t* a, uint64_t R2_mod_q, uint64_t n,
                                         uint64_t modulus, int r, uint64_t neg_inv_mod, void* stream);
int hexl_b200_eltwise_montgomery_form_out(uint64_t* result, const uint64_t* a, uint64_t n, uint64_t modulus, int r,
                                          uint64_t neg_inv_mod, void* stream);

/* ---- SEAL-shaped composites built on the hot path (hexl/include/hexl/experimental/seal/)
 * `moduli`, `modswitch_factors` and the array `k_switch_keys` itself are small HOST
 * arrays; the coefficient buffers (result, operands, t_target, every k_switch_keys[j])
 * are all device pointers or all host pointers. */
/* NTT cache, GetNTT(N, modulus): ntt-cache.hpp:27-53.  Returns a retained handle
 * shared by every caller; release it with hexl_b200_ntt_release. */
int hexl_b200_ntt_get_cached(hexl_b200_ntt** out, uint64_t degree, uint64_t q);
/* DyadicMultiply, dyadic-multiply.hpp:26: (x0*y0, x0*y1 + x1*y0, x1*y1) per modulus;
 * operands hold 2 polynomials x num_moduli x n, result 3; result may alias an operand. */
int hexl_b200_dyadic_multiply(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                              uint64_t n, const uint64_t* moduli, uint64_t num_moduli, void* stream);
/* KeySwitch, key-switch.hpp:34 (CKKS): result (key_component_count x decomp x n) is
 * updated in place; t_target_iter_ptr holds decomp x n digits in NTT form;
 * k_switch_keys[j] holds key_component_count x key_modulus_size x n. */
int hexl_b200_key_switch(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
                         uint64_t decomp_modulus_size, uint64_t key_modulus_size, uint64_t rns_modulus_size,
                         uint64_t key_component_count, const uint64_t* moduli,
                         const uint64_t* const* k_switch_keys, const uint64_t* modswitch_factors,
                         void* stream);

/* Key-switch keys resident on the GPU.  The reference keeps the keys in caller memory and reads them on every
 * call (key-switch.hpp:34-39); a host caller of hexl_b200_key_switch therefore pays decomp x key_component_count
 * x key_modulus_size x n words of PCIe traffic per call.  hexl_b200_keys_upload copies the `decomp` key buffers
 * (host or device pointers, the layout KeySwitch takes) once to the current device -- to every device listed with
 * hexl_b200_set_host_devices when that was called -- and hexl_b200_key_switch_resident runs `batch` key switches
 * against them: ciphertext c uses result + c * key_component_count * decomp * n and t_target + c * decomp * n.
 * Host buffers are pipelined (copies of one ciphertext under the kernels of its neighbours) and split across the
 * devices holding the keys; device buffers run on `stream` on their own device. */
typedef struct hexl_b200_keys hexl_b200_keys;
int hexl_b200_keys_upload(hexl_b200_keys** out, const uint64_t* const* k_switch_keys, uint64_t n,
                          uint64_t decomp_modulus_size, uint64_t key_modulus_size, uint64_t key_component_count);
/* The same keys SHARDED BY RNS MODULUS over the devices of hexl_b200_set_host_devices (one shard per listed device; a
 * device listed twice carries two shards): shard s keeps only the key slices of its moduli.  ONE key switch then runs on
 * all shards at once -- the decomposed digits are all-gathered and the special prime's part is broadcast with P2P stores
 * over NVLink from the kernels that produce them (the exchange of key-switch-internal.cpp:60-131,134-198) -- which cuts
 * the latency of a single switch;
 * hexl_b200_key_switch_resident takes such a handle with HOST result / t_target buffers. */
int hexl_b200_keys_upload_sharded(hexl_b200_keys** out, const uint64_t* const* k_switch_keys, uint64_t n,
                                  uint64_t decomp_modulus_size, uint64_t key_modulus_size,
                                  uint64_t key_component_count);
void hexl_b200_keys_release(hexl_b200_keys* keys);
int hexl_b200_key_switch_resident(uint64_t* result, const uint64_t* t_target_iter_ptr, uint64_t n,
                                  uint64_t decomp_modulus_size, uint64_t key_modulus_size, uint64_t rns_modulus_size,
                                  uint64_t key_component_count, const uint64_t* moduli, const hexl_b200_keys* keys,
                                  const uint64_t* modswitch_factors, uint64_t batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEXL_B200_H */
