/* TEST INFRASTRUCTURE ONLY -- see hexl_oracle.c.  Parity status: PINNED
 * (checked against the reference's golden vectors in tests/golden/ and against
 * oracle/_ref, the compiled reference itself; tests/test_oracle_*.py). */
#ifndef HEXL_ORACLE_H
#define HEXL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* number theory */
uint64_t orc_multiply_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t orc_add_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t orc_sub_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t orc_pow_mod(uint64_t base, uint64_t exp, uint64_t q);
uint64_t orc_inverse_mod(uint64_t x, uint64_t q);
uint64_t orc_reverse_bits(uint64_t x, uint64_t bit_width);
int orc_is_prime(uint64_t n);
int orc_is_primitive_root(uint64_t root, uint64_t degree, uint64_t q);
uint64_t orc_minimal_primitive_root(uint64_t degree, uint64_t q);
int orc_generate_primes(uint64_t* out, uint64_t num, uint64_t bits, int prefer_small,
                        uint64_t ntt_size);
uint64_t orc_multiply_factor(uint64_t operand, unsigned bit_shift, uint64_t q);

/* NTT tables, laid out exactly as the reference's getters return them */
void orc_ntt_tables(uint64_t n, uint64_t q, uint64_t root, uint64_t* w, uint64_t* w_precon,
                    uint64_t* inv_w, uint64_t* inv_w_precon);

/* transforms on `batch` back-to-back polynomials; tables from orc_ntt_tables */
void orc_ntt_forward(uint64_t* result, const uint64_t* operand, uint64_t n, uint64_t q,
                     const uint64_t* w, const uint64_t* w_precon, uint64_t in_mf,
                     uint64_t out_mf, uint64_t batch, int threads);
void orc_ntt_inverse(uint64_t* result, const uint64_t* operand, uint64_t n, uint64_t q,
                     const uint64_t* inv_w, const uint64_t* inv_w_precon, uint64_t in_mf,
                     uint64_t out_mf, uint64_t batch, int threads);
void orc_ntt_forward_textbook(uint64_t* operand, uint64_t n, uint64_t q, const uint64_t* w);
void orc_ntt_inverse_textbook(uint64_t* operand, uint64_t n, uint64_t q,
                              const uint64_t* inv_w);

/* element-wise ops */
void orc_eltwise_add_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                         uint64_t q);
void orc_eltwise_add_mod_scalar(uint64_t* r, const uint64_t* a, uint64_t b, uint64_t n,
                                uint64_t q);
void orc_eltwise_sub_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                         uint64_t q);
void orc_eltwise_sub_mod_scalar(uint64_t* r, const uint64_t* a, uint64_t b, uint64_t n,
                                uint64_t q);
void orc_eltwise_mult_mod(uint64_t* r, const uint64_t* a, const uint64_t* b, uint64_t n,
                          uint64_t q, uint64_t in_mf);
void orc_eltwise_fma_mod(uint64_t* r, const uint64_t* a, uint64_t b, const uint64_t* c,
                         uint64_t n, uint64_t q, uint64_t in_mf);
void orc_eltwise_reduce_mod(uint64_t* r, const uint64_t* a, uint64_t n, uint64_t q,
                            uint64_t in_mf, uint64_t out_mf);
void orc_eltwise_cmp_add(uint64_t* r, const uint64_t* a, uint64_t n, int cmp, uint64_t bound,
                         uint64_t diff);
void orc_eltwise_cmp_sub_mod(uint64_t* r, const uint64_t* a, uint64_t n, uint64_t q, int cmp,
                             uint64_t bound, uint64_t diff);

/* Montgomery-form helpers (number-theory.hpp:269-336; eltwise-reduce-mod-avx512.hpp:156-352, BitShift 64) */
uint64_t orc_hensel_lemma_2adic_root(uint32_t r, uint64_t q);
uint64_t orc_montgomery_reduce(uint64_t T_hi, uint64_t T_lo, uint64_t q, int r, uint64_t inv_mod);
void orc_eltwise_mont_reduce_mod(uint64_t* res, const uint64_t* a, const uint64_t* b, uint64_t n, uint64_t q, int r,
                                 uint64_t inv_mod);
void orc_eltwise_montgomery_form_in(uint64_t* res, const uint64_t* a, uint64_t R2_mod_q, uint64_t n, uint64_t q, int r,
                                    uint64_t inv_mod);
void orc_eltwise_montgomery_form_out(uint64_t* res, const uint64_t* a, uint64_t n, uint64_t q, int r, uint64_t inv_mod);

/* SEAL-shaped composites (reference: hexl/experimental/seal/) */
void orc_dyadic_multiply(uint64_t* result, const uint64_t* operand1, const uint64_t* operand2,
                         uint64_t n, const uint64_t* moduli, uint64_t num_moduli);
void orc_key_switch(uint64_t* result, const uint64_t* t_target, uint64_t n, uint64_t decomp_modulus_size,
                    uint64_t key_modulus_size, uint64_t rns_modulus_size, uint64_t key_component_count,
                    const uint64_t* moduli, const uint64_t* const* k_switch_keys,
                    const uint64_t* modswitch_factors);

#ifdef __cplusplus
}
#endif
#endif
