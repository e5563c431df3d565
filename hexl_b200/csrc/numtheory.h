// Host number theory used by the C ABI (see numtheory.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace hexl_b200 {
namespace nt {
uint64_t mul_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t add_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t sub_mod(uint64_t x, uint64_t y, uint64_t q);
uint64_t pow_mod(uint64_t base, uint64_t exp, uint64_t q);
uint64_t inverse_mod(uint64_t x, uint64_t q);
uint64_t reverse_bits(uint64_t x, uint64_t bit_width);
bool is_prime(uint64_t n);
bool is_primitive_root(uint64_t root, uint64_t degree, uint64_t q);
uint64_t generate_primitive_root(uint64_t degree, uint64_t q);
uint64_t minimal_primitive_root(uint64_t degree, uint64_t q);
uint64_t multiply_factor(uint64_t operand, uint64_t shift, uint64_t q);
// q*x = -1 mod 2^r (Newton iteration on the 2-adic inverse; same value as the reference's bit-by-bit Hensel lift)
uint64_t neg_inverse_mod_pow2(uint32_t r, uint64_t q);
// T * 2^-r mod q for T = T_hi*2^64 + T_lo < q * 2^r (REDC)
uint64_t montgomery_reduce(uint64_t T_hi, uint64_t T_lo, uint64_t q, int r, uint64_t inv_mod);
std::vector<uint64_t> generate_primes(size_t num, size_t bits, bool prefer_small, size_t ntt_size);
}  // namespace nt
}  // namespace hexl_b200
