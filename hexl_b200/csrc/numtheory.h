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
std::vector<uint64_t> generate_primes(size_t num, size_t bits, bool prefer_small, size_t ntt_size);
}  // namespace nt
}  // namespace hexl_b200
