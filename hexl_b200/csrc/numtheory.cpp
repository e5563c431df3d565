// Host-side number theory behind the C ABI: the scalar helpers an NTT object
// needs for its one-off table construction, and the public helper functions of
// hexl/include/hexl/number-theory/number-theory.hpp.  Written from the
// mathematical definitions; semantics (argument order, edge cases) follow the
// reference functions cited at each definition.
#include "numtheory.h"

namespace hexl_b200 {
namespace nt {

typedef unsigned __int128 u128;

// (x*y) mod q, any x, y, q != 0          [MultiplyMod, number-theory.cpp:44-52]
uint64_t mul_mod(uint64_t x, uint64_t y, uint64_t q) {
  return static_cast<uint64_t>((static_cast<u128>(x) * y) % q);
}

// inputs < q                             [AddUIntMod / SubUIntMod, :61-73]
uint64_t add_mod(uint64_t x, uint64_t y, uint64_t q) {
  uint64_t s = x + y;
  return s >= q ? s - q : s;
}
uint64_t sub_mod(uint64_t x, uint64_t y, uint64_t q) { return x >= y ? x - y : x + (q - y); }

// base^exp mod q                         [PowMod, :76-87]
uint64_t pow_mod(uint64_t base, uint64_t exp, uint64_t q) {
  uint64_t acc = 1 % q;
  base %= q;
  while (exp) {
    if (exp & 1) acc = mul_mod(acc, base, q);
    base = mul_mod(base, base, q);
    exp >>= 1;
  }
  return acc;
}

// x^-1 mod q by the extended Euclidean algorithm; q == 1 gives 0
//                                        [InverseMod, :13-42]
uint64_t inverse_mod(uint64_t x, uint64_t q) {
  if (q == 1) return 0;
  __int128 r0 = q, r1 = x % q, t0 = 0, t1 = 1;
  while (r1 > 1) {
    __int128 k = r0 / r1;
    __int128 r2 = r0 - k * r1, t2 = t0 - k * t1;
    r0 = r1, r1 = r2;
    t0 = t1, t1 = t2;
  }
  if (t1 < 0) t1 += q;
  return static_cast<uint64_t>(t1);
}

// reverse the low bit_width bits         [ReverseBits, :150-163]
uint64_t reverse_bits(uint64_t x, uint64_t bit_width) {
  if (bit_width == 0) return 0;
  uint64_t r = 0;
  for (uint64_t b = 0; b < bit_width; ++b)
    if ((x >> b) & 1) r |= 1ull << (bit_width - 1 - b);
  return r;
}

// deterministic Miller-Rabin for 64-bit n with the first twelve primes as bases
//                                        [IsPrime, :166-212]
bool is_prime(uint64_t n) {
  static const uint64_t witnesses[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return false;
  for (uint64_t p : witnesses) {
    if (n == p) return true;
    if (n % p == 0) return false;
  }
  uint64_t odd = n - 1;
  int twos = 0;
  while (!(odd & 1)) odd >>= 1, ++twos;
  for (uint64_t a : witnesses) {
    uint64_t y = pow_mod(a, odd, n);
    if (y == 1 || y == n - 1) continue;
    bool composite = true;
    for (int k = 1; k < twos && composite; ++k) {
      y = mul_mod(y, y, n);
      if (y == n - 1) composite = false;
    }
    if (composite) return false;
  }
  return true;
}

// root has order exactly `degree` (a power of two) iff root^(degree/2) == -1
//                                        [IsPrimitiveRoot, :91-102]
bool is_primitive_root(uint64_t root, uint64_t degree, uint64_t q) {
  if (root == 0) return false;
  return pow_mod(root, degree / 2, q) == q - 1;
}

// some primitive degree-th root: g^((q-1)/degree) for the first g that works.
// The reference draws g at random (GeneratePrimitiveRoot, :106-124); any valid
// root satisfies its contract.
uint64_t generate_primitive_root(uint64_t degree, uint64_t q) {
  const uint64_t cofactor = (q - 1) / degree;
  for (uint64_t g = 2; g < q; ++g) {
    uint64_t cand = pow_mod(g, cofactor, q);
    if (is_primitive_root(cand, degree, q)) return cand;
    if (g > 4096) break;  // q is not a suitable prime
  }
  return 0;
}

// the smallest primitive degree-th root: the odd powers of any one primitive
// root are all of them            [MinimalPrimitiveRoot, :128-148]
uint64_t minimal_primitive_root(uint64_t degree, uint64_t q) {
  uint64_t g = generate_primitive_root(degree, q);
  if (!g) return 0;
  const uint64_t g2 = mul_mod(g, g, q);
  uint64_t best = g, cur = g;
  for (uint64_t k = 1; k < degree / 2; ++k) {
    cur = mul_mod(cur, g2, q);
    if (cur < best) best = cur;
  }
  return best;
}

// floor(operand * 2^shift / q), low 64 bits   [MultiplyFactor, number-theory.hpp:29-40]
uint64_t multiply_factor(uint64_t operand, uint64_t shift, uint64_t q) {
  return static_cast<uint64_t>((static_cast<u128>(operand) << shift) / q);
}

// primes p == 1 (mod 2*ntt_size) inside (2^bits, 2^(bits+1)), scanning upward
// from 2^bits + 1 or downward from the top     [GeneratePrimes, number-theory.cpp:214-261]
std::vector<uint64_t> generate_primes(size_t num, size_t bits, bool prefer_small, size_t ntt_size) {
  std::vector<uint64_t> out;
  if (num == 0 || bits >= 63 || ntt_size == 0) return out;
  const int64_t lo = (int64_t(1) << bits) + 1, hi = (int64_t(1) << (bits + 1)) - 1;
  const int64_t step = 2 * static_cast<int64_t>(ntt_size);
  int64_t cand = prefer_small ? lo : hi - (hi % step) + 1;
  while (prefer_small ? cand < hi : cand > lo) {
    if (is_prime(static_cast<uint64_t>(cand))) {
      out.push_back(static_cast<uint64_t>(cand));
      if (out.size() == num) break;
    }
    cand += prefer_small ? step : -step;
  }
  return out;
}

// HenselLemma2adicRoot of the reference (number-theory.hpp:303-336) lifts one bit per step; the 2-adic Newton
// iteration x <- x (2 - q x) doubles the number of correct bits per step and lands on the same unique value.
uint64_t neg_inverse_mod_pow2(uint32_t r, uint64_t q) {
  uint64_t x = q;                       // q * q = 1 mod 8 for odd q: 3 correct bits
  for (int i = 0; i < 5; ++i) x *= 2 - q * x;  // 6, 12, 24, 48, 96 bits
  const uint64_t mask = r >= 64 ? ~0ull : ((1ull << r) - 1);
  return (0 - x) & mask;
}

uint64_t montgomery_reduce(uint64_t T_hi, uint64_t T_lo, uint64_t q, int r, uint64_t inv_mod) {
  typedef unsigned __int128 u128;
  const uint64_t mask = (1ull << r) - 1;
  const uint64_t m = ((T_lo & mask) * inv_mod) & mask;
  const u128 t = (((u128)T_hi << 64) | T_lo) + (u128)m * q;
  const uint64_t s = (uint64_t)(t >> r);
  return s >= q ? s - q : s;
}

}  // namespace nt
}  // namespace hexl_b200
