// A reference-style caller: written against the intel::hexl public API only
// (the scenarios of the reference's example/example.cpp:27-144 plus the NTT
// allocator/copy semantics of test/test-ntt.cpp:117-200), compiled against
// include/hexl/hexl.hpp and linked to libhexl_b200.so.  With a GPU it runs and
// checks the known answers; tests/test_host_api.py at least compiles and links it.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "hexl/hexl.hpp"

using namespace intel::hexl;

static int failures = 0;
static void Expect(const std::vector<uint64_t>& got, const std::vector<uint64_t>& want, const char* what) {
  if (got != want) {
    std::printf("MISMATCH in %s\n", what);
    ++failures;
  }
}

struct CountingAllocator {
  void* get(size_t bytes) {
    ++allocations;
    return std::malloc(bytes);
  }
  void put(void* p) { std::free(p); }
  static size_t allocations;
};
size_t CountingAllocator::allocations = 0;

namespace intel {
namespace hexl {
template <>
struct NTT::AllocatorAdapter<CountingAllocator> : public AllocatorInterface<NTT::AllocatorAdapter<CountingAllocator>> {
  explicit AllocatorAdapter(CountingAllocator&& a_) : a(std::move(a_)) {}
  void* allocate(size_t bytes_count) { return a.get(bytes_count); }
  void deallocate(void* p, size_t) { a.put(p); }
  CountingAllocator a;
};
}  // namespace hexl
}  // namespace intel

int main(int argc, char** argv) {
  const bool run = argc > 1;  // without arguments: link check + host-only API
  // host-side API (no GPU needed)
  if (MinimalPrimitiveRoot(8, 1234565441ULL) != 249725733ULL) ++failures;
  if (GeneratePrimes(1, 60, true, 1024).size() != 1) ++failures;
  if (Not(CMPINT::LT) != CMPINT::NLT || Not(CMPINT::TRUE) != CMPINT::FALSE) ++failures;
  {
    NTT ntt(4, 0xffffffffffc0001ULL);
    if (ntt.GetRootOfUnityPower(2) != 178930308976060547ULL) ++failures;
    if (ntt.GetAVX512RootOfUnityPowers().size() < 4) ++failures;
    NTT copy = ntt;  // copyable
    NTT assigned;
    assigned = NTT(8, 769);  // move-assignable (test/include/test/test-ntt-util.hpp:29)
    if (assigned.GetDegree() != 8 || copy.GetModulus() != ntt.GetModulus()) ++failures;
    CountingAllocator ca;
    NTT with_alloc(8, 769, std::move(ca));
    if (CountingAllocator::allocations == 0) ++failures;
    std::allocator<int> sa;
    (void)sa;
  }
  if (run) {
    {
      std::vector<uint64_t> op1{1, 2, 3, 4, 5, 6, 7, 8}, op2{1, 3, 5, 7, 2, 4, 6, 8};
      EltwiseAddMod(op1.data(), op1.data(), op2.data(), op1.size(), 10);
      Expect(op1, {2, 5, 8, 1, 7, 0, 3, 6}, "EltwiseAddMod vector-vector");
    }
    {
      std::vector<uint64_t> op1{1, 2, 3, 4, 5, 6, 7, 8};
      EltwiseAddMod(op1.data(), op1.data(), uint64_t(3), op1.size(), 10);
      Expect(op1, {4, 5, 6, 7, 8, 9, 0, 1}, "EltwiseAddMod vector-scalar");
    }
    {
      std::vector<uint64_t> op1{1, 2, 3, 4, 5, 6, 7, 8};
      EltwiseCmpAdd(op1.data(), op1.data(), op1.size(), CMPINT::NLE, 3, 5);
      Expect(op1, {1, 2, 3, 9, 10, 11, 12, 13}, "EltwiseCmpAdd");
    }
    {
      std::vector<uint64_t> op1{1, 2, 3, 4, 5, 6, 7};
      EltwiseCmpSubMod(op1.data(), op1.data(), op1.size(), 10, CMPINT::NLE, 4, 5);
      Expect(op1, {1, 2, 3, 4, 0, 1, 2}, "EltwiseCmpSubMod");
    }
    {
      std::vector<uint64_t> arg1{1, 2, 3, 4, 5, 6, 7, 8, 9};
      EltwiseFMAMod(arg1.data(), arg1.data(), 1, nullptr, arg1.size(), 769, 1);
      Expect(arg1, {1, 2, 3, 4, 5, 6, 7, 8, 9}, "EltwiseFMAMod");
    }
    {
      std::vector<uint64_t> op1{2, 4, 3, 2}, op2{2, 1, 2, 0};
      EltwiseMultMod(op1.data(), op1.data(), op2.data(), op1.size(), 769, 1);
      Expect(op1, {4, 4, 6, 0}, "EltwiseMultMod");
    }
    {
      std::vector<uint64_t> arg{1, 2, 3, 4, 5, 6, 7, 8};
      const auto want = arg;
      NTT ntt(8, 769);
      ntt.ComputeForward(arg.data(), arg.data(), 1, 1);
      ntt.ComputeInverse(arg.data(), arg.data(), 1, 1);
      Expect(arg, want, "NTT round trip");
    }
    {
      std::vector<uint64_t> arg{1, 2, 3, 4, 5, 6, 7, 8}, result(8, 0);
      EltwiseReduceMod(result.data(), arg.data(), arg.size(), 5, 2, 1);
      Expect(result, {1, 2, 3, 4, 0, 1, 2, 3}, "EltwiseReduceMod");
    }
    {
      // the 32-point known answer of test/test-ntt.cpp:391-403
      std::vector<uint64_t> in{401, 203, 221, 352, 487, 151, 405, 356, 343, 424, 635, 757, 457, 280, 624, 353,
                               496, 353, 624, 280, 457, 757, 635, 424, 343, 356, 405, 151, 487, 352, 221, 203};
      std::vector<uint64_t> want(32), out(32);
      for (int i = 0; i < 32; ++i) want[i] = i + 1;
      NTT ntt(32, 769);
      ntt.ComputeForward(out.data(), in.data(), 1, 1);
      Expect(out, want, "NTT N=32 known answer");
    }
    {
      // the same known answer on AlignedVector64 buffers whose storage comes from the
      // GPU-aware allocator strategies: unified memory (worked on in place, read back by
      // the host right after the call) and pinned host memory (staged at full PCIe rate)
      const uint64_t in[32] = {401, 203, 221, 352, 487, 151, 405, 356, 343, 424, 635, 757, 457, 280, 624, 353,
                               496, 353, 624, 280, 457, 757, 635, 424, 343, 356, 405, 151, 487, 352, 221, 203};
      NTT ntt(32, 769);
      AllocatorStrategyPtr strategies[2] = {std::make_shared<b200::ManagedStrategy>(),
                                            std::make_shared<b200::PinnedStrategy>()};
      for (auto& strat : strategies) {
        AlignedVector64<uint64_t> v(in, in + 32, AlignedAllocator<uint64_t, 64>(strat));
        ntt.ComputeForward(v.data(), v.data(), 1, 1);
        for (int i = 0; i < 32; ++i)
          if (v[i] != uint64_t(i + 1)) {
            std::printf("allocator-strategy buffer: forward[%d] = %llu\n", i, (unsigned long long)v[i]);
            ++failures;
            break;
          }
        ntt.ComputeInverse(v.data(), v.data(), 1, 1);
        for (int i = 0; i < 32; ++i)
          if (v[i] != in[i]) {
            ++failures;
            break;
          }
      }
    }
    {
      // RNS batch in one call: two moduli, two polynomials each, round trip
      NTT a(32, 769), b(32, 193);
      const NTT* list[2] = {&a, &b};
      std::vector<uint64_t> v(4 * 32), orig;
      for (size_t i = 0; i < v.size(); ++i) v[i] = (i * 37 + 5) % (i < 64 ? 769 : 193);
      orig = v;
      NTT::ComputeForwardMulti(list, 2, v.data(), v.data(), 1, 1, 2);
      std::vector<uint64_t> first(32);
      a.ComputeForward(first.data(), orig.data(), 1, 1);
      if (!std::equal(first.begin(), first.end(), v.begin())) ++failures;
      NTT::ComputeInverseMulti(list, 2, v.data(), v.data(), 1, 1, 2);
      Expect(v, orig, "multi-modulus round trip");
    }
    {
      // Montgomery form in and out (test/test-eltwise-reduce-mod-avx512.cpp:38-64, r = 46) and a Montgomery product
      const uint64_t modulus = 67280421310725ULL;
      const int r = 46;
      std::vector<uint64_t> in{0, 67280421310000, 25040294381203, 340231313, 769231483400, 90032324, 120042353, 1530};
      std::vector<uint64_t> out(in.size()), prod(in.size()), plain(in.size());
      const uint64_t R_reduced = (1ULL << r) % modulus;
      const uint64_t R2 = MultiplyMod(R_reduced, R_reduced, modulus);
      const uint64_t inv_mod = HenselLemma2adicRoot(r, modulus);
      if (inv_mod != 62463730494515ULL) ++failures;  // test/test-avx512-util.cpp:417
      EltwiseMontgomeryFormIn(out.data(), in.data(), R2, in.size(), modulus, r, inv_mod);
      EltwiseMontReduceMod(prod.data(), out.data(), in.data(), in.size(), modulus, r, inv_mod);  // (aR)(a)/R = a*a
      EltwiseMultMod(plain.data(), in.data(), in.data(), in.size(), modulus, 1);
      Expect(prod, plain, "Montgomery product");
      EltwiseMontgomeryFormOut(out.data(), out.data(), in.size(), modulus, r, inv_mod);
      Expect(out, in, "Montgomery form in/out");
    }
    {
      // KeySwitch against keys uploaded once equals KeySwitch with the keys in caller memory
      const uint64_t n = 64, decomp = 2, kms = 3, rns = 3, kcc = 2;
      std::vector<uint64_t> moduli = GeneratePrimes(kms, 40, true, n);
      std::vector<std::vector<uint64_t>> keys(decomp, std::vector<uint64_t>(kcc * kms * n));
      for (uint64_t j = 0; j < decomp; ++j)
        for (uint64_t k = 0; k < kcc; ++k)
          for (uint64_t i = 0; i < kms; ++i)
            for (uint64_t l = 0; l < n; ++l) keys[j][(k * kms + i) * n + l] = (j * 1315423911ULL + k * 2654435761ULL + i * 97 + l * l + 3) % moduli[i];
      const uint64_t* key_ptrs[2] = {keys[0].data(), keys[1].data()};
      std::vector<uint64_t> t(decomp * n), res(kcc * decomp * n), res2, modswitch(decomp);
      for (uint64_t j = 0; j < decomp; ++j) {
        modswitch[j] = InverseMod(moduli[kms - 1] % moduli[j], moduli[j]);
        for (uint64_t l = 0; l < n; ++l) t[j * n + l] = (l * 7919 + j) % moduli[j];
      }
      for (uint64_t k = 0; k < kcc; ++k)
        for (uint64_t i = 0; i < decomp; ++i)
          for (uint64_t l = 0; l < n; ++l) res[(k * decomp + i) * n + l] = (l * 104729 + k * 31 + i) % moduli[i];
      res2 = res;
      KeySwitch(res.data(), t.data(), n, decomp, kms, rns, kcc, moduli.data(), key_ptrs, modswitch.data());
      b200::KeySwitchKeys resident(key_ptrs, n, decomp, kms, kcc);
      KeySwitch(res2.data(), t.data(), n, decomp, kms, rns, kcc, moduli.data(), resident, modswitch.data());
      Expect(res2, res, "KeySwitch with resident keys");
    }
    bool threw = false;
    try {
      std::vector<uint64_t> a{1, 2};
      EltwiseMultMod(a.data(), a.data(), a.data(), 2, 769, 3);  // bad input_mod_factor
    } catch (const std::runtime_error&) {
      threw = true;
    }
    if (!threw) ++failures;
  }
  std::printf(failures ? "FAILED (%d)\n" : "ok%.0d\n", failures);
  return failures ? 1 : 0;
}
