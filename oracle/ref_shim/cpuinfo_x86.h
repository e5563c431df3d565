// Stand-in for google/cpu_features' cpuinfo_x86.h (pinned 32b49eb in the
// reference: cmake/third-party/cpu-features/CMakeLists.txt.in:10-11; not
// vendored, no network).  The reference's only use is
// hexl/util/cpu-features.hpp:10,23-32, which reads five feature bits.
// This is test infrastructure for the oracle build, not product code.
#pragma once
namespace cpu_features {
struct X86Features {
  int avx512f, avx512dq, avx512vl, avx512ifma, avx512vbmi2;
};
struct X86Info {
  X86Features features;
};
inline X86Info GetX86Info() {
  X86Info info;
  __builtin_cpu_init();
  info.features.avx512f = __builtin_cpu_supports("avx512f");
  info.features.avx512dq = __builtin_cpu_supports("avx512dq");
  info.features.avx512vl = __builtin_cpu_supports("avx512vl");
  info.features.avx512ifma = __builtin_cpu_supports("avx512ifma");
  info.features.avx512vbmi2 = __builtin_cpu_supports("avx512vbmi2");
  return info;
}
}  // namespace cpu_features
