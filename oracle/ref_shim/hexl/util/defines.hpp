// Build-configuration header the reference's CMake would generate from
// hexl/include/hexl/util/defines.hpp.in:6-13 (configure_file writes it into the
// source tree, which is read-only here).  Written by hand for the oracle build:
// GNU toolchain, release mode (no HEXL_DEBUG).
#pragma once
#define HEXL_USE_GNU
#define HEXL_UNUSED(x) (void)(x)
