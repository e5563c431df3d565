// Negacyclic NTT over Z_q[X]/(X^N + 1) for sm_100a.
//
// What is computed is exactly the reference's transform
// (hexl/ntt/ntt-radix-2.cpp:17-261 forward, :330-519 inverse; butterflies
// hexl/ntt/ntt-default.hpp:28-42,112-125): Cooley-Tukey forward from natural to
// bit-reversed order, Gentleman-Sande inverse back with N^-1 folded into the
// last stage, Harvey lazy ranges ([0,4q) forward, [0,2q) inverse), Shoup
// twiddles.  HOW it is computed is B200-shaped:
//
//   * The transform of size N = 2^n is the binary tree of its butterfly groups:
//     node 1 is the stage-0 group, node k has children 2k, 2k+1, and the twiddle
//     of node k is table[k] (for the forward table that IS the reference's
//     bit-reversed power layout).  A sub-transform rooted at node b over a
//     contiguous block of S elements uses node (b << s) + i in its stage s.
//   * "Row" kernel: one CTA owns a contiguous block of C = 2^c <= 16384
//     coefficients (a whole polynomial when N <= C, else one of N/C rows rooted
//     at node N/C + r).  Each thread holds 16 coefficients in registers and runs
//     4 butterfly stages per pass with no data movement; passes are separated by
//     a bank-conflict-free (XOR-swizzled) shared-memory transpose.  Global
//     loads/stores are fully coalesced and touch each coefficient exactly once.
//   * "Column" kernel (N > C only): the top log2(N/C) stages pair coefficients
//     C or more apart.  Each thread owns one column of R <= 32 coefficients
//     (stride S/R), keeps them in registers for log2(R) stages, twiddles staged
//     once per CTA in shared memory (they are the same for every column).
//   * Tiny N (< 16): one radix-2 stage per launch straight on global memory.
//
// No tensor cores: this is 64-bit integer modular arithmetic (IMAD-bound).
#include <cstdio>
#include <cstdlib>

#include "internal.h"

namespace hexl_b200 {
namespace {

// ----------------------------------------------------------------- butterflies

// Harvey forward butterfly, X,Y in [0,4q) -> [0,4q)   (ntt-default.hpp:28-42)
__device__ __forceinline__ void fwd_bfly(u64& X, u64& Y, const Twiddle w, u64 q, u64 two_q) {
  u64 tx = csub(X, two_q);
  u64 T = shoup_lazy(Y, w.w, w.wp, q);
  X = tx + T;
  Y = tx + two_q - T;
}

// Harvey inverse butterfly, X,Y in [0,2q) -> [0,2q)   (ntt-default.hpp:112-125)
__device__ __forceinline__ void inv_bfly(u64& X, u64& Y, const Twiddle w, u64 q, u64 two_q) {
  u64 s = X + Y;
  u64 d = X + two_q - Y;
  X = csub(s, two_q);
  Y = shoup_lazy(d, w.w, w.wp, q);
}

// Last inverse stage with N^-1 folded in (ntt-radix-2.cpp:484-509)
__device__ __forceinline__ void inv_bfly_last(u64& X, u64& Y, const Twiddle inv_n,
                                              const Twiddle inv_n_w, u64 q, u64 two_q) {
  u64 s = csub(X + Y, two_q);
  u64 d = X + two_q - Y;
  X = shoup_lazy(s, inv_n.w, inv_n.wp, q);
  Y = shoup_lazy(d, inv_n_w.w, inv_n_w.wp, q);
}

__device__ __forceinline__ Twiddle ld_tw(const Twiddle* p) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
  Twiddle t;
  t.w = v.x;
  t.wp = v.y;
  return t;
}

// forward output range: [0,4q) -> [0,q) when out_mf == 1  (ntt-radix-2.cpp:254-260)
__device__ __forceinline__ u64 fwd_out(u64 v, u64 q, u64 two_q, int out_mf) {
  return out_mf == 1 ? csub(csub(v, two_q), q) : v;
}
// inverse output range: [0,2q) -> [0,q) when out_mf == 1  (ntt-radix-2.cpp:511-518)
__device__ __forceinline__ u64 inv_out(u64 v, u64 q, int out_mf) {
  return out_mf == 1 ? csub(v, q) : v;
}

// --------------------------------------------------------------- row kernel
// Shared-memory index swizzle for 64-bit elements: XOR the 8-byte-bank index
// (low 4 bits) with the next 4 bits.  Conflict-free (per half-warp) for every
// access pattern of the passes below; a bijection inside each aligned block of
// 16 elements.
__device__ __forceinline__ unsigned swz(unsigned j) { return j ^ ((j >> 4) & 15u); }

// Coefficient index held in register slot e of thread u when the 4 register
// bits sit at bit position LB of the row-local index.
template <int LB>
__device__ __forceinline__ unsigned reg_index(unsigned u, int e) {
  return ((u >> LB) << (LB + 4)) | ((unsigned)e << LB) | (u & ((1u << LB) - 1u));
}

// Butterfly stages on row-local index bits HB..LOB (all inside [LB, LB+3]).
// FWD: bits descend (CT).  INV: bits ascend (GS).
template <int LOGC, int LB, int HB, int LOB, bool FWD>
__device__ __forceinline__ void reg_stages(u64 (&v)[16], unsigned u, u64 base,
                                           const Twiddle* __restrict__ tw, u64 q, u64 two_q,
                                           bool fold, Twiddle inv_n, Twiddle inv_n_w) {
#pragma unroll
  for (int step = 0; step <= HB - LOB; ++step) {
    const int beta = FWD ? HB - step : LOB + step;  // index bit of this stage
    const int eb = beta - LB;                       // register bit
    const int sp = LOGC - 1 - beta;                 // stage number inside the row
    const u64 node0 = (base << sp) + ((u64)(u >> LB) << (LB + 3 - beta));
    if (!FWD && sp == 0 && fold) {
      // root stage of the whole transform: one group, N^-1 folded in
#pragma unroll
      for (int l = 0; l < (1 << eb); ++l) inv_bfly_last(v[l], v[l | (1 << eb)], inv_n, inv_n_w, q, two_q);
    } else {
#pragma unroll
      for (int g = 0; g < (8 >> eb); ++g) {
        const Twiddle w = ld_tw(tw + node0 + g);
#pragma unroll
        for (int l = 0; l < (1 << eb); ++l) {
          const int e = (g << (eb + 1)) | l;
          if (FWD)
            fwd_bfly(v[e], v[e | (1 << eb)], w, q, two_q);
          else
            inv_bfly(v[e], v[e | (1 << eb)], w, q, two_q);
        }
      }
    }
  }
}

template <int LB_FROM, int LB_TO>
__device__ __forceinline__ void smem_exchange(u64 (&v)[16], u64* srow, unsigned u) {
#pragma unroll
  for (int e = 0; e < 16; ++e) srow[swz(reg_index<LB_FROM>(u, e))] = v[e];
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = srow[swz(reg_index<LB_TO>(u, e))];
  __syncthreads();
}

// Forward passes after pass 0: register bits move down by 4 per pass, clamped at 0.
template <int LOGC, int PASS>
__device__ __forceinline__ void fwd_passes(u64 (&v)[16], u64* srow, unsigned u, u64 base,
                                           const Twiddle* tw, u64 q, u64 two_q) {
  constexpr int PREV_LB = (LOGC - 4 * PASS) > 0 ? (LOGC - 4 * PASS) : 0;
  constexpr int HB = LOGC - 4 * PASS - 1;  // highest index bit not yet processed
  if constexpr (HB >= 0) {
    constexpr int LB = (HB - 3) > 0 ? (HB - 3) : 0;
    smem_exchange<PREV_LB, LB>(v, srow, u);
    reg_stages<LOGC, LB, HB, LB, true>(v, u, base, tw, q, two_q, false, Twiddle{}, Twiddle{});
    fwd_passes<LOGC, PASS + 1>(v, srow, u, base, tw, q, two_q);
  }
}

// Inverse passes: mirror image.  PASS counts down; pass P-1 is done first.
template <int LOGC, int PASS>
__device__ __forceinline__ void inv_passes(u64 (&v)[16], u64* srow, unsigned u, u64 base,
                                           const Twiddle* tw, u64 q, u64 two_q, bool fold,
                                           Twiddle inv_n, Twiddle inv_n_w) {
  // forward pass PASS handled bits HB..LB; the inverse handles the same bits ascending
  constexpr int HB = LOGC - 4 * PASS - 1;
  constexpr int LB = (HB - 3) > 0 ? (HB - 3) : 0;
  reg_stages<LOGC, LB, HB, LB, false>(v, u, base, tw, q, two_q, fold, inv_n, inv_n_w);
  if constexpr (PASS > 0) {
    constexpr int NHB = LOGC - 4 * (PASS - 1) - 1;
    constexpr int NLB = (NHB - 3) > 0 ? (NHB - 3) : 0;
    smem_exchange<LB, NLB>(v, srow, u);
    inv_passes<LOGC, PASS - 1>(v, srow, u, base, tw, q, two_q, fold, inv_n, inv_n_w);
  }
}

template <int LOGC>
struct RowCfg {
  static constexpr int C = 1 << LOGC;
  static constexpr int T = C / 16;                        // threads per row
  static constexpr int ROWS = T >= 256 ? 1 : 256 / T;     // rows per CTA
  static constexpr int THREADS = T * ROWS;
  static constexpr int PASSES = (LOGC + 3) / 4;
  static constexpr size_t SMEM = (size_t)ROWS * C * sizeof(u64);
};

// One CTA = ROWS rows of C contiguous coefficients.  rows_per_poly = N / C.
template <int LOGC>
__global__ void __launch_bounds__(RowCfg<LOGC>::THREADS)
    ntt_row_fwd(u64* result, const u64* operand, const Twiddle* __restrict__ tw, u64 q,
                u64 total_rows, unsigned rows_per_poly, int out_mf) {
  using Cfg = RowCfg<LOGC>;
  extern __shared__ __align__(16) u64 smem[];
  const unsigned row_local = threadIdx.x / Cfg::T, u = threadIdx.x % Cfg::T;
  u64 row = (u64)blockIdx.x * Cfg::ROWS + row_local;
  const bool active = row < total_rows;
  if (!active) row = total_rows - 1;  // keep barriers uniform; stores are masked
  const u64 base = (u64)rows_per_poly + (row % rows_per_poly);
  const u64 two_q = q << 1;
  const u64* in = operand + row * Cfg::C;
  u64* out = result + row * Cfg::C;
  u64* srow = smem + (size_t)row_local * Cfg::C;

  u64 v[16];
  constexpr int LB0 = LOGC - 4;  // pass 0: register bits are the top 4 index bits
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = __ldcs(in + reg_index<LB0>(u, e));
  reg_stages<LOGC, LB0, LOGC - 1, LB0, true>(v, u, base, tw, q, two_q, false, Twiddle{}, Twiddle{});
  fwd_passes<LOGC, 1>(v, srow, u, base, tw, q, two_q);
  // registers now hold 16 consecutive coefficients per thread (LB = 0)
  if constexpr (LOGC > 4) {
#pragma unroll
    for (int e = 0; e < 16; ++e) srow[swz(reg_index<0>(u, e))] = v[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = srow[swz(reg_index<LB0>(u, e))];
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < 16; ++e) __stcs(out + reg_index<LB0>(u, e), fwd_out(v[e], q, two_q, out_mf));
  }
}

template <int LOGC>
__global__ void __launch_bounds__(RowCfg<LOGC>::THREADS)
    ntt_row_inv(u64* result, const u64* operand, const Twiddle* __restrict__ tw, u64 q,
                u64 total_rows, unsigned rows_per_poly, int out_mf, int fold, Twiddle inv_n,
                Twiddle inv_n_w) {
  using Cfg = RowCfg<LOGC>;
  extern __shared__ __align__(16) u64 smem[];
  const unsigned row_local = threadIdx.x / Cfg::T, u = threadIdx.x % Cfg::T;
  u64 row = (u64)blockIdx.x * Cfg::ROWS + row_local;
  const bool active = row < total_rows;
  if (!active) row = total_rows - 1;
  const u64 base = (u64)rows_per_poly + (row % rows_per_poly);
  const u64 two_q = q << 1;
  const u64* in = operand + row * Cfg::C;
  u64* out = result + row * Cfg::C;
  u64* srow = smem + (size_t)row_local * Cfg::C;

  u64 v[16];
  constexpr int LB0 = LOGC - 4;
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = __ldcs(in + reg_index<LB0>(u, e));
  if constexpr (LOGC > 4) {
    // coalesced load layout -> 16 consecutive coefficients per thread
#pragma unroll
    for (int e = 0; e < 16; ++e) srow[swz(reg_index<LB0>(u, e))] = v[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = srow[swz(reg_index<0>(u, e))];
    __syncthreads();
  }
  inv_passes<LOGC, Cfg::PASSES - 1>(v, srow, u, base, tw, q, two_q, fold != 0, inv_n, inv_n_w);
  // last pass left the registers in the coalesced layout (LB = LOGC-4)
  if (active) {
    const bool final_out = fold != 0;  // only the kernel holding the root stage reduces
#pragma unroll
    for (int e = 0; e < 16; ++e)
      __stcs(out + reg_index<LB0>(u, e), final_out ? inv_out(v[e], q, out_mf) : v[e]);
  }
}

// ------------------------------------------------------------- column kernel
// Sub-blocks of S = 2^log_s contiguous coefficients, each rooted at tree node
// (N/S) + block_index.  A thread owns column c of one sub-block: R coefficients
// at stride S/R, and runs the sub-block's first log2(R) stages (forward) or last
// log2(R) stages (inverse) on them in registers.
template <int LOGR, bool FWD>
__global__ void __launch_bounds__(256)
    ntt_col(u64* result, const u64* operand, const Twiddle* __restrict__ tw, u64 q, int log_n,
            int log_s, u64 total_cols, int out_mf, int fold, Twiddle inv_n, Twiddle inv_n_w) {
  constexpr int R = 1 << LOGR;
  __shared__ Twiddle stw[R];
  const int log_cols = log_s - LOGR;                 // columns per sub-block (log2)
  const u64 g0 = (u64)blockIdx.x * blockDim.x;       // first column of this CTA
  // blockDim.x divides the columns of a sub-block, so the CTA shares one root node
  const u64 blk = g0 >> log_cols;                    // sub-block index over the whole batch
  const u64 blocks_per_poly = 1ull << (log_n - log_s);
  const u64 base = blocks_per_poly + (blk & (blocks_per_poly - 1));
  for (int l = threadIdx.x; l < R; l += blockDim.x) {
    if (l == 0) continue;                            // local node l = 2^s + i
    const int s = 31 - __clz(l);
    stw[l] = ld_tw(tw + (base << s) + (l - (1 << s)));
  }
  __syncthreads();
  const u64 g = g0 + threadIdx.x;
  if (g >= total_cols) return;
  const u64 c = g & ((1ull << log_cols) - 1);
  const u64 off = (blk << log_s) + c;
  const u64 two_q = q << 1;
  u64 v[R];
#pragma unroll
  for (int e = 0; e < R; ++e) v[e] = __ldcs(operand + off + ((u64)e << log_cols));
#pragma unroll
  for (int step = 0; step < LOGR; ++step) {
    const int s = FWD ? step : LOGR - 1 - step;      // stage inside the sub-block
    const int eb = LOGR - 1 - s;                     // register bit
    if (!FWD && fold && s == 0 && log_s == log_n) {
#pragma unroll
      for (int l = 0; l < (1 << eb); ++l) inv_bfly_last(v[l], v[l | (1 << eb)], inv_n, inv_n_w, q, two_q);
    } else {
#pragma unroll
      for (int gi = 0; gi < (1 << s); ++gi) {
        const Twiddle w = stw[(1 << s) + gi];
#pragma unroll
        for (int l = 0; l < (1 << eb); ++l) {
          const int e = (gi << (eb + 1)) | l;
          if (FWD)
            fwd_bfly(v[e], v[e | (1 << eb)], w, q, two_q);
          else
            inv_bfly(v[e], v[e | (1 << eb)], w, q, two_q);
        }
      }
    }
  }
  const bool final_out = !FWD && fold && log_s == log_n;
#pragma unroll
  for (int e = 0; e < R; ++e)
    __stcs(result + off + ((u64)e << log_cols), final_out ? inv_out(v[e], q, out_mf) : v[e]);
}

// --------------------------------------------------------- tiny-N stage kernel
// One radix-2 stage per launch on global memory; used for N < 16.
template <bool FWD>
__global__ void ntt_stage_simple(u64* result, const u64* src, const Twiddle* __restrict__ tw,
                                 u64 q, int log_n, int s /*stage: m = 2^s groups*/,
                                 u64 total_bflies, int out_mf, int last, Twiddle inv_n,
                                 Twiddle inv_n_w) {
  const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_bflies) return;
  const int log_t = log_n - 1 - s;
  const u64 half = 1ull << (log_n - 1);
  const u64 poly = g >> (log_n - 1), k = g & (half - 1);
  const u64 i = k >> log_t, jj = k & ((1ull << log_t) - 1);
  const u64 j = (poly << log_n) + (i << (log_t + 1)) + jj;
  const u64 two_q = q << 1;
  u64 X = src[j], Y = src[j + (1ull << log_t)];
  if (FWD) {
    fwd_bfly(X, Y, ld_tw(tw + (1ull << s) + i), q, two_q);
    if (last) {
      X = fwd_out(X, q, two_q, out_mf);
      Y = fwd_out(Y, q, two_q, out_mf);
    }
  } else if (last) {
    inv_bfly_last(X, Y, inv_n, inv_n_w, q, two_q);
    X = inv_out(X, q, out_mf);
    Y = inv_out(Y, q, out_mf);
  } else {
    inv_bfly(X, Y, ld_tw(tw + (1ull << s) + i), q, two_q);
  }
  result[j] = X;
  result[j + (1ull << log_t)] = Y;
}

// --------------------------------------------------------------- host side

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// log2 of the row length used for a transform of size 2^log_n
int pick_row_log(int log_n) {
  static const int max_row = [] {
    int v = env_int("HEXL_B200_MAX_ROW_LOG", 14);
    return v < 4 ? 4 : (v > 14 ? 14 : v);
  }();
  static const int split_row = [] {
    int v = env_int("HEXL_B200_SPLIT_ROW_LOG", 12);
    return v < 4 ? 4 : (v > 14 ? 14 : v);
  }();
  if (log_n <= max_row) return log_n;
  return split_row;
}

template <int LOGC>
cudaError_t launch_row(bool fwd, const NttDeviceTables& t, u64* result, const u64* operand,
                       u64 batch, int out_mf, int fold, cudaStream_t stream) {
  using Cfg = RowCfg<LOGC>;
  const unsigned rows_per_poly = (unsigned)(t.n >> LOGC);
  const u64 total_rows = batch * rows_per_poly;
  const unsigned grid = (unsigned)((total_rows + Cfg::ROWS - 1) / Cfg::ROWS);
  if (fwd) {
    if (Cfg::SMEM > 48 * 1024) {  // per-device attribute: set on every launch (cheap)
      cudaError_t e = cudaFuncSetAttribute(ntt_row_fwd<LOGC>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
    }
    ntt_row_fwd<LOGC><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(result, operand, t.fwd, t.q,
                                                                 total_rows, rows_per_poly, out_mf);
  } else {
    if (Cfg::SMEM > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ntt_row_inv<LOGC>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
    }
    ntt_row_inv<LOGC><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(
        result, operand, t.inv, t.q, total_rows, rows_per_poly, out_mf, fold, t.inv_n, t.inv_n_w);
  }
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_row_dyn(int log_c, bool fwd, const NttDeviceTables& t, u64* result,
                           const u64* operand, u64 batch, int out_mf, int fold,
                           cudaStream_t stream) {
  switch (log_c) {
#define ROW_CASE(L) \
  case L: return launch_row<L>(fwd, t, result, operand, batch, out_mf, fold, stream);
    ROW_CASE(4) ROW_CASE(5) ROW_CASE(6) ROW_CASE(7) ROW_CASE(8) ROW_CASE(9) ROW_CASE(10)
    ROW_CASE(11) ROW_CASE(12) ROW_CASE(13) ROW_CASE(14)
#undef ROW_CASE
  }
  return cudaErrorInvalidValue;
}

template <int LOGR>
cudaError_t launch_col(bool fwd, const NttDeviceTables& t, u64* result, const u64* operand,
                       u64 batch, int log_s, int out_mf, int fold, cudaStream_t stream) {
  const u64 total_cols = (batch << t.log_n) >> LOGR;
  const u64 cols_per_block = 1ull << (log_s - LOGR);
  const unsigned threads = (unsigned)(cols_per_block < 256 ? cols_per_block : 256);
  const unsigned grid = (unsigned)((total_cols + threads - 1) / threads);
  if (fwd)
    ntt_col<LOGR, true><<<grid, threads, 0, stream>>>(result, operand, t.fwd, t.q, t.log_n, log_s,
                                                      total_cols, out_mf, fold, t.inv_n, t.inv_n_w);
  else
    ntt_col<LOGR, false><<<grid, threads, 0, stream>>>(result, operand, t.inv, t.q, t.log_n, log_s,
                                                       total_cols, out_mf, fold, t.inv_n, t.inv_n_w);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_col_dyn(int log_r, bool fwd, const NttDeviceTables& t, u64* result,
                           const u64* operand, u64 batch, int log_s, int out_mf, int fold,
                           cudaStream_t stream) {
  switch (log_r) {
    case 1: return launch_col<1>(fwd, t, result, operand, batch, log_s, out_mf, fold, stream);
    case 2: return launch_col<2>(fwd, t, result, operand, batch, log_s, out_mf, fold, stream);
    case 3: return launch_col<3>(fwd, t, result, operand, batch, log_s, out_mf, fold, stream);
    case 4: return launch_col<4>(fwd, t, result, operand, batch, log_s, out_mf, fold, stream);
    case 5: return launch_col<5>(fwd, t, result, operand, batch, log_s, out_mf, fold, stream);
  }
  return cudaErrorInvalidValue;
}

// Split the top (log_n - log_c) stages into column passes of at most 5 stages,
// as even as possible, larger first.
int plan_col_passes(int top_stages, int out[8]) {
  if (top_stages <= 0) return 0;
  const int passes = (top_stages + 4) / 5;
  int left = top_stages;
  for (int p = 0; p < passes; ++p) {
    out[p] = (left + (passes - p) - 1) / (passes - p);
    left -= out[p];
  }
  return passes;
}

cudaError_t simple_transform(bool fwd, const NttDeviceTables& t, u64* result, const u64* operand,
                             int out_mf, u64 batch, cudaStream_t stream) {
  const u64 total = batch << (t.log_n - 1);
  const unsigned threads = 128, grid = (unsigned)((total + threads - 1) / threads);
  const u64* src = operand;
  for (int k = 0; k < t.log_n; ++k) {
    const int s = fwd ? k : t.log_n - 1 - k;
    const int last = k == t.log_n - 1;
    if (fwd)
      ntt_stage_simple<true><<<grid, threads, 0, stream>>>(result, src, t.fwd, t.q, t.log_n, s, total,
                                                           out_mf, last, t.inv_n, t.inv_n_w);
    else
      ntt_stage_simple<false><<<grid, threads, 0, stream>>>(result, src, t.inv, t.q, t.log_n, s, total,
                                                            out_mf, last, t.inv_n, t.inv_n_w);
    count_launch();
    src = result;
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_ntt_forward(const NttDeviceTables& t, u64* result, const u64* operand,
                               int /*in_mf*/, int out_mf, u64 batch, cudaStream_t stream) {
  if (batch == 0) return cudaSuccess;
  if (t.log_n < 4) return simple_transform(true, t, result, operand, out_mf, batch, stream);
  const int log_c = pick_row_log(t.log_n);
  int radices[8];
  const int ncol = plan_col_passes(t.log_n - log_c, radices);
  const u64* src = operand;
  int log_s = t.log_n;
  for (int p = 0; p < ncol; ++p) {
    cudaError_t e = launch_col_dyn(radices[p], true, t, result, src, batch, log_s, out_mf, 0, stream);
    if (e != cudaSuccess) return e;
    log_s -= radices[p];
    src = result;
  }
  return launch_row_dyn(log_c, true, t, result, src, batch, out_mf, 0, stream);
}

cudaError_t launch_ntt_inverse(const NttDeviceTables& t, u64* result, const u64* operand,
                               int /*in_mf*/, int out_mf, u64 batch, cudaStream_t stream) {
  if (batch == 0) return cudaSuccess;
  if (t.log_n < 4) return simple_transform(false, t, result, operand, out_mf, batch, stream);
  const int log_c = pick_row_log(t.log_n);
  int radices[8];
  const int ncol = plan_col_passes(t.log_n - log_c, radices);
  // the kernel that contains the root stage folds N^-1 and applies out_mf
  cudaError_t e = launch_row_dyn(log_c, false, t, result, operand, batch, out_mf, ncol == 0, stream);
  if (e != cudaSuccess) return e;
  // column passes in reverse: innermost (smallest sub-blocks) first
  int log_s = log_c;
  for (int p = ncol - 1; p >= 0; --p) {
    log_s += radices[p];
    e = launch_col_dyn(radices[p], false, t, result, result, batch, log_s, out_mf, p == 0, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace hexl_b200
