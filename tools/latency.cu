// Host-side cost of one C-ABI call on device pointers (what a caller issuing many small
// NTTs pays per call), and the launch+execute latency of a single small transform.
//   nvcc -O2 -std=c++17 -I include -o tools/bin/latency tools/latency.cu -L hexl_b200/lib -lhexl_b200 -Xlinker -rpath=$PWD/hexl_b200/lib
#include <chrono>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "hexl_b200.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

__global__ void empty_kernel() {}

int main() {
  cudaStream_t s;
  cudaStreamCreate(&s);
  {  // the floor under any one-call-one-sync pattern on this box, and what pointer classification costs
    for (int i = 0; i < 100; ++i) empty_kernel<<<1, 32, 0, s>>>();
    cudaStreamSynchronize(s);
    double t0 = now();
    for (int i = 0; i < 2000; ++i) {
      empty_kernel<<<1, 32, 0, s>>>();
      cudaStreamSynchronize(s);
    }
    const double t_floor = (now() - t0) / 2000;
    t0 = now();
    for (int i = 0; i < 5000; ++i) empty_kernel<<<1, 32, 0, s>>>();
    const double t_issue = (now() - t0) / 5000;
    cudaStreamSynchronize(s);
    void* d = nullptr;
    cudaMalloc(&d, 4096);
    cudaPointerAttributes a;
    t0 = now();
    for (int i = 0; i < 20000; ++i) cudaPointerGetAttributes(&a, d);
    const double t_attr = (now() - t0) / 20000;
    std::printf("floor: empty kernel launch+sync %.2f us, launch alone %.2f us, cudaPointerGetAttributes %.3f us\n",
                t_floor * 1e6, t_issue * 1e6, t_attr * 1e6);
    cudaFree(d);
  }
  for (int logn : {10, 12, 14, 16}) {
    const uint64_t n = 1ull << logn;
    uint64_t q = 0;
    hexl_b200_generate_primes(&q, 1, 50, 1, n);
    hexl_b200_ntt* h = nullptr;
    if (hexl_b200_ntt_create(&h, n, q)) { std::printf("create failed: %s\n", hexl_b200_last_error()); return 1; }
    uint64_t *a, *b;
    cudaMalloc(&a, n * 8);
    cudaMalloc(&b, n * 8);
    cudaMemset(a, 0, n * 8);
    for (int i = 0; i < 100; ++i) hexl_b200_ntt_forward(h, b, a, 1, 1, 1, s);
    cudaStreamSynchronize(s);
    const int iters = 5000;
    double t0 = now();
    for (int i = 0; i < iters; ++i) hexl_b200_ntt_forward(h, b, a, 1, 1, 1, s);
    double t_issue = now() - t0;
    cudaStreamSynchronize(s);
    double t_all = now() - t0;
    t0 = now();
    for (int i = 0; i < 1000; ++i) {
      hexl_b200_ntt_forward(h, b, a, 1, 1, 1, s);
      cudaStreamSynchronize(s);
    }
    double t_sync = (now() - t0) / 1000;
    t0 = now();
    for (int i = 0; i < iters; ++i) hexl_b200_eltwise_mult_mod(b, a, a, n, q, 1, s);
    double t_elt = now() - t0;
    cudaStreamSynchronize(s);
    uint64_t *ha = (uint64_t*)hexl_b200_host_alloc(n * 8), *hb2 = (uint64_t*)hexl_b200_host_alloc(n * 8);
    for (uint64_t i = 0; i < n; ++i) ha[i] = i % q;
    for (int i = 0; i < 20; ++i) hexl_b200_ntt_forward(h, hb2, ha, 1, 1, 1, nullptr);
    t0 = now();
    for (int i = 0; i < 500; ++i) hexl_b200_ntt_forward(h, hb2, ha, 1, 1, 1, nullptr);
    const double t_host = (now() - t0) / 500;
    hexl_b200_host_free(ha);
    hexl_b200_host_free(hb2);
    std::printf("N=2^%d  host-pointer (pinned) forward NTT, complete on return: %.2f us\n", logn, t_host * 1e6);
    std::printf("N=2^%d  forward NTT: host issue %.2f us/call, back-to-back throughput %.2f us/call, call+sync %.2f us;  MultMod host issue %.2f us/call\n",
                logn, t_issue / iters * 1e6, t_all / iters * 1e6, t_sync * 1e6, t_elt / iters * 1e6);
    hexl_b200_ntt_release(h);
    cudaFree(a);
    cudaFree(b);
  }
  return 0;
}
