// load_under_stores.cu — how long does a global READ take while the SM is streaming WRITES, and does it matter which
// unit issues it?  (Context: profile r1h of the fused eval kernel — 92 % of its traffic is writes at 5.3 TB/s — shows
// the consumer of the one-joint-ahead input prefetch holding 10 % of all stall samples although the load was issued a
// whole joint (~5 000 cycles) earlier, and an L2 prefetch several joints ahead made the kernel slower, not faster.)
//
// Every warp of a persistent grid (148 blocks x 8 warps, like the eval kernel) repeats: issue S coalesced 256-byte
// stores to a streaming output region, then time ONE 256-byte read of a line nobody touched before
//   mode 0: ld.global (LSU path), latency = clock64 from issue to first use
//   mode 1: cp.async.bulk 256 B -> shared memory + mbarrier wait (copy-engine path)
//   mode 2: ld.global issued BEFORE the S stores of the iteration, consumed after them (what the eval kernel does)
//   mode 3: cp.async.bulk issued before the S stores, waited for after them
// Prints the mean / p50 / p99 latency in cycles per mode and S.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o experiments/load_under_stores experiments/load_under_stores.cu
//   gpurun -- ./experiments/load_under_stores
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(const double* __restrict__ in, double* __restrict__ out, int iters, int S,
                                                unsigned* lat, double* sink) {
  __shared__ __align__(16) double stage[8][32];
  __shared__ __align__(8) unsigned long long bars[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[warp])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const unsigned bar = smem_u32(&bars[warp]);
  unsigned phase = 0;
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double* src = in + ((long long)it * nw + gw) * 32;           // a fresh 256-byte block per warp and iteration
    double* dst = out + (((long long)it * nw + gw) * S) * 32 + lane;   // S fresh 256-byte blocks
    long long t0 = 0, t1 = 0;
    double v = 0.0;
    if (MODE == 2) { t0 = clock64(); v = src[lane]; }
    if (MODE == 3) {
      t0 = clock64();
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];" ::"r"(
                         smem_u32(&stage[warp][0])),
                     "l"(src), "r"(bar)
                     : "memory");
      }
    }
    for (int s = 0; s < S; ++s) dst[(long long)s * 32] = (double)(it + s) + acc;
    if (MODE == 0) { t0 = clock64(); v = src[lane]; }
    if (MODE == 1) {
      t0 = clock64();
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];" ::"r"(
                         smem_u32(&stage[warp][0])),
                     "l"(src), "r"(bar)
                     : "memory");
      }
    }
    if (MODE == 1 || MODE == 3) {
      mbar_wait(bar, phase);
      phase ^= 1u;
      v = stage[warp][lane];
      __syncwarp();
    }
    acc += v;  // first use
    asm volatile("" ::"d"(acc) : "memory");
    t1 = clock64();
    if (lane == 0) lat[(long long)it * nw + gw] = (unsigned)(t1 - t0);
  }
  if (acc == 123.456) sink[0] = acc;
}

int main() {
  const int grid = 148, iters = 400;
  const long long nw = (long long)grid * 8;
  const int Smax = 64;
  double *in, *out, *sink;
  unsigned* lat;
  cudaMalloc(&in, sizeof(double) * 32 * nw * iters);
  cudaMalloc(&out, sizeof(double) * 32 * nw * iters * Smax);
  cudaMalloc(&sink, 8);
  cudaMalloc(&lat, sizeof(unsigned) * nw * iters);
  cudaMemset(in, 0, sizeof(double) * 32 * nw * iters);
  std::vector<unsigned> h(nw * iters);
  const int Ss[] = {0, 8, 24, 40, 64};
  printf("mode  S   mean    p50    p99   GB/s(write)\n");
  for (int mode = 0; mode < 4; ++mode)
    for (int S : Ss) {
      // flush L2 so that the reads go to HBM
      cudaMemset(out, 0, sizeof(double) * 32 * nw * iters * Smax);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<grid, 256>>>(in, out, iters, S, lat, sink);
      if (mode == 1) probe<1><<<grid, 256>>>(in, out, iters, S, lat, sink);
      if (mode == 2) probe<2><<<grid, 256>>>(in, out, iters, S, lat, sink);
      if (mode == 3) probe<3><<<grid, 256>>>(in, out, iters, S, lat, sink);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h.data(), lat, sizeof(unsigned) * nw * iters, cudaMemcpyDeviceToHost);
      std::vector<unsigned> v(h.begin() + nw * 20, h.end());  // skip the ramp-up
      std::sort(v.begin(), v.end());
      double mean = 0;
      for (unsigned x : v) mean += x;
      mean /= v.size();
      printf("%4d %3d %7.0f %6u %6u   %7.0f\n", mode, S, mean, v[v.size() / 2], v[v.size() * 99 / 100],
             (double)S * 256 * nw * iters / (ms * 1e6));
    }
  return 0;
}
