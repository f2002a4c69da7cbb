// tmem_scratch.cu — feasibility + cost of using Tensor Memory (tcgen05.ld/st) as per-thread scratch storage.
// Each thread owns one TMEM lane (lane quarter of its warp) and a column range; doubles are stored as 2 columns.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tm_st2(uint32_t addr, double v) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tm_ld2(uint32_t addr) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tm_ld8d(uint32_t addr, double* v) {  // 8 doubles = 16 columns, one instruction
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(addr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tm_st8d(uint32_t addr, const double* v) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(256) tmem_kernel(int ndoubles, int iters, long long* cyc, int* errors) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                     (uint64_t)__cvta_generic_to_shared(&tbase)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tbase;
  // lane quarter of this warp in bits 31:16, column offset in bits 15:0; warps 4..7 take the upper 256 columns
  const uint32_t my = base + (((uint32_t)(warp % 4) * 32u) << 16) + (uint32_t)(warp / 4) * 256u;
  const double tag = threadIdx.x * 1000.0 + blockIdx.x * 1.0e6;
  // 1) correctness: write ndoubles doubles, read them back (dynamic column index)
  for (int k = 0; k < ndoubles; ++k) tm_st2(my + 2 * k, tag + k);
  tm_wait_st();
  int bad = 0;
  for (int k = ndoubles - 1; k >= 0; --k) {
    const double v = tm_ld2(my + 2 * k);
    if (v != tag + k) ++bad;
  }
  // 8-double vectors
  double v8[8];
  for (int k = 0; k + 8 <= ndoubles; k += 8) {
    tm_ld8d(my + 2 * k, v8);
    for (int e = 0; e < 8; ++e) { if (v8[e] != tag + k + e) ++bad; v8[e] = -v8[e]; }
    tm_st8d(my + 2 * k, v8);
  }
  tm_wait_st();
  for (int k = 0; k < (ndoubles / 8) * 8; ++k)
    if (tm_ld2(my + 2 * k) != -(tag + k)) ++bad;
  if (bad) atomicAdd(errors, bad);
  // 2) cost: dependent read-modify-write chain on one slot (latency), then a streaming pass (throughput)
  long long t0 = clock64();
  double acc = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double v = tm_ld2(my + 2 * (it & 63));
    acc += v;
    tm_st2(my + 2 * (it & 63), acc);
    tm_wait_st();
  }
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) {
    tm_ld8d(my + 16 * (it & 7), v8);
    for (int e = 0; e < 8; ++e) v8[e] += 1.0;
    tm_st8d(my + 16 * (it & 7), v8);
    tm_wait_st();
  }
  long long t2 = clock64();
  // pure loads, 8 doubles each, independent
  double s = 0;
  for (int it = 0; it < iters; ++it) {
    tm_ld8d(my + 16 * (it & 7), v8);
    s += v8[0] + v8[7];
  }
  long long t3 = clock64();
  if (threadIdx.x % 32 == 0) {
    cyc[(blockIdx.x * 8 + warp) * 4 + 0] = t1 - t0;
    cyc[(blockIdx.x * 8 + warp) * 4 + 1] = t2 - t1;
    cyc[(blockIdx.x * 8 + warp) * 4 + 2] = t3 - t2;
    cyc[(blockIdx.x * 8 + warp) * 4 + 3] = (long long)(acc + s);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
}

// shared-memory equivalent of the same access pattern for comparison
__global__ void __launch_bounds__(256) smem_kernel(int iters, long long* cyc) {
  extern __shared__ double sm[];
  double* my = sm + (threadIdx.x / 32) * 128 * 32 + threadIdx.x % 32;
  for (int k = 0; k < 128; ++k) my[k * 32] = k;
  long long t0 = clock64();
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
    acc += my[(it & 63) * 32];
    my[(it & 63) * 32] = acc;
  }
  long long t1 = clock64();
  double v8[8];
  for (int it = 0; it < iters; ++it) {
    for (int e = 0; e < 8; ++e) v8[e] = my[((it & 7) * 8 + e) * 32] + 1.0;
    for (int e = 0; e < 8; ++e) my[((it & 7) * 8 + e) * 32] = v8[e];
  }
  long long t2 = clock64();
  if (threadIdx.x % 32 == 0) {
    cyc[(blockIdx.x * 8 + threadIdx.x / 32) * 4 + 0] = t1 - t0;
    cyc[(blockIdx.x * 8 + threadIdx.x / 32) * 4 + 1] = t2 - t1;
    cyc[(blockIdx.x * 8 + threadIdx.x / 32) * 4 + 3] = (long long)acc + (long long)v8[0];
  }
}

int main() {
  long long* cyc; int* err;
  cudaMallocManaged(&cyc, 148 * 8 * 4 * sizeof(long long));
  cudaMallocManaged(&err, sizeof(int));
  const int iters = 2000;
  for (int warps : {1, 4, 6, 8}) {
    *err = 0;
    for (int i = 0; i < 148 * 8 * 4; ++i) cyc[i] = 0;
    tmem_kernel<<<148, warps * 32>>>(128, iters, cyc, err);
    cudaError_t e = cudaDeviceSynchronize();
    printf("TMEM warps/SM=%d: %s errors=%d  rmw-chain %.1f cyc/iter  ld8+st8 %.1f cyc/iter  ld8 %.1f cyc/iter\n", warps,
           cudaGetErrorString(e), *err, cyc[0] / (double)iters, cyc[1] / (double)iters, cyc[2] / (double)iters);
    if (e != cudaSuccess) return 1;
    cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 128 * 32 * 8);
    smem_kernel<<<148, warps * 32, warps * 128 * 32 * 8>>>(iters, cyc);
    e = cudaDeviceSynchronize();
    printf("SMEM warps/SM=%d: %s  rmw-chain %.1f cyc/iter  ld8+st8 %.1f cyc/iter\n", warps, cudaGetErrorString(e),
           cyc[0] / (double)iters, cyc[1] / (double)iters);
  }
  return 0;
}
