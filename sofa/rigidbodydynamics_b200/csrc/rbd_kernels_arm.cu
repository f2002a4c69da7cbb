// rbd_kernels_arm.cu — the fused eval of fixed-base serial chains with a compile-time joint count (arm_instance,
// rbd_arm_body.h): joint table in the kernel-parameter constant bank, both passes unrolled, A columns in registers.
#include <cuda_runtime.h>

#include "rbd_kernels.h"
#include "rbd_sweeps.h"

namespace rbd {

extern __shared__ __align__(16) double rbd_smem[];

template <int NJ, int TILE>
__global__ void __launch_bounds__(256, 1) eval_arm_kernel(const __grid_constant__ chain::ArmParams<NJ> m, const EvalArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const chain::Stack S = chain::make_stack(rbd_smem, threadIdx.x, RBD_ARM_LEVEL * (NJ - 1));
  // tensor memory: 6 NJ doubles per thread (the subspace columns during the forward pass); the base address lands in the
  // word behind the stacks
  unsigned* const tm_base_p = reinterpret_cast<unsigned*>(rbd_smem + (size_t)blockDim.x * RBD_ARM_LEVEL * (NJ - 1));
  constexpr int kTmDoubles = 6 * NJ;
  constexpr int kTmCols = 2 * kTmDoubles * 2 <= 32 ? 32 : 2 * kTmDoubles * 2 <= 64 ? 64 : 2 * kTmDoubles * 2 <= 128 ? 128 : 256;  // 8 warps: 2 per lane quarter
  static_assert(2 * kTmDoubles * 2 <= 256, "tensor-memory share of a thread");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                     (unsigned long long)__cvta_generic_to_shared(tm_base_p)),
                 "r"(kTmCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  chain::TmStack T;
  T.base = *tm_base_p + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 2u * (unsigned)kTmDoubles;
  const long long ntiles = (a.n + 31) / 32;
  for (long long t = (long long)blockIdx.x * wpb + warp; t < ntiles; t += (long long)gridDim.x * wpb) {
    long long inst = t * 32 + lane;
    if (inst >= a.n) inst = a.n - 1;  // lanes past the end recompute the last instance (warp-collective tensor-memory accesses)
    chain::arm_instance<NJ, TILE>(m, chain::make_field_t<TILE>(const_cast<double*>(a.q), inst, NJ, a.tile),
                                  chain::make_field_t<TILE>(const_cast<double*>(a.v), inst, NJ, a.tile),
                                  chain::make_field_t<TILE>(const_cast<double*>(a.a), inst, NJ, a.tile),
                                  chain::make_field_t<TILE>(a.oMi, inst, 12 * NJ, a.tile),
                                  chain::make_field_t<TILE>(a.J, inst, 6 * NJ, a.tile),
                                  chain::make_field_t<TILE>(a.M, inst, NJ * (NJ + 1) / 2, a.tile),
                                  chain::make_field_t<TILE>(a.tau, inst, NJ, a.tile), S, T);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_base_p), "r"(kTmCols) : "memory");
}

#ifndef RBD_ARM_PERSISTENT
#define RBD_ARM_PERSISTENT 1
#endif
template <int NJ, int TILE>
static cudaError_t launch_arm_t(const DevModel& hm, const EvalArgs& a, int num_sms, cudaStream_t st) {
  chain::ArmParams<NJ> P;
  for (int i = 0; i < NJ; ++i) P.joint[i] = hm.joint[i + 1];
  for (int k = 0; k < 3; ++k) P.gravity[k] = hm.gravity[k];
  // Block size: the time of a wave of blocks grows almost linearly with the warps in it (measured on Panda: one wave of
  // 8-warp blocks 18 us, of 12-warp blocks 25 us: t(w) ~ c + k w with c / k ~ 2.5), so a batch of a few waves is fastest
  // with the block size that wastes the least of its last wave: N = 65 536 = 2 048 tiles on 148 SMs is two exact waves
  // of 7-warp blocks (2 x 148 x 7 = 2 072), against 1.73 waves of 8-warp blocks.
  const long long ntiles = (a.n + 31) / 32;
  int wpb = 8;
  {
    double best = 0;
    for (int w = 8; w >= 4; --w) {
      const long long waves = ((ntiles + w - 1) / w + num_sms - 1) / num_sms;
      const double cost = (double)waves * (2.5 + w);
      if (w == 8 || cost < best - 1e-9) { best = cost; wpb = w; }
    }
  }
  const int block = wpb * 32;
  const size_t smem = (size_t)block * RBD_ARM_LEVEL * (NJ - 1) * sizeof(double) + 16;  // + the tensor-memory base word
  static thread_local bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(eval_arm_kernel<NJ, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(eval_arm_kernel<NJ, TILE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  // one block per SM; its warps stride over the tiles (a second wave of blocks would pay block launch, tensor-memory
  // allocation and pipeline fill again)
  long long grid = (ntiles + wpb - 1) / wpb;
  if (RBD_ARM_PERSISTENT && grid > num_sms) grid = num_sms;
  note_kernel(reinterpret_cast<const void*>(eval_arm_kernel<NJ, TILE>));
  eval_arm_kernel<NJ, TILE><<<(unsigned)grid, block, smem, st>>>(P, a);
  return cudaGetLastError();
}

// true when `hm` is a chain this file has an instantiation for and the call asks for what it computes (M and tau; oMi
// and J optional); *err receives the launch status
bool launch_eval_arm(const DevModel& hm, const EvalArgs& a, int num_sms, cudaStream_t st, cudaError_t* err) {
  if (hm.eval_variant != RBD_EVAL_CHAIN || !a.M || !a.tau || !a.q || !a.v || !a.a || a.tau_acc || !a.gravity_on) return false;
  if (a.n <= 0) { *err = cudaSuccess; return true; }
  // large batches (persistent regime of launch_eval: every resident warp gets >= 8 tiles) run the generic chain sweep in
  // 12-warp blocks: measured N = 2^20 Panda 0.320 ms against 0.338 ms here (8 warps per SM at 242 registers)
  if ((a.n + 255) / 256 > 8LL * num_sms) return false;
  {  // ... and so do the batches that fit ONE wave of 12-warp blocks but not one of 8-warp blocks (N = 49 152: 24.8 vs 26.9 us)
    const long long tiles = (a.n + 31) / 32;
    if (tiles > 8LL * num_sms && tiles <= 12LL * num_sms) return false;
  }
  const int nj = hm.njoints - 1;
#define RBD_ARM_CASE(N)                                                                    \
  case N:                                                                                  \
    *err = a.tile == 32 ? launch_arm_t<N, 32>(hm, a, num_sms, st) : launch_arm_t<N, 0>(hm, a, num_sms, st); \
    return true;
  switch (nj) {
    RBD_ARM_CASE(6)
    RBD_ARM_CASE(7)
    default: return false;
  }
#undef RBD_ARM_CASE
}

}  // namespace rbd
