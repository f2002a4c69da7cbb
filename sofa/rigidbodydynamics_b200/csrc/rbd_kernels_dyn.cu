// rbd_kernels_dyn.cu — forward-dynamics kernels on the depth-first storage plan (aba2_instance, rbd_dyn_body.h).
//
// Same launch shape as the fused eval: one persistent block per SM, its warps stride over the batch's 32-instance
// tiles; the joint table is staged once per block into shared memory; level records of the sweep live in shared memory
// and — through tcgen05.ld / tcgen05.st, one private column range per thread — in tensor memory (plan_aba2).
#include <cuda_runtime.h>

#include "rbd_kernels.h"
#include "rbd_sweeps.h"

namespace rbd {

extern __shared__ __align__(16) double rbd_smem[];

namespace {
struct GlobalRec {
  double* p;  // this lane's column of its warp's [E][32] record
  __device__ __forceinline__ double ld(int e) const { return p[e * 32]; }
  __device__ __forceinline__ void st(int e, double v) const { p[e * 32] = v; }
  __device__ __forceinline__ GlobalRec at(int e) const { return GlobalRec{p + e * 32}; }  // base of a joint's record
};
// ... the same with the L2 eviction priority "evict_last" on every access (cold level records of drnea_instance: written
// once, read back tens of microseconds later under a 3 TB/s store stream that would otherwise push them out to HBM), and a
// += that does not read back (red.global.add.f64: nothing waits on it)
struct KeepRec {
  double* p;
  unsigned long long pol;
  __device__ __forceinline__ double ld(int e) const {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p + e * 32), "l"(pol));
    return v;
  }
  __device__ __forceinline__ void st(int e, double v) const {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p + e * 32), "d"(v), "l"(pol) : "memory");
  }
  __device__ __forceinline__ void add(int e, double v) const {
    asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p + e * 32), "d"(v), "l"(pol) : "memory");
  }
  __device__ __forceinline__ const double* ptr(int e) const { return p + e * 32; }
  __device__ __forceinline__ KeepRec at(int e) const { return KeepRec{p + e * 32, pol}; }
};
}  // namespace

// Block prologue shared by the kernels of this file: stage the joint table (its last two words are spare: the
// tensor-memory base address lands there), allocate the block's tensor-memory columns, hand every thread its slices.
struct DynCtx {
  const DevModel* m;
  Stack S;
  TmStack T;
  unsigned* tm_base_p;
  int tm_cols;
};
__device__ __forceinline__ DynCtx dyn_prologue(const DevModel* __restrict__ gm, const int model_words) {
  DynCtx c;
  c.tm_base_p = reinterpret_cast<unsigned*>(rbd_smem + model_words - 2);
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);
    for (int w = threadIdx.x; w < model_words - 2; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  c.m = reinterpret_cast<const DevModel*>(rbd_smem);
  const int warp = threadIdx.x >> 5;
  c.tm_cols = c.m->plan.tmem_cols;
  c.T.base = 0;
  if (c.tm_cols > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                       (unsigned long long)__cvta_generic_to_shared(c.tm_base_p)),
                   "r"(c.tm_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.T.base = *c.tm_base_p + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 2u * (unsigned)c.m->plan.tmem_doubles;
  }
  c.S = make_stack(rbd_smem + model_words, threadIdx.x, c.m->plan.smem_doubles);
  return c;
}
__device__ __forceinline__ void dyn_epilogue(const DynCtx& c) {
  if (c.tm_cols > 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*c.tm_base_p), "r"(c.tm_cols) : "memory");
  }
}

template <int TILE>
__global__ void __launch_bounds__(256, 1) aba2_kernel(const DevModel* __restrict__ gm, const int model_words, const AbaArgs a) {
  const DynCtx c = dyn_prologue(gm, model_words);
  const DevModel& m = *c.m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpb + warp;
  const GlobalRec X{a.scratch + gw * (long long)aba2_scratch_doubles(m) * 32 + lane};
  const long long ntiles = (a.n + 31) / 32;
  for (long long t = gw; t < ntiles; t += (long long)gridDim.x * wpb) {
    long long inst = t * 32 + lane;
    if (inst >= a.n) inst = a.n - 1;  // lanes past the end recompute the last instance (warp-collective tensor-memory accesses)
    aba2_instance(m, make_field_t<TILE>(const_cast<double*>(a.q), inst, m.nq, a.tile),
                  make_field_t<TILE>(const_cast<double*>(a.v), inst, m.nv, a.tile),
                  make_field_t<TILE>(const_cast<double*>(a.tau), inst, m.nv, a.tile),
                  make_field_t<TILE>(a.ddq, inst, m.nv, a.tile), X, c.S, c.T);
  }
  dyn_epilogue(c);
}

// Inverse of the mass matrix (minv_instance): same launch shape; a.Minv = packed upper triangle, nv (nv + 1) / 2 fields
template <int TILE>
__global__ void __launch_bounds__(256, 1) minv_kernel(const DevModel* __restrict__ gm, const int model_words, const MinvArgs a) {
  const DynCtx c = dyn_prologue(gm, model_words);
  const DevModel& m = *c.m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpb + warp;
  const GlobalRec X{a.scratch + gw * (long long)(RBD_MINV_REC * (m.njoints - 1) + RBD_MINV_ROOT_REC) * 32 + lane};
  const long long ntiles = (a.n + 31) / 32;
  const int mf = m.nv * (m.nv + 1) / 2;
  for (long long t = gw; t < ntiles; t += (long long)gridDim.x * wpb) {
    long long inst = t * 32 + lane;
    const int live = inst < a.n ? 1 : 0;
    if (!live) inst = a.n - 1;
    minv_instance(m, make_field_t<TILE>(const_cast<double*>(a.q), inst, m.nq, a.tile),
                  make_field_t<TILE>(a.Minv, inst, mf, a.tile), X, c.S, c.T, live);
  }
  dyn_epilogue(c);
}

// Derivatives of inverse dynamics (drnea_instance, rbd_deriv_body.h): same launch shape; every level record is on chip,
// the only global traffic is q, v, a in and the two nv x nv matrices (+ optional packed M) out
template <int TILE, bool SPARSE, bool WITH_M>
__global__ void __launch_bounds__(256, 1) drnea_kernel(const DevModel* __restrict__ gm, const int model_words, const DrneaArgs a) {
  const DynCtx c = dyn_prologue(gm, model_words);
  const DevModel& m = *c.m;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpb + warp;
  const long long ntiles = (a.n + 31) / 32;
  const int nv = m.nv;
  unsigned long long pol;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  const KeepRec X{a.scratch + gw * (long long)m.slot_total[0] * 32 + lane, pol};
  InRing inr;  // asynchronous input ring: this thread's column behind its level records
  inr.p = c.S.p + m.plan.fbase * RBD_WARP;
  inr.ps = (unsigned)__cvta_generic_to_shared(inr.p);
  for (long long t = gw; t < ntiles; t += (long long)gridDim.x * wpb) {
    long long inst = t * 32 + lane;
    if (inst >= a.n) inst = a.n - 1;  // lanes past the end recompute the last instance (same values to the same addresses)
    const int K = SPARSE ? m.slot_total[1] : nv * nv;
    drnea_instance<SPARSE, WITH_M>(m, make_field_t<TILE>(const_cast<double*>(a.q), inst, m.nq, a.tile),
                   make_field_t<TILE>(const_cast<double*>(a.v), inst, nv, a.tile),
                   make_field_t<TILE>(const_cast<double*>(a.a), inst, nv, a.tile),
                   make_field_t<TILE>(a.dtau_dq, inst, K, a.tile), make_field_t<TILE>(a.dtau_dv, inst, K, a.tile),
                   make_field_t<TILE>(a.M, inst, nv * (nv + 1) / 2, a.tile), X, c.S, c.T, inr);
  }
  dyn_epilogue(c);
}

template <class K>
static cudaError_t dyn_attr(K kern) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_aba2(const DevModel& hm, const DevModel* d_model, const AbaArgs& a, int num_sms, int* scratch_warps,
                        cudaStream_t st) {
  const int mw = (int)(aba2_model_bytes(hm) / 8);
  const long long ntiles = (a.n + 31) / 32;
  const int W = hm.plan.warps, wpb = pick_warps_per_block(W, ntiles, num_sms), block = wpb * 32;
  long long grid = (ntiles + wpb - 1) / wpb;
  if (grid > num_sms) grid = num_sms;  // one block per SM, warps stride over the tiles
  if (grid < 1) grid = 1;
  if (scratch_warps) *scratch_warps = num_sms * W;
  if (!a.scratch || a.n <= 0) return cudaSuccess;
  const size_t smem = (size_t)mw * 8 + (size_t)W * 32 * hm.plan.smem_doubles * 8;
  static thread_local bool attr_done[2] = {false, false};
  cudaError_t e = cudaSuccess;
  if (a.tile == 32) {
    if (!attr_done[1]) { e = dyn_attr(aba2_kernel<32>); if (e != cudaSuccess) return e; attr_done[1] = true; }
    note_kernel(reinterpret_cast<const void*>(aba2_kernel<32>));
    aba2_kernel<32><<<(unsigned)grid, block, smem, st>>>(d_model, mw, a);
  } else {
    if (!attr_done[0]) { e = dyn_attr(aba2_kernel<0>); if (e != cudaSuccess) return e; attr_done[0] = true; }
    note_kernel(reinterpret_cast<const void*>(aba2_kernel<0>));
    aba2_kernel<0><<<(unsigned)grid, block, smem, st>>>(d_model, mw, a);
  }
  return cudaGetLastError();
}

cudaError_t launch_minv(const DevModel& hm, const DevModel* d_model, const MinvArgs& a, int num_sms, int* scratch_warps,
                        cudaStream_t st) {
  const int mw = (int)(aba2_model_bytes(hm) / 8);
  const long long ntiles = (a.n + 31) / 32;
  const int W = hm.plan.warps, wpb = pick_warps_per_block(W, ntiles, num_sms), block = wpb * 32;
  long long grid = (ntiles + wpb - 1) / wpb;
  if (grid > num_sms) grid = num_sms;
  if (grid < 1) grid = 1;
  if (scratch_warps) *scratch_warps = num_sms * W;
  if (!a.scratch || a.n <= 0) return cudaSuccess;
  const size_t smem = (size_t)mw * 8 + (size_t)W * 32 * hm.plan.smem_doubles * 8;
  static thread_local bool attr_done[2] = {false, false};
  cudaError_t e = cudaSuccess;
  if (a.tile == 32) {
    if (!attr_done[1]) { e = dyn_attr(minv_kernel<32>); if (e != cudaSuccess) return e; attr_done[1] = true; }
    note_kernel(reinterpret_cast<const void*>(minv_kernel<32>));
    minv_kernel<32><<<(unsigned)grid, block, smem, st>>>(d_model, mw, a);
  } else {
    if (!attr_done[0]) { e = dyn_attr(minv_kernel<0>); if (e != cudaSuccess) return e; attr_done[0] = true; }
    note_kernel(reinterpret_cast<const void*>(minv_kernel<0>));
    minv_kernel<0><<<(unsigned)grid, block, smem, st>>>(d_model, mw, a);
  }
  return cudaGetLastError();
}

}  // namespace rbd

namespace rbd {
cudaError_t launch_drnea(const DevModel& hm, const DevModel* d_model, const DrneaArgs& a, int num_sms, int* scratch_warps,
                         cudaStream_t st) {
  const int mw = (int)(aba2_model_bytes(hm) / 8);
  const long long ntiles = (a.n + 31) / 32;
  const int W = hm.plan.warps, wpb = pick_warps_per_block(W, ntiles, num_sms), block = wpb * 32;
  long long grid = (ntiles + wpb - 1) / wpb;
  if (grid > num_sms) grid = num_sms;  // one block per SM, warps stride over the tiles
  if (scratch_warps) *scratch_warps = num_sms * W;
  if (!a.scratch || a.n <= 0) return cudaSuccess;
  const size_t smem = (size_t)mw * 8 + (size_t)W * 32 * hm.plan.smem_doubles * 8;
  static thread_local bool attr_done[6] = {false, false, false, false, false, false};
  cudaError_t e = cudaSuccess;
#define RBD_DR_LAUNCH(T_, S_, M_, slot_)                                                                          \
  do {                                                                                                            \
    if (!attr_done[slot_]) { e = dyn_attr(drnea_kernel<T_, S_, M_>); if (e != cudaSuccess) return e; attr_done[slot_] = true; } \
    note_kernel(reinterpret_cast<const void*>(drnea_kernel<T_, S_, M_>));                                         \
    drnea_kernel<T_, S_, M_><<<(unsigned)grid, block, smem, st>>>(d_model, mw, a);                                \
  } while (0)
  const bool with_m = a.M != nullptr && !a.sparse;
  if (a.tile == 32) {
    if (a.sparse) RBD_DR_LAUNCH(32, true, false, 3); else if (with_m) RBD_DR_LAUNCH(32, false, true, 5); else RBD_DR_LAUNCH(32, false, false, 1);
  } else {
    if (a.sparse) RBD_DR_LAUNCH(0, true, false, 2); else if (with_m) RBD_DR_LAUNCH(0, false, true, 4); else RBD_DR_LAUNCH(0, false, false, 0);
  }
#undef RBD_DR_LAUNCH
  return cudaGetLastError();
}
}  // namespace rbd
