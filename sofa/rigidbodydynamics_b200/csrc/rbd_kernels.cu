// rbd_kernels.cu — sm_100a kernels of the batched articulated-body back end.
//
// Design (DESIGN.md has the numbers):
//  * one thread = one robot instance; a warp's 32 instances sit side by side in every tiled-SoA field, so each
//    global load/store instruction moves 32 consecutive doubles (two full 128-byte lines);
//  * the joint table is a __grid_constant__ parameter -> constant bank, broadcast to the warp;
//  * per-instance sweep state (depth-indexed stack + branch slots, rbd_sweeps.h) lives in shared memory laid
//    out [slot][thread] -> bank-conflict free, no __syncthreads anywhere: threads never exchange data;
//  * FP64 CUDA-core arithmetic only: the path is a tree recursion of 3x3/6x6 products, not a dense
//    contraction, so tensor cores (tcgen05) do not apply (BASELINE north-star);
//  * the block size is chosen at run time so that resident warps per SM fill the 227 KB of shared memory.
#include <cuda_runtime.h>

#include "rbd_kernels.h"
#include "rbd_sweeps.h"

namespace rbd {

extern __shared__ __align__(16) double rbd_smem[];

// --------------------------------------------------------------------------------------------------------
// K7 / K5 / K6: fused FK + joint Jacobian (+ CRBA) (+ RNEA)
// The joint table is staged ONCE PER BLOCK from global memory into shared memory (coalesced copy), and the sweep
// reads it from there: every lane of a warp reads the same record at the same time, which shared memory serves as
// a single broadcast wavefront at a fixed ~30-cycle latency.  (Round-1 profile r1b: the same table read through
// the constant bank with run-time indices cost ~40 stall samples per dependent ISETP — the 8.7 KB table does not
// stay in the small constant L1 while three warps walk it at different phases.)
//
// Tensor memory: when the plan says so (m.plan.tmem_cols > 0) warp 0 allocates the block's columns, every thread
// derives its private slice (lane quarter of its warp, column range of its warp group), and warp 0 frees them
// after the whole block is done.  Lanes past the end of the batch recompute the last instance instead of
// exiting: tcgen05.ld/st are warp-collective (.sync.aligned) and the duplicate stores write identical values.
#define RBD_KERNEL_BOUNDS_0(MINB) __launch_bounds__((MINB == 1 && DO_TAU && !DO_M) ? 384 : 256, MINB)
#define RBD_KERNEL_BOUNDS_1(MINB) __maxnreg__(RBD_CHAIN_REGS)
#define RBD_KERNEL_BOUNDS_2(MINB) __maxnreg__(RBD_FFROOT_REGS)
#ifndef RBD_FFROOT_REGS
#define RBD_FFROOT_REGS 168
#endif
#define RBD_KERNEL_BOUNDS(MAXT, MINB) RBD_KERNEL_BOUNDS_##MAXT(MINB)
#ifndef RBD_CHAIN_REGS
#define RBD_CHAIN_REGS 168
#endif
#define RBD_DEFINE_EVAL_KERNEL(NAME, NSQ, REAL, MINB, MAXT) \
template <bool DO_M, bool DO_TAU, int TILE>                                                                                                             \
__global__ void RBD_KERNEL_BOUNDS(MAXT, MINB) NAME(const DevModelT<REAL>* __restrict__ gm, const EvalArgsT<REAL> a,                                            \
                                                   const int model_words) {                                                                             \
 /* the last two staged words are spare: the tensor-memory base address lands there (no static shared memory, so */                                     \
 /* the block can ask for the full 227 KB dynamically) */                                                                                               \
  unsigned* const tm_base_p = reinterpret_cast<unsigned*>(rbd_smem + model_words - 2);                                                                  \
  {                                                                                                                                                     \
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);                                                                    \
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);                                                                          \
    for (int w = threadIdx.x; w < model_words - 2; w += blockDim.x) dst[w] = src[w];                                                                    \
  }                                                                                                                                                     \
  __syncthreads();                                                                                                                                      \
  const DevModelT<REAL>& m = *reinterpret_cast<const DevModelT<REAL>*>(rbd_smem);                                                                       \
  const int warp = threadIdx.x >> 5;                                                                                                                    \
  const int tm_cols = m.plan.tmem_cols;                                                                                                                 \
  NSQ::TmStack T;                                                                                                                                       \
  T.base = 0;                                                                                                                                           \
  if (tm_cols > 0) {                                                                                                                                    \
    if (warp == 0) {                                                                                                                                    \
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(                                                           \
                       (unsigned long long)__cvta_generic_to_shared(tm_base_p)),                                                                        \
                   "r"(tm_cols)                                                                                                                         \
                   : "memory");                                                                                                                         \
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");                                                          \
    }                                                                                                                                                   \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                                                                    \
    __syncthreads();                                                                                                                                    \
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");                                                                                     \
    T.base = *tm_base_p + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * (unsigned)(sizeof(REAL) / 4) * (unsigned)m.plan.tmem_doubles;  \
  }                                                                                                                                                     \
 /* persistent: the block's warps stride over the batch's 32-instance tiles independently (no barrier inside), */                                       \
 /* so the staging, the tensor-memory allocation and the end-of-block wait are paid once per SM, not per tile */                                        \
  const NSQ::Stack S = NSQ::make_stack(reinterpret_cast<REAL*>(rbd_smem + model_words), threadIdx.x, m.plan.smem_doubles);                              \
  const long long ntiles = (a.n + 31) / 32;                                                                                                             \
  const int wpb = blockDim.x >> 5;                                                                                                                      \
 /* asynchronous input ring: per warp [in_slots * 3][32] elements behind the stacks (tile-32 batches only) */                                           \
  NSQ::InRing inr;                                                                                                                                      \
  if (TILE != 0 && m.plan.in_slots > 0)                                                                                                                 \
    inr.p = reinterpret_cast<REAL*>(rbd_smem + model_words) + (size_t)blockDim.x * m.plan.smem_doubles +                                                \
            (size_t)warp * (3 * m.plan.in_slots * 32) + (threadIdx.x & 31);                                                                             \
  if (inr.p) inr.ps = (unsigned)__cvta_generic_to_shared(inr.p);                                                                                        \
  for (long long t = (long long)blockIdx.x * wpb + warp; t < ntiles; t += (long long)gridDim.x * wpb) {                                                 \
    long long inst = t * 32 + (threadIdx.x & 31);                                                                                                       \
    NSQ::EvalIOT<TILE> io;                                                                                                                              \
    io.tau_live = inst < a.n ? 1 : 0;                                                                                                                   \
    if (inst >= a.n) inst = a.n - 1;                                                                                                                    \
    io.q = NSQ::make_field_t<TILE>(const_cast<REAL*>(a.q), inst, m.nq, a.tile);                                                                         \
    io.v = NSQ::make_field_t<TILE>(const_cast<REAL*>(a.v), inst, m.nv, a.tile);                                                                         \
    io.a = NSQ::make_field_t<TILE>(const_cast<REAL*>(a.a), inst, m.nv, a.tile);                                                                         \
    io.oMi = NSQ::make_field_t<TILE>(a.oMi, inst, 12 * (m.njoints - 1), a.tile);                                                                        \
    io.J = NSQ::make_field_t<TILE>(a.J, inst, 6 * m.nv, a.tile);                                                                                        \
    io.M = NSQ::make_field_t<TILE>(a.M, inst, m.mfields, a.tile);                                                                                       \
    io.tau = NSQ::make_field_t<TILE>(a.tau, inst, m.nv, a.tile);                                                                                        \
    io.Mtile = (TILE == 32 && a.M && (reinterpret_cast<unsigned long long>(a.M) & 15) == 0)                                                             \
                   ? a.M + t * (long long)m.mfields * 32 : nullptr;                                                                                     \
    io.tau_scale = a.tau_scale; io.tau_acc = a.tau_acc; io.gravity_on = a.gravity_on;                                                                  \
    io.next = (TILE != 0 && (t + (long long)gridDim.x * wpb + 1) * 32 <= a.n) ? (long long)gridDim.x * wpb : 0; /* full tiles only */                   \
    NSQ::eval_instance<DO_M, DO_TAU, TILE>(m, io, S, T, inr);                                                                                           \
  }                                                                                                                                                     \
  if (tm_cols > 0) {                                                                                                                                    \
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                                                                                    \
    __syncthreads();                                                                                                                                    \
    if (warp == 0)                                                                                                                                      \
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_base_p), "r"(tm_cols) : "memory");                                 \
  }                                                                                                                                                     \
}

RBD_DEFINE_EVAL_KERNEL(eval_kernel, rbd, double, 1, 0)
// the same for models with RX / RY / RZ / free-flyer joints only (rbd::simple: rarer joint types not compiled in)
RBD_DEFINE_EVAL_KERNEL(eval_kernel_simple, rbd::simple, double, 1, 0)
// FP32: 128 registers per thread so that two 8-warp blocks share an SM (the planner counts on it)
RBD_DEFINE_EVAL_KERNEL(eval_kernel_f32, rbd::f32, float, 2, 0)
// fixed-base serial chains (rbd::chain): compiled for blocks of up to RBD_CHAIN_THREADS threads
#ifndef RBD_CHAIN_THREADS
#define RBD_CHAIN_THREADS 448
#endif
RBD_DEFINE_EVAL_KERNEL(eval_kernel_chain, rbd::chain, double, 1, 1)
RBD_DEFINE_EVAL_KERNEL(eval_kernel_ffroot, rbd::ffroot, double, 1, 2)
RBD_DEFINE_EVAL_KERNEL(eval_kernel_ffroot8, rbd::ffroot8, double, 1, 0)

// The mapping kernels stage the joint table the same way (one coalesced copy per block, broadcast reads after).
// The cached configuration of the last apply (workspace) is always tile-32.
__device__ __forceinline__ const DevModel& stage_model(const DevModel* __restrict__ gm, int model_words) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);
  for (int w = threadIdx.x; w < model_words; w += blockDim.x) dst[w] = src[w];
  __syncthreads();
  return *reinterpret_cast<const DevModel*>(rbd_smem);
}

// The per-output tables of the vector mapping kernels go to shared memory too: read through the (small, with the
// maximum shared-memory carve-out) L1 they were evicted by the streaming wrench / pose traffic and every dependent
// sorted_k -> wrench address chain paid an L2 round trip (r1e: applyJT at 19 % of the HBM roofline).
// Words (8 bytes) needed behind the staged joint table for n_out outputs:
// (table_words, rbd_dev_model.h)
__device__ __forceinline__ DevTables stage_tables(const DevTables& g, int n_out, double* dst) {
  double* pl = dst;
  int* k = reinterpret_cast<int*>(dst + 12 * n_out);
  int* id = k + (n_out + 1) / 2 * 2;
  for (int w = threadIdx.x; w < 12 * n_out; w += blockDim.x) pl[w] = g.sorted_pl[w];
  for (int w = threadIdx.x; w < n_out; w += blockDim.x) { k[w] = g.sorted_k[w]; id[w] = g.sorted_ident[w]; }
  __syncthreads();
  DevTables t = g;
  t.sorted_pl = pl; t.sorted_k = k; t.sorted_ident = id;
  return t;
}

// K1: apply
template <int TILE>
__global__ void __launch_bounds__(256) apply_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                    const DevTables tb, const MapArgs a, const int slots) {
  const DevModel& m = stage_model(gm, model_words);
  const DevTables tbs = stage_tables(tb, m.n_out, rbd_smem + model_words);
  const int tw = table_words(m.n_out);
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= a.n) return;
  const int off = a.root ? 7 : 0;
  SplitQT<FieldT<TILE>> q{make_field_t<TILE>(const_cast<double*>(a.root), inst, 7, a.tile),
                          make_field_t<TILE>(const_cast<double*>(a.joints), inst, m.nq - off, a.tile), off};
  Stack S = make_stack(rbd_smem + model_words + tw, threadIdx.x, slots);
  MapRing ring;
  ring.p = rbd_smem + model_words + tw + (size_t)blockDim.x * slots + (size_t)(threadIdx.x >> 5) * (RBD_MAP_SLOTS * 32) +
           (threadIdx.x & 31);
  ring.ps = (unsigned)__cvta_generic_to_shared(ring.p);
  apply_instance(m, tbs, q, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32),
                 make_field_t<TILE>(a.out, inst, 7 * m.n_out, a.tile), S, ring);
}

// K2: applyJ
template <int TILE>
__global__ void __launch_bounds__(256) applyJ_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                     const DevTables tb, const MapArgs a, const int slots) {
  const DevModel& m = stage_model(gm, model_words);
  const DevTables tbs = stage_tables(tb, m.n_out, rbd_smem + model_words);
  const int tw = table_words(m.n_out);
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= a.n) return;
  const int off = a.root ? 6 : 0;
  SplitQT<FieldT<TILE>> dq{make_field_t<TILE>(const_cast<double*>(a.root), inst, 6, a.tile),
                           make_field_t<TILE>(const_cast<double*>(a.joints), inst, m.nv - off, a.tile), off};
  Stack S = make_stack(rbd_smem + model_words + tw, threadIdx.x, slots);
  MapRing ring;
  ring.p = rbd_smem + model_words + tw + (size_t)blockDim.x * slots +
           (size_t)(threadIdx.x >> 5) * (2 * RBD_MAP_SLOTS * 32) + (threadIdx.x & 31);
  ring.ps = (unsigned)__cvta_generic_to_shared(ring.p);
  applyJ_instance(m, tbs, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), dq,
                  make_field_t<TILE>(a.out, inst, 6 * m.n_out, a.tile), S, ring);
}

// K3: applyJT (vector).  a.joints / a.root are the ACCUMULATED outputs here, a.in the wrenches.
template <int TILE>
__global__ void __launch_bounds__(256) applyJT_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                      const DevTables tb, const MapArgs a, const int slots) {
  const DevModel& m = stage_model(gm, model_words);
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= a.n) return;
  const int off = a.root ? 6 : 0;
  DenseWrenchSrcT<FieldT<TILE>> src{make_field_t<TILE>(const_cast<double*>(a.in), inst, 6 * m.n_out, a.tile), &tb};
  SplitAccumSinkT<FieldT<TILE>> sink{make_field_t<TILE>(const_cast<double*>(a.root), inst, 6, a.tile),
                                     make_field_t<TILE>(const_cast<double*>(a.joints), inst, m.nv - off, a.tile), off};
  Stack S = make_stack(rbd_smem + model_words, threadIdx.x, slots);
  applyJT_instance(m, tb, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), src, sink, S);
}

// K3 on the storage plan (shared + tensor memory, persistent warps): same prologue / epilogue as eval_kernel.
#ifndef RBD_JT_MINB
#define RBD_JT_MINB 1
#endif
template <int TILE>
__global__ void __launch_bounds__(256, RBD_JT_MINB) applyJT_plan_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                              const DevTables tb, const MapArgs a) {
  unsigned* const tm_base_p = reinterpret_cast<unsigned*>(rbd_smem + model_words - 2);
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);
    for (int w = threadIdx.x; w < model_words - 2; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const DevModel& m = *reinterpret_cast<const DevModel*>(rbd_smem);
  const int warp = threadIdx.x >> 5;
  const int tm_cols = m.plan.tmem_cols;
  TmStack T;
  T.base = 0;
  if (tm_cols > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                       (unsigned long long)__cvta_generic_to_shared(tm_base_p)),
                   "r"(tm_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    T.base = *tm_base_p + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 2u * (unsigned)m.plan.tmem_doubles;
  }
  const DevTables tbs = stage_tables(tb, m.n_out, rbd_smem + model_words);
  const Stack S = make_stack(rbd_smem + model_words + table_words(m.n_out), threadIdx.x, m.plan.smem_doubles);
  const long long ntiles = (a.n + 31) / 32;
  const int wpb = blockDim.x >> 5;
  const int off = a.root ? 6 : 0;
  for (long long t = (long long)blockIdx.x * wpb + warp; t < ntiles; t += (long long)gridDim.x * wpb) {
    long long inst = t * 32 + (threadIdx.x & 31);
    // lanes past the end recompute the last instance but must not accumulate twice: their sink is disabled
    const bool live = inst < a.n;
    if (!live) inst = a.n - 1;
    DenseWrenchSrcT<FieldT<TILE>> src{make_field_t<TILE>(const_cast<double*>(a.in), inst, 6 * m.n_out, a.tile), &tbs};
    SplitAccumSinkT<FieldT<TILE>> sink{
        make_field_t<TILE>(live ? const_cast<double*>(a.root) : nullptr, inst, 6, a.tile),
        make_field_t<TILE>(live ? const_cast<double*>(a.joints) : nullptr, inst, m.nv - off, a.tile), off};
    applyJT_instance_plan(m, tbs, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), src, sink, S, T);
  }
  if (tm_cols > 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_base_p), "r"(tm_cols) : "memory");
  }
}

// K3, ring variant: full tile-32 batches.  The wrench of mapped output k of a warp's 32 instances is one contiguous
// 1 536-byte block of the input; lane 0 of every warp streams those blocks with cp.async.bulk into the warp's ring of
// `ns` slots in shared memory, always as far ahead of the sweep as the ring allows and across tile boundaries, each
// piece (<= gmax outputs of one joint) completing on its own mbarrier.  The sweep (applyJT_instance_ring) waits for a
// piece, reads it with conflict-free shared-memory loads and hands the slots back.  Nothing of the 5.6 KB/instance
// wrench stream passes through a register before it is used, and no load of the stream can stall the sweep for
// longer than the copy engine is behind.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RBD_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RBD_DONE;\n"
      "bra RBD_WAIT;\n"
      "RBD_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// one lane of the (converged) warp: the compiler knows a single thread runs the guarded code, so operands of the
// copy-engine instructions move to uniform registers without a per-value waterfall loop (the `lane == 0` test of the
// first ring kernel cost ~25 instructions per copy for exactly that loop)
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct WarpRing {
  JtProg prog;
  const double* ring;  // this lane's column of stage 0
  const double* cur;   // ... of the acquired stage
  const char* in;      // wrench array (bytes)
  long long tile_bytes, ntiles, stride;
  unsigned ring_s, bar_s;
  int nst, np, lane;
  int c_piece, stage;
  unsigned phase;
  long long p_tile;
  int p_piece;

  // fetch the producer cursor's piece into stage `st` (warp-uniform bookkeeping, lane 0 talks to the copy engine)
  __device__ __forceinline__ void issue(int st) {
    if (p_tile < ntiles) {
      if (elect_one()) {
        const int4 pe = *reinterpret_cast<const int4*>(prog.piece(p_piece));
        const unsigned bar = bar_s + 8u * (unsigned)st;
        mbar_expect_tx(bar, (unsigned)pe.y);
        const char* src = in + p_tile * tile_bytes;
        const unsigned dst = ring_s + (unsigned)st * (unsigned)JtRing::kStageBytes;
        const int2* cp = reinterpret_cast<const int2*>(prog.copy(pe.z));
#pragma unroll 1
        for (int c = 0; c < pe.w; ++c) {
          const int2 ce = cp[c];
          bulk_g2s(dst + ((unsigned)ce.y & 0xffffu), src + ce.x, (unsigned)ce.y >> 16, bar);
        }
      }
      if (++p_piece == np) { p_piece = 0; p_tile += stride; }
    }
  }
  __device__ __forceinline__ int acquire() {
    const int cnt = prog.piece(c_piece)[0] >> 16;
    c_piece = c_piece + 1 == np ? 0 : c_piece + 1;
    cur = ring + stage * (JtRing::kStageBytes / 8);
    mbar_wait(bar_s + 8u * (unsigned)stage, phase);
    return cnt;
  }
  template <int J>
  __device__ __forceinline__ void ld(double* w) const {
#pragma unroll
    for (int e = 0; e < 6; ++e) w[e] = cur[J * (JtRing::kSlotBytes / 8) + e * 32];
  }
  // ... with the slot index at run time (rolled wrench loops)
  __device__ __forceinline__ void ldr(int j, double* w) const {
    const double* c = cur + j * (JtRing::kSlotBytes / 8);
#pragma unroll
    for (int e = 0; e < 6; ++e) w[e] = c[e * 32];
  }
  // the piece has been consumed: its stage receives the piece `nst` ahead
  __device__ __forceinline__ void release(int) {
    __syncwarp();  // every lane has read the stage before lane 0 lets the copy engine overwrite it
    issue(stage);
    if (++stage == nst) { stage = 0; phase ^= 1u; }
  }
};

// W: warps per block (JtRing::kWarps, or kWarpsSmall = 12 when the state leaves room: <= 170 registers);
// MODE: 0 = depth-stack sweep, 1 = forward-projection sweep (applyJT_instance_fwd), 2 = applyDJT (applyDJT_instance_fwd)
template <int W, int MODE>
__global__ void __launch_bounds__(W * 32, 1) applyJT_ring_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                                 const MapArgs a, const int nst) {
  unsigned* const tm_base_p = reinterpret_cast<unsigned*>(rbd_smem + model_words - 2);
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);
    for (int w = threadIdx.x; w < model_words - 2; w += blockDim.x) dst[w] = src[w];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(rbd_smem + model_words);
  if (lane < nst) mbar_init(smem_u32(bars + warp * JtRing::kMaxStages + lane), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const DevModel& m = *reinterpret_cast<const DevModel*>(rbd_smem);
  double* const stacks = reinterpret_cast<double*>(bars + W * JtRing::kMaxStages);
  double* const rings = stacks + (size_t)blockDim.x * m.plan.smem_doubles;
  const int tm_cols = m.plan.tmem_cols;
  TmStack T;
  T.base = 0;
  if (tm_cols > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"(
                       (unsigned long long)__cvta_generic_to_shared(tm_base_p)),
                   "r"(tm_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    T.base = *tm_base_p + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 2u * (unsigned)m.plan.tmem_doubles;
  }
  const JtProg prog{col_prog(m)};
  const int np = prog.npieces();
  const long long ntiles = a.n / 32;  // full tiles only (the host sends a partial last tile down the register path)
  const int wpb = blockDim.x >> 5;
  if (np > 0) {
    const Stack S = make_stack(stacks, threadIdx.x, m.plan.smem_doubles);
    const int off = a.root ? 6 : 0;
    WarpRing ring;
    ring.prog = prog;
    ring.ring = rings + (size_t)warp * nst * (JtRing::kStageBytes / 8) + lane;
    ring.cur = ring.ring;
    ring.in = reinterpret_cast<const char*>(a.in);
    ring.tile_bytes = (long long)6 * m.n_out * 256;
    ring.ntiles = ntiles; ring.stride = (long long)gridDim.x * wpb;
    ring.ring_s = smem_u32(rings + (size_t)warp * nst * (JtRing::kStageBytes / 8));
    ring.bar_s = smem_u32(bars + warp * JtRing::kMaxStages);
    ring.nst = nst; ring.np = np; ring.lane = lane;
    ring.c_piece = 0; ring.stage = 0; ring.phase = 0u;
    ring.p_tile = (long long)blockIdx.x * wpb + warp; ring.p_piece = 0;
    for (int st = 0; st < nst; ++st) ring.issue(st);
    const double* const opl = prog.opl();
    for (long long t = (long long)blockIdx.x * wpb + warp; t < ntiles; t += (long long)gridDim.x * wpb) {
      const long long inst = t * 32 + lane;
      SplitRedSinkT<FieldT<32>> sink{make_field_t<32>(const_cast<double*>(a.root), inst, 6, 32),
                                     make_field_t<32>(const_cast<double*>(a.joints), inst, m.nv - off, 32), off};
      if constexpr (MODE == 2) {
        const int doff = a.dq_root ? 6 : 0;
        SplitQT<FieldT<32>> dq{make_field_t<32>(const_cast<double*>(a.dq_root), inst, 6, 32),
                               make_field_t<32>(const_cast<double*>(a.dq_joints), inst, m.nv - doff, 32), doff};
        applyDJT_instance_fwd(m, opl, prog.anc(), make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), dq, ring, sink,
                              S, T, a.kfac);
      } else if constexpr (MODE == 1)
        applyJT_instance_fwd(m, opl, prog.anc(), make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), ring, sink, S, T);
      else
        applyJT_instance_ring(m, opl, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), ring, sink, S, T);
    }
  }
  if (tm_cols > 0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_base_p), "r"(tm_cols) : "memory");
  }
}

// The ring kernel for fixed-base serial chains of NJ joints (chain::applyJT_arm_instance, rbd_arm_body.h): same blob, ring
// program and shared-memory layout as applyJT_ring_kernel<planW, 0> (planW = the block size the plan was made for), but
// the sweep is unrolled over the joints, reads their constants from the kernel-parameter constant bank and keeps its whole
// state in registers: no tensor memory, no level records.
template <int NJ>
__global__ void __launch_bounds__(256, 1) applyJT_arm_kernel(const __grid_constant__ chain::ArmParams<NJ> P,
                                                             const DevModel* __restrict__ gm, const int model_words,
                                                             const MapArgs a, const int nst, const int planW) {
  {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gm);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(rbd_smem);
    for (int w = threadIdx.x; w < model_words - 2; w += blockDim.x) dst[w] = src[w];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* const bars = reinterpret_cast<unsigned long long*>(rbd_smem + model_words);
  if (lane < nst) mbar_init(smem_u32(bars + warp * JtRing::kMaxStages + lane), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const DevModel& m = *reinterpret_cast<const DevModel*>(rbd_smem);
  double* const rings = reinterpret_cast<double*>(bars + planW * JtRing::kMaxStages);
  const JtProg prog{col_prog(m)};
  const int np = prog.npieces();
  const long long ntiles = a.n / 32;
  const int wpb = blockDim.x >> 5;
  if (np <= 0) return;
  WarpRing ring;
  ring.prog = prog;
  ring.ring = rings + (size_t)warp * nst * (JtRing::kStageBytes / 8) + lane;
  ring.cur = ring.ring;
  ring.in = reinterpret_cast<const char*>(a.in);
  ring.tile_bytes = (long long)6 * m.n_out * 256;
  ring.ntiles = ntiles; ring.stride = (long long)gridDim.x * wpb;
  ring.ring_s = smem_u32(rings + (size_t)warp * nst * (JtRing::kStageBytes / 8));
  ring.bar_s = smem_u32(bars + warp * JtRing::kMaxStages);
  ring.nst = nst; ring.np = np; ring.lane = lane;
  ring.c_piece = 0; ring.stage = 0; ring.phase = 0u;
  ring.p_tile = (long long)blockIdx.x * wpb + warp; ring.p_piece = 0;
  for (int st = 0; st < nst; ++st) ring.issue(st);
  const double* const opl = prog.opl();
  for (long long t = (long long)blockIdx.x * wpb + warp; t < ntiles; t += (long long)gridDim.x * wpb) {
    const long long inst = t * 32 + lane;
    SplitRedSinkT<FieldT<32>> sink{make_field_t<32>(nullptr, inst, 6, 32), make_field_t<32>(const_cast<double*>(a.joints), inst, NJ, 32), 0};
    chain::applyJT_arm_instance<NJ>(P, opl, make_field_t<32>(a.qcache, a.cache_inst0 + inst, NJ, 32), ring, sink);
  }
}

// applyDJT (geometric stiffness, optional extension): one thread per instance, shared-memory stack
template <int TILE>
__global__ void __launch_bounds__(256) applyDJT_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                       const DevTables tb, const DjtArgs a, const int slots) {
  const DevModel& m = stage_model(gm, model_words);
  const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= a.n) return;
  const int off = a.dq_root ? 6 : 0;
  SplitQT<FieldT<TILE>> dq{make_field_t<TILE>(const_cast<double*>(a.dq_root), inst, 6, a.tile),
                           make_field_t<TILE>(const_cast<double*>(a.dq_joints), inst, m.nv - off, a.tile), off};
  const int ooff = a.out_root ? 6 : 0;
  SplitRedSinkT<FieldT<TILE>> sink{make_field_t<TILE>(a.out_root, inst, 6, a.tile),
                                   make_field_t<TILE>(a.out_joints, inst, m.nv - ooff, a.tile), ooff};
  Stack S = make_stack(rbd_smem + model_words, threadIdx.x, slots);
  applyDJT_instance(m, tb, make_field_t<32>(a.qcache, a.cache_inst0 + inst, m.nq, 32), dq,
                    make_field_t<TILE>(const_cast<double*>(a.in), inst, 6 * m.n_out, a.tile), sink, S, a.kfac);
}

// K4: applyJT (constraint rows): one thread per row
__global__ void __launch_bounds__(256) applyJT_rows_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                           const DevTables tb, const RowArgs a, const int slots) {
  const DevModel& m = stage_model(gm, model_words);
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.nrows) return;
  const int b = a.row_ptr[r], e = a.row_ptr[r + 1];
  double n2 = 0.0;
  RowWrenchSrc src{a.col + b, a.val + 6 * (long long)b, e - b, &tb};
  RowSink sink{a.out_rows + r * m.nv, &n2};
  Stack S = make_stack(rbd_smem + model_words, threadIdx.x, slots);
  applyJT_instance(m, tb, make_field_t<32>(a.qcache, a.row_instance[r], m.nq, 32), src, sink, S);
  a.keep[r] = n2 > 1e-12 ? 1 : 0;  // kMinTorqueSqrd, reference KinematicChainMapping.inl:43-45,587,605
}

// instance-major record of the Jacobian cache: N consecutive doubles from an even offset with 16-byte loads (a lane's
// record is contiguous; 6 nv and 12 (njoints-1) are even, so every column / placement starts 16-byte aligned)
struct AosRec {
  const double* p;
  template <int N>
  __device__ __forceinline__ void ldn(int k, double* v) const {
    const double2* q2 = reinterpret_cast<const double2*>(p + k);
#pragma unroll
    for (int e = 0; e < N / 2; ++e) { const double2 t = __ldg(q2 + e); v[2 * e] = t.x; v[2 * e + 1] = t.y; }
  }
};

// K4 on cached Jacobians (rows_cached_row): one thread per row; a warp's 32 rows are accumulated in a padded
// shared-memory tile [nv][33] (conflict-free both ways) and written out with consecutive addresses — the dense rows
// are row-major in the caller's array, which one thread per row can only write with a stride of nv doubles.
__global__ void __launch_bounds__(256) applyJT_rows_cached_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                                  const DevTables tb, const RowArgs a,
                                                                  const double* __restrict__ jc,
                                                                  const double* __restrict__ oc) {
  const DevModel& m = stage_model(gm, model_words);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = m.nv;
  double* const st = rbd_smem + model_words + (size_t)warp * nv * 33;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < a.nrows;
  for (int c = 0; c < nv; ++c) st[c * 33 + lane] = 0.0;
  if (live) {
    const int b = a.row_ptr[r], e = a.row_ptr[r + 1];
    const long long inst = a.row_instance[r];
    // the cache is instance-major: a lane of this kernel reads whole Jacobian columns / placements of ITS instance
    // (48 / 96 contiguous bytes), not one field of 32 neighbouring instances
    const AosRec Jc{jc + inst * (long long)(6 * nv)};
    const AosRec Oc{oc + inst * (long long)(12 * (m.njoints - 1))};
    rows_cached_row(m, tb, Jc, Oc, a.col + b, a.val + 6 * (long long)b, e - b,
                    [&](int c, double t) { st[c * 33 + lane] += t; });
    double n2 = 0.0;
    for (int c = 0; c < nv; ++c) { const double t = st[c * 33 + lane]; n2 += t * t; }
    a.keep[r] = n2 > 1e-12 ? 1 : 0;  // kMinTorqueSqrd, reference KinematicChainMapping.inl:43-45,587,605
  }
  __syncwarp();
  const long long row0 = (long long)blockIdx.x * blockDim.x + warp * 32;
  for (int idx = lane; idx < 32 * nv; idx += 32) {
    const int rw = idx / nv, c = idx - rw * nv;
    if (row0 + rw < a.nrows) a.out_rows[(row0 + rw) * nv + c] = st[c * 33 + rw];
  }
}

// The same rows with one WARP per row (A/B variant, off: see RBD_ROWS_WARP): lane c owns dof c (and c + 32, c + 64), so the Jacobian columns on the
// path of a non-zero's output are 32 consecutive 48-byte records of ONE instance — a coalesced 1.5 KB read with 32
// loads in flight per warp — instead of one lane walking to the root through the joint table four columns at a time
// (the thread-per-row kernel above: 24 % of the HBM roofline, 10 long-scoreboard stall cycles per issue, r1i / r2).
// Which dofs lie on the path of a joint is a bit mask per joint, built once per block from the staged table.  The wrench
// shift and the per-dof accumulation over the non-zeros run in the same order as rows_cached_row (same dot6): the
// dense row is bit-identical; the squared norm of the 1e-12 filter is summed in dof order through shuffles.
__global__ void __launch_bounds__(256) applyJT_rows_warp_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                                const DevTables tb, const RowArgs a,
                                                                const double* __restrict__ jc,
                                                                const double* __restrict__ oc) {
  const DevModel& m = stage_model(gm, model_words);
  unsigned* const mask = reinterpret_cast<unsigned*>(rbd_smem + model_words);  // [njoints][3]: dofs 0..95 on the path to the root
  const int nj = m.njoints, nv = m.nv;
  for (int j = threadIdx.x; j < nj; j += blockDim.x) {
    unsigned b0 = 0, b1 = 0, b2 = 0;
    for (int k = j; k != 0; k = m.joint[k].parent)
      for (int e = 0; e < m.joint[k].nv; ++e) {
        const int c = m.joint[k].idx_v + e;
        if (c < 32) b0 |= 1u << c; else if (c < 64) b1 |= 1u << (c - 32); else b2 |= 1u << (c - 64);
      }
    mask[3 * j] = b0; mask[3 * j + 1] = b1; mask[3 * j + 2] = b2;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < a.nrows; r += (long long)gridDim.x * wpb) {
    const int b = a.row_ptr[r], e = a.row_ptr[r + 1];
    const long long inst = a.row_instance[r];
    const double* const Ji = jc + inst * (long long)(6 * nv);
    const double* const Oi = oc + inst * (long long)(12 * (nj - 1));
    double acc[3] = {0.0, 0.0, 0.0};
    for (int j = b; j < e; ++j) {
      const int k = a.col[j];
      const int pj = tb.outk_parent[k];
      if (pj == 0) continue;  // attached to the universe: empty support, zero Jacobian
      const double* pl = tb.outk_pl + 12 * k;
      const double* w = a.val + 6 * (long long)j;
      double Rp[12];
      AosRec{Oi}.ldn<12>(12 * (pj - 1), Rp);
      const double px = Rp[0] * pl[9] + Rp[1] * pl[10] + Rp[2] * pl[11] + Rp[9];
      const double py = Rp[3] * pl[9] + Rp[4] * pl[10] + Rp[5] * pl[11] + Rp[10];
      const double pz = Rp[6] * pl[9] + Rp[7] * pl[10] + Rp[8] * pl[11] + Rp[11];
      double F[6], tx, ty, tz;
      RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);
      F[0] = w[0]; F[1] = w[1]; F[2] = w[2];
      F[3] = w[3] + tx; F[4] = w[4] + ty; F[5] = w[5] + tz;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int c = lane + 32 * t;
        if (c < nv && ((mask[3 * pj + t] >> lane) & 1u)) {
          double A[6];
          AosRec{Ji}.ldn<6>(6 * c, A);
          acc[t] += dot6(A, F);
        }
      }
    }
    double* const row = a.out_rows + r * nv;
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (lane + 32 * t < nv) row[lane + 32 * t] = acc[t];
    double n2 = 0.0;  // in dof order, as one thread would sum it
    for (int c = 0; c < nv; ++c) {
      const double v = __shfl_sync(0xffffffffu, c < 32 ? acc[0] : c < 64 ? acc[1] : acc[2], c & 31);
      n2 += v * v;
    }
    if (lane == 0) a.keep[r] = n2 > 1e-12 ? 1 : 0;  // kMinTorqueSqrd, reference KinematicChainMapping.inl:43-45,587,605
  }
}

// getFrameJacobian(LWA) of (instance, output) pairs from the Jacobian cache: one thread per pair writes the columns on
// the path to the root (the launcher zeroed the array: off-path columns stay 0, as pinocchio leaves them)
__global__ void __launch_bounds__(256) frame_jacobian_kernel(const DevModel* __restrict__ gm, const int model_words,
                                                             const DevTables tb, const long long npairs,
                                                             const int* __restrict__ instance,
                                                             const int* __restrict__ out_index,
                                                             double* __restrict__ J, const double* __restrict__ jc,
                                                             const double* __restrict__ oc) {
  const DevModel& m = stage_model(gm, model_words);
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= npairs) return;
  const long long inst = instance[r];
  const AosRec Jc{jc + inst * (long long)(6 * m.nv)};
  const AosRec Oc{oc + inst * (long long)(12 * (m.njoints - 1))};
  double* const Jr = J + r * (long long)(6 * m.nv);
  frame_jacobian_cached(m, tb, Jc, Oc, out_index[r], [&](int c, const double* o) {
    double2* q2 = reinterpret_cast<double2*>(Jr + 6 * c);  // 6 nv even, cudaMalloc'ed or 16-byte aligned by contract
    q2[0] = make_double2(o[0], o[1]); q2[1] = make_double2(o[2], o[3]); q2[2] = make_double2(o[4], o[5]);
  });
}

// Forward dynamics (aba_instance): persistent warps over the batch's 32-instance tiles, one scratch record per warp
struct GlobalScratch {
  double* p;  // this lane's column of its warp's [E][32] record
  __device__ __forceinline__ double ld(int e) const { return p[e * 32]; }
  __device__ __forceinline__ void st(int e, double v) const { p[e * 32] = v; }
};
template <int TILE>
__global__ void __launch_bounds__(256, 1) aba_kernel(const DevModel* __restrict__ gm, const int model_words, const AbaArgs a) {
  const DevModel& m = stage_model(gm, model_words);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const long long gw = (long long)blockIdx.x * wpb + warp;
  const GlobalScratch X{a.scratch + gw * (long long)aba_scratch_doubles(m) * 32 + lane};
  const long long ntiles = (a.n + 31) / 32;
  for (long long t = gw; t < ntiles; t += (long long)gridDim.x * wpb) {
    const long long inst = t * 32 + lane;
    if (inst >= a.n) continue;  // (no warp-collective operation in this sweep)
    aba_instance(m, make_field_t<TILE>(const_cast<double*>(a.q), inst, m.nq, a.tile),
                 make_field_t<TILE>(const_cast<double*>(a.v), inst, m.nv, a.tile),
                 make_field_t<TILE>(const_cast<double*>(a.tau), inst, m.nv, a.tile),
                 make_field_t<TILE>(a.ddq, inst, m.nv, a.tile), X);
  }
}

// Assembled mapping Jacobian: one thread per (instance, selected output) writes that block row's non-zero columns
__global__ void __launch_bounds__(256) getJ_kernel(const DevModel* __restrict__ gm, const int model_words, const DevTables tb,
                                                   const long long n, const int n_sel, const int* __restrict__ sel,
                                                   const int* __restrict__ blk_ptr, double* __restrict__ val,
                                                   const double* __restrict__ jc, const double* __restrict__ oc) {
  const DevModel& m = stage_model(gm, model_words);
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n * n_sel) return;
  const long long inst = r / n_sel;
  const int s = (int)(r - inst * n_sel);
  const AosRec Jc{jc + inst * (long long)(6 * m.nv)};
  const AosRec Oc{oc + inst * (long long)(12 * (m.njoints - 1))};
  getJ_block_cached(m, tb, Jc, Oc, sel[s], val + inst * (long long)(6 * blk_ptr[n_sel]) + 6 * (long long)blk_ptr[s]);
}

// K8: layout adapters.  32 instances x 32 fields per block through a padded shared tile, so both the
// instance-major side and the tiled-SoA side are accessed with consecutive addresses.
__global__ void aos_to_soa_kernel(long long n, int K, long long tile, const double* __restrict__ aos,
                                  double* __restrict__ soa) {
  __shared__ double t[32][33];
  const long long i0 = (long long)blockIdx.x * 32;  // instance blocks on x (no 65535 limit)
  const int k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long i = i0 + r;
    const int k = k0 + threadIdx.x;
    if (i < n && k < K) t[r][threadIdx.x] = aos[i * K + k];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r;
    const long long i = i0 + threadIdx.x;
    if (i < n && k < K) soa[(i / tile) * (long long)K * tile + (long long)k * tile + (i % tile)] = t[threadIdx.x][r];
  }
}
__global__ void soa_to_aos_kernel(long long n, int K, long long tile, const double* __restrict__ soa,
                                  double* __restrict__ aos) {
  __shared__ double t[32][33];
  const long long i0 = (long long)blockIdx.x * 32;  // instance blocks on x (no 65535 limit)
  const int k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r;
    const long long i = i0 + threadIdx.x;
    if (i < n && k < K) t[threadIdx.x][r] = soa[(i / tile) * (long long)K * tile + (long long)k * tile + (i % tile)];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long i = i0 + r;
    const int k = k0 + threadIdx.x;
    if (i < n && k < K) aos[i * K + k] = t[r][threadIdx.x];
  }
}

// tile-32 fast path of the adapters: one block moves the [KC fields] x [32 instances] slab of one instance tile through
// a padded shared tile — 16 loads in flight per thread before the barrier (the 32 x 32 kernels above keep 4 and pay a
// block launch per 8 KB: K = 741 ran at 0.39 of the HBM roofline, profiles/r1j_launches_condensed.txt).  In the tiled
// layout the slab is contiguous (field k of the tile at k * 32), on the instance-major side every instance owns KC
// consecutive doubles.
#define RBD_TR_KC 128
template <bool TO_AOS>
__global__ void __launch_bounds__(256) transpose32_kernel(long long n, int K, const double* __restrict__ src, double* __restrict__ dst) {
  __shared__ double t[RBD_TR_KC][33];
  // 1-D grid, field chunk fastest: the blocks that are resident together then cover whole tiles — 32 K contiguous
  // doubles on either side — instead of one 1 KB piece of many tiles (DRAM page locality: K = 741 ran at 0.75 of the
  // roofline with the chunk index on grid.y, K = 128, where a block's piece IS the whole tile, at 0.93)
  const int nchunks = (K + RBD_TR_KC - 1) / RBD_TR_KC;
  const long long tile = blockIdx.x / nchunks;
  const int k0 = (int)(blockIdx.x % nchunks) * RBD_TR_KC;
  const int kc = min(RBD_TR_KC, K - k0);
  const long long i0 = tile * 32;
  const int ni = (int)min(32LL, n - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (TO_AOS) {
    const double* s = src + tile * (long long)K * 32 + (long long)k0 * 32;  // [kc][32] contiguous
#pragma unroll
    for (int j = 0; j < RBD_TR_KC * 32 / 256; ++j) {
      const int idx = tid + 256 * j, k = idx >> 5;
      if (k < kc) t[k][idx & 31] = s[idx];
    }
    __syncthreads();
    double* d = dst + i0 * K + k0;
    for (int r = warp; r < ni; r += 8)
#pragma unroll
      for (int j = 0; j < RBD_TR_KC / 32; ++j) {
        const int k = lane + 32 * j;
        if (k < kc) d[(long long)r * K + k] = t[k][r];
      }
  } else {
    const double* s = src + i0 * K + k0;
    for (int r = warp; r < ni; r += 8)
#pragma unroll
      for (int j = 0; j < RBD_TR_KC / 32; ++j) {
        const int k = lane + 32 * j;
        if (k < kc) t[k][r] = s[(long long)r * K + k];
      }
    __syncthreads();
    double* d = dst + tile * (long long)K * 32 + (long long)k0 * 32;
#pragma unroll
    for (int j = 0; j < RBD_TR_KC * 32 / 256; ++j) {
      const int idx = tid + 256 * j, k = idx >> 5;
      if (k < kc && (idx & 31) < ni) d[idx] = t[k][idx & 31];
    }
  }
}

// The same slab with 16-byte accesses on all four legs (global load, shared store, shared load, global store): the
// kernel above issues four memory instructions per 8 bytes and stalls on the memory-instruction queue (mio / lg throttle
// 21 + 5 stall cycles per issue, profiles/r2_soa_to_aos_ncu_full.txt).  Every thread moves 2 x 2 blocks — fields (k, k+1)
// of instances (i, i+1): on the tiled side a block is two 16-byte pieces of two field rows, in the shared tile (kept
// instance-major, row stride RBD_TRV_S doubles) two 16-byte pieces of two instance rows.  On the instance-major side
// row r of the slab starts at element (i0 + r) K + k0, whose parity alternates from row to row when K is odd: row r of the
// shared tile is shifted by that parity, so that a pair which is 16-byte aligned in global memory is 16-byte aligned in
// shared memory too (a lone head / tail element per row moves as a scalar, and the 2 x 2 blocks of the shifted rows use
// two 8-byte shared accesses instead of one 16-byte access).  Lane mapping inside a warp: instance pair ip = (lane & 3) +
// 4 (lane >> 3), field pair bit (lane >> 2) & 1 — a quarter warp's eight 16-byte shared accesses then hit eight different
// 16-byte bank groups.  Needs 16-byte aligned base pointers (the launcher checks).
#define RBD_TRV_KC 128
#define RBD_TRV_S 130
#ifndef RBD_TR_VEC
#define RBD_TR_VEC 1 /* A/B knob: 0 = the 8-byte kernel for every launch */
#endif
#define RBD_TRV_MINK 96 /* narrower arrays leave most lanes of the row phase idle: the 8-byte kernel is faster there (K = 38) */
static_assert(RBD_TRV_KC == RBD_TR_KC, "both tile-32 adapters share the launch grid");
template <bool TO_AOS>
__global__ void __launch_bounds__(256) transpose32v_kernel(long long n, int K, const double* __restrict__ src, double* __restrict__ dst) {
  __shared__ __align__(16) double t[32 * RBD_TRV_S];
  const int nchunks = (K + RBD_TRV_KC - 1) / RBD_TRV_KC;  // 1-D grid, field chunk fastest (see transpose32_kernel)
  const long long tile = blockIdx.x / nchunks;
  const int k0 = (int)(blockIdx.x % nchunks) * RBD_TRV_KC;
  const int kc = min(RBD_TRV_KC, K - k0);
  const long long i0 = tile * 32;
  const int ni = (int)min(32LL, n - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int par_even = k0 & 1, par_odd = (K + k0) & 1;  // parity of (i0 + r) K + k0 for even / odd r (i0 K is even)
  const int ip = (lane & 3) + 4 * (lane >> 3);           // instance pair 0 .. 15
  const int kpbit = (lane >> 2) & 1;
  const int i = 2 * ip;
  double* const te = t + i * RBD_TRV_S + par_even;       // row i (even), element k at te[k]
  double* const to = t + (i + 1) * RBD_TRV_S + par_odd;  // row i + 1
  if (TO_AOS) {
    const double* s = src + tile * (long long)K * 32 + (long long)k0 * 32 + i;  // + k * 32: fields k of instances i, i + 1
    double2 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 2 * (2 * warp + kpbit + 16 * j);
      a[j] = make_double2(0.0, 0.0); b[j] = a[j];
      if (k < kc) a[j] = *reinterpret_cast<const double2*>(s + (long long)k * 32);
      if (k + 1 < kc) b[j] = *reinterpret_cast<const double2*>(s + (long long)(k + 1) * 32);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 2 * (2 * warp + kpbit + 16 * j);
      if (k < kc) {
        if (par_even == 0) *reinterpret_cast<double2*>(te + k) = make_double2(a[j].x, b[j].x);
        else { te[k] = a[j].x; te[k + 1] = b[j].x; }
        if (par_odd == 0) *reinterpret_cast<double2*>(to + k) = make_double2(a[j].y, b[j].y);
        else { to[k] = a[j].y; to[k + 1] = b[j].y; }
      }
    }
    __syncthreads();
    for (int r = warp; r < ni; r += 8) {
      const int par = (r & 1) ? par_odd : par_even;
      const double* tr = t + r * RBD_TRV_S + par;
      double* gd = dst + (i0 + r) * K + k0;
      if (par && lane == 0) gd[0] = tr[0];
      const int npairs = (kc - par) >> 1;
      for (int m = lane; m < npairs; m += 32)
        *reinterpret_cast<double2*>(gd + par + 2 * m) = *reinterpret_cast<const double2*>(tr + par + 2 * m);
      if (((kc - par) & 1) && lane == 31) gd[kc - 1] = tr[kc - 1];
    }
  } else {
    {
      // all of this thread's loads first (4 rows x 2 pairs + head / tail scalars), then the shared stores
      double2 v[4][2];
      double hd[4], tl[4];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = warp + 8 * rr;
        const int par = (r & 1) ? par_odd : par_even;
        const double* gs = src + (i0 + r) * K + k0;
        const int npairs = (kc - par) >> 1;
        hd[rr] = tl[rr] = 0.0;
#pragma unroll
        for (int mm = 0; mm < 2; ++mm) {
          const int m = lane + 32 * mm;
          v[rr][mm] = make_double2(0.0, 0.0);
          if (r < ni && m < npairs) v[rr][mm] = *reinterpret_cast<const double2*>(gs + par + 2 * m);
        }
        if (r < ni && par && lane == 0) hd[rr] = gs[0];
        if (r < ni && ((kc - par) & 1) && lane == 31) tl[rr] = gs[kc - 1];
      }
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = warp + 8 * rr;
        const int par = (r & 1) ? par_odd : par_even;
        double* tr = t + r * RBD_TRV_S + par;
        const int npairs = (kc - par) >> 1;
#pragma unroll
        for (int mm = 0; mm < 2; ++mm) {
          const int m = lane + 32 * mm;
          if (r < ni && m < npairs) *reinterpret_cast<double2*>(tr + par + 2 * m) = v[rr][mm];
        }
        if (r < ni && par && lane == 0) tr[0] = hd[rr];
        if (r < ni && ((kc - par) & 1) && lane == 31) tr[kc - 1] = tl[rr];
      }
    }
    __syncthreads();
    double* d = dst + tile * (long long)K * 32 + (long long)k0 * 32 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 2 * (2 * warp + kpbit + 16 * j);
      if (k < kc && i < ni) {
        double2 x, y;
        if (par_even == 0) x = *reinterpret_cast<const double2*>(te + k);
        else { x.x = te[k]; x.y = te[k + 1]; }
        if (i + 1 < ni) {
          if (par_odd == 0) y = *reinterpret_cast<const double2*>(to + k);
          else { y.x = to[k]; y.y = to[k + 1]; }
          *reinterpret_cast<double2*>(d + (long long)k * 32) = make_double2(x.x, y.x);
          if (k + 1 < kc) *reinterpret_cast<double2*>(d + (long long)(k + 1) * 32) = make_double2(x.y, y.y);
        } else {  // the batch ends on an even instance: the padding lane of the last tile is left alone
          d[(long long)k * 32] = x.x;
          if (k + 1 < kc) d[(long long)(k + 1) * 32] = x.y;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------------
// launch helpers
static const int kMaxSmemOptin = 227 * 1024;

// The __global__ function this thread launched last (rbd_last_kernel_name: bench.py reads the name of the timed
// instantiation back from the launch instead of deriving it from the model).
static thread_local const void* g_last_kernel = nullptr;
#define RBD_NOTE(K) (g_last_kernel = reinterpret_cast<const void*>(K))
void note_kernel(const void* fn) { g_last_kernel = fn; }
const char* last_kernel_mangled() {
  const char* name = nullptr;
  if (!g_last_kernel || cudaFuncGetName(&name, g_last_kernel) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return name;
}

template <class Kern>
static cudaError_t pick_block(Kern kern, int smem_doubles_per_thread, int forced, LaunchCfg* cfg,
                              size_t block_fixed_bytes = 0) {
  // opt in to the full 227 KB of dynamic shared memory once per kernel
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  const size_t per_thread = (size_t)smem_doubles_per_thread * sizeof(double);
  int best_threads = 0, best_block = 0, best_blocks_per_sm = 0;
  const int cand[] = {32, 64, 96, 128, 160, 192, 224, 256};
  for (int c : cand) {
    if (forced > 0 && c != forced) continue;
    const size_t smem = block_fixed_bytes + per_thread * c;
    if (smem > (size_t)kMaxSmemOptin) continue;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, c, smem);
    if (e != cudaSuccess) return e;
    // prefer more resident threads; on ties the LARGER block (its warps spread over the four SM sub-partitions)
    if (nb * c >= best_threads) {
      best_threads = nb * c; best_block = c; best_blocks_per_sm = nb;
    }
  }
  if (best_block == 0) return cudaErrorInvalidConfiguration;
  cfg->block = best_block;
  cfg->smem_bytes = block_fixed_bytes + per_thread * best_block;
  cfg->blocks_per_sm = best_blocks_per_sm;
  return cudaSuccess;
}

#define RBD_GRID(n, block) (unsigned)(((n) + (block)-1) / (block))

// variant index: bit 2 = compile-time tile of 32, bit 1 = CRBA, bit 0 = RNEA
#define RBD_EVAL_DISPATCH(KERN, v, CALL)             \
  switch (v) {                                       \
    case 7: CALL((KERN<true, true, 32>)); break;     \
    case 6: CALL((KERN<true, false, 32>)); break;    \
    case 5: CALL((KERN<false, true, 32>)); break;    \
    case 4: CALL((KERN<false, false, 32>)); break;   \
    case 3: CALL((KERN<true, true, 0>)); break;      \
    case 2: CALL((KERN<true, false, 0>)); break;     \
    case 1: CALL((KERN<false, true, 0>)); break;     \
    default: CALL((KERN<false, false, 0>)); break;   \
  }

// `hm` is the host copy of the planned model (hm.plan chosen by plan_eval), `d_model` its staged blob on the device.
#define RBD_DEFINE_EVAL_LAUNCH(FN, KERN, REAL, TAG)                                                                   \
  cudaError_t FN(const DevModelT<REAL>& hm, const DevModelT<REAL>* d_model, const EvalArgsT<REAL>& a, LaunchCfg* cache, \
                 int num_sms, cudaStream_t st) {                                                                   \
    const bool do_m = a.M != nullptr, do_tau = a.tau != nullptr;                                                   \
    const int v = (a.tile == 32 ? 4 : 0) | (do_m ? 2 : 0) | (do_tau ? 1 : 0);                                      \
    LaunchCfg& cfg = cache[v];                                                                                     \
    const int model_words = (int)(eval_model_bytes(hm) / 8); /* staged joint table + 2 spare words */              \
    int warps = hm.plan.warps;                                                                                     \
    if (warps > 8 && a.n > 0) {                                                                                    \
      /* 12-warp plans (chain / ffroot kernels; the storage layout holds for any smaller block): a batch that needs */ \
      /* as many waves of 12-warp blocks as of 8-warp blocks runs the latter — shorter waves (Panda, N = 65 536) */ \
      const long long tiles = (a.n + 31) / 32;                                                                     \
      const long long ww = ((tiles + warps - 1) / warps + num_sms - 1) / num_sms;                                  \
      const long long w8 = ((tiles + 7) / 8 + num_sms - 1) / num_sms;                                              \
      if (ww >= w8) warps = 8;                                                                                     \
    }                                                                                                              \
    const int block = warps * 32;                                                                                  \
    const size_t smem = (size_t)model_words * 8 + (size_t)block * (hm.plan.smem_doubles + 3 * hm.plan.in_slots) * sizeof(REAL); \
    cudaError_t e = cudaSuccess;                                                                                   \
    if (cfg.block != block || cfg.smem_bytes != smem || cfg.tag != TAG) {                                          \
      RBD_EVAL_DISPATCH(KERN, v, RBD_ATTR)                                                                         \
      if (e != cudaSuccess) return e;                                                                              \
      cfg.block = block; cfg.smem_bytes = smem; cfg.blocks_per_sm = hm.plan.blocks_per_sm; cfg.tag = TAG;          \
    }                                                                                                              \
    if (a.n <= 0) return cudaSuccess;                                                                              \
    unsigned grid = RBD_GRID(a.n, block);                                                                          \
    const unsigned resident = (unsigned)(num_sms * (hm.plan.blocks_per_sm > 0 ? hm.plan.blocks_per_sm : 1));       \
    /* persistent blocks (one wave, warps stride over the tiles) once every warp gets >= 8 tiles; smaller */      \
    /* batches run one block per `block` instances so that the last wave is not half empty */                     \
    if (grid > 8 * resident) grid = resident;                                                                      \
    RBD_EVAL_DISPATCH(KERN, v, RBD_LAUNCH)                                                                         \
    return cudaGetLastError();                                                                                     \
  }
#define RBD_ATTR(K)                                                                                 \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);         \
  if (e == cudaSuccess)                                                                             \
    e = cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
#define RBD_LAUNCH(K) RBD_NOTE(K); K<<<grid, block, smem, st>>>(d_model, a, model_words)
RBD_DEFINE_EVAL_LAUNCH(launch_eval, eval_kernel, double, RBD_EVAL_GENERIC)
RBD_DEFINE_EVAL_LAUNCH(launch_eval_simple, eval_kernel_simple, double, RBD_EVAL_SIMPLE)
RBD_DEFINE_EVAL_LAUNCH(launch_eval_f32, eval_kernel_f32, float, 100)
RBD_DEFINE_EVAL_LAUNCH(launch_eval_chain, eval_kernel_chain, double, RBD_EVAL_CHAIN)
RBD_DEFINE_EVAL_LAUNCH(launch_eval_ffroot, eval_kernel_ffroot, double, RBD_EVAL_FFROOT)
RBD_DEFINE_EVAL_LAUNCH(launch_eval_ffroot8, eval_kernel_ffroot8, double, RBD_EVAL_FFROOT8)
#undef RBD_ATTR
#undef RBD_LAUNCH
cudaError_t launch_eval_any(const DevModel& hm, const DevModel* d_model, const EvalArgs& a, LaunchCfg* cache,
                            int num_sms, cudaStream_t st) {
  switch (hm.plan.kvariant) {
    case RBD_EVAL_FFROOT: return launch_eval_ffroot(hm, d_model, a, cache, num_sms, st);
    case RBD_EVAL_FFROOT8: return launch_eval_ffroot8(hm, d_model, a, cache, num_sms, st);
    case RBD_EVAL_CHAIN: return launch_eval_chain(hm, d_model, a, cache, num_sms, st);
    case RBD_EVAL_SIMPLE: return launch_eval_simple(hm, d_model, a, cache, num_sms, st);
    default: return launch_eval(hm, d_model, a, cache, num_sms, st);
  }
}

// mapping kernels: cfg[0] = run-time tile, cfg[1] = tile 32
static inline int map_model_words(const DevModel& m) { return (int)((dev_model_bytes(m) + 15) / 16) * 2; }

#define RBD_MAP_LAUNCH(KERN, n_items, ring_doubles) /* ring_doubles: input ring behind the stacks, per instance */ \
  const int mw = map_model_words(m);                                                                          \
  const size_t fixed = ((size_t)mw + table_words(m.n_out)) * 8;                                               \
  const int v = a_tile == 32 ? 1 : 0;                                                                         \
  LaunchCfg& c = cfg[v];                                                                                      \
  if (c.block == 0) {                                                                                         \
    cudaError_t e = v ? pick_block(KERN<32>, smem_doubles + (ring_doubles), 0, &c, fixed)                     \
                      : pick_block(KERN<0>, smem_doubles + (ring_doubles), 0, &c, fixed);                     \
    if (e != cudaSuccess) return e;                                                                           \
  }                                                                                                           \
  if ((n_items) <= 0) return cudaSuccess;                                                                     \
  if (v) { RBD_NOTE(KERN<32>); KERN<32><<<RBD_GRID(n_items, c.block), c.block, c.smem_bytes, st>>>(d_model, mw, tb, a, smem_doubles); } \
  else { RBD_NOTE(KERN<0>); KERN<0><<<RBD_GRID(n_items, c.block), c.block, c.smem_bytes, st>>>(d_model, mw, tb, a, smem_doubles); } \
  return cudaGetLastError()

cudaError_t launch_apply(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                         int smem_doubles, LaunchCfg* cfg, cudaStream_t st) {
  const long long a_tile = a.tile;
  RBD_MAP_LAUNCH(apply_kernel, a.n, RBD_MAP_SLOTS);
}
cudaError_t launch_applyJ(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                          int smem_doubles, LaunchCfg* cfg, cudaStream_t st) {
  const long long a_tile = a.tile;
  RBD_MAP_LAUNCH(applyJ_kernel, a.n, 2 * RBD_MAP_SLOTS);
}
cudaError_t launch_applyJT(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                           int smem_doubles, LaunchCfg* cfg, cudaStream_t st) {
  const long long a_tile = a.tile;
  RBD_MAP_LAUNCH(applyJT_kernel, a.n, 0);
}
// `hm` = host copy of the model planned with plan_eval(..., jt = true), `d_blob` = its staged blob on the device
cudaError_t launch_applyJT_plan(const DevModel& hm, const DevModel* d_blob, const DevTables& tb, const MapArgs& a,
                                LaunchCfg* cfg2, int num_sms, cudaStream_t st) {
  const int v = a.tile == 32 ? 1 : 0;
  LaunchCfg& c = cfg2[v];
  const int model_words = (int)(eval_model_bytes(hm) / 8);
  const int block = hm.plan.warps * 32;
  const size_t smem = ((size_t)model_words + table_words(hm.n_out)) * 8 + (size_t)block * hm.plan.smem_doubles * 8;
  if (c.block != block || c.smem_bytes != smem) {
    cudaError_t e = v ? cudaFuncSetAttribute(applyJT_plan_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin)
                      : cudaFuncSetAttribute(applyJT_plan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);
    if (e != cudaSuccess) return e;
    e = v ? cudaFuncSetAttribute(applyJT_plan_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)
          : cudaFuncSetAttribute(applyJT_plan_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    c.block = block; c.smem_bytes = smem; c.blocks_per_sm = hm.plan.blocks_per_sm;
  }
  if (a.n <= 0) return cudaSuccess;
  unsigned grid = RBD_GRID(a.n, block);
  const unsigned resident = (unsigned)(num_sms * (hm.plan.blocks_per_sm > 0 ? hm.plan.blocks_per_sm : 1));
  if (grid > 8 * resident) grid = resident;  // persistent only when every warp gets >= 8 tiles (else: wave tail)
  if (v) { RBD_NOTE(applyJT_plan_kernel<32>); applyJT_plan_kernel<32><<<grid, block, smem, st>>>(d_blob, model_words, tb, a); }
  else { RBD_NOTE(applyJT_plan_kernel<0>); applyJT_plan_kernel<0><<<grid, block, smem, st>>>(d_blob, model_words, tb, a); }
  return cudaGetLastError();
}

// `hm` = host copy of the model planned with plan_jt_ring, `d_blob` = its staged blob (build_jt_blob), nst = ring
// stages per warp.  a.n must be a multiple of 32 and a.tile == 32; a.in 16-byte aligned.
template <class K>
static cudaError_t jt_ring_attr(K kern) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
cudaError_t launch_applyJT_ring(const DevModel& hm, const DevModel* d_blob, const MapArgs& a, int nst, LaunchCfg* cfg,
                                int num_sms, cudaStream_t st, bool use_arm) {
  const int model_words = (int)(eval_model_bytes(hm) / 8);
  const int W = hm.plan.warps;
  const int block = W * 32;
  // 0: stack / 8 warps, 1: stack / 12, 2: forward / 12, 3: applyDJT / 8
  const int variant = hm.plan.jt_fwd == 2 ? 3 : hm.plan.jt_fwd == 1 ? 2 : (W == JtRing::kWarpsSmall ? 1 : 0);
  const size_t smem = ((size_t)model_words + (size_t)W * JtRing::kMaxStages) * 8 + (size_t)block * hm.plan.smem_doubles * 8 +
                      (size_t)W * nst * JtRing::kStageBytes;
  if (cfg->block != block || cfg->smem_bytes != smem) {
    cudaError_t e = variant == 3 ? jt_ring_attr(applyJT_ring_kernel<JtRing::kWarps, 2>)
                  : variant == 2 ? jt_ring_attr(applyJT_ring_kernel<JtRing::kWarpsSmall, 1>)
                  : variant == 1 ? jt_ring_attr(applyJT_ring_kernel<JtRing::kWarpsSmall, 0>)
                                 : jt_ring_attr(applyJT_ring_kernel<JtRing::kWarps, 0>);
    if (e != cudaSuccess) return e;
    cfg->block = block; cfg->smem_bytes = smem; cfg->blocks_per_sm = 1;
  }
  if (a.n <= 0) return cudaSuccess;
  const long long ntiles = a.n / 32;
  // fixed-base serial chains of 6 / 7 joints (stack variants only: the ring program is the same): the unrolled sweep
  if (use_arm && hm.eval_variant == RBD_EVAL_CHAIN && hm.plan.jt_fwd == 0 && !a.root && (hm.njoints - 1 == 6 || hm.njoints - 1 == 7)) {
    const int wa = pick_warps_per_block(W < 8 ? W : 8, ntiles, num_sms);
    unsigned ga = (unsigned)((ntiles + wa - 1) / wa);
    if (ga > (unsigned)num_sms) ga = (unsigned)num_sms;
    static thread_local bool attr_done[2] = {false, false};
    if (hm.njoints - 1 == 7) {
      if (!attr_done[1]) { cudaError_t e = jt_ring_attr(applyJT_arm_kernel<7>); if (e != cudaSuccess) return e; attr_done[1] = true; }
      chain::ArmParams<7> P;
      for (int i = 0; i < 7; ++i) P.joint[i] = hm.joint[i + 1];
      for (int k = 0; k < 3; ++k) P.gravity[k] = hm.gravity[k];
      RBD_NOTE(applyJT_arm_kernel<7>);
      applyJT_arm_kernel<7><<<ga, wa * 32, smem, st>>>(P, d_blob, model_words, a, nst, W);
    } else {
      if (!attr_done[0]) { cudaError_t e = jt_ring_attr(applyJT_arm_kernel<6>); if (e != cudaSuccess) return e; attr_done[0] = true; }
      chain::ArmParams<6> P;
      for (int i = 0; i < 6; ++i) P.joint[i] = hm.joint[i + 1];
      for (int k = 0; k < 3; ++k) P.gravity[k] = hm.gravity[k];
      RBD_NOTE(applyJT_arm_kernel<6>);
      applyJT_arm_kernel<6><<<ga, wa * 32, smem, st>>>(P, d_blob, model_words, a, nst, W);
    }
    return cudaGetLastError();
  }
  const int wpb = pick_warps_per_block(W, ntiles, num_sms);  // <= W, the size the plan and the tensor-memory shares were made for
  const int lblock = wpb * 32;
  unsigned grid = (unsigned)((ntiles + wpb - 1) / wpb);
  if (grid > (unsigned)num_sms) grid = (unsigned)num_sms;
  if (variant == 3) { RBD_NOTE((applyJT_ring_kernel<JtRing::kWarps, 2>)); applyJT_ring_kernel<JtRing::kWarps, 2><<<grid, lblock, smem, st>>>(d_blob, model_words, a, nst); }
  else if (variant == 2) { RBD_NOTE((applyJT_ring_kernel<JtRing::kWarpsSmall, 1>)); applyJT_ring_kernel<JtRing::kWarpsSmall, 1><<<grid, lblock, smem, st>>>(d_blob, model_words, a, nst); }
  else if (variant == 1) { RBD_NOTE((applyJT_ring_kernel<JtRing::kWarpsSmall, 0>)); applyJT_ring_kernel<JtRing::kWarpsSmall, 0><<<grid, lblock, smem, st>>>(d_blob, model_words, a, nst); }
  else { RBD_NOTE((applyJT_ring_kernel<JtRing::kWarps, 0>)); applyJT_ring_kernel<JtRing::kWarps, 0><<<grid, lblock, smem, st>>>(d_blob, model_words, a, nst); }
  return cudaGetLastError();
}

cudaError_t launch_applyDJT(const DevModel& m, const DevModel* d_model, const DevTables& tb, const DjtArgs& a,
                            LaunchCfg* cfg2, cudaStream_t st) {
  const int mw = map_model_words(m);
  const int slots = RBD_DJT_LEVEL * m.max_depth + RBD_DJT_BRANCH * m.nbranch;
  const int v = a.tile == 32 ? 1 : 0;
  LaunchCfg& c = cfg2[v];
  if (c.block == 0) {
    cudaError_t e = v ? pick_block(applyDJT_kernel<32>, slots, 0, &c, (size_t)mw * 8)
                      : pick_block(applyDJT_kernel<0>, slots, 0, &c, (size_t)mw * 8);
    if (e != cudaSuccess) return e;
  }
  if (a.n <= 0) return cudaSuccess;
  if (v) { RBD_NOTE(applyDJT_kernel<32>); applyDJT_kernel<32><<<RBD_GRID(a.n, c.block), c.block, c.smem_bytes, st>>>(d_model, mw, tb, a, slots); }
  else { RBD_NOTE(applyDJT_kernel<0>); applyDJT_kernel<0><<<RBD_GRID(a.n, c.block), c.block, c.smem_bytes, st>>>(d_model, mw, tb, a, slots); }
  return cudaGetLastError();
}

cudaError_t launch_applyJT_rows(const DevModel& m, const DevModel* d_model, const DevTables& tb, const RowArgs& a,
                                int smem_doubles, LaunchCfg* cfg, cudaStream_t st) {
  const int mw = map_model_words(m);
  if (cfg->block == 0) {
    cudaError_t e = pick_block(applyJT_rows_kernel, smem_doubles, 0, cfg, (size_t)mw * 8);
    if (e != cudaSuccess) return e;
  }
  if (a.nrows <= 0) return cudaSuccess;
  RBD_NOTE(applyJT_rows_kernel);
  applyJT_rows_kernel<<<RBD_GRID(a.nrows, cfg->block), cfg->block, cfg->smem_bytes, st>>>(d_model, mw, tb, a, smem_doubles);
  return cudaGetLastError();
}

// rows on the Jacobian cache: `jc` / `oc` = world joint Jacobian and oMi of the applied instances, instance-major
cudaError_t launch_applyJT_rows_cached(const DevModel& m, const DevModel* d_model, const DevTables& tb, const RowArgs& a,
                                       const double* jc, const double* oc, int num_sms, cudaStream_t st) {
  const int mw = map_model_words(m);
  const int block = 256;
#ifndef RBD_ROWS_WARP
#define RBD_ROWS_WARP 0 /* 1 = one warp per row (applyJT_rows_warp_kernel): coalesced column reads, but measured 2.3x - 3.3x
                           SLOWER (A/B s3t, 2 rows x 2 non-zeros per instance: Talos 2 M rows 2.19 vs 0.97 ms, Solo 0.47 vs
                           0.15, Panda 0.147 vs 0.045): the chain row_ptr -> col -> parent -> R|p / J is three dependent
                           memory round trips per ROW, and a warp then has one row in flight where the thread-per-row
                           kernel has 32 */
#endif
  if (RBD_ROWS_WARP && m.nv <= 96) {
    if (a.nrows <= 0) return cudaSuccess;
    const size_t smem = (size_t)mw * 8 + (size_t)m.njoints * 3 * sizeof(unsigned) + 8;
    long long grid = (a.nrows + block / 32 - 1) / (block / 32);
    const long long cap = (long long)num_sms * 8 * 16;  // 8 resident blocks per SM, >= 16 rows per warp: the prologue amortised
    if (grid > cap) grid = cap;
    RBD_NOTE(applyJT_rows_warp_kernel);
    applyJT_rows_warp_kernel<<<(unsigned)grid, block, smem, st>>>(d_model, mw, tb, a, jc, oc);
    return cudaGetLastError();
  }
  const size_t smem = (size_t)mw * 8 + (size_t)(block / 32) * m.nv * 33 * 8;
  cudaError_t e = cudaFuncSetAttribute(applyJT_rows_cached_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);
  if (e != cudaSuccess) return e;
  if (a.nrows <= 0) return cudaSuccess;
  RBD_NOTE(applyJT_rows_cached_kernel);
  applyJT_rows_cached_kernel<<<RBD_GRID(a.nrows, block), block, smem, st>>>(d_model, mw, tb, a, jc, oc);
  return cudaGetLastError();
}

cudaError_t launch_frame_jacobian(const DevModel& m, const DevModel* d_model, const DevTables& tb, long long npairs,
                                  const int* instance, const int* out_index, double* J, const double* jc, const double* oc,
                                  cudaStream_t st) {
  if (npairs <= 0) return cudaSuccess;
  const int mw = map_model_words(m);
  cudaError_t e = cudaMemsetAsync(J, 0, (size_t)npairs * 6 * m.nv * sizeof(double), st);
  if (e != cudaSuccess) return e;
  RBD_NOTE(frame_jacobian_kernel);
  frame_jacobian_kernel<<<RBD_GRID(npairs, 256), 256, (size_t)mw * 8, st>>>(d_model, mw, tb, npairs, instance, out_index, J, jc, oc);
  return cudaGetLastError();
}

cudaError_t launch_aba(const DevModel& m, const DevModel* d_model, const AbaArgs& a, int num_sms, int* scratch_warps,
                       cudaStream_t st) {
  const int mw = map_model_words(m);
  const int block = 256, wpb = block / 32;
  const long long ntiles = (a.n + 31) / 32;
  long long grid = (ntiles + wpb - 1) / wpb;
  if (grid > num_sms) grid = num_sms;  // one block per SM, warps stride over the tiles
  if (grid < 1) grid = 1;
  if (scratch_warps) *scratch_warps = num_sms * wpb;
  if (!a.scratch || a.n <= 0) return cudaSuccess;
  if (a.tile == 32) { RBD_NOTE(aba_kernel<32>); aba_kernel<32><<<(unsigned)grid, block, (size_t)mw * 8, st>>>(d_model, mw, a); }
  else { RBD_NOTE(aba_kernel<0>); aba_kernel<0><<<(unsigned)grid, block, (size_t)mw * 8, st>>>(d_model, mw, a); }
  return cudaGetLastError();
}

cudaError_t launch_getJ(const DevModel& m, const DevModel* d_model, const DevTables& tb, long long n, int n_sel, const int* sel,
                        const int* blk_ptr, double* val, const double* jc, const double* oc, cudaStream_t st) {
  if (n <= 0 || n_sel <= 0) return cudaSuccess;
  const int mw = map_model_words(m);
  RBD_NOTE(getJ_kernel);
  getJ_kernel<<<RBD_GRID(n * n_sel, 256), 256, (size_t)mw * 8, st>>>(d_model, mw, tb, n, n_sel, sel, blk_ptr, val, jc, oc);
  return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(long long n, int K, long long tile, const double* aos, double* soa, cudaStream_t st) {
  if (n <= 0 || K <= 0) return cudaSuccess;
  if (tile == 32) {
    const long long nb32 = ((n + 31) / 32) * ((K + RBD_TR_KC - 1) / RBD_TR_KC);
    if (nb32 > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const unsigned g32 = (unsigned)nb32;
    if (RBD_TR_VEC && K >= RBD_TRV_MINK && ((reinterpret_cast<unsigned long long>(aos) | reinterpret_cast<unsigned long long>(soa)) & 15) == 0)
      transpose32v_kernel<false><<<g32, 256, 0, st>>>(n, K, aos, soa);
    else
      transpose32_kernel<false><<<g32, 256, 0, st>>>(n, K, aos, soa);
    return cudaGetLastError();
  }
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((K + 31) / 32)), block(32, 8);
  aos_to_soa_kernel<<<grid, block, 0, st>>>(n, K, tile, aos, soa);
  return cudaGetLastError();
}
cudaError_t launch_soa_to_aos(long long n, int K, long long tile, const double* soa, double* aos, cudaStream_t st) {
  if (n <= 0 || K <= 0) return cudaSuccess;
  if (tile == 32) {
    const long long nb32 = ((n + 31) / 32) * ((K + RBD_TR_KC - 1) / RBD_TR_KC);
    if (nb32 > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const unsigned g32 = (unsigned)nb32;
    if (RBD_TR_VEC && K >= RBD_TRV_MINK && ((reinterpret_cast<unsigned long long>(aos) | reinterpret_cast<unsigned long long>(soa)) & 15) == 0)
      transpose32v_kernel<true><<<g32, 256, 0, st>>>(n, K, soa, aos);
    else
      transpose32_kernel<true><<<g32, 256, 0, st>>>(n, K, soa, aos);
    return cudaGetLastError();
  }
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((K + 31) / 32)), block(32, 8);
  soa_to_aos_kernel<<<grid, block, 0, st>>>(n, K, tile, soa, aos);
  return cudaGetLastError();
}

}  // namespace rbd
