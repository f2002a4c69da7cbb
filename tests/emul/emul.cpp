// emul.cpp — CPU harness around the kernel bodies in sofa/rigidbodydynamics_b200/csrc/rbd_sweeps.h.
// TEST INFRASTRUCTURE ONLY: lets `pytest -m "not gpu"` exercise the exact per-thread code the sm_100a kernels
// run (same templates, same stack/branch-slot layout, same tiled addressing), with shared memory emulated by a
// heap array.  It is never loaded by the product and is not a fallback.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../sofa/rigidbodydynamics_b200/csrc/rbd_model_build.h"
#include "../../sofa/rigidbodydynamics_b200/csrc/rbd_sweeps.h"

using namespace rbd;

static std::string g_err;
extern "C" const char* emul_last_error() { return g_err.c_str(); }

struct Ctx {
  DevModel m;
  HostTables ht;
  DevTables tb;
};
static int make_ctx(const rbd_model_desc* d, Ctx* c) {
  int rc = build_dev_model(d, &c->m, &c->ht, &g_err);
  if (rc) return rc;
  c->tb.sorted_k = c->ht.sorted_k.data(); c->tb.sorted_pl = c->ht.sorted_pl.data();
  c->tb.sorted_ident = c->ht.sorted_ident.data();
  c->tb.outk_parent = c->ht.outk_parent.data(); c->tb.outk_pl = c->ht.outk_pl.data();
  return 0;
}

extern "C" void emul_sincos(long long n, const double* x, double* s, double* c) {
  for (long long i = 0; i < n; ++i) sincos_(x[i], s + i, c + i);
}

extern "C" int emul_model_info(const rbd_model_desc* d, int* out /*[8]*/) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  out[0] = c.m.max_depth; out[1] = c.m.nbranch; out[6] = smem_applyJT(c.m); out[7] = c.m.has_ff_root;
  DevModel t = c.m;
  if (!plan_eval(&t, true, true, true, 0, true)) return -100;
  out[2] = t.plan.smem_doubles; out[3] = t.plan.tmem_doubles; out[4] = t.plan.warps * t.plan.blocks_per_sm;
  t = c.m;
  if (!plan_eval(&t, true, true, false)) return -101;
  out[5] = t.plan.warps * t.plan.blocks_per_sm;
  return 0;
}

// The per-instance eval of one namespace (rbd = generic, rbd::simple, rbd::chain, rbd::ffroot: the sweep compiled with
// different joint-type knowledge, rbd_sweeps.h) over a whole batch with the storage plan `pl` of the staged blob `bm`.
#define EMUL_DEFINE_EVAL_RUN(FN, NSQ)                                                                                  \
  template <int TILE>                                                                                                  \
  static void FN##_one(const DevModel& m, long long i, long long tile, double* q, double* v, double* a, double* oMi,   \
                       double* J, double* M, double* tau, const NSQ::Stack& S, const NSQ::TmStack& T, NSQ::InRing& inr, \
                       long long next) {                                                                               \
    const bool do_m = M != nullptr, do_tau = tau != nullptr;                                                           \
    NSQ::EvalIOT<TILE> io;                                                                                             \
    io.next = next;                                                                                                    \
    io.q = NSQ::make_field_t<TILE>(q, i, m.nq, tile); io.v = NSQ::make_field_t<TILE>(v, i, m.nv, tile);                \
    io.a = NSQ::make_field_t<TILE>(a, i, m.nv, tile);                                                                  \
    io.oMi = NSQ::make_field_t<TILE>(oMi, i, 12 * (m.njoints - 1), tile);                                              \
    io.J = NSQ::make_field_t<TILE>(J, i, 6 * m.nv, tile);                                                              \
    io.M = NSQ::make_field_t<TILE>(M, i, m.mfields, tile); io.tau = NSQ::make_field_t<TILE>(tau, i, m.nv, tile);       \
    if (do_m && do_tau) NSQ::eval_instance<true, true, TILE>(m, io, S, T, inr);                                        \
    else if (do_m) NSQ::eval_instance<true, false, TILE>(m, io, S, T, inr);                                            \
    else if (do_tau) NSQ::eval_instance<false, true, TILE>(m, io, S, T, inr);                                          \
    else NSQ::eval_instance<false, false, TILE>(m, io, S, T, inr);                                                     \
  }                                                                                                                    \
  static void FN(const DevModel& bm, long long n, long long tile, double* q, double* v, double* a, double* oMi,        \
                 double* J, double* M, double* tau) {                                                                  \
    const EvalPlan& pl = bm.plan;                                                                                      \
    const int bt = pl.warps * 32;                                                                                      \
    std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300); /* poison */            \
    std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);                         \
    /* asynchronous input ring (tile-32 batches, when the plan has room): emulated per lane, persisting from one tile */ \
    /* to the next as it does for a warp that strides over the tiles (here: stride 1, a single "warp") */             \
    const bool use_ring = tile == 32 && pl.in_slots > 0;                                                               \
    std::vector<double> ringbuf((size_t)3 * kEvalInSlots * 32 * 32, 1e300);                                            \
    std::vector<NSQ::InRing> rings(32);                                                                                \
    for (int l = 0; l < 32; ++l) rings[l].p = use_ring ? ringbuf.data() + (size_t)l * 3 * kEvalInSlots * 32 : nullptr; \
    for (long long b0 = 0; b0 < n; b0 += bt)                                                                           \
      for (int t = 0; t < bt && b0 + t < n; ++t) {                                                                     \
        const long long i = b0 + t;                                                                                    \
        NSQ::Stack S = NSQ::make_stack(smem.data(), t, pl.smem_doubles);                                               \
        NSQ::TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};                         \
        const long long next = (use_ring && (i / 32 + 2) * 32 <= n) ? 1 : 0; /* the next tile is a full one */         \
        if (tile == 32) FN##_one<32>(bm, i, tile, q, v, a, oMi, J, M, tau, S, T, rings[i % 32], next);                 \
        else FN##_one<0>(bm, i, tile, q, v, a, oMi, J, M, tau, S, T, rings[0], 0);                                     \
      }                                                                                                                \
  }
EMUL_DEFINE_EVAL_RUN(run_eval_generic, rbd)
EMUL_DEFINE_EVAL_RUN(run_eval_simple, rbd::simple)
EMUL_DEFINE_EVAL_RUN(run_eval_chain, rbd::chain)
EMUL_DEFINE_EVAL_RUN(run_eval_ffroot, rbd::ffroot)
EMUL_DEFINE_EVAL_RUN(run_eval_ffroot8, rbd::ffroot8)

// tile == 32 runs the compile-time-tile instantiation the GPU fast path uses.  use_tmem selects the storage plan
// (tensor memory is emulated by a second poisoned array); block = warps * 32 pins the planner's block size (0 = auto).
extern "C" int emul_eval_ex(const rbd_model_desc* d, long long n, long long tile, int block, int use_tmem, double* q,
                            double* v, double* a, double* oMi, double* J, double* M, double* tau, int sparse_m,
                            int variant);
extern "C" int emul_eval(const rbd_model_desc* d, long long n, long long tile, int block, int use_tmem, double* q,
                         double* v, double* a, double* oMi, double* J, double* M, double* tau) {
  return emul_eval_ex(d, n, tile, block, use_tmem, q, v, a, oMi, J, M, tau, 0, -1);
}
// support-only pattern of M: col_ptr [nv + 1], row_idx [nnz]; returns nnz
extern "C" int emul_mass_pattern(const rbd_model_desc* d, int* col_ptr, int* row_idx) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  return mass_support_pattern(c.m, col_ptr, row_idx);
}
// which instantiation the planner picks for this model and output set (RBD_EVAL_*), and the warps per SM of its plan
extern "C" int emul_eval_variant(const rbd_model_desc* d, int do_m, int do_tau, int* warps_per_sm) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, do_m != 0, do_tau != 0, true, 0, true)) return -100;
  if (warps_per_sm) *warps_per_sm = c.m.plan.warps * c.m.plan.blocks_per_sm;
  return c.m.plan.kvariant;
}
// fused eval of a 6- / 7-joint serial chain on the unrolled arm sweep (rbd::chain::arm_instance); -102 when the model is
// not such a chain
template <int NJ>
static void run_arm(const DevModel& dm, long long n, long long tile, double* q, double* v, double* a, double* oMi, double* J,
                    double* M, double* tau) {
  namespace ch = rbd::chain;
  ch::ArmParams<NJ> P;
  for (int i = 0; i < NJ; ++i) P.joint[i] = dm.joint[i + 1];
  for (int k = 0; k < 3; ++k) P.gravity[k] = dm.gravity[k];
  std::vector<double> smem((size_t)RBD_ARM_LEVEL * (NJ - 1) * 32, 1e300);
  for (long long i = 0; i < n; ++i) {
    std::fill(smem.begin(), smem.end(), 1e300);
    ch::Stack S = ch::make_stack(smem.data(), (int)(i % 32), RBD_ARM_LEVEL * (NJ - 1));
    std::vector<double> tmem((size_t)6 * NJ, 1e300);
    ch::TmStack T{tmem.data()};
    ch::arm_instance<NJ, 0>(P, ch::make_field(q, i, NJ, tile), ch::make_field(v, i, NJ, tile), ch::make_field(a, i, NJ, tile),
                            ch::make_field(oMi, i, 12 * NJ, tile), ch::make_field(J, i, 6 * NJ, tile),
                            ch::make_field(M, i, NJ * (NJ + 1) / 2, tile), ch::make_field(tau, i, NJ, tile), S, T);
  }
}
extern "C" int emul_eval_arm(const rbd_model_desc* d, long long n, long long tile, double* q, double* v, double* a,
                             double* oMi, double* J, double* M, double* tau) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (c.m.eval_variant != RBD_EVAL_CHAIN) { g_err = "not a fixed-base serial chain of RX / RY / RZ joints"; return -102; }
  switch (c.m.njoints - 1) {
    case 6: run_arm<6>(c.m, n, tile, q, v, a, oMi, J, M, tau); return 0;
    case 7: run_arm<7>(c.m, n, tile, q, v, a, oMi, J, M, tau); return 0;
    default: g_err = "no arm instantiation for this joint count"; return -102;
  }
}

// the storage plan of the fused eval: {warps, blocks_per_sm, smem_doubles, tmem_doubles, tmem_cols, in_slots, kvariant,
// dynamic shared-memory bytes of a full block, ysplit, fsplit}
extern "C" int emul_eval_plan(const rbd_model_desc* d, int do_m, int do_tau, int* out /*[10]*/) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, do_m != 0, do_tau != 0, true, 0, true)) return -100;
  const EvalPlan& p = c.m.plan;
  out[0] = p.warps; out[1] = p.blocks_per_sm; out[2] = p.smem_doubles; out[3] = p.tmem_doubles; out[4] = p.tmem_cols;
  out[5] = p.in_slots; out[6] = p.kvariant;
  out[7] = (int)(eval_model_bytes(c.m) + (size_t)p.warps * 32 * (p.smem_doubles + 3 * p.in_slots) * 8);
  out[8] = p.ysplit; out[9] = p.fsplit;
  return 0;
}
// sparse_m != 0: M holds the support-only fields (rbd_eval_soa_sparse).  variant: -1 = the generic sweep (namespace
// rbd) on the planner's plan, -2 = the instantiation the planner picked (what the GPU launches), >= 0 = that
// instantiation (RBD_EVAL_*; the model must qualify) on a plan made for it.
extern "C" int emul_eval_ex(const rbd_model_desc* d, long long n, long long tile, int block, int use_tmem, double* q,
                            double* v, double* a, double* oMi, double* J, double* M, double* tau, int sparse_m,
                            int variant) {
  if ((tau != nullptr) != (M != nullptr) && (oMi || J)) {
    // as the library does (rbd_capi.cu eval_on): the single-algorithm instantiations have no oMi / J stores compiled in
    int rc2 = emul_eval_ex(d, n, tile, block, use_tmem, q, nullptr, nullptr, oMi, J, nullptr, nullptr, 0, variant);
    if (rc2 == 0) rc2 = emul_eval_ex(d, n, tile, block, use_tmem, q, v, a, nullptr, nullptr, M, tau, sparse_m, variant);
    return rc2;
  }
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const bool do_m = M != nullptr, do_tau = tau != nullptr;
  if (variant >= 0) {
    if (variant > c.m.eval_variant && !(variant == RBD_EVAL_SIMPLE && c.m.eval_variant >= RBD_EVAL_SIMPLE) &&
        !(variant == RBD_EVAL_FFROOT8 && c.m.eval_variant == RBD_EVAL_FFROOT)) {
      g_err = "the model does not qualify for this instantiation"; return -102;
    }
    if (variant == RBD_EVAL_SIMPLE && c.m.eval_variant < RBD_EVAL_SIMPLE) { g_err = "not a simple-joint model"; return -102; }
    if (variant >= RBD_EVAL_CHAIN && variant != c.m.eval_variant && !(variant == RBD_EVAL_FFROOT8 && c.m.eval_variant == RBD_EVAL_FFROOT)) { g_err = "the model does not qualify for this instantiation"; return -102; }
  }
  if (!plan_eval(&c.m, do_m, do_tau, use_tmem != 0, block / 32, oMi != nullptr)) { g_err = "plan_eval failed"; return -100; }
  // the sweep reads the model through the staged blob (planned DevModel prefix + ColProg table), as the kernel does
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m, sparse_m != 0);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const int kv = variant == -2 ? bm.plan.kvariant : variant;
  switch (kv) {
    case RBD_EVAL_FFROOT8: run_eval_ffroot8(bm, n, tile, q, v, a, oMi, J, M, tau); break;
    case RBD_EVAL_FFROOT: run_eval_ffroot(bm, n, tile, q, v, a, oMi, J, M, tau); break;
    case RBD_EVAL_CHAIN: run_eval_chain(bm, n, tile, q, v, a, oMi, J, M, tau); break;
    case RBD_EVAL_SIMPLE: run_eval_simple(bm, n, tile, q, v, a, oMi, J, M, tau); break;
    default: run_eval_generic(bm, n, tile, q, v, a, oMi, J, M, tau); break;
  }
  return 0;
}

// Mass / ForceField products (rbd_addMDx_soa / rbd_addForce_soa): the RNEA sweep with absent v or a, gravity switch
// and accumulation out <- out + scale * rnea(...)
extern "C" int emul_mass_force(const rbd_model_desc* d, long long n, long long tile, double* q, double* v, double* a,
                               double* out, double scale, int gravity_on) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, false, true, true, 0, false)) { g_err = "plan_eval failed"; return -100; }
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const EvalPlan& pl = bm.plan;
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    EvalIO io;
    io.q = make_field(q, i, bm.nq, tile); io.v = make_field(v, i, bm.nv, tile); io.a = make_field(a, i, bm.nv, tile);
    io.oMi = make_field(nullptr, i, 1, tile); io.J = io.oMi; io.M = io.oMi;
    io.tau = make_field(out, i, bm.nv, tile);
    io.tau_scale = scale; io.tau_acc = 1; io.gravity_on = gravity_on;
    InRing inr;
    eval_instance<false, true, 0>(bm, io, S, T, inr);
  }
  return 0;
}

// FP32 mode of the fused evaluation (namespace rbd::f32): float I/O, float state, the same sweep and storage plan
template <int TILE>
static void eval_one_f32(const DevModelT<float>& m, long long i, long long tile, float* q, float* v, float* a, float* oMi,
                         float* J, float* M, float* tau, const f32::Stack& S, const f32::TmStack& T) {
  const bool do_m = M != nullptr, do_tau = tau != nullptr;
  f32::EvalIOT<TILE> io;
  f32::InRing inr;  // register prefetch (the FP32 plan leaves no room for the ring)
  io.q = f32::make_field_t<TILE>(q, i, m.nq, tile); io.v = f32::make_field_t<TILE>(v, i, m.nv, tile);
  io.a = f32::make_field_t<TILE>(a, i, m.nv, tile);
  io.oMi = f32::make_field_t<TILE>(oMi, i, 12 * (m.njoints - 1), tile); io.J = f32::make_field_t<TILE>(J, i, 6 * m.nv, tile);
  io.M = f32::make_field_t<TILE>(M, i, m.mfields, tile); io.tau = f32::make_field_t<TILE>(tau, i, m.nv, tile);
  if (do_m && do_tau) f32::eval_instance<true, true, TILE>(m, io, S, T, inr);
  else if (do_m) f32::eval_instance<true, false, TILE>(m, io, S, T, inr);
  else if (do_tau) f32::eval_instance<false, true, TILE>(m, io, S, T, inr);
  else f32::eval_instance<false, false, TILE>(m, io, S, T, inr);
}

extern "C" int emul_eval_f32(const rbd_model_desc* d, long long n, long long tile, int block, int use_tmem, float* q,
                             float* v, float* a, float* oMi, float* J, float* M, float* tau) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const bool do_m = M != nullptr, do_tau = tau != nullptr;
  if (!plan_eval(&c.m, do_m, do_tau, use_tmem != 0, block / 32, oMi != nullptr, false, B200Limits::kRegsPerThread, 4)) {
    g_err = "plan_eval failed"; return -100;
  }
  std::vector<DevModelT<float>> fm(1);
  const std::vector<unsigned long long> blob = build_eval_blob_f32(&c.m, &fm[0]);
  const DevModelT<float>& bm = *reinterpret_cast<const DevModelT<float>*>(blob.data());
  const EvalPlan& pl = bm.plan;
  const int bt = pl.warps * 32;
  std::vector<float> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e30f);
  std::vector<float> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e30f);
  for (long long b0 = 0; b0 < n; b0 += bt)
    for (int t = 0; t < bt && b0 + t < n; ++t) {
      const long long i = b0 + t;
      f32::Stack S = f32::make_stack(smem.data(), t, pl.smem_doubles);
      f32::TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
      if (tile == 32) eval_one_f32<32>(bm, i, tile, q, v, a, oMi, J, M, tau, S, T);
      else eval_one_f32<0>(bm, i, tile, q, v, a, oMi, J, M, tau, S, T);
    }
  return 0;
}

extern "C" int emul_apply(const rbd_model_desc* d, long long n, long long tile, int block, double* q_joints,
                          double* root, double* qcache, long long cache_tile, double* out) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const int sd = smem_apply(c.m);
  std::vector<double> smem((size_t)(sd > 0 ? sd : 1) * ((block + 31) / 32) * 32, 1e300);
  const int off = (c.m.has_ff_root && root) ? 7 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % block);
    SplitQ sq{make_field(root, i, 7, tile), make_field(q_joints, i, c.m.nq - off, tile), off};
    Stack S = make_stack(smem.data(), t, sd);
    apply_instance(c.m, c.tb, sq, make_field(qcache, i, c.m.nq, cache_tile), make_field(out, i, 7 * c.m.n_out, tile), S);
  }
  return 0;
}

extern "C" int emul_applyJ(const rbd_model_desc* d, long long n, long long tile, int block, double* qcache,
                           long long cache_tile, double* dq_joints, double* root_vel, double* out) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const int sd = smem_applyJ(c.m);
  std::vector<double> smem((size_t)(sd > 0 ? sd : 1) * ((block + 31) / 32) * 32, 1e300);
  const int off = (c.m.has_ff_root && root_vel) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % block);
    SplitQ dq{make_field(root_vel, i, 6, tile), make_field(dq_joints, i, c.m.nv - off, tile), off};
    Stack S = make_stack(smem.data(), t, sd);
    applyJ_instance(c.m, c.tb, make_field(qcache, i, c.m.nq, cache_tile), dq, make_field(out, i, 6 * c.m.n_out, tile), S);
  }
  return 0;
}

extern "C" int emul_applyJT(const rbd_model_desc* d, long long n, long long tile, int block, double* qcache,
                            long long cache_tile, double* in, double* out_tau, double* out_root) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const int sd = smem_applyJT(c.m);
  std::vector<double> smem((size_t)(sd > 0 ? sd : 1) * ((block + 31) / 32) * 32, 1e300);
  const int off = (c.m.has_ff_root && out_root) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % block);
    DenseWrenchSrc src{make_field(in, i, 6 * c.m.n_out, tile), &c.tb};
    SplitAccumSink sink{make_field(out_root, i, 6, tile), make_field(out_tau, i, c.m.nv - off, tile), off};
    Stack S = make_stack(smem.data(), t, sd);
    applyJT_instance(c.m, c.tb, make_field(qcache, i, c.m.nq, cache_tile), src, sink, S);
  }
  return 0;
}

// applyJT on the shared + tensor memory storage plan (the batched K3 kernel's body)
extern "C" int emul_applyJT_plan(const rbd_model_desc* d, long long n, long long tile, int use_tmem, double* qcache,
                                 long long cache_tile, double* in, double* out_tau, double* out_root) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, false, true, use_tmem != 0, 0, false, true)) { g_err = "plan_eval failed"; return -100; }
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const EvalPlan& pl = bm.plan;
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const int off = (bm.has_ff_root && out_root) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    DenseWrenchSrc src{make_field(in, i, 6 * bm.n_out, tile), &c.tb};
    SplitAccumSink sink{make_field(out_root, i, 6, tile), make_field(out_tau, i, bm.nv - off, tile), off};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    applyJT_instance_plan(bm, c.tb, make_field(qcache, i, bm.nq, cache_tile), src, sink, S, T);
  }
  return 0;
}

// applyJT, ring variant (the batched K3 kernel's body on full tile-32 batches): storage plan and ring program of
// plan_jt_ring, the wrench ring replaced by a reader that follows the program's copy list, the reduction sink by +=
// allow_fwd: 0 pins the depth-stack sweep, 1 = the planner's automatic choice, 2 = forward projection wherever it fits;
// *variant receives 1 when the forward-projection body ran
extern "C" int emul_applyJT_ring(const rbd_model_desc* d, long long n, long long tile, int use_tmem, double* qcache,
                                 long long cache_tile, double* in, double* out_tau, double* out_root, int* ring_stages,
                                 int allow_fwd, int* variant) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  std::vector<int> prog;
  const bool arm = allow_fwd == 3;  // 3: the unrolled chain sweep on the stack variant's ring program (6- / 7-joint chains)
  if (arm) allow_fwd = 0;
  if (arm && !(c.m.eval_variant == RBD_EVAL_CHAIN && (c.m.njoints - 1 == 6 || c.m.njoints - 1 == 7))) return 1;
  const int nst = plan_jt_ring(&c.m, c.ht, use_tmem != 0, &prog, allow_fwd);
  if (variant) *variant = nst > 0 ? c.m.plan.jt_fwd : 0;
  if (ring_stages) *ring_stages = nst;
  if (nst == 0) return 1;  // the kernel would not run this model: nothing to emulate
  const std::vector<unsigned long long> blob = build_jt_blob(&c.m, prog);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const EvalPlan& pl = bm.plan;
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const JtProg jp{col_prog(bm)};
  if (jp.npieces() == 0) return 0;  // no wrench reaches a moving joint: the accumulated outputs stay as they are
  const int off = (bm.has_ff_root && out_root) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    DirectRingT<Field> ring{make_field(in, i, 6 * bm.n_out, tile), jp, 0, nullptr};
    SplitRedSinkT<Field> sink{make_field(out_root, i, 6, tile), make_field(out_tau, i, bm.nv - off, tile), off};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    if (arm) {  // the unrolled chain sweep (chain::applyJT_arm_instance): joint table as the kernel parameter carries it
      if (bm.njoints - 1 == 7) {
        chain::ArmParams<7> P;
        for (int k = 0; k < 7; ++k) P.joint[k] = bm.joint[k + 1];
        chain::applyJT_arm_instance<7>(P, jp.opl(), make_field(qcache, i, bm.nq, cache_tile), ring, sink);
      } else {
        chain::ArmParams<6> P;
        for (int k = 0; k < 6; ++k) P.joint[k] = bm.joint[k + 1];
        chain::applyJT_arm_instance<6>(P, jp.opl(), make_field(qcache, i, bm.nq, cache_tile), ring, sink);
      }
    } else if (bm.plan.jt_fwd) applyJT_instance_fwd(bm, jp.opl(), jp.anc(), make_field(qcache, i, bm.nq, cache_tile), ring, sink, S, T);
    else applyJT_instance_ring(bm, jp.opl(), make_field(qcache, i, bm.nq, cache_tile), ring, sink, S, T);
  }
  return 0;
}

// constraint rows on cached Jacobians (rows_cached_row): the cache (world joint Jacobian + oMi of every instance) is
// filled by the FK + Jacobian instantiation of the eval sweep from the cached configuration, as the library does
extern "C" int emul_applyJT_rows_cached(const rbd_model_desc* d, long long n, double* qcache, long long cache_tile,
                                        long long nrows, const int* row_ptr, const int* row_instance, const int* col,
                                        const double* val, double* out_rows, int* keep) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, false, false, true, 0, true)) { g_err = "plan_eval failed"; return -100; }
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const int nj = bm.njoints, nv = bm.nv;
  const long long nt = (n + cache_tile - 1) / cache_tile;
  std::vector<double> J((size_t)nt * cache_tile * 6 * nv + 1, 1e300), O((size_t)nt * cache_tile * 12 * (nj - 1) + 1, 1e300);
  std::vector<double> smem(64, 1e300), tmem(64, 1e300);
  for (long long i = 0; i < n; ++i) {
    Stack S = make_stack(smem.data(), 0, 1);
    TmStack T{tmem.data()};
    InRing inr;
    run_eval_generic_one<0>(bm, i, cache_tile, qcache, nullptr, nullptr, O.data(), J.data(), nullptr, nullptr, S, T, inr, 0);
  }
  for (long long r = 0; r < nrows; ++r) {
    double* row = out_rows + (size_t)r * nv;
    for (int k = 0; k < nv; ++k) row[k] = 0.0;
    const Field Jc = make_field(J.data(), row_instance[r], 6 * nv, cache_tile);
    const Field Oc = make_field(O.data(), row_instance[r], 12 * (nj - 1), cache_tile);
    rows_cached_row(bm, c.tb, Jc, Oc, col + row_ptr[r], val + 6 * (size_t)row_ptr[r], row_ptr[r + 1] - row_ptr[r],
                    [&](int cc, double t) { row[cc] += t; });
    double n2 = 0.0;
    for (int k = 0; k < nv; ++k) n2 += row[k] * row[k];
    keep[r] = n2 > 1e-12 ? 1 : 0;
  }
  return 0;
}

// dense LOCAL_WORLD_ALIGNED frame Jacobian of (instance, output) pairs from the same cache (frame_jacobian_cached)
extern "C" int emul_frame_jacobians(const rbd_model_desc* d, long long n, double* qcache, long long cache_tile,
                                    long long npairs, const int* instance, const int* out_index, double* Jout) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!plan_eval(&c.m, false, false, true, 0, true)) { g_err = "plan_eval failed"; return -100; }
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const int nj = bm.njoints, nv = bm.nv;
  const long long nt = (n + cache_tile - 1) / cache_tile;
  std::vector<double> J((size_t)nt * cache_tile * 6 * nv + 1, 1e300), O((size_t)nt * cache_tile * 12 * (nj - 1) + 1, 1e300);
  std::vector<double> smem(64, 1e300), tmem(64, 1e300);
  for (long long i = 0; i < n; ++i) {
    Stack S = make_stack(smem.data(), 0, 1);
    TmStack T{tmem.data()};
    InRing inr;
    run_eval_generic_one<0>(bm, i, cache_tile, qcache, nullptr, nullptr, O.data(), J.data(), nullptr, nullptr, S, T, inr, 0);
  }
  for (long long p = 0; p < npairs; ++p) {
    double* Jr = Jout + (size_t)p * 6 * nv;
    for (int k = 0; k < 6 * nv; ++k) Jr[k] = 0.0;
    const Field Jc = make_field(J.data(), instance[p], 6 * nv, cache_tile);
    const Field Oc = make_field(O.data(), instance[p], 12 * (nj - 1), cache_tile);
    frame_jacobian_cached(bm, c.tb, Jc, Oc, out_index[p], [&](int cc, const double* o) {
      for (int e = 0; e < 6; ++e) Jr[6 * cc + e] = o[e];
    });
  }
  return 0;
}

// forward dynamics (aba_instance): the scratch record is a poisoned heap array per lane
struct EmulScratch {
  double* p;
  double ld(int e) const { return p[e]; }
  void st(int e, double v) const { p[e] = v; }
  EmulScratch at(int e) const { return EmulScratch{p + e}; }
  void add(int e, double v) const { p[e] += v; }
  const double* ptr(int e) const { return p + e; }
};
extern "C" int emul_aba(const rbd_model_desc* d, long long n, long long tile, double* q, double* v, double* tau, double* ddq) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  for (int i = 1; i < c.m.njoints; ++i)
    if (c.m.joint[i].type == JT_FF && !(i == 1 && c.m.joint[i].parent == 0)) { g_err = "free-flyer below the root"; return -5; }
  const int E = aba_scratch_doubles(c.m);
  std::vector<double> scr((size_t)(E > 0 ? E : 1));
  for (long long i = 0; i < n; ++i) {
    std::fill(scr.begin(), scr.end(), 1e300);
    EmulScratch X{scr.data()};
    aba_instance(c.m, make_field(q, i, c.m.nq, tile), make_field(v, i, c.m.nv, tile), make_field(tau, i, c.m.nv, tile),
                 make_field(ddq, i, c.m.nv, tile), X);
  }
  return 0;
}

// forward dynamics, fused variant (aba2_instance): level records on the plan of plan_aba2 (shared + "tensor" memory,
// poisoned), the pass-3 record a poisoned heap array per lane.  warps: 0 = the planner's choice, > 0 pins the block size
// (smaller blocks move levels from tensor to shared memory and back: exercises both spaces).  Returns -100 when the
// plan does not fit.  plan_out (optional): {warps, smem_doubles, tmem_doubles, tmem_cols}
extern "C" int emul_aba2(const rbd_model_desc* d, long long n, long long tile, int use_tmem, int warps, double* q, double* v,
                         double* tau, double* ddq, int* plan_out) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  for (int i = 1; i < c.m.njoints; ++i)
    if (c.m.joint[i].type == JT_FF && !(i == 1 && c.m.joint[i].parent == 0)) { g_err = "free-flyer below the root"; return -5; }
  if (!plan_aba2(&c.m, use_tmem != 0, B200Limits::kMaxWarpsPerBlock, warps)) { g_err = "plan_aba2: nothing fits"; return -100; }
  const EvalPlan& pl = c.m.plan;
  if (plan_out) { plan_out[0] = pl.warps; plan_out[1] = pl.smem_doubles; plan_out[2] = pl.tmem_doubles; plan_out[3] = pl.tmem_cols; }
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const int E = aba2_scratch_doubles(c.m);
  std::vector<double> scr((size_t)(E > 0 ? E : 1));
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    std::fill(scr.begin(), scr.end(), 1e300);
    // poison this thread's level records too: nothing may survive from the previous instance
    for (int e = 0; e < pl.smem_doubles; ++e) smem[(size_t)(t / 32) * pl.smem_doubles * 32 + (size_t)e * 32 + t % 32] = 1e300;
    for (int e = 0; e < pl.tmem_doubles; ++e) tmem[(size_t)t * pl.tmem_doubles + e] = 1e300;
    EmulScratch X{scr.data()};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    aba2_instance(c.m, make_field(q, i, c.m.nq, tile), make_field(v, i, c.m.nv, tile), make_field(tau, i, c.m.nv, tile),
                  make_field(ddq, i, c.m.nv, tile), X, S, T);
  }
  return 0;
}

// inverse of the mass matrix (minv_instance): packed upper triangle, nv (nv + 1) / 2 fields; storage as emul_aba2
extern "C" int emul_minv(const rbd_model_desc* d, long long n, long long tile, int use_tmem, int warps, double* q, double* Mv,
                         int* plan_out) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  for (int i = 1; i < c.m.njoints; ++i)
    if (c.m.joint[i].type == JT_FF && !(i == 1 && c.m.joint[i].parent == 0)) { g_err = "free-flyer below the root"; return -5; }
  if (!plan_minv(&c.m, use_tmem != 0, B200Limits::kMaxWarpsPerBlock, warps)) { g_err = "plan_minv: nothing fits"; return -100; }
  const EvalPlan& pl = c.m.plan;
  if (plan_out) { plan_out[0] = pl.warps; plan_out[1] = pl.smem_doubles; plan_out[2] = pl.tmem_doubles; plan_out[3] = pl.tmem_cols; }
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const int E = RBD_MINV_REC * (c.m.njoints - 1) + RBD_MINV_ROOT_REC;
  std::vector<double> scr((size_t)E);
  const int mf = c.m.nv * (c.m.nv + 1) / 2;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    std::fill(scr.begin(), scr.end(), 1e300);
    for (int e = 0; e < pl.smem_doubles; ++e) smem[(size_t)(t / 32) * pl.smem_doubles * 32 + (size_t)e * 32 + t % 32] = 1e300;
    for (int e = 0; e < pl.tmem_doubles; ++e) tmem[(size_t)t * pl.tmem_doubles + e] = 1e300;
    EmulScratch X{scr.data()};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    minv_instance(c.m, make_field(q, i, c.m.nq, tile), make_field(Mv, i, mf, tile), X, S, T, 1);
  }
  return 0;
}

// block size of one launch of a persistent kernel planned for W warps (rbd_dev_model.h)
extern "C" int emul_pick_wpb(int W, long long ntiles, int num_sms) { return pick_warps_per_block(W, ntiles, num_sms); }

// derivatives of inverse dynamics (drnea_instance): DQ, DV nv x nv row-major fields, M (optional) packed upper triangle;
// level records on the plan of plan_drnea, storage as emul_aba2 (poisoned between instances)
// sparse != 0: DQ / DV in the support-only format (pattern: row_ptr [nv + 1], col_idx, either may be null; *nnz its size)
extern "C" int emul_drnea_pattern(const rbd_model_desc* d, int* row_ptr, int* col_idx, int* nnz) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  *nnz = drnea_support_pattern(c.m, row_ptr, col_idx);
  return 0;
}
extern "C" int emul_drnea(const rbd_model_desc* d, long long n, long long tile, int use_tmem, int warps, double* q, double* v,
                          double* a, double* DQ, double* DV, double* M, int* plan_out, int sparse) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  for (int i = 1; i < c.m.njoints; ++i)
    if (c.m.joint[i].type == JT_FF && !(i == 1 && c.m.joint[i].parent == 0)) { g_err = "free-flyer below the root"; return -5; }
  if (!plan_drnea(&c.m, use_tmem != 0, B200Limits::kMaxWarpsPerBlock, warps)) { g_err = "plan_drnea: nothing fits"; return -100; }
  const EvalPlan& pl = c.m.plan;
  if (plan_out) { plan_out[0] = pl.warps; plan_out[1] = pl.smem_doubles; plan_out[2] = pl.tmem_doubles; plan_out[3] = pl.tmem_cols; }
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const int nv = c.m.nv;
  std::vector<double> scr((size_t)(c.m.slot_total[0] > 0 ? c.m.slot_total[0] : 1));
  if (plan_out) plan_out[3] = c.m.slot_total[0];  // (scratch doubles per instance in place of the column count)
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    std::fill(scr.begin(), scr.end(), 1e300);
    for (int e = 0; e < pl.smem_doubles; ++e) smem[(size_t)(t / 32) * pl.smem_doubles * 32 + (size_t)e * 32 + t % 32] = 1e300;
    for (int e = 0; e < pl.tmem_doubles; ++e) tmem[(size_t)t * pl.tmem_doubles + e] = 1e300;
    EmulScratch X{scr.data()};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    const int K = sparse ? c.m.slot_total[1] : nv * nv;
    if (sparse)
      drnea_instance<true, false>(c.m, make_field(q, i, c.m.nq, tile), make_field(v, i, nv, tile), make_field(a, i, nv, tile),
                                  make_field(DQ, i, K, tile), make_field(DV, i, K, tile), make_field(nullptr, i, 1, tile), X, S, T, InRing());
    else if (M)
      drnea_instance<false, true>(c.m, make_field(q, i, c.m.nq, tile), make_field(v, i, nv, tile), make_field(a, i, nv, tile),
                                  make_field(DQ, i, K, tile), make_field(DV, i, K, tile), make_field(M, i, nv * (nv + 1) / 2, tile), X, S, T, InRing());
    else
      drnea_instance<false, false>(c.m, make_field(q, i, c.m.nq, tile), make_field(v, i, nv, tile), make_field(a, i, nv, tile),
                                   make_field(DQ, i, K, tile), make_field(DV, i, K, tile), make_field(nullptr, i, 1, tile), X, S, T, InRing());
  }
  return 0;
}

// assembled mapping Jacobian (getJ_block_cached) of the selected outputs for every instance; also returns the pattern
// (blk_ptr [n_sel + 1], cols): val [n][6 blk_ptr[n_sel]]
extern "C" int emul_getJ(const rbd_model_desc* d, long long n, double* qcache, long long cache_tile, int n_sel,
                         const int* sel, int* blk_ptr, int* cols, double* val) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  if (!getJ_pattern(c.m, c.ht, n_sel, sel, blk_ptr, cols)) { g_err = "bad selection"; return -1; }
  if (!val) return 0;
  if (!plan_eval(&c.m, false, false, true, 0, true)) { g_err = "plan_eval failed"; return -100; }
  const std::vector<unsigned long long> blob = build_eval_blob(&c.m);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const int nj = bm.njoints, nv = bm.nv;
  const long long nt = (n + cache_tile - 1) / cache_tile;
  std::vector<double> J((size_t)nt * cache_tile * 6 * nv + 1, 1e300), O((size_t)nt * cache_tile * 12 * (nj - 1) + 1, 1e300);
  std::vector<double> smem(64, 1e300), tmem(64, 1e300);
  for (long long i = 0; i < n; ++i) {
    Stack S = make_stack(smem.data(), 0, 1);
    TmStack T{tmem.data()};
    InRing inr;
    run_eval_generic_one<0>(bm, i, cache_tile, qcache, nullptr, nullptr, O.data(), J.data(), nullptr, nullptr, S, T, inr, 0);
    const Field Jc = make_field(J.data(), i, 6 * nv, cache_tile);
    const Field Oc = make_field(O.data(), i, 12 * (nj - 1), cache_tile);
    for (int s = 0; s < n_sel; ++s)
      getJ_block_cached(bm, c.tb, Jc, Oc, sel ? sel[s] : s, val + (size_t)i * 6 * blk_ptr[n_sel] + 6 * (size_t)blk_ptr[s]);
  }
  return 0;
}

// applyDJT (geometric stiffness, optional extension): out += kfac * d/de [J(q (+) e dq)^T in]
extern "C" int emul_applyDJT(const rbd_model_desc* d, long long n, long long tile, double* qcache, long long cache_tile,
                             double* dq_joints, double* dq_root, double* in, double kfac, double* out_tau, double* out_root) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const int sd = RBD_DJT_LEVEL * c.m.max_depth + RBD_DJT_BRANCH * c.m.nbranch;
  std::vector<double> smem((size_t)(sd > 0 ? sd : 1) * 32, 1e300);
  const int off = (c.m.has_ff_root && dq_root) ? 6 : 0;
  const int ooff = (c.m.has_ff_root && out_root) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    SplitQ dq{make_field(dq_root, i, 6, tile), make_field(dq_joints, i, c.m.nv - off, tile), off};
    SplitRedSinkT<Field> sink{make_field(out_root, i, 6, tile), make_field(out_tau, i, c.m.nv - ooff, tile), ooff};
    Stack S = make_stack(smem.data(), (int)(i % 32), sd);
    applyDJT_instance(c.m, c.tb, make_field(qcache, i, c.m.nq, cache_tile), dq, make_field(in, i, 6 * c.m.n_out, tile), sink, S,
                      kfac);
  }
  return 0;
}

// applyDJT on the ring kernel's forward-projection sweep (applyDJT_instance_fwd; plan_djt_ring); returns 1 when the
// planner leaves the model to the shared-memory kernel
extern "C" int emul_applyDJT_ring(const rbd_model_desc* d, long long n, long long tile, double* qcache, long long cache_tile,
                                  double* dq_joints, double* dq_root, double* in, double kfac, double* out_tau,
                                  double* out_root) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  std::vector<int> prog;
  const int nst = plan_djt_ring(&c.m, c.ht, true, &prog);
  if (nst == 0) return 1;
  const std::vector<unsigned long long> blob = build_jt_blob(&c.m, prog);
  const DevModel& bm = *reinterpret_cast<const DevModel*>(blob.data());
  const EvalPlan& pl = bm.plan;
  const int bt = pl.warps * 32;
  std::vector<double> smem((size_t)(pl.smem_doubles > 0 ? pl.smem_doubles : 1) * bt, 1e300);
  std::vector<double> tmem((size_t)(pl.tmem_doubles > 0 ? pl.tmem_doubles : 1) * bt, 1e300);
  const JtProg jp{col_prog(bm)};
  if (jp.npieces() == 0) return 0;
  const int off = (bm.has_ff_root && dq_root) ? 6 : 0;
  const int ooff = (bm.has_ff_root && out_root) ? 6 : 0;
  for (long long i = 0; i < n; ++i) {
    const int t = (int)(i % bt);
    DirectRingT<Field> ring{make_field(in, i, 6 * bm.n_out, tile), jp, 0, nullptr};
    SplitQ dq{make_field(dq_root, i, 6, tile), make_field(dq_joints, i, bm.nv - off, tile), off};
    SplitRedSinkT<Field> sink{make_field(out_root, i, 6, tile), make_field(out_tau, i, bm.nv - ooff, tile), ooff};
    Stack S = make_stack(smem.data(), t, pl.smem_doubles);
    TmStack T{tmem.data() + (size_t)t * (pl.tmem_doubles > 0 ? pl.tmem_doubles : 1)};
    applyDJT_instance_fwd(bm, jp.opl(), jp.anc(), make_field(qcache, i, bm.nq, cache_tile), dq, ring, sink, S, T, kfac);
  }
  return 0;
}

// plan of the applyJT ring kernel for this model: out = {ring stages (0 = register path), forward projection?,
// warps per block, tensor-memory doubles, shared-memory doubles, shared-memory bytes of the block}
extern "C" int emul_jt_ring_plan(const rbd_model_desc* d, int use_tmem, int fwd_mode, int* out /*[6]*/) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  std::vector<int> prog;
  const int nst = plan_jt_ring(&c.m, c.ht, use_tmem != 0, &prog, fwd_mode);
  out[0] = nst;
  if (nst == 0) { out[1] = out[2] = out[3] = out[4] = out[5] = 0; return 0; }
  const EvalPlan& pl = c.m.plan;
  out[1] = pl.jt_fwd; out[2] = pl.warps; out[3] = pl.tmem_doubles; out[4] = pl.smem_doubles;
  out[5] = (int)(eval_model_bytes(c.m) + (size_t)pl.warps * JtRing::kMaxStages * 8 + (size_t)pl.warps * 32 * pl.smem_doubles * 8 +
                 (size_t)pl.warps * nst * JtRing::kStageBytes);
  return 0;
}

extern "C" int emul_applyJT_rows(const rbd_model_desc* d, long long nrows, int block, double* qcache,
                                 long long cache_tile, const int* row_ptr, const int* row_instance, const int* col,
                                 const double* val, double* out_rows, int* keep) {
  Ctx c;
  int rc = make_ctx(d, &c);
  if (rc) return rc;
  const int sd = smem_applyJT(c.m);
  std::vector<double> smem((size_t)(sd > 0 ? sd : 1) * ((block + 31) / 32) * 32, 1e300);
  for (long long r = 0; r < nrows; ++r) {
    const int t = (int)(r % block);
    double n2 = 0.0;
    RowWrenchSrc src{col + row_ptr[r], val + 6 * (size_t)row_ptr[r], row_ptr[r + 1] - row_ptr[r], &c.tb};
    RowSink sink{out_rows + (size_t)r * c.m.nv, &n2};
    Stack S = make_stack(smem.data(), t, sd);
    applyJT_instance(c.m, c.tb, make_field(qcache, row_instance[r], c.m.nq, cache_tile), src, sink, S);
    keep[r] = n2 > 1e-12 ? 1 : 0;
  }
  return 0;
}
