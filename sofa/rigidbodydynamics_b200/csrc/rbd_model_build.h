// rbd_model_build.h — host-side validation of an rbd_model_desc and construction of the device tables.
// Plain C++ (no CUDA) so the CPU unit tests of the kernel logic (tests/emul) build the same structures.
// Plays the role of KinematicChainMapping::setModel / setBodyCoMFrames (reference KinematicChainMapping.inl:304-317):
// everything derived from the model once, nothing per step.
#pragma once
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../../include/rbd_b200.h"
#include "rbd_dev_model.h"

namespace rbd {

struct HostTables {
  std::vector<int> sorted_k, outk_parent, sorted_ident;
  std::vector<double> sorted_pl, outk_pl;
};

// Column programs of M for one planned model (offA values come from plan_eval): see RBD_PROG_* in rbd_dev_model.h.
// Fills joint[*].prog_off / nzseg, the place of every joint's columns in the M output (RBD_MBASE / RBD_MOWN),
// m->mfields and m->prog_len; returns the table.  sparse_m = the support-only M format (structural zeros not stored).
inline std::vector<int> build_col_prog(DevModel* m, bool sparse_m = false) {
  std::vector<int> prog;
  const int nj = m->njoints;
  int mcount = 0;  // support-only entries before the current joint
  for (int k = 1; k < nj; ++k) {
    DevJoint& jk = m->joint[k];
    jk.prog_off = (int)prog.size();
    std::vector<int> chain;  // ancestors, root first
    for (int a = jk.parent; a != 0; a = m->joint[a].parent) chain.insert(chain.begin(), a);
    int nanc = 0;  // dofs of the ancestors
    for (int a : chain) {
      prog.push_back(RBD_PROG_ANC(sparse_m ? nanc : m->joint[a].idx_v, m->joint[a].type == JT_FF ? 1 : 0, m->joint[a].offA));
      nanc += m->joint[a].nv;
    }
    // zero segments: rows below the joint's first dof that no ancestor owns
    int nz = 0;
    if (!sparse_m) {
      int r = 0;
      for (int a : chain) {
        const int iv = m->joint[a].idx_v;
        if (iv > r) { prog.push_back(RBD_PROG_ZSEG(r, iv - r)); ++nz; }
        r = iv + m->joint[a].nv;
      }
      if (jk.idx_v > r) { prog.push_back(RBD_PROG_ZSEG(r, jk.idx_v - r)); ++nz; }
    }
    jk.nzseg = nz;
    RBD_MBASE(jk) = sparse_m ? mcount : jk.idx_v * (jk.idx_v + 1) / 2;
    RBD_MOWN(jk) = sparse_m ? nanc : jk.idx_v;
    mcount += jk.nv * nanc + jk.nv * (jk.nv + 1) / 2;
  }
  m->mfields = sparse_m ? mcount : m->nv * (m->nv + 1) / 2;
  m->prog_len = (int)prog.size();
  m->prog_words_off = (int)((dev_model_bytes(*m) + 15) / 16 * 2);
  return prog;
}

// Pattern of the assembled mapping Jacobian restricted to n_sel mapped outputs (sel == nullptr: all of them, in order):
// block row s holds the dofs on the path from output sel[s]'s parent joint to the root, ascending.  blk_ptr
// [n_sel + 1] counts COLUMNS (6 values each); cols [blk_ptr[n_sel]] may be null.  Returns false on a bad index.
inline bool getJ_pattern(const DevModel& m, const HostTables& tb, int n_sel, const int* sel, int* blk_ptr, int* cols) {
  int nnz = 0;
  for (int s = 0; s < n_sel; ++s) {
    const int k = sel ? sel[s] : s;
    if (k < 0 || k >= m.n_out) return false;
    blk_ptr[s] = nnz;
    std::vector<int> chain;
    for (int a = tb.outk_parent[k]; a != 0; a = m.joint[a].parent) chain.insert(chain.begin(), a);
    for (int a : chain)
      for (int e = 0; e < m.joint[a].nv; ++e) { if (cols) cols[nnz] = m.joint[a].idx_v + e; ++nnz; }
  }
  blk_ptr[n_sel] = nnz;
  return true;
}

// Support-only pattern of M (upper triangle, by columns): entry (r, c), r <= c, is structurally non-zero iff the joint
// owning dof r is the joint owning dof c or one of its ancestors.  col_ptr [nv + 1], row_idx [nnz] (either may be null);
// returns nnz.  The order is exactly the order of the fields the support-only eval writes (build_col_prog above).
inline int mass_support_pattern(const DevModel& m, int* col_ptr, int* row_idx) {
  int nnz = 0;
  for (int k = 1; k < m.njoints; ++k) {
    const DevJoint& jk = m.joint[k];
    std::vector<int> chain;
    for (int a = jk.parent; a != 0; a = m.joint[a].parent) chain.insert(chain.begin(), a);
    for (int cc = 0; cc < jk.nv; ++cc) {
      if (col_ptr) col_ptr[jk.idx_v + cc] = nnz;
      for (int a : chain)
        for (int e = 0; e < m.joint[a].nv; ++e) { if (row_idx) row_idx[nnz] = m.joint[a].idx_v + e; ++nnz; }
      for (int e = 0; e <= cc; ++e) { if (row_idx) row_idx[nnz] = jk.idx_v + e; ++nnz; }
    }
  }
  if (col_ptr) col_ptr[m.nv] = nnz;
  return nnz;
}

// The blob the eval kernel stages into shared memory: planned DevModel prefix (16-byte rounded) + ColProg table.
inline std::vector<unsigned long long> build_eval_blob(DevModel* m, bool sparse_m = false) {
  const std::vector<int> prog = build_col_prog(m, sparse_m);
  std::vector<unsigned long long> blob(eval_model_bytes(*m) / 8 - 2, 0ull);
  memcpy(blob.data(), m, dev_model_bytes(*m));
  if (!prog.empty()) memcpy(blob.data() + m->prog_words_off, prog.data(), prog.size() * sizeof(int));
  return blob;
}
// ... and its FP32 twin: the same planned model with float constants (`f` receives the converted model)
inline std::vector<unsigned long long> build_eval_blob_f32(DevModel* m, DevModelT<float>* f, bool sparse_m = false) {
  const std::vector<int> prog = build_col_prog(m, sparse_m);
  dev_model_to_f32(*m, f);
  f->prog_words_off = (int)((dev_model_bytes(*f) + 15) / 16 * 2);
  std::vector<unsigned long long> blob(eval_model_bytes(*f) / 8 - 2, 0ull);
  memcpy(blob.data(), f, dev_model_bytes(*f));
  if (!prog.empty()) memcpy(blob.data() + f->prog_words_off, prog.data(), prog.size() * sizeof(int));
  return blob;
}

inline int joint_nq_of(int t) { return t == JT_FF ? 7 : (t >= JT_RUBX && t <= JT_RUBU) ? 2 : (t == JT_UNIVERSE ? 0 : 1); }
inline int joint_nv_of(int t) { return t == JT_FF ? 6 : (t == JT_UNIVERSE ? 0 : 1); }

// returns RBD_OK or a negative rbd_status; `err` explains.
inline int build_dev_model(const rbd_model_desc* d, DevModel* m, HostTables* tb, std::string* err) {
  auto fail = [&](int code, const std::string& s) { *err = s; return code; };
  if (!d || !m || !tb) return fail(RBD_ERR_INVALID_ARG, "null argument");
  if (!d->parents || !d->jtype || !d->axis || !d->placement || !d->idx_q || !d->idx_v || !d->inertia)
    return fail(RBD_ERR_INVALID_ARG, "descriptor has null joint arrays");
  if (d->njoints < 1) return fail(RBD_ERR_MODEL, "njoints < 1");
  if (d->njoints > RBD_MAX_JOINTS)
    return fail(RBD_ERR_UNSUPPORTED, "njoints exceeds RBD_MAX_JOINTS (" + std::to_string(RBD_MAX_JOINTS) +
                                         "): the joint table no longer fits the staged-constant path");
  if (d->nframes < 0 || d->n_out < 0) return fail(RBD_ERR_MODEL, "negative frame counts");
  if ((d->nframes > 0 && (!d->frame_parent || !d->frame_placement)) || (d->n_out > 0 && !d->out_frames))
    return fail(RBD_ERR_INVALID_ARG, "descriptor has null frame arrays");
  const int nj = d->njoints;
  memset(m, 0, sizeof(*m));
  if (d->parents[0] != 0 || d->jtype[0] != JT_UNIVERSE) return fail(RBD_ERR_MODEL, "joint 0 must be the universe");
  int iq = 0, iv = 0;
  std::vector<int> depth(nj, 0);
  for (int i = 0; i < nj; ++i) {
    DevJoint& j = m->joint[i];
    const int t = d->jtype[i];
    if (i > 0) {
      if (t < JT_RX || t > JT_FF) return fail(RBD_ERR_UNSUPPORTED, "unknown joint type at joint " + std::to_string(i));
      const int p = d->parents[i];
      if (p < 0 || p >= i) return fail(RBD_ERR_MODEL, "parents[i] < i violated at joint " + std::to_string(i));
      // DFS pre-order: the parent is joint i-1 or one of its ancestors
      int k = i - 1;
      while (k != p && k != 0) k = d->parents[k];
      if (k != p) return fail(RBD_ERR_MODEL, "joints are not numbered in DFS pre-order at joint " + std::to_string(i));
      if (d->idx_q[i] != iq || d->idx_v[i] != iv)
        return fail(RBD_ERR_MODEL, "idx_q/idx_v are not cumulative at joint " + std::to_string(i));
      depth[i] = depth[p] + 1;
      m->joint[p].nchild++;
    }
    memcpy(j.pl, d->placement + 12 * i, 12 * sizeof(double));
    memcpy(j.axis, d->axis + 3 * i, 3 * sizeof(double));
    memcpy(j.Y, d->inertia + 10 * i, 10 * sizeof(double));
    if (t == JT_RX || t == JT_PX || t == JT_RUBX) { j.axis[0] = 1; j.axis[1] = 0; j.axis[2] = 0; }
    if (t == JT_RY || t == JT_PY || t == JT_RUBY) { j.axis[0] = 0; j.axis[1] = 1; j.axis[2] = 0; }
    if (t == JT_RZ || t == JT_PZ || t == JT_RUBZ) { j.axis[0] = 0; j.axis[1] = 0; j.axis[2] = 1; }
    j.parent = d->parents[i];
    j.type = t;
    j.idx_q = i ? iq : 0;
    j.idx_v = i ? iv : 0;
    j.nv = joint_nv_of(t);
    j.bidx = -1;
    iq += joint_nq_of(t);
    iv += joint_nv_of(t);
  }
  if (iq != d->nq || iv != d->nv) return fail(RBD_ERR_MODEL, "nq/nv do not match the joint list");
  m->njoints = nj; m->nq = d->nq; m->nv = d->nv; m->n_out = d->n_out;
  memcpy(m->gravity, d->gravity, sizeof(m->gravity));
  m->has_ff_root = (nj > 1 && d->jtype[1] == JT_FF && d->parents[1] == 0) ? 1 : 0;
  m->nq_joints = m->has_ff_root ? d->nq - 7 : d->nq;
  m->nv_joints = m->has_ff_root ? d->nv - 6 : d->nv;
  // stack slots by depth; branch slots for joints with >= 2 children.  The full layout (fused eval with CRBA) only
  // needs slots down to the deepest joint that has children: leaves are emitted from registers.
  int maxd = 0, maxd_internal = 0;
  for (int i = 1; i < nj; ++i) {
    maxd = std::max(maxd, depth[i]);
    if (m->joint[i].nchild > 0) maxd_internal = std::max(maxd_internal, depth[i]);
  }
  m->max_depth = maxd;
  {
    bool simple = true, chain = nj > 1, ffroot = nj > 1 && d->jtype[1] == JT_FF && d->parents[1] == 0;
    for (int i = 1; i < nj; ++i) {
      const int t = d->jtype[i];
      if (!(t == JT_RX || t == JT_RY || t == JT_RZ || t == JT_FF)) simple = false;
      if (t == JT_FF || d->parents[i] != i - 1) chain = false;
      if (i >= 2 && (t == JT_FF || d->parents[i] < 1)) ffroot = false;
    }
    m->eval_variant = !simple ? RBD_EVAL_GENERIC : chain ? RBD_EVAL_CHAIN : ffroot ? RBD_EVAL_FFROOT : RBD_EVAL_SIMPLE;
  }
  m->max_depth_internal = maxd_internal;
  std::vector<int> aw(maxd + 1, 6);
  for (int i = 1; i < nj; ++i) aw[depth[i]] = std::max(aw[depth[i]], RBD_SLOT_A(d->jtype[i]));
  std::vector<int> off_full(maxd + 2, 0), off_lite(maxd + 2, 0);
  for (int dd = 1; dd <= maxd; ++dd) {
    off_full[dd + 1] = off_full[dd] + aw[dd] + 16;  // A | f(6) | Y(10)
    off_lite[dd + 1] = off_lite[dd] + aw[dd] + 6;   // A | f(6)
  }
  m->slot_total[0] = off_full[maxd_internal + 1];
  m->slot_total[1] = off_lite[maxd + 1];
  int nb = 0;
  for (int i = 1; i < nj; ++i) {
    m->joint[i].slot[0] = off_full[depth[i]];
    m->joint[i].slot[1] = off_lite[depth[i]];
    m->joint[i].depth = depth[i];
    m->joint[i].offA = 0;
    m->joint[i].boff = 0;
    const double* pl = m->joint[i].pl;
    m->joint[i].pl_rot_identity = (pl[0] == 1.0 && pl[1] == 0.0 && pl[2] == 0.0 && pl[3] == 0.0 && pl[4] == 1.0 &&
                                   pl[5] == 0.0 && pl[6] == 0.0 && pl[7] == 0.0 && pl[8] == 1.0) ? 1 : 0;
    if (m->joint[i].nchild >= 2) m->joint[i].bidx = nb++;
  }
  m->nbranch = nb;
  for (int i = 0; i < nj; ++i) m->joint[i].msub = m->joint[i].Y[0];
  for (int i = nj - 1; i >= 1; --i) m->joint[m->joint[i].parent].msub += m->joint[i].msub;
  // mapped outputs grouped by parent joint (stable: keeps the reference's output order within a joint)
  tb->sorted_k.clear(); tb->sorted_pl.clear(); tb->sorted_ident.clear(); tb->outk_parent.assign(d->n_out, 0); tb->outk_pl.assign(12 * (size_t)d->n_out, 0.0);
  for (int k = 0; k < d->n_out; ++k) {
    const int f = d->out_frames[k];
    if (f < 0 || f >= d->nframes) return fail(RBD_ERR_MODEL, "out_frames entry out of range");
    const int pj = d->frame_parent[f];
    if (pj < 0 || pj >= nj) return fail(RBD_ERR_MODEL, "frame parent joint out of range");
    tb->outk_parent[k] = pj;
    memcpy(&tb->outk_pl[12 * (size_t)k], d->frame_placement + 12 * (size_t)f, 12 * sizeof(double));
  }
  for (int j = 0; j < nj; ++j) {
    m->joint[j].out_begin = (int)tb->sorted_k.size();
    for (int k = 0; k < d->n_out; ++k)
      if (tb->outk_parent[k] == j) {
        tb->sorted_k.push_back(k);
        {
          const double* pl = &tb->outk_pl[12 * (size_t)k];
          tb->sorted_ident.push_back((pl[0] == 1.0 && pl[1] == 0.0 && pl[2] == 0.0 && pl[3] == 0.0 && pl[4] == 1.0 &&
                                      pl[5] == 0.0 && pl[6] == 0.0 && pl[7] == 0.0 && pl[8] == 1.0) ? 1 : 0);
        }
        tb->sorted_pl.insert(tb->sorted_pl.end(), &tb->outk_pl[12 * (size_t)k], &tb->outk_pl[12 * (size_t)k] + 12);
      }
    m->joint[j].out_end = (int)tb->sorted_k.size();
  }
  return RBD_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Storage plan of the fused eval kernel for one variant (do_m = CRBA, do_tau = RNEA).
//
// Per-instance sweep state (doubles), D = number of depth levels that hold joints with children:
//   A-stack   sum of A widths over the D levels (6, free-flyer level 12)        shared memory, read by every column
//   f-stack   6 D          (RNEA subtree forces)                                  shared memory
//   Y-stack   10 D         (CRBA composite inertias)                              shared or tensor memory
//   branch    per joint with >= 2 children: (R|p) 12 [not for a free-flyer: its A record already is R|p] + v|a 12
//             (RNEA)                                                              shared or tensor memory, per record
// B200 budget per SM: 227 KB shared memory (minus the staged joint table per block), 512 tensor-memory columns of
// 128 lanes x 32 bit: a warp owns the 32 lanes of quarter (warp % 4), so a thread owns 512 / ceil(W / 4) columns =
// 256 / ceil(W / 4) doubles in a block of W warps.  The plan maximises resident warps per SM.
struct B200Limits {
  static constexpr int kSmemBytes = 227 * 1024;
  static constexpr int kSmemReservedPerBlock = 2048;  // 1 KB driver reservation per block + allocation granularity slack
  static constexpr int kTmemCols = 512;
  static constexpr int kRegsPerSM = 65536;
  static constexpr int kRegsPerThread = 255;  // upper bound of what ptxas may give the eval kernels (launch_bounds 256)
  static constexpr int kMaxWarpsPerBlock = 8;
  static constexpr int kMaxWarpsPerBlockRnea = 12;  // eval_kernel<false, true, *>: __launch_bounds__(384, 1)
  static constexpr int kRegsPerThreadRnea = 170;
  static constexpr int kRegsPerThreadWide = 168;  // eval_kernel_chain / eval_kernel_ffroot: __maxnreg__(168)
};

static constexpr int kEvalInSlots = 2;  // == RBD_INR_SLOTS (rbd_sweeps_body.h)

inline int pow2_cols(int cols) {
  int c = 32;
  while (c < cols) c *= 2;
  return c;
}

// Fills m->plan and m->joint[*].boff.  forced_warps > 0 pins the block size.  Returns false when nothing fits.
inline bool plan_eval_cfg(DevModel* m, bool do_m, bool do_tau, bool use_tmem, int forced_warps, bool omi_out, bool jt,
                          int regs_per_thread, int elem_bytes, int wide_warps);
inline bool plan_eval(DevModel* m, bool do_m, bool do_tau, bool use_tmem, int forced_warps = 0, bool omi_out = false,
                      bool jt = false, int regs_per_thread = B200Limits::kRegsPerThread, int elem_bytes = 8) {
  // Serial chains and free-flyer-rooted trees have 168-register instantiations of the CRBA sweeps (rbd::chain,
  // rbd::ffroot): try blocks of up to 12 warps first and keep that plan when it really puts more than 8 warps on an SM
  // (small robots: Panda, Solo); otherwise (Talos: the sweep state caps the SM at 8 warps) the standard plan runs on
  // the 255-register kernels, which need no spills.
  if (elem_bytes == 8 && !jt && do_m && forced_warps == 0 && m->eval_variant >= RBD_EVAL_CHAIN &&
      regs_per_thread == B200Limits::kRegsPerThread) {
    DevModel t = *m;
    if (plan_eval_cfg(&t, do_m, do_tau, use_tmem, 0, omi_out, false, B200Limits::kRegsPerThreadWide, 8,
                      B200Limits::kMaxWarpsPerBlockRnea) &&
        t.plan.warps * t.plan.blocks_per_sm > B200Limits::kMaxWarpsPerBlock) {
      *m = t;
      m->plan.kvariant = m->eval_variant;
      return true;
    }
  }
  const bool ok = plan_eval_cfg(m, do_m, do_tau, use_tmem, forced_warps, omi_out, jt, regs_per_thread, elem_bytes, 0);
#ifndef RBD_USE_FFROOT8
#define RBD_USE_FFROOT8 1
#endif
  if (ok) m->plan.kvariant = m->eval_variant >= RBD_EVAL_SIMPLE ? RBD_EVAL_SIMPLE : RBD_EVAL_GENERIC;
  // a floating-base robot on the standard (<= 8-warp) plan: the ffroot sweep with the full register budget — the fused
  // CRBA + RNEA variant only (A/B r2: Talos 2.29 -> see DESIGN.md §4.1a)
  if (ok && RBD_USE_FFROOT8 && elem_bytes == 8 && !jt && do_m && do_tau && forced_warps == 0 && m->eval_variant == RBD_EVAL_FFROOT &&
      regs_per_thread == B200Limits::kRegsPerThread)
    m->plan.kvariant = RBD_EVAL_FFROOT8;
#ifndef RBD_RNEA_FFROOT
#define RBD_RNEA_FFROOT 1
#endif
  // the RNEA-only instantiation (rnea, Mass / ForceField products) of a floating-base robot: the ffroot sweep too (168
  // registers, the same 12-warp blocks as the 170-register simple kernel)
  if (ok && RBD_RNEA_FFROOT && elem_bytes == 8 && !jt && !do_m && do_tau && forced_warps == 0 && m->eval_variant == RBD_EVAL_FFROOT &&
      regs_per_thread == B200Limits::kRegsPerThread)
    m->plan.kvariant = RBD_EVAL_FFROOT;
  return ok;
}
// wide_warps > 0: blocks of up to that many warps at regs_per_thread registers (the caller vouches for the kernel)
inline bool plan_eval_cfg(DevModel* m, bool do_m, bool do_tau, bool use_tmem, int forced_warps, bool omi_out, bool jt,
                          int regs_per_thread, int elem_bytes, int wide_warps) {
  // elem_bytes: 8 = the FP64 product, 4 = the optional FP32 mode (state elements are floats: half the shared-memory
  // bytes, one tensor-memory column each)
  // jt: the applyJT sweep — stacks of the (no CRBA, RNEA) variant, but branch records hold R|p only (no v|a)
  const int nj = m->njoints;
  const bool need_stack = do_m || do_tau;
  const int D = need_stack ? m->max_depth_internal : 0;
  int totalA = 0;
  if (need_stack) {
    std::vector<int> aw(D + 1, 6);
    for (int i = 1; i < nj; ++i)
      if (m->joint[i].nchild > 0) aw[m->joint[i].depth] = std::max(aw[m->joint[i].depth], RBD_SLOT_A(m->joint[i].type));
    std::vector<int> pre(D + 2, 0);
    for (int dd = 1; dd <= D; ++dd) pre[dd + 1] = pre[dd] + aw[dd];
    totalA = pre[D + 1];
    for (int i = 1; i < nj; ++i) m->joint[i].offA = m->joint[i].depth <= D ? pre[m->joint[i].depth] : 0;
  }
  {
    // sizes the ColProg table for the staged-blob budget (the packed-M program is the longer one; its length does not
    // depend on the plan).  On a copy: build_col_prog writes the M placement into the joints' slot words, which the
    // caller's model may still need as stack slots.
    DevModel t = *m;
    (void)build_col_prog(&t);
    m->prog_len = t.prog_len; m->prog_words_off = t.prog_words_off;
  }
  // the applyJT kernel also stages the per-output tables behind the blob
  size_t model_bytes = eval_model_bytes(*m) + (jt ? (size_t)table_words(m->n_out) * 8 : 0);
  if (elem_bytes == 4) {
    DevModelT<float> probe;
    probe.njoints = m->njoints; probe.prog_len = m->prog_len;
    model_bytes = eval_model_bytes(probe);
  }
  const int cols_per_elem = elem_bytes / 4;
  EvalPlan best{};
  std::vector<int> best_boff, bj;
  for (int i = 1; i < nj; ++i)
    if (m->joint[i].nchild >= 2) bj.push_back(i);
  int best_warps_sm = 0;
  // The RNEA-only instantiation (rnea, Mass / ForceField products) is compiled for blocks of up to 12 warps (<= 170
  // registers): its state is smaller, and three warps per scheduler hide more of the issue latency that bounds it.
  const bool wide = elem_bytes == 8 && !jt && do_tau && !do_m && regs_per_thread == B200Limits::kRegsPerThread;
  const int max_warps = wide_warps > 0 ? wide_warps : wide ? B200Limits::kMaxWarpsPerBlockRnea : B200Limits::kMaxWarpsPerBlock;
  if (wide) regs_per_thread = B200Limits::kRegsPerThreadRnea;
  for (int W = 1; W <= max_warps; ++W) {
    if (forced_warps > 0 && W != forced_warps) continue;
    const int cap = use_tmem ? (B200Limits::kTmemCols / ((W + 3) / 4)) / cols_per_elem : 0;  // elements per thread
    // per block size: the configuration with the most resident blocks; among equals the cheapest one (everything
    // saved, fewest f-stack levels in tensor memory)
    int best_c_w = 0;
    // cheapest configuration first: everything saved; then drop the v|a record of branch joints attached to the
    // universe (recomputed from the A record and the inputs: costs a dozen global re-reads per return)
    // ...; then re-read R|p of the other branch joints from the oMi output this thread wrote (when oMi is requested)
    const int ncfg = jt ? 1 : 1 + (do_tau ? 1 : 0) + ((do_tau && omi_out) ? 1 : 0);
    for (int cfg = 0; cfg < ncfg; ++cfg) {
      const int drop_root_va = cfg >= 1, rp_from_omi = cfg >= 2;
      std::vector<int> bw(bj.size()), bflag(bj.size(), 0);
      for (size_t k = 0; k < bj.size(); ++k) {
        const DevJoint& jb = m->joint[bj[k]];
        // R|p: not for a free-flyer (its A record already is R|p)
        const bool ff_dedupe = need_stack && jb.type == JT_FF;
        const bool nova = drop_root_va && need_stack && jb.parent == 0;
        const bool rpomi = rp_from_omi && !ff_dedupe;
        bw[k] = ((ff_dedupe || rpomi) ? 0 : 12) + ((do_tau && !nova && !jt) ? 12 : 0);
        bflag[k] = (nova ? RBD_BOFF_NOVA : 0) | (rpomi ? RBD_BOFF_RPOMI : 0);
      }
      // greedy: Y-stack first (one vector access per visit), then the widest branch records that still fit, then
      // — only as many as shared memory needs to get rid of — the deepest f-stack levels
      int t_base = 0;
      EvalPlan pl{};
      std::vector<int> boff(bj.size(), 0);
      const int s_base = totalA;
      // Y-stack: whole if it fits the thread's tensor-memory share, else its deepest levels (long serial chains)
      const int ky = do_m ? std::min(D, cap / 9) : 0;  // levels in tensor memory
      pl.ysplit = do_m ? D - ky : 0; pl.ytbase = t_base; pl.y_in_tmem = ky > 0 ? 1 : 0;
      t_base += 9 * ky;
      const int ysz_s = do_m ? 9 * (D - ky) : 0;       // levels left in shared memory
      std::vector<int> order(bj.size());
      for (size_t k = 0; k < order.size(); ++k) order[k] = (int)k;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bw[a] > bw[b]; });
      std::vector<char> b_tm(bj.size(), 0);
      for (int k : order)
        if (bw[k] > 0 && t_base + bw[k] <= cap) { boff[k] = t_base | RBD_BOFF_TMEM; t_base += bw[k]; b_tm[k] = 1; }
      const int kmax = do_tau ? std::min(D, (cap - t_base) / 6) : 0;
      for (int kf = 0; kf <= kmax; ++kf) {
        int s_used = s_base, t_used = t_base;
        pl.fbase = s_used; pl.fsplit = do_tau ? D - kf : 0; s_used += do_tau ? 6 * (D - kf) : 0;
        pl.ftbase = t_used; t_used += 6 * kf;
        pl.ybase = s_used; s_used += ysz_s;
        for (size_t k = 0; k < bj.size(); ++k)
          if (!b_tm[k]) { boff[k] = s_used; s_used += bw[k]; }
        pl.smem_doubles = s_used; pl.tmem_doubles = t_used; pl.warps = W;
        pl.tmem_cols = t_used > 0 ? pow2_cols(cols_per_elem * t_used * ((W + 3) / 4)) : 0;
        const size_t block_bytes = model_bytes + (size_t)W * 32 * s_used * elem_bytes;
        if (block_bytes + B200Limits::kSmemReservedPerBlock > (size_t)B200Limits::kSmemBytes) continue;
        int c = (int)(B200Limits::kSmemBytes / (block_bytes + B200Limits::kSmemReservedPerBlock));
        if (pl.tmem_cols > 0) c = std::min(c, B200Limits::kTmemCols / pl.tmem_cols);
        c = std::min(c, B200Limits::kRegsPerSM / (regs_per_thread * W * 32));
        c = std::min(c, 32);
        if (c < 1) continue;
        pl.blocks_per_sm = c;
        if (c <= best_c_w) continue;  // not better than an earlier (cheaper) configuration of this block size
        best_c_w = c;
        // more resident warps wins; on ties the LARGER block: the warps of one block spread over the four SM
        // sub-partitions (warp % 4), whereas many one-warp blocks were measured 3x slower (r1c: 15.8 ms vs 5.2 ms)
        if (W * c >= best_warps_sm) {
          best_warps_sm = W * c; best = pl; best_boff = boff;
          for (size_t k = 0; k < bj.size(); ++k) best_boff[k] |= bflag[k];
        }
      }
    }
  }
  if (best_warps_sm == 0) return false;
  // asynchronous input ring (tile-32 batches): 3 elements per slot and instance, if the SM still has room for it next
  // to the plan (228 KB per SM, 1 KB of it reserved per resident block; 227 KB at most per block)
  best.in_slots = 0;
  if (!jt && nj - 1 >= kEvalInSlots) {
    const size_t blk = model_bytes + (size_t)best.warps * 32 * best.smem_doubles * elem_bytes +
                       (size_t)best.warps * 32 * 3 * kEvalInSlots * elem_bytes;
    if (blk <= (size_t)B200Limits::kSmemBytes && (size_t)best.blocks_per_sm * (blk + 1024) <= (size_t)228 * 1024)
      best.in_slots = kEvalInSlots;
  }
  m->plan = best;
  for (size_t k = 0; k < bj.size(); ++k) m->joint[bj[k]].boff = best_boff[k];
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Storage plan and ring program of the applyJT ring kernel (applyJT_instance_ring).  Placement is by record kind, fixed
// at compile time in the kernel: per depth level of 1-dof joints with children one record [F | A] (12 doubles) and per
// 1-dof branch joint one R|p record (12) in TENSOR memory; the level record of a free-flyer [R|p | F] (18) in shared
// memory.  What remains of the SM's shared memory becomes the per-warp wrench ring (stages of 4 wrench blocks).
// Fills m->plan (warps, smem_doubles, tmem_doubles, tmem_cols), joint[*].offA (level record), joint[*].boff (branch
// record) and m->prog_len / prog_words_off for the ring program `prog` (layout: rbd_sweeps_body.h, JtProg), which takes
// the place of the ColProg table in the staged blob.  Returns the number of ring stages per warp, 0 when the kernel
// cannot run this model (tensor memory off or too small for the tree's depth: the caller keeps the register path).
inline int plan_jt_ring_w(DevModel* m, const HostTables& tb, bool use_tmem, std::vector<int>* prog, int W, int fwd);
static constexpr int kJtMaxDepth = 12;  // == RBD_JT_MAXD (rbd_sweeps_body.h)
// fwd_mode: 0 = never the forward projection, 1 = automatic (order below), 2 = forward projection first (tests)
inline int plan_jt_ring(DevModel* m, const HostTables& tb, bool use_tmem, std::vector<int>* prog, int fwd_mode = 1) {
  // forward projection with 12 warps per SM for shallow trees (applyJT_instance_fwd); else the stack variant with 12
  // warps when the state fits a third of the tensor memory, else with 8
  // order: stack sweep with 12 warps (small robots: fewest instructions AND three warps per scheduler), forward
  // projection with 12 warps (shallow trees whose stack state needs the 8-warp share: Talos), stack sweep with 8
  const DevModel keep = *m;
  int nst = fwd_mode == 2 ? plan_jt_ring_w(m, tb, use_tmem, prog, JtRing::kWarpsSmall, 1) : 0;
  if (nst == 0) { *m = keep; nst = plan_jt_ring_w(m, tb, use_tmem, prog, JtRing::kWarpsSmall, 0); }
  if (nst == 0 && fwd_mode == 1) { *m = keep; nst = plan_jt_ring_w(m, tb, use_tmem, prog, JtRing::kWarpsSmall, 1); }
  if (nst == 0) { *m = keep; nst = plan_jt_ring_w(m, tb, use_tmem, prog, JtRing::kWarps, 0); }
  return nst;
}
// applyDJT on the ring kernel: 8 warps (the sweep carries dV, dA, dF on top of applyJT's state: 226 registers, a
// 12-warp build spills); 0 = the shared-memory kernel keeps the job
inline int plan_djt_ring(DevModel* m, const HostTables& tb, bool use_tmem, std::vector<int>* prog) {
  return plan_jt_ring_w(m, tb, use_tmem, prog, JtRing::kWarps, 2);
}
// fwd: 0 = depth-stack sweep, 1 = forward projection (applyJT), 2 = forward projection of applyDJT (level records
// [A | dA] in tensor memory, branch records [R|p | dV] and the free-flyer root's [R|p | v_b] in shared memory)
inline int plan_jt_ring_w(DevModel* m, const HostTables& tb, bool use_tmem, std::vector<int>* prog, int W, int fwd) {
  const int nj = m->njoints;
  const int D = m->max_depth_internal;
  if (!use_tmem || nj < 2) return 0;
  const int cap = B200Limits::kTmemCols / ((W + 3) / 4) / 2;
  std::vector<int> lvl_t(D + 1, -1), lvl_s(D + 1, -1);
  int t_used = 0, s_used = 0;
  if (fwd) {
    // forward projection: A column of level L at tensor-memory offset 6 L (compile-time in the kernel), branch R|p
    // records behind; a free-flyer only as a root joint (level 0), its R|p in shared memory
    if (m->max_depth > kJtMaxDepth) return 0;
    for (int i = 1; i < nj; ++i)
      if (m->joint[i].type == JT_FF && m->joint[i].parent != 0) return 0;
    const int lw = fwd == 2 ? 12 : 6;  // level record width
    t_used = lw * D;
    bool any_ff = false;
    for (int i = 1; i < nj; ++i) {
      DevJoint& j = m->joint[i];
      j.offA = 0; j.boff = 0;
      if (j.type == JT_FF) { any_ff = true; continue; }
      if (j.nchild > 0) j.offA = lw * (j.depth - 1);
    }
    if (any_ff) s_used = fwd == 2 ? 18 : 12;  // a root free-flyer's record sits at shared-memory offset 0
    for (int i = 1; i < nj; ++i) {
      DevJoint& j = m->joint[i];
      if (j.type == JT_FF || j.nchild < 2) continue;
      if (fwd == 2) { j.boff = s_used; s_used += 18; }   // [R|p | dV] in shared memory
      else { j.boff = t_used; t_used += 12; }
    }
  } else
  for (int i = 1; i < nj; ++i) {
    DevJoint& j = m->joint[i];
    j.offA = 0; j.boff = 0;
    if (j.nchild == 0) continue;
    if (j.type != JT_FF) {
      if (lvl_t[j.depth] < 0) { lvl_t[j.depth] = t_used; t_used += 12; }
      j.offA = lvl_t[j.depth];
      if (j.nchild >= 2) { j.boff = t_used; t_used += 12; }
    } else {
      if (lvl_s[j.depth] < 0) { lvl_s[j.depth] = s_used; s_used += 18; }
      j.offA = lvl_s[j.depth];
    }
  }
  if (t_used > cap) return 0;
  // ring program
  std::vector<int> pieces, copies;
  const int gmax = 4;  // RBD_JT_GMAX
  for (int i = 1; i < nj; ++i)
    for (int o = m->joint[i].out_begin; o < m->joint[i].out_end;) {
      const int cnt = std::min(gmax, m->joint[i].out_end - o);
      const int c0 = (int)copies.size() / 2;
      for (int j = 0; j < cnt;) {
        int r = 1;
        while (j + r < cnt && tb.sorted_k[o + j + r] == tb.sorted_k[o + j] + r) ++r;
        copies.push_back(6 * tb.sorted_k[o + j] * 256);                                   // source offset in the tile's block
        copies.push_back((j * JtRing::kSlotBytes) | ((r * JtRing::kSlotBytes) << 16));    // stage offset | bytes
        j += r;
      }
      pieces.push_back(o | (cnt << 16));
      pieces.push_back(cnt * JtRing::kSlotBytes);
      pieces.push_back(c0);
      pieces.push_back((int)copies.size() / 2 - c0);
      o += cnt;
    }
  const int np = (int)pieces.size() / 4, nc = (int)copies.size() / 2;
  const int off_opl = (4 + 4 * np + 2 * nc + 1) / 2 * 2;
  const int off_anc = off_opl + 6 * m->n_out;
  prog->assign((size_t)off_anc + (fwd ? (size_t)nj * kJtMaxDepth : 0), 0);
  (*prog)[0] = np; (*prog)[1] = nc; (*prog)[2] = off_opl; (*prog)[3] = fwd ? off_anc : 0;
  if (fwd)
    for (int i = 1; i < nj; ++i)
      for (int a = i; a != 0; a = m->joint[a].parent) (*prog)[(size_t)off_anc + (size_t)i * kJtMaxDepth + (m->joint[a].depth - 1)] = a;
  if (np) memcpy(prog->data() + 4, pieces.data(), pieces.size() * sizeof(int));
  if (nc) memcpy(prog->data() + 4 + 4 * np, copies.data(), copies.size() * sizeof(int));
  for (int o = 0; o < m->n_out; ++o)
    memcpy(prog->data() + off_opl + 6 * (size_t)o, &tb.sorted_pl[12 * (size_t)o + 9], 3 * sizeof(double));
  m->prog_len = (int)prog->size();
  m->prog_words_off = (int)((dev_model_bytes(*m) + 15) / 16 * 2);
  EvalPlan pl{};
  pl.smem_doubles = s_used; pl.tmem_doubles = t_used; pl.warps = W; pl.blocks_per_sm = 1;
  pl.tmem_cols = t_used > 0 ? pow2_cols(2 * t_used * ((W + 3) / 4)) : 0;
  pl.jt_fwd = fwd;
  m->plan = pl;
  const size_t fixed = eval_model_bytes(*m) + (size_t)W * JtRing::kMaxStages * 8 + (size_t)W * 32 * s_used * 8 +
                       B200Limits::kSmemReservedPerBlock;
  if (fixed >= (size_t)B200Limits::kSmemBytes) return 0;
  int nst = (int)(((size_t)B200Limits::kSmemBytes - fixed) / ((size_t)W * JtRing::kStageBytes));
  nst = std::min(nst, (int)JtRing::kMaxStages);
  return nst >= JtRing::kMinStages ? nst : 0;
}
// the blob the ring kernel stages: planned DevModel prefix + ring program
inline std::vector<unsigned long long> build_jt_blob(const DevModel* m, const std::vector<int>& prog) {
  std::vector<unsigned long long> blob(eval_model_bytes(*m) / 8 - 2, 0ull);
  memcpy(blob.data(), m, dev_model_bytes(*m));
  if (!prog.empty()) memcpy(blob.data() + m->prog_words_off, prog.data(), prog.size() * sizeof(int));
  return blob;
}

// ---------------------------------------------------------------------------------------------------------------
// Storage plan of the fused forward-dynamics sweep (aba2_instance, rbd_dyn_body.h): one record per DEPTH LEVEL that
// holds a joint with children — 15 doubles (one child) or 27 + 18 (>= 2 children: accumulator + restart record), the
// maximum over the joints at that depth.  Levels are placed whole, largest first, into the thread's share of shared
// memory and, when that is full, of tensor memory; the plan takes the largest block (<= max_warps warps, one block per
// SM) for which every level fits.  Fills m->plan (warps, smem_doubles, tmem_doubles, tmem_cols), joint[*].offA (level
// record) and joint[*].boff (restart record), bit RBD_BOFF_TMEM = tensor memory.  Returns false when nothing fits (very
// deep trees): the caller keeps the global-memory sweep (aba_instance).
inline bool plan_levels(DevModel* m, bool use_tmem, int size_one_child, int size_branch, int restart_off, int max_warps,
                        int forced_warps, size_t model_bytes);
inline bool plan_aba2(DevModel* m, bool use_tmem, int max_warps = B200Limits::kMaxWarpsPerBlock, int forced_warps = 0) {
  return plan_levels(m, use_tmem, 15, 27 + 18, 27, max_warps, forced_warps, aba2_model_bytes(*m));
}
// ... and of the M^-1 sweep (minv_instance): one child 9 (own inertia), >= 2 children I^A accumulator (21) + restart R | p
// (12) in phase 1, the block's accelerations (6 RBD_MINV_CB) in phase 2; DevJoint::slot[0] <- end of the joint's subtree
inline bool plan_minv(DevModel* m, bool use_tmem, int max_warps = B200Limits::kMaxWarpsPerBlock, int forced_warps = 0) {
  const int nj = m->njoints;
  for (int i = nj - 1; i >= 1; --i) m->joint[i].slot[0] = i + 1;
  for (int i = nj - 1; i >= 1; --i) {  // pre-order: a joint's subtree ends where its last child's subtree ends
    const int p = m->joint[i].parent;
    if (p > 0) m->joint[p].slot[0] = std::max(m->joint[p].slot[0], m->joint[i].slot[0]);
  }
  return plan_levels(m, use_tmem, 9, std::max(21 + 12, 6 * RBD_MINV_CB), 21, max_warps, forced_warps, aba2_model_bytes(*m));
}
// ... and of the inverse-dynamics derivatives (drnea_instance, rbd_deriv_body.h): record sizes depend on the joint kind
// (free-flyer root: R | p instead of J; a 1-dof joint with >= 2 children keeps R | p for the restart behind its accumulator)
inline bool plan_levels_sized(DevModel* m, bool use_tmem, const std::vector<int>& jsize, const std::vector<int>& jrestart,
                              int max_warps, int forced_warps, size_t model_bytes, int extra_smem = 0);
static constexpr int kDrneaSlots = 3;  // == RBD_DR_SLOTS (rbd_deriv_body.h): input ring of 3 doubles per slot and instance,
                                       // behind the level records in shared memory (plan.fbase = its offset, plan.in_slots)
inline bool plan_drnea(DevModel* m, bool use_tmem, int max_warps = B200Limits::kMaxWarpsPerBlock, int forced_warps = 0) {
  std::vector<int> jsize(m->njoints, 0), jrst(m->njoints, 0);
  for (int i = 1; i < m->njoints; ++i)
    if (m->joint[i].nchild > 0) jsize[i] = m->joint[i].type == JT_FF ? 24 : 18;  // RBD_DR_HOT_FF / RBD_DR_HOT
  if (!plan_levels_sized(m, use_tmem, jsize, jrst, max_warps, forced_warps, aba2_model_bytes(*m), 3 * kDrneaSlots + 9)) return false;
  // cold records (global scratch, per thread): one per depth level, as wide as the widest joint of the level needs
  const int D = m->max_depth_internal;
  std::vector<int> csize(D + 2, 0), coff(D + 2, 0);
  for (int i = 1; i < m->njoints; ++i)
    if (m->joint[i].nchild > 0) csize[m->joint[i].depth] = std::max(csize[m->joint[i].depth], m->joint[i].nchild >= 2 ? 42 : 9);
  for (int d = 1; d <= D; ++d) coff[d + 1] = coff[d] + csize[d];
  for (int i = 1; i < m->njoints; ++i) m->joint[i].boff = m->joint[i].nchild > 0 ? coff[m->joint[i].depth] : 0;
  m->slot_total[0] = coff[D + 1];  // doubles of scratch per instance in flight
  // support-only output format: row base / ancestor dofs / row length of every joint (RBD_DR_ROWBASE / _NANC / _ROWLEN)
  {
    const int nj = m->njoints;
    std::vector<int> nvsub(nj, 0), nanc(nj, 0);
    for (int i = nj - 1; i >= 1; --i) { nvsub[i] += m->joint[i].nv; nvsub[m->joint[i].parent] += nvsub[i]; }
    int base = 0;
    for (int i = 1; i < nj; ++i) {
      DevJoint& j = m->joint[i];
      nanc[i] = j.parent ? nanc[j.parent] + m->joint[j.parent].nv : 0;
      j.slot[0] = base; j.slot[1] = nanc[i]; j.out_begin = nanc[i] + nvsub[i];
      base += j.nv * j.out_begin;
    }
    m->slot_total[1] = base;  // entries of the support pattern
  }
  return true;
}
// Support pattern of dtau_dq / dtau_dv, compressed by rows (the order of the fields the SPARSE sweep writes): row of dof r =
// the dofs of the ancestors of r's joint (root first), then the dofs [idx_v, idx_v + nv_subtree) of the joint's subtree.
// row_ptr [nv + 1], col_idx [nnz] (either may be null); returns nnz.
inline int drnea_support_pattern(const DevModel& m, int* row_ptr, int* col_idx) {
  const int nj = m.njoints;
  std::vector<int> nvsub(nj, 0);
  for (int i = nj - 1; i >= 1; --i) { nvsub[i] += m.joint[i].nv; nvsub[m.joint[i].parent] += nvsub[i]; }
  int nnz = 0;
  for (int i = 1; i < nj; ++i) {
    const DevJoint& j = m.joint[i];
    std::vector<int> chain;
    for (int a = j.parent; a != 0; a = m.joint[a].parent) chain.insert(chain.begin(), a);
    for (int e = 0; e < j.nv; ++e) {
      if (row_ptr) row_ptr[j.idx_v + e] = nnz;
      for (int a : chain)
        for (int k = 0; k < m.joint[a].nv; ++k) { if (col_idx) col_idx[nnz] = m.joint[a].idx_v + k; ++nnz; }
      for (int c = j.idx_v; c < j.idx_v + nvsub[i]; ++c) { if (col_idx) col_idx[nnz] = c; ++nnz; }
    }
  }
  if (row_ptr) row_ptr[m.nv] = nnz;
  return nnz;
}
inline bool plan_levels(DevModel* m, bool use_tmem, int size_one_child, int size_branch, int restart_off, int max_warps,
                        int forced_warps, size_t model_bytes) {
  std::vector<int> jsize(m->njoints, 0), jrst(m->njoints, restart_off);
  for (int i = 1; i < m->njoints; ++i)
    if (m->joint[i].nchild > 0) jsize[i] = m->joint[i].nchild >= 2 ? size_branch : size_one_child;
  return plan_levels_sized(m, use_tmem, jsize, jrst, max_warps, forced_warps, model_bytes);
}
inline bool plan_levels_sized(DevModel* m, bool use_tmem, const std::vector<int>& jsize, const std::vector<int>& jrestart,
                              int max_warps, int forced_warps, size_t model_bytes, int extra_smem) {
  const int nj = m->njoints;
  const int D = m->max_depth_internal;
  std::vector<int> size(D + 1, 0);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& j = m->joint[i];
    if (j.nchild <= 0) continue;
    if (j.depth > D) return false;
    size[j.depth] = std::max(size[j.depth], jsize[i]);
  }
  std::vector<int> order;
  for (int d = 1; d <= D; ++d)
    if (size[d] > 0) order.push_back(d);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return size[a] > size[b]; });
  for (int W = max_warps; W >= 1; --W) {
    if (forced_warps > 0 && W != forced_warps) continue;
    if (model_bytes + B200Limits::kSmemReservedPerBlock >= (size_t)B200Limits::kSmemBytes) return false;
    const int cap_s = (int)(((size_t)B200Limits::kSmemBytes - model_bytes - B200Limits::kSmemReservedPerBlock) / ((size_t)W * 32 * 8)) - extra_smem;
    const int cap_t = use_tmem ? (B200Limits::kTmemCols / ((W + 3) / 4)) / 2 : 0;
    std::vector<int> off(D + 1, 0);
    int s_used = 0, t_used = 0;
    bool ok = true;
    for (int d : order) {
      if (s_used + size[d] <= cap_s) { off[d] = s_used; s_used += size[d]; }
      else if (t_used + size[d] <= cap_t) { off[d] = t_used | RBD_BOFF_TMEM; t_used += size[d]; }
      else { ok = false; break; }
    }
    if (!ok) continue;
    EvalPlan pl{};
    pl.smem_doubles = s_used + extra_smem; pl.tmem_doubles = t_used; pl.warps = W; pl.blocks_per_sm = 1;
    pl.fbase = s_used; pl.in_slots = extra_smem > 0 ? 1 : 0;  // (input ring + staging behind the level records: drnea only)
    pl.tmem_cols = t_used > 0 ? pow2_cols(2 * t_used * ((W + 3) / 4)) : 0;
    m->plan = pl;
    for (int i = 1; i < nj; ++i) {
      DevJoint& j = m->joint[i];
      j.offA = 0; j.boff = 0;
      if (j.nchild <= 0) continue;
      j.offA = off[j.depth];
      if (j.nchild >= 2) j.boff = (off[j.depth] & RBD_BOFF_TMEM) | ((off[j.depth] & RBD_BOFF_MASK) + jrestart[i]);
    }
    return true;
  }
  return false;
}

// shared-memory doubles per instance of the mapping kernels
// (branch slots; the kernels add their input ring behind them: kMapSlots (apply) / 2 kMapSlots (applyJ) per instance)
static constexpr int kMapSlots = 4;  // == RBD_MAP_SLOTS (rbd_sweeps_body.h)
inline int smem_apply(const DevModel& m) { return m.nbranch * RBD_BRANCH_FK; }
inline int smem_applyJ(const DevModel& m) { return m.nbranch * RBD_BRANCH_J; }
inline int smem_applyJT(const DevModel& m) { return m.slot_total[1] + m.nbranch * RBD_BRANCH_FK; }

}  // namespace rbd
