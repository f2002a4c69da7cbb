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
  std::vector<int> sorted_k, outk_parent;
  std::vector<double> sorted_pl, outk_pl;
};

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
    const double* pl = m->joint[i].pl;
    m->joint[i].pl_rot_identity = (pl[0] == 1.0 && pl[1] == 0.0 && pl[2] == 0.0 && pl[3] == 0.0 && pl[4] == 1.0 &&
                                   pl[5] == 0.0 && pl[6] == 0.0 && pl[7] == 0.0 && pl[8] == 1.0) ? 1 : 0;
    if (m->joint[i].nchild >= 2) m->joint[i].bidx = nb++;
  }
  m->nbranch = nb;
  // mapped outputs grouped by parent joint (stable: keeps the reference's output order within a joint)
  tb->sorted_k.clear(); tb->sorted_pl.clear(); tb->outk_parent.assign(d->n_out, 0); tb->outk_pl.assign(12 * (size_t)d->n_out, 0.0);
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
        tb->sorted_pl.insert(tb->sorted_pl.end(), &tb->outk_pl[12 * (size_t)k], &tb->outk_pl[12 * (size_t)k] + 12);
      }
    m->joint[j].out_end = (int)tb->sorted_k.size();
  }
  return RBD_OK;
}

// shared-memory doubles per instance of each kernel
inline int smem_eval(const DevModel& m, bool do_m, bool do_tau) {
  if (!do_m && !do_tau) return m.nbranch * RBD_BRANCH_FK;
  return m.slot_total[do_m ? 0 : 1] + m.nbranch * (do_tau ? RBD_BRANCH_EVAL : RBD_BRANCH_FK);
}
inline int smem_apply(const DevModel& m) { return m.nbranch * RBD_BRANCH_FK; }
inline int smem_applyJ(const DevModel& m) { return m.nbranch * RBD_BRANCH_J; }
inline int smem_applyJT(const DevModel& m) { return m.slot_total[1] + m.nbranch * RBD_BRANCH_FK; }

}  // namespace rbd
