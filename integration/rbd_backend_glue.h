// rbd_backend_glue.h — SOFA-free helpers between the plugin's KinematicChainMapping and the C ABI of the B200 back end
// (include/rbd_b200.h).  Everything the patched component needs that is not SOFA code lives here, so that it compiles
// — and is compiled, tests/test_integration_patch.py — without SOFA or pinocchio: the model flattening is a template
// over the few members of pinocchio::Model it reads (reference call sites in brackets):
//   Model::njoints, nq, nv, nframes, parents, idx_qs, idx_vs, joints[i].shortname(), jointPlacements[i], inertias[i],
//   frames[f].parentJoint / .placement, gravity.linear()
//     [URDFModelLoader.cpp:126-184 builds them; KinematicChainMapping.inl:304-317 receives them]
// integration/apply_backend_patch.py splices the calls into KinematicChainMapping.{h,inl}.
#pragma once
#include <stdint.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "rbd_b200.h"

namespace rbd_glue {

// pinocchio JointModel::shortname() -> RBD_JOINT_* (-1: not supported by the back end).  The unaligned kinds also need
// their axis, which the caller copies from the joint model (see axisOf in KinematicChainMapping.backend.inl).
inline int rbdJointCode(const std::string& shortname) {
  static const struct { const char* name; int code; } table[] = {
      {"JointModelRX", RBD_JOINT_RX}, {"JointModelRY", RBD_JOINT_RY}, {"JointModelRZ", RBD_JOINT_RZ},
      {"JointModelRevoluteUnaligned", RBD_JOINT_RU},
      {"JointModelPX", RBD_JOINT_PX}, {"JointModelPY", RBD_JOINT_PY}, {"JointModelPZ", RBD_JOINT_PZ},
      {"JointModelPrismaticUnaligned", RBD_JOINT_PU},
      {"JointModelRUBX", RBD_JOINT_RUBX}, {"JointModelRUBY", RBD_JOINT_RUBY}, {"JointModelRUBZ", RBD_JOINT_RUBZ},
      {"JointModelRevoluteUnboundedUnaligned", RBD_JOINT_RUBU},
      {"JointModelFreeFlyer", RBD_JOINT_FF}};
  for (const auto& e : table)
    if (shortname == e.name) return e.code;
  return -1;
}
inline bool rbdJointNeedsAxis(int code) { return code == RBD_JOINT_RU || code == RBD_JOINT_PU || code == RBD_JOINT_RUBU; }

// pinocchio::SE3 -> 12 doubles, R row-major then p (the descriptor's placement format)
template <class SE3>
inline void packSE3(const SE3& M, double* out12) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out12[3 * r + c] = M.rotation()(r, c);
  for (int k = 0; k < 3; ++k) out12[9 + k] = M.translation()[k];
}

// The flattened model: owns the arrays the rbd_model_desc points to.
struct Descriptor {
  std::vector<int32_t> parents, jtype, idx_q, idx_v, frame_parent, out_frames;
  std::vector<double> axis, placement, inertia, frame_placement;
  rbd_model_desc d{};
};

// Flattens a pinocchio::Model plus the two frame tables of the mapping (bodyCoMFrames then extraFrames: the order of the
// mapped outputs, KinematicChainMapping.inl:382-395).  axisOf(jointModel, double[3]) fills the axis of an unaligned
// joint.  Returns false and explains in *err when the model holds a joint kind the back end does not implement.
template <class Model, class FrameIndexVec, class AxisFn>
inline bool fillDescriptor(const Model& M, const FrameIndexVec& bodyCoMFrames, const FrameIndexVec& extraFrames,
                           AxisFn axisOf, Descriptor* D, std::string* err) {
  const int nj = (int)M.njoints, nf = (int)M.nframes;
  D->parents.assign(nj, 0); D->jtype.assign(nj, RBD_JOINT_UNIVERSE); D->idx_q.assign(nj, 0); D->idx_v.assign(nj, 0);
  D->axis.assign(3 * (size_t)nj, 0.0); D->placement.assign(12 * (size_t)nj, 0.0); D->inertia.assign(10 * (size_t)nj, 0.0);
  for (int i = 0; i < nj; ++i) {
    D->parents[i] = (int32_t)M.parents[i];
    if (i > 0) {
      const int code = rbdJointCode(M.joints[i].shortname());
      if (code < 0) {
        if (err) *err = "joint " + std::to_string(i) + " (" + M.joints[i].shortname() + ") is not supported by rbd_b200";
        return false;
      }
      D->jtype[i] = code;
      D->idx_q[i] = (int32_t)M.idx_qs[i]; D->idx_v[i] = (int32_t)M.idx_vs[i];
      if (rbdJointNeedsAxis(code)) axisOf(M.joints[i], &D->axis[3 * (size_t)i]);
    }
    packSE3(M.jointPlacements[i], &D->placement[12 * (size_t)i]);
    const auto& Y = M.inertias[i];
    double* y = &D->inertia[10 * (size_t)i];
    y[0] = Y.mass();
    for (int k = 0; k < 3; ++k) y[1 + k] = Y.lever()[k];
    for (int k = 0; k < 6; ++k) y[4 + k] = Y.inertia().data()[k];  // Symmetric3: xx, xy, yy, xz, yz, zz
  }
  D->frame_parent.assign(nf, 0); D->frame_placement.assign(12 * (size_t)nf, 0.0);
  for (int f = 0; f < nf; ++f) {
    D->frame_parent[f] = (int32_t)M.frames[f].parentJoint;
    packSE3(M.frames[f].placement, &D->frame_placement[12 * (size_t)f]);
  }
  D->out_frames.clear();
  for (const auto f : bodyCoMFrames) D->out_frames.push_back((int32_t)f);
  for (const auto f : extraFrames) D->out_frames.push_back((int32_t)f);
  rbd_model_desc& d = D->d;
  d.njoints = nj; d.nq = (int32_t)M.nq; d.nv = (int32_t)M.nv;
  d.parents = D->parents.data(); d.jtype = D->jtype.data(); d.axis = D->axis.data(); d.placement = D->placement.data();
  d.idx_q = D->idx_q.data(); d.idx_v = D->idx_v.data(); d.inertia = D->inertia.data();
  d.nframes = nf; d.frame_parent = D->frame_parent.data(); d.frame_placement = D->frame_placement.data();
  for (int k = 0; k < 3; ++k) d.gravity[k] = M.gravity.linear()[k];
  d.n_out = (int32_t)D->out_frames.size(); d.out_frames = D->out_frames.data();
  return true;
}

// One robot's handles (what replaces the pinocchio::Data member, KinematicChainMapping.h:121).
struct Backend {
  rbd_model* model = nullptr;
  rbd_workspace* ws = nullptr;
  rbd_model_info info{};
  Backend() = default;
  Backend(const Backend&) = delete;
  Backend& operator=(const Backend&) = delete;
  ~Backend() { reset(); }
  void reset() {
    if (ws) rbd_workspace_destroy(ws);
    if (model) rbd_model_destroy(model);
    ws = nullptr; model = nullptr;
  }
  // returns RBD_OK or the failing call's status (message: rbd_last_error())
  int create(const rbd_model_desc& d, int device, int64_t capacity) {
    reset();
    int rc = rbd_model_create(&d, &model);
    if (rc == RBD_OK) rc = rbd_model_get_info(model, &info);
    if (rc == RBD_OK) rc = rbd_workspace_create(model, device, capacity, nullptr, &ws);
    if (rc != RBD_OK) reset();
    return rc;
  }
};

// Many robots of one scene sharing one workspace: mappings built from the same
// pinocchio::Model join the group keyed by that model's address (the loader shares the shared_ptr with its mapping,
// URDFModelLoader.cpp:96,268; a scene-level component would hand the same pointer to every robot) and get a slot.
// The group's capacity is fixed when it is created (robots expected in the scene).
class SharedBatch {
 public:
  struct Group {
    Backend be;
    rbd_batch* batch = nullptr;
    int64_t capacity = 0, used = 0;
    ~Group() { if (batch) rbd_batch_destroy(batch); }
  };
  // joins (or creates) the group of `key`; *slot receives this robot's slot.  nullptr on failure (rbd_last_error()).
  static std::shared_ptr<Group> join(const void* key, const rbd_model_desc& d, int device, int64_t capacity, int64_t* slot) {
    std::lock_guard<std::mutex> lock(mutex());
    auto& reg = registry();
    std::shared_ptr<Group> g = reg[key].lock();
    if (!g) {
      g = std::make_shared<Group>();
      if (g->be.create(d, device, capacity) != RBD_OK) return nullptr;
      if (rbd_batch_create(g->be.ws, capacity, &g->batch) != RBD_OK) return nullptr;
      g->capacity = capacity;
      reg[key] = g;
    }
    if (g->used >= g->capacity) return nullptr;
    *slot = g->used++;
    return g;
  }

 private:
  static std::mutex& mutex() { static std::mutex m; return m; }
  static std::map<const void*, std::weak_ptr<Group>>& registry() { static std::map<const void*, std::weak_ptr<Group>> r; return r; }
};

// SOFA MatrixDeriv (compressed rows of Deriv values) flattened for rbd_applyJT_rows_host
// (KinematicChainMapping.inl:557-583 walks the same structure): one addRow per constraint line, one addEntry per column.
struct RowsCSR {
  std::vector<int32_t> row_ptr{0}, row_instance, col, row_index;
  std::vector<double> val, rows;
  std::vector<int32_t> keep;
  void clear() { row_ptr.assign(1, 0); row_instance.clear(); col.clear(); row_index.clear(); val.clear(); }
  void addEntry(int32_t mappedOutput, const double* wrench6) {
    col.push_back(mappedOutput);
    val.insert(val.end(), wrench6, wrench6 + 6);
  }
  void endRow(int32_t constraintIndex, int32_t instance = 0) {
    row_ptr.push_back((int32_t)col.size()); row_instance.push_back(instance); row_index.push_back(constraintIndex);
  }
  int64_t nrows() const { return (int64_t)row_index.size(); }
  // runs the back end; afterwards row(r) is the dense nv-wide torque row of line r and kept(r) tells whether the
  // reference would have emitted it (squared norm > kMinTorqueSqrd, KinematicChainMapping.inl:43-45,587,605)
  int run(rbd_workspace* ws, int nv) {
    rows.assign((size_t)nrows() * (size_t)nv, 0.0); keep.assign((size_t)nrows(), 0);
    if (nrows() == 0) return RBD_OK;
    return rbd_applyJT_rows_host(ws, nrows(), row_ptr.data(), row_instance.data(), col.data(), val.data(), rows.data(),
                                 keep.data());
  }
  const double* row(int64_t r, int nv) const { return rows.data() + (size_t)r * (size_t)nv; }
  bool kept(int64_t r) const { return keep[(size_t)r] != 0; }
};

}  // namespace rbd_glue
