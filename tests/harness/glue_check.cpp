// glue_check.cpp — compiles integration/rbd_backend_glue.h (the SOFA-free part of the patched KinematicChainMapping)
// standalone against include/rbd_b200.h, with a minimal stand-in for the few members of pinocchio::Model the glue
// reads, flattens a two-joint arm + free-flyer variant, creates the model through the C ABI and prints what it got.
// TEST INFRASTRUCTURE (tests/test_integration_patch.py).
#include <array>
#include <cstdio>
#include <string>
#include <vector>

#include "../../integration/rbd_backend_glue.h"

namespace mock {  // the shape of pinocchio's API, nothing of its code
struct Mat3 { double m[9]; double operator()(int r, int c) const { return m[3 * r + c]; } };
struct SE3 {
  Mat3 R{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; std::array<double, 3> p{{0, 0, 0}};
  const Mat3& rotation() const { return R; }
  const std::array<double, 3>& translation() const { return p; }
};
struct Sym3 { double d[6]; const double* data() const { return d; } };
struct Inertia {
  double m = 0; std::array<double, 3> c{{0, 0, 0}}; Sym3 I{{0, 0, 0, 0, 0, 0}};
  double mass() const { return m; }
  const std::array<double, 3>& lever() const { return c; }
  const Sym3& inertia() const { return I; }
};
struct Joint { std::string name; std::array<double, 3> axis{{0, 0, 0}}; std::string shortname() const { return name; } };
struct Frame { int parentJoint = 0; SE3 placement; };
struct Motion { std::array<double, 3> l{{0, 0, -9.81}}; const std::array<double, 3>& linear() const { return l; } };
struct Model {
  int njoints = 0, nq = 0, nv = 0, nframes = 0;
  std::vector<int> parents, idx_qs, idx_vs;
  std::vector<Joint> joints;
  std::vector<SE3> jointPlacements;
  std::vector<Inertia> inertias;
  std::vector<Frame> frames;
  Motion gravity;
};
}  // namespace mock

static mock::Model make_model(bool free_flyer, bool unsupported) {
  mock::Model M;
  auto add = [&](int parent, const std::string& kind, int nq, int nv, double px) {
    M.parents.push_back(parent); M.idx_qs.push_back(M.nq); M.idx_vs.push_back(M.nv);
    mock::Joint j; j.name = kind; if (kind == "JointModelRevoluteUnaligned") j.axis = {{0.6, 0.0, 0.8}};
    M.joints.push_back(j);
    mock::SE3 pl; pl.p[0] = px; M.jointPlacements.push_back(pl);
    mock::Inertia Y; if (parent >= 0 && kind != "universe") { Y.m = 1.5; Y.c = {{0.1, 0.0, 0.05}}; Y.I = {{0.01, 0.001, 0.02, 0.002, 0.003, 0.03}}; }
    M.inertias.push_back(Y);
    M.nq += nq; M.nv += nv; M.njoints++;
  };
  add(0, "universe", 0, 0, 0.0);
  int base = 0;
  if (free_flyer) { add(0, "JointModelFreeFlyer", 7, 6, 0.0); base = 1; }
  add(base, "JointModelRZ", 1, 1, 0.0);
  add(base + 1, unsupported ? "JointModelSpherical" : "JointModelRevoluteUnaligned", 1, 1, 0.5);
  for (int f = 0; f < M.njoints + 2; ++f) {
    mock::Frame fr; fr.parentJoint = f < M.njoints ? f : M.njoints - 1; fr.placement.p[2] = 0.01 * f; M.frames.push_back(fr);
  }
  M.nframes = (int)M.frames.size();
  return M;
}

int main() {
  auto axisOf = [](const mock::Joint& j, double* a) { for (int k = 0; k < 3; ++k) a[k] = j.axis[k]; };
  for (int ff = 0; ff < 2; ++ff) {
    mock::Model M = make_model(ff != 0, false);
    std::vector<int> com, extra;
    for (int j = 0; j < M.njoints; ++j) com.push_back(j);
    extra.push_back(M.nframes - 1); extra.push_back(M.nframes - 2);
    rbd_glue::Descriptor D; std::string why;
    if (!rbd_glue::fillDescriptor(M, com, extra, axisOf, &D, &why)) { std::printf("FILL_FAILED %s\n", why.c_str()); return 1; }
    std::printf("DESC ff=%d njoints=%d nq=%d nv=%d n_out=%d jtype=", ff, D.d.njoints, D.d.nq, D.d.nv, D.d.n_out);
    for (int t : D.jtype) std::printf("%d,", t);
    std::printf(" axis_last=%.1f,%.1f,%.1f out=", D.axis[3 * (D.d.njoints - 1)], D.axis[3 * (D.d.njoints - 1) + 1], D.axis[3 * (D.d.njoints - 1) + 2]);
    for (int f : D.out_frames) std::printf("%d,", f);
    std::printf(" pl_last_px=%.2f inertia1=%.3f,%.3f grav_z=%.2f\n", D.placement[12 * (D.d.njoints - 1) + 9], D.inertia[10 * (D.d.njoints - 1)],
                D.inertia[10 * (D.d.njoints - 1) + 9], D.d.gravity[2]);
    rbd_model* model = nullptr;
    int rc = rbd_model_create(&D.d, &model);
    rbd_model_info I{};
    if (rc == RBD_OK) rbd_model_get_info(model, &I);
    std::printf("MODEL rc=%d has_ff=%d nq_joints=%d nv_joints=%d n_out=%d mass_support=%d\n", rc, I.has_freeflyer_root, I.nq_joints, I.nv_joints,
                I.n_out, I.mass_support_fields);
    if (model) rbd_model_destroy(model);
    rbd_glue::Backend be;
    rc = be.create(D.d, 0, 4);
    std::printf("BACKEND rc=%d ws=%s model=%s\n", rc, be.ws ? "set" : "null", be.model ? "set" : "null");
    if (rc == RBD_OK) {  // a device is present: one apply + one constraint row through the glue's CSR helper
      const rbd_model_info& J = be.info;
      std::vector<double> q((size_t)J.nq_joints, 0.3), root = {0, 0, 0, 0, 0, 0, 1}, out((size_t)7 * J.n_out, 0.0);
      rc = rbd_apply_host(be.ws, 1, q.data(), J.has_freeflyer_root ? root.data() : nullptr, out.data());
      rbd_glue::RowsCSR R;
      const double w[6] = {1, 0, 0, 0, 0, 1};
      R.addEntry(J.n_out - 1, w); R.endRow(7);
      int rc2 = R.run(be.ws, J.nv);
      std::printf("APPLY rc=%d rows_rc=%d kept=%d row0=%.6f\n", rc, rc2, (int)R.kept(0), R.row(0, J.nv)[J.nv - 1]);
    }
  }
  mock::Model bad = make_model(false, true);
  std::vector<int> com, extra;
  rbd_glue::Descriptor D; std::string why;
  const bool ok = rbd_glue::fillDescriptor(bad, com, extra, axisOf, &D, &why);
  std::printf("UNSUPPORTED ok=%d why=%s\n", (int)ok, why.c_str());
  std::printf("CODES %d %d %d\n", rbd_glue::rbdJointCode("JointModelRX"), rbd_glue::rbdJointCode("JointModelFreeFlyer"),
              rbd_glue::rbdJointCode("JointModelPlanar"));
  return 0;
}
