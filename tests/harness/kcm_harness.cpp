// kcm_harness.cpp — a compiled C++ caller of the C ABI, the way the patched reference component links it
// (INTEGRATION.md): no Python, no torch, only include/rbd_b200.h + librbd_b200.so.  It replays the call order of
// KinematicChainMapping for one robot instance held in SOFA's own memory layouts —
//   init: setModel -> rbd_model_create / rbd_workspace_create          (KinematicChainMapping.inl:304-311)
//   step: apply -> applyJ -> applyJT -> applyJT(MatrixDeriv)           (KinematicChainMapping.inl:338-639)
// — and prints every output as text; tests/test_gpu_harness.py compares them with the oracle.
// Model: a planar 2R arm built by hand (what URDFModelLoader would hand over for a two-link URDF).
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "../../include/rbd_b200.h"

struct Rigid3Coord { double p[3]; double q[4]; };  // sofa RigidCoord<3,double>: center, orientation (x y z w)
struct Rigid3Deriv { double v[3]; double w[3]; };  // sofa RigidDeriv<3,double>: vCenter, vOrientation
static_assert(sizeof(Rigid3Coord) == 7 * sizeof(double) && sizeof(Rigid3Deriv) == 6 * sizeof(double), "SOFA layouts");

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != RBD_OK) { std::printf("ERROR %d %s\n", rc_, rbd_last_error()); return 2; } \
  } while (0)

int main(int argc, char** argv) {
  const double q1 = argc > 1 ? std::atof(argv[1]) : 0.3, q2 = argc > 2 ? std::atof(argv[2]) : -0.7;
  // universe + 2 revolute-Z joints; link length 0.5; each link: m = 1, CoM at (0.25, 0, 0), I = diag(0.01, 0.02, 0.02)
  const int nj = 3;
  int32_t parents[nj] = {0, 0, 1}, jtype[nj] = {RBD_JOINT_UNIVERSE, RBD_JOINT_RZ, RBD_JOINT_RZ};
  int32_t idx_q[nj] = {0, 0, 1}, idx_v[nj] = {0, 0, 1};
  double axis[3 * nj] = {0, 0, 0, 0, 0, 1, 0, 0, 1};
  double I12[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  double placement[12 * nj];
  for (int i = 0; i < nj; ++i) for (int k = 0; k < 12; ++k) placement[12 * i + k] = I12[k];
  placement[12 * 2 + 9] = 0.5;  // joint 2 sits at the end of link 1
  double inertia[10 * nj] = {0};
  for (int i = 1; i < nj; ++i) {
    double* y = inertia + 10 * i;
    y[0] = 1.0; y[1] = 0.25; y[4] = 0.01; y[6] = 0.02; y[9] = 0.02;
  }
  // frames: universe, joint1, link1, joint2, link2, tool (fixed at the tip), then the loader's CoM frames (one per joint)
  const int nf0 = 6, nf = nf0 + nj;
  int32_t fparent[nf] = {0, 1, 1, 2, 2, 2, 0, 1, 2};
  double fplace[12 * nf];
  for (int f = 0; f < nf; ++f) for (int k = 0; k < 12; ++k) fplace[12 * f + k] = I12[k];
  fplace[12 * 5 + 9] = 0.5;                                  // tool
  for (int j = 0; j < nj; ++j) for (int k = 0; k < 3; ++k) fplace[12 * (nf0 + j) + 9 + k] = inertia[10 * j + 1 + k];
  std::vector<int32_t> out_frames;
  for (int j = 0; j < nj; ++j) out_frames.push_back(nf0 + j);  // bodyCoMFrames (KinematicChainMapping.inl:382-386)
  for (int f = 0; f < nf0; ++f) out_frames.push_back(f);       // extraFrames   (KinematicChainMapping.inl:389-395)

  rbd_model_desc d{};
  d.njoints = nj; d.nq = 2; d.nv = 2; d.parents = parents; d.jtype = jtype; d.axis = axis; d.placement = placement;
  d.idx_q = idx_q; d.idx_v = idx_v; d.inertia = inertia; d.nframes = nf; d.frame_parent = fparent;
  d.frame_placement = fplace; d.gravity[0] = 0; d.gravity[1] = 0; d.gravity[2] = -9.81;
  d.n_out = (int32_t)out_frames.size(); d.out_frames = out_frames.data();

  rbd_model* model = nullptr; rbd_workspace* ws = nullptr;
  CHECK(rbd_model_create(&d, &model));
  CHECK(rbd_workspace_create(model, 0, 1, nullptr, &ws));
  rbd_model_info info;
  CHECK(rbd_model_get_info(model, &info));
  const int n_out = info.n_out;

  // SOFA state vectors of the one robot
  std::vector<double> in_q = {q1, q2}, in_dq = {0.4, -1.1};                 // VecCoord<Vec1>, VecDeriv<Vec1>
  std::vector<Rigid3Coord> out_pos(n_out); std::vector<Rigid3Deriv> out_vel(n_out), in_force(n_out);
  for (int k = 0; k < n_out; ++k) for (int e = 0; e < 3; ++e) { in_force[k].v[e] = 0.1 * (k + 1) * (e + 1); in_force[k].w[e] = -0.05 * (k + e); }
  std::vector<double> out_tau = {1.0, 2.0};                                   // applyJT accumulates into it

  // applyJ before apply must be refused (the reference relies on the Jacobians cached by apply)
  std::printf("noapply %d\n", rbd_applyJ_host(ws, 1, in_dq.data(), nullptr, reinterpret_cast<double*>(out_vel.data())));
  CHECK(rbd_apply_host(ws, 1, in_q.data(), nullptr, reinterpret_cast<double*>(out_pos.data())));
  CHECK(rbd_applyJ_host(ws, 1, in_dq.data(), nullptr, reinterpret_cast<double*>(out_vel.data())));
  CHECK(rbd_applyJT_host(ws, 1, reinterpret_cast<const double*>(in_force.data()), out_tau.data(), nullptr));
  // one constraint row touching two outputs (MatrixDeriv), plus an all-zero row that must be dropped
  int32_t row_ptr[3] = {0, 2, 3}, row_inst[2] = {0, 0}, col[3] = {n_out - 1, 1, 2};
  double val[18] = {1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0.5, 0, 0, 0, 0, 0, 0};
  double rows[4]; int32_t keep[2];
  CHECK(rbd_applyJT_rows_host(ws, 2, row_ptr, row_inst, col, val, rows, keep));
  // Mass / ForceField quantities
  double v[2] = {0.4, -1.1}, a[2] = {0.2, 0.3}, oMi[24], J[12], M[3], tau[2];
  CHECK(rbd_eval_host(ws, 1, in_q.data(), v, a, oMi, J, M, tau));
  // ... and their matrix-free products on SOFA's own vectors (Mass::addMDx, ForceField::addForce), accumulated in place
  double res[2] = {10.0, 20.0}, fbias[2] = {0.0, 0.0}, fgrav[2] = {1.0, -1.0};
  CHECK(rbd_addMDx_host(ws, 1, in_q.data(), a, 2.0, res));          // res += 2 M(q) a
  CHECK(rbd_addForce_host(ws, 1, in_q.data(), v, fbias));           // f -= C(q, v) v + g(q)
  CHECK(rbd_addForce_host(ws, 1, in_q.data(), nullptr, fgrav));     // f -= g(q)

  for (int k = 0; k < n_out; ++k)
    std::printf("pos %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k, out_pos[k].p[0], out_pos[k].p[1], out_pos[k].p[2],
                out_pos[k].q[0], out_pos[k].q[1], out_pos[k].q[2], out_pos[k].q[3]);
  for (int k = 0; k < n_out; ++k)
    std::printf("vel %d %.17g %.17g %.17g %.17g %.17g %.17g\n", k, out_vel[k].v[0], out_vel[k].v[1], out_vel[k].v[2],
                out_vel[k].w[0], out_vel[k].w[1], out_vel[k].w[2]);
  std::printf("tauJT %.17g %.17g\n", out_tau[0], out_tau[1]);
  std::printf("rows %d %.17g %.17g %d %.17g %.17g\n", keep[0], rows[0], rows[1], keep[1], rows[2], rows[3]);
  std::printf("M %.17g %.17g %.17g\n", M[0], M[1], M[2]);
  std::printf("tau %.17g %.17g\n", tau[0], tau[1]);
  std::printf("mdx %.17g %.17g\n", res[0], res[1]);
  std::printf("fbias %.17g %.17g\n", fbias[0], fbias[1]);
  std::printf("fgrav %.17g %.17g\n", fgrav[0], fgrav[1]);
  std::printf("launches %lld\n", (long long)rbd_workspace_launch_count(ws));
  rbd_workspace_destroy(ws);
  rbd_model_destroy(model);
  std::printf("OK\n");
  return 0;
}
