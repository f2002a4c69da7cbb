// rbd_kernels.h — launch interface between the C ABI (rbd_capi.cu) and the kernels (rbd_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "rbd_dev_model.h"

namespace rbd {

template <class real>
struct EvalArgsT {
  long long n, tile;
  const real *q, *v, *a;  // with tau requested, v or a may be null (zeros): Mass / ForceField products
  real *oMi, *J, *M, *tau;
  real tau_scale = real(1);  // tau_acc != 0: tau <- tau + tau_scale * rnea(...)
  int tau_acc = 0;
  int gravity_on = 1;
};
using EvalArgs = EvalArgsT<double>;
struct MapArgs {
  long long n, tile;
  const double* joints;  // apply: q_joints | applyJ: dq_joints | applyJT: out_tau (accumulated)
  const double* root;    // apply: root pose (7) | applyJ: root velocity (6) | applyJT: out_root (accumulated) | nullptr
  const double* in;      // applyJT: wrenches (n_out*6)
  double* out;           // apply: n_out*7 | applyJ: n_out*6
  double* qcache;        // workspace copy of the assembled configuration (nq fields, tile 32)
  long long cache_inst0;
  // applyDJT on the ring kernel only: direction (as applyJ's inputs) and factor; joints / root are the accumulated outputs
  const double* dq_joints = nullptr;
  const double* dq_root = nullptr;
  double kfac = 1.0;
};
struct DjtArgs {  // applyDJT: geometric stiffness (optional extension; the reference's applyDJT is a no-op)
  long long n, tile;
  const double *dq_joints, *dq_root;  // direction (like applyJ's inputs)
  const double* in;                   // wrenches (n_out*6), like applyJT's input
  double *out_joints, *out_root;      // accumulated: += kfac * dtau
  double kfac;
  double* qcache;
  long long cache_inst0;
};
struct MinvArgs {  // inverse of the joint-space mass matrix: q -> packed upper triangle of M^-1
  long long n, tile;
  const double* q;
  double* Minv;
  double* scratch;  // (RBD_MINV_REC * (njoints - 1) + RBD_MINV_ROOT_REC) * 32 doubles per resident warp
};
struct DrneaArgs {  // derivatives of inverse dynamics: q, v, a -> dtau_dq, dtau_dv (nv x nv row-major fields) [, M packed upper]
  long long n, tile;
  const double *q, *v, *a;
  double *dtau_dq, *dtau_dv, *M;
  double* scratch;  // planned_model.slot_total[0] * 32 doubles per resident warp (cold level records)
  int sparse = 0;   // 1: dtau_dq / dtau_dv hold the structurally non-zero entries only (planned_model.slot_total[1] fields)
};
struct AbaArgs {  // forward dynamics (articulated-body algorithm): q, v, tau -> ddq
  long long n, tile;
  const double *q, *v, *tau;
  double* ddq;
  double* scratch;  // aba_scratch_doubles(model) * 32 doubles per resident warp
};
struct RowArgs {
  long long nrows;
  const int *row_ptr, *row_instance, *col;
  const double* val;
  double* out_rows;
  int* keep;
  double* qcache;  // tile 32
};
struct LaunchCfg {
  int block = 0, forced = 0, blocks_per_sm = 0;
  int tag = -1;  // which __global__ function the cached attributes were set on (eval: RBD_EVAL_* variant)
  size_t smem_bytes = 0;
};

cudaError_t launch_eval(const DevModel& planned_host_model, const DevModel* d_eval_blob, const EvalArgs& a,
                        LaunchCfg* cache8, int num_sms, cudaStream_t st);
// the same for models whose joints are all RX / RY / RZ / free-flyer (DevModel::simple_joints)
cudaError_t launch_eval_simple(const DevModel& planned_host_model, const DevModel* d_eval_blob, const EvalArgs& a,
                               LaunchCfg* cache8, int num_sms, cudaStream_t st);
// picks the instantiation the plan was made for (planned_host_model.plan.kvariant: generic, simple, chain, ffroot)
cudaError_t launch_eval_any(const DevModel& planned_host_model, const DevModel* d_eval_blob, const EvalArgs& a,
                            LaunchCfg* cache8, int num_sms, cudaStream_t st);
// optional FP32 mode of the fused evaluation (float I/O and state, tolerance 1e-4)
cudaError_t launch_eval_f32(const DevModelT<float>& planned_host_model, const DevModelT<float>* d_eval_blob,
                            const EvalArgsT<float>& a, LaunchCfg* cache8, int num_sms, cudaStream_t st);
// mapping kernels: `m` = host copy, `d_model` = device copy of the same (unplanned) joint table; cfg2[0] run-time
// tile, cfg2[1] tile 32
cudaError_t launch_apply(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                         int smem_doubles, LaunchCfg* cfg2, cudaStream_t st);
cudaError_t launch_applyJ(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                          int smem_doubles, LaunchCfg* cfg2, cudaStream_t st);
cudaError_t launch_applyJT(const DevModel& m, const DevModel* d_model, const DevTables& tb, const MapArgs& a,
                           int smem_doubles, LaunchCfg* cfg2, cudaStream_t st);
cudaError_t launch_applyJT_plan(const DevModel& planned_host_model, const DevModel* d_blob, const DevTables& tb,
                                const MapArgs& a, LaunchCfg* cfg2, int num_sms, cudaStream_t st);
// ring variant of the vector applyJT (full tiles of 32, a.tile == 32, a.n % 32 == 0, a.in 16-byte aligned):
// `planned_host_model` planned with plan_jt_ring, `d_blob` built by build_jt_blob, nst = ring stages per warp
// use_arm: fixed-base serial chains of 6 / 7 joints run applyJT_arm_kernel (unrolled sweep, state in registers)
cudaError_t launch_applyJT_ring(const DevModel& planned_host_model, const DevModel* d_blob, const MapArgs& a, int nst,
                                LaunchCfg* cfg, int num_sms, cudaStream_t st, bool use_arm = false);
cudaError_t launch_applyDJT(const DevModel& m, const DevModel* d_model, const DevTables& tb, const DjtArgs& a,
                            LaunchCfg* cfg2, cudaStream_t st);
cudaError_t launch_applyJT_rows(const DevModel& m, const DevModel* d_model, const DevTables& tb, const RowArgs& a,
                                int smem_doubles, LaunchCfg* cfg, cudaStream_t st);
cudaError_t launch_applyJT_rows_cached(const DevModel& m, const DevModel* d_model, const DevTables& tb, const RowArgs& a,
                                       const double* jcache, const double* omicache, int num_sms, cudaStream_t st);
// dense getFrameJacobian(LWA) of (instance, output) pairs from the Jacobian cache; J [npairs][nv][6], 16-byte aligned
cudaError_t launch_frame_jacobian(const DevModel& m, const DevModel* d_model, const DevTables& tb, long long npairs,
                                  const int* instance, const int* out_index, double* J, const double* jcache,
                                  const double* omicache, cudaStream_t st);
// forward dynamics; *scratch_warps receives the number of warp records the launch uses when a.scratch == nullptr
// (sizing call: nothing is launched)
cudaError_t launch_aba(const DevModel& m, const DevModel* d_model, const AbaArgs& a, int num_sms, int* scratch_warps,
                       cudaStream_t st);
// the fused variant (aba2_instance; rbd_kernels_dyn.cu): `planned_host_model` planned with plan_aba2, `d_model` its device
// copy; a.scratch holds RBD_ABA2_REC * (njoints - 1) * 32 doubles per resident warp (sizing call as above)
cudaError_t launch_aba2(const DevModel& planned_host_model, const DevModel* d_model, const AbaArgs& a, int num_sms,
                        int* scratch_warps, cudaStream_t st);
// derivatives of inverse dynamics (drnea_instance): `planned_host_model` planned with plan_drnea; sizing call as launch_aba
cudaError_t launch_drnea(const DevModel& planned_host_model, const DevModel* d_model, const DrneaArgs& a, int num_sms,
                         int* scratch_warps, cudaStream_t st);
// inverse of the mass matrix (minv_instance): `planned_host_model` planned with plan_minv; sizing call as launch_aba
cudaError_t launch_minv(const DevModel& planned_host_model, const DevModel* d_model, const MinvArgs& a, int num_sms,
                        int* scratch_warps, cudaStream_t st);
// fused eval of fixed-base serial chains with a compile-time joint count (rbd_kernels_arm.cu): returns false when `m` is
// not such a chain (6 or 7 RX / RY / RZ joints) or the call does not ask for M and tau; else launches and sets *err
bool launch_eval_arm(const DevModel& m, const EvalArgs& a, int num_sms, cudaStream_t st, cudaError_t* err);
// records the __global__ function a launch_* of another translation unit started (rbd_last_kernel_name)
void note_kernel(const void* fn);
// assembled mapping Jacobian: block rows of the n_sel selected outputs (device arrays sel [n_sel], blk_ptr [n_sel + 1])
// of n instances, val [n][6 blk_ptr[n_sel]]
cudaError_t launch_getJ(const DevModel& m, const DevModel* d_model, const DevTables& tb, long long n, int n_sel, const int* sel,
                        const int* blk_ptr, double* val, const double* jc, const double* oc, cudaStream_t st);
cudaError_t launch_aos_to_soa(long long n, int K, long long tile, const double* aos, double* soa, cudaStream_t st);
cudaError_t launch_soa_to_aos(long long n, int K, long long tile, const double* soa, double* aos, cudaStream_t st);

// mangled name of the __global__ function the calling thread launched last through the functions above (the layout
// adapters are not recorded), or nullptr
const char* last_kernel_mangled();

}  // namespace rbd
