/*
 * rbd_b200.h — C ABI of the B200-native batched articulated-body dynamics back end
 *              for the Sofa.RigidBodyDynamics plugin.
 *
 * This is the drop-in boundary.  It replaces, for N robot instances at once, the
 * pinocchio calls and the per-frame Eigen loops inside the reference's
 *   src/sofa/RigidBodyDynamics/KinematicChainMapping.inl   (short: KCM.inl)
 * protected workers, and adds the Mass/ForceField quantities the north-star names (CRBA, RNEA).
 * Each entry point below cites the reference interface it stands behind (file:line, relative to the
 * reference root).  INTEGRATION.md shows the patched KCM.inl that binds them.
 *
 * Conventions (identical to the reference unless stated):
 *   - spatial 6-vectors are [linear(3); angular(3)]                      (Conversions.h:80-93)
 *   - rigid poses are [x y z qx qy qz qw]                                (Conversions.h:34-37,58-71)
 *   - free-flyer configuration q = [x y z qx qy qz qw], velocity = 6-vector in the joint (body) frame
 *   - SE3 placements in descriptors / oMi outputs: 12 doubles, R row-major (9) then p (3)
 *   - inertia records: [m, cx, cy, cz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz]      (I about the CoM; pinocchio order)
 *   - joint 0 is pinocchio's "universe"; kSkipUniverse = 0 (Types.h:32) so it is also output body 0
 *
 * Memory layouts:
 *   *_soa entry points take DEVICE pointers in batch-strided ("tiled SoA") layout: an array with K fields
 *   per instance stores field k of instance i at
 *        (i / tile) * K * tile  +  k * tile  +  (i % tile)            [doubles]
 *   tile >= n gives the plain SoA x[k*tile + i]; tile = 32..256 gives AoSoA.  tile must be >= 1.
 *   *_host entry points take HOST pointers in the reference's own instance-major ("SOFA AoS") layout,
 *   i.e. exactly the bytes of N concatenated sofa VecCoord/VecDeriv vectors; copies to and from the
 *   device happen inside the call.
 *
 * Error convention (replaces ComponentState::Invalid + msg_error, KCM.inl:114-139): every function
 * returns 0 on success or a negative rbd_status; nothing throws or aborts across this boundary.
 * rbd_last_error() returns a thread-local description of the last failure.
 *
 * Threading (KCM is single-threaded and not re-entrant, KCM.h:120-121): calls on one workspace must be
 * serialised by the caller; distinct workspaces are independent.  Multi-GPU = one workspace per device.
 *
 * There is NO CPU fallback behind this ABI: without a CUDA device every compute call fails with
 * RBD_ERR_CUDA.
 */
#ifndef RBD_B200_H
#define RBD_B200_H

#include <stdint.h>

#if defined(_WIN32)
#  define RBD_API __declspec(dllexport)
#else
#  define RBD_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RBD_B200_VERSION 100 /* 0.1.0 */

typedef enum rbd_status {
  RBD_OK = 0,
  RBD_ERR_INVALID_ARG = -1, /* null pointer, negative size, n > capacity ...                          */
  RBD_ERR_MODEL = -2,       /* descriptor violates a model invariant (parents[i] < i, pre-order ...)   */
  RBD_ERR_CUDA = -3,        /* CUDA runtime error (message in rbd_last_error)                         */
  RBD_ERR_NO_APPLY = -4,    /* applyJ / applyJT called before apply (reference relies on the cached
                               Jacobians of the previous apply, KCM.inl:365-372)                       */
  RBD_ERR_UNSUPPORTED = -5, /* model too large for the staged-constant path, unknown joint type ...    */
  RBD_ERR_ALLOC = -6
} rbd_status;

/* Joint type codes (pinocchio joint models reachable from URDF). */
enum {
  RBD_JOINT_UNIVERSE = 0,
  RBD_JOINT_RX = 1, RBD_JOINT_RY = 2, RBD_JOINT_RZ = 3, RBD_JOINT_RU = 4,         /* revolute             */
  RBD_JOINT_PX = 5, RBD_JOINT_PY = 6, RBD_JOINT_PZ = 7, RBD_JOINT_PU = 8,         /* prismatic            */
  RBD_JOINT_RUBX = 9, RBD_JOINT_RUBY = 10, RBD_JOINT_RUBZ = 11, RBD_JOINT_RUBU = 12, /* q = (cos, sin)     */
  RBD_JOINT_FF = 13                                                               /* free-flyer           */
};

/*
 * Model descriptor: the constants of pinocchio::Model the hot path reads.  Stands for the
 * std::shared_ptr<pinocchio::Model> handed over by KinematicChainMapping::setModel (KCM.inl:304-311)
 * plus the two frame tables (setBodyCoMFrames KCM.inl:313-317; m_extraFrames KCM.h:95), as produced by
 * URDFModelLoader::load() (URDFModelLoader.cpp:126-184, 263-268).  All arrays are copied by
 * rbd_model_create; the caller keeps ownership.
 */
typedef struct rbd_model_desc {
  int32_t njoints;                /* model.njoints, universe included                                    */
  int32_t nq, nv;                 /* model.nq, model.nv                                                  */
  const int32_t* parents;         /* [njoints]   model.parents; parents[0] = 0, parents[i] < i           */
  const int32_t* jtype;           /* [njoints]   RBD_JOINT_*                                             */
  const double* axis;             /* [njoints*3] unit joint axis in the joint frame                      */
  const double* placement;        /* [njoints*12] model.jointPlacements                                  */
  const int32_t* idx_q;           /* [njoints]   first configuration index of each joint                 */
  const int32_t* idx_v;           /* [njoints]   first velocity index of each joint                      */
  const double* inertia;          /* [njoints*10] model.inertias                                         */
  int32_t nframes;                /* model.nframes after the CoM frames were appended                    */
  const int32_t* frame_parent;    /* [nframes]   frames[f].parentJoint                                   */
  const double* frame_placement;  /* [nframes*12] frames[f].placement                                    */
  double gravity[3];              /* model.gravity.linear()                                              */
  int32_t n_out;                  /* njoints + nframes0 (mappedDof size, URDFModelLoader.cpp:277)        */
  const int32_t* out_frames;      /* [n_out] frame index of every mapped output:
                                     bodyCoMFrames then extraFrames (KCM.inl:382-395)                    */
} rbd_model_desc;

typedef struct rbd_model rbd_model;         /* immutable, shareable between workspaces                   */
typedef struct rbd_workspace rbd_workspace; /* per device / per calling thread; owns device buffers      */

typedef struct rbd_model_info {
  int32_t njoints, nq, nv, nframes, n_out;
  int32_t has_freeflyer_root;   /* 1 when joint 1 is a free-flyer attached to the universe               */
  int32_t nq_joints, nv_joints; /* sizes of SOFA input1: nq (-7), nv (-6)  (URDFModelLoader.cpp:200)      */
  int32_t omi_fields;           /* 12*(njoints-1)                                                        */
  int32_t jac_fields;           /* 6*nv                                                                  */
  int32_t mass_fields;          /* nv*(nv+1)/2                                                           */
  int32_t max_depth;            /* longest joint chain                                                   */
  int32_t eval_smem_doubles;    /* shared-memory doubles per instance used by the fused eval kernel      */
  int64_t eval_bytes;           /* algorithmic bytes per eval                          */
  int32_t eval_tmem_doubles;    /* tensor-memory doubles per instance used by the fused eval kernel      */
  int32_t eval_warps_per_sm;    /* resident warps per SM the storage plan of the fused eval reaches      */
  int32_t mass_support_fields;  /* structurally non-zero entries of M's upper triangle (rbd_model_mass_pattern)  */
} rbd_model_info;

/* ---------------------------------------------------------------------------------------------------- */
/* library                                                                                              */
RBD_API int rbd_version(void);
RBD_API const char* rbd_last_error(void);
RBD_API void rbd_clear_error(void);
RBD_API int rbd_device_count(int* count); /* RBD_ERR_CUDA when no driver / device is present            */
/* Demangled name of the __global__ function the calling thread launched last through this library ("" = none yet; the
 * layout adapters are not recorded).  Thread-local like rbd_last_error; bench.py reports the kernel it timed with it. */
RBD_API const char* rbd_last_kernel_name(void);

/* Measurement probe (no model, no workspace): peak FP64 throughput of `device` from dependency-free DFMA chains on every
 * SM, in TFLOP/s (2 flops per DFMA), best of a few repetitions — the measured FP64 roofline denominator bench.py
 * reports next to the HBM one ("measured DFMA peak").  Takes a few milliseconds.                  */
RBD_API int rbd_probe_fp64_peak(int device, double* tflops);

/* ---------------------------------------------------------------------------------------------------- */
/* model: replaces setModel / setBodyCoMFrames / m_extraFrames (KCM.inl:304-317, KCM.h:90-95)            */
RBD_API int rbd_model_create(const rbd_model_desc* desc, rbd_model** out);
RBD_API void rbd_model_destroy(rbd_model* model);
RBD_API int rbd_model_get_info(const rbd_model* model, rbd_model_info* info);
/* Support pattern of the joint-space mass matrix: M(r, c), r <= c, is structurally non-zero iff the joint that owns dof
 * r is the joint that owns dof c or one of its ancestors (pinocchio's crba fills exactly these entries of data.M's
 * upper triangle and leaves the rest 0).  Compressed by columns: col_ptr [nv + 1], row_idx [mass_support_fields], rows
 * ascending within a column; either pointer may be NULL.  Entry e of the pattern is field e of the `Mval` arrays of
 * the *_sparse entry points below — the layout a sparse direct solver (the reference's example scenes use
 * SparseLDLSolver, examples/sofa/RigidBodyDynamics/LoadRobot/robot.py:48) assembles from without scanning zeros.  */
RBD_API int rbd_model_mass_pattern(const rbd_model* model, int32_t* col_ptr, int32_t* row_idx);

/* ---------------------------------------------------------------------------------------------------- */
/* workspace: replaces the pinocchio::Data owned by the mapping (KCM.h:121, KCM.inl:310).  `capacity` is  */
/* the largest batch the workspace will see; `stream` is a cudaStream_t (NULL = default stream).          */
RBD_API int rbd_workspace_create(const rbd_model* model, int device, int64_t capacity, void* stream,
                                 rbd_workspace** out);
RBD_API void rbd_workspace_destroy(rbd_workspace* ws);
RBD_API int rbd_workspace_set_stream(rbd_workspace* ws, void* stream);
RBD_API int rbd_workspace_synchronize(rbd_workspace* ws);
/* Number of kernels this workspace has launched since creation (bench.py's "gpu_launches").              */
RBD_API int64_t rbd_workspace_launch_count(const rbd_workspace* ws);
/* Block size override for the fused eval kernel (0 = automatic).  Tuning knob, not part of the contract. */
RBD_API int rbd_workspace_set_eval_block(rbd_workspace* ws, int threads);
/* Tensor memory as sweep-state storage of the fused eval kernel: 1 = on (default), 0 = shared memory only.
 * Tuning / A-B knob, not part of the contract: results are identical either way.                         */
RBD_API int rbd_workspace_set_eval_tmem(rbd_workspace* ws, int enable);
/* Ring variant of the vector applyJT (full tiles of a tile-32 batch: the wrench stream is staged in shared memory by
 * bulk asynchronous copies): 1 = on (default), 0 = always the register path; out (may be NULL) receives the ring stages
 * per warp the plan found for this model (0 = the ring kernel cannot run it).  Tuning / A-B knob: same results.    */
RBD_API int rbd_workspace_set_applyJT_ring(rbd_workspace* ws, int enable, int32_t* ring_slots);
/* Constraint-row applyJT on cached Jacobians (default 1): the first row call after an apply fills a cache of the
 * applied instances' world joint Jacobian and joint placements (what the reference's apply leaves in pinocchio's
 * data.J / data.oMi, KinematicChainMapping.inl:365-373) — 8 (6 nv + 12 (njoints-1)) bytes per instance of workspace
 * capacity, allocated on first use — and every row only reads it; 0 = per-row sweep, no cache.  Same results.      */
RBD_API int rbd_workspace_set_rows_cache(rbd_workspace* ws, int enable);
/* Fused eval (with M and tau requested) and vector applyJT (full tiles of a tile-32 batch) of fixed-base serial chains of
 * 6 or 7 RX / RY / RZ joints (arms): 1 (default) = the sweeps compiled for that joint count (unrolled, joint table in the
 * kernel-parameter constant bank, subspace columns in registers); 0 = the generic instantiations (A/B knob).  Same
 * results up to rounding.                                                                                         */
RBD_API int rbd_workspace_set_eval_arm(rbd_workspace* ws, int enable);
/* Forward dynamics (rbd_aba_*): 1 (default) = the fused sweep — passes 1 + 2 of the articulated-body algorithm on on-chip
 * level records (shared + tensor memory), only the 20 doubles per joint pass 3 needs go through global memory — whenever
 * its storage plan fits the SM; 0 = always the first version with its whole per-joint state in a global-memory record
 * (A/B knob; very deep trees take that path anyway).  Same results up to rounding.                              */
RBD_API int rbd_workspace_set_aba_fused(rbd_workspace* ws, int enable);
/* Storage plan the fused eval kernel uses on this workspace for the output set (with_M, with_tau):
 * out[0] warps per block, [1] blocks per SM, [2] shared-memory doubles per instance, [3] tensor-memory doubles per
 * instance, [4] tensor-memory columns per block, [5] 1 if the Y-stack lives in tensor memory, [6] shared-memory
 * bytes per block, [7] reserved.  Introspection for bench.py / DESIGN.md.                                 */
RBD_API int rbd_workspace_get_eval_plan(rbd_workspace* ws, int with_M, int with_tau, int32_t out[8]);
/* ... and of the FP32 mode (elements are floats: [2], [3] count floats).                                    */
RBD_API int rbd_workspace_get_eval_plan_f32(rbd_workspace* ws, int with_M, int with_tau, int32_t out[8]);

/* ---------------------------------------------------------------------------------------------------- */
/* Mapping path, device SoA.                                                                            */

/* apply — KinematicChainMapping::apply worker (KCM.inl:338-396): q -> world poses of the n_out mapped
 * bodies/frames, [x y z qx qy qz qw] each (Eigen's rotation->quaternion branch order, Conversions.h:60).
 *   q_joints : nq_joints fields    root_pose : 7 fields or NULL (required iff has_freeflyer_root)
 *   out      : n_out*7 fields (overwritten)
 * Caches the assembled configuration in the workspace for applyJ / applyJT (KCM.inl:365-372).           */
RBD_API int rbd_apply_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q_joints,
                          const double* root_pose, double* out);

/* applyJ — KCM.inl:398-448: out[k] = J_k(LOCAL_WORLD_ALIGNED) * dq for every mapped output k,
 * evaluated matrix-free.  The root velocity is used verbatim as the free-flyer joint velocity, as the
 * reference does (KCM.inl:417-418).   dq_joints: nv_joints fields; root_vel: 6 fields or NULL;
 * out: n_out*6 fields (overwritten).                                                                    */
RBD_API int rbd_applyJ_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* dq_joints,
                           const double* root_vel, double* out);

/* applyJT (vector) — KCM.inl:450-513: tau += sum_k J_k^T in[k].   in: n_out*6 fields;
 * out_tau: nv_joints fields, ACCUMULATED (+=) as the reference does (KCM.inl:498-511);
 * out_root: 6 fields accumulated, or NULL.                                                              */
RBD_API int rbd_applyJT_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* in, double* out_tau,
                            double* out_root);

/* applyDJT — OPTIONAL EXTENSION.  The reference's applyDJT is an explicit no-op (KinematicChainMapping.h:73-80), i.e.
 * the mapping contributes no geometric stiffness; a drop-in keeps it a no-op (the host mirror does).  This entry point
 * is the real term for callers who want it:  out += kfactor * d/de [ J(q (+) e dq)^T in ] at e = 0  — how the
 * generalized forces of the FIXED world-aligned wrenches `in` (n_out*6 fields, as in applyJT) change along the
 * direction dq (dq_joints / root_vel as in applyJ; the root velocity in the free-flyer's joint frame).  Uses the
 * configuration cached by the last apply; out_tau / out_root are ACCUMULATED; root_vel and out_root go together.
 * A free-flyer is supported as the root joint only.                                                              */
RBD_API int rbd_applyDJT_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* dq_joints,
                             const double* root_vel, const double* in, double kfactor, double* out_tau,
                             double* out_root);
RBD_API int rbd_applyDJT_host(rbd_workspace* ws, int64_t n, const double* dq_joints, const double* root_vel,
                              const double* in, double kfactor, double* out_tau, double* out_root);

/* applyJT (MatrixDeriv) — KCM.inl:515-639: one dense nv-wide row per constraint row.
 *   nrows rows in CSR form: row r owns non-zeros [row_ptr[r], row_ptr[r+1]); row_instance[r] selects the
 *   robot instance (its cached configuration); col[j] is the mapped-output index (< n_out); val is
 *   nnz x 6 instance-major ([lin; ang], Conversions.h:86-93).  All five arrays are DEVICE pointers.
 *   out_rows: nrows x nv, row-major (root wrench first when has_freeflyer_root), overwritten.
 *   keep[r] = 1 iff ||row||^2 > 1e-12 (kMinTorque^2, KCM.inl:43-45,587,605): the caller must only emit
 *   kept rows into the SOFA MatrixDeriv.                                                                */
/* Preconditions of the _dev variant (the _host variant checks them and returns RBD_ERR_INVALID_ARG): row_ptr[0] == 0,
 * row_ptr non-decreasing, 0 <= col[j] < n_out, 0 <= row_instance[r] < the batch size of the last apply.             */
RBD_API int rbd_applyJT_rows_dev(rbd_workspace* ws, int64_t nrows, const int32_t* row_ptr,
                                 const int32_t* row_instance, const int32_t* col, const double* val,
                                 double* out_rows, int32_t* keep);

/* getFrameJacobian — pinocchio::getFrameJacobian(model, data, frame, LOCAL_WORLD_ALIGNED, J) as the reference
 * calls it for every mapped output before each J dq / J^T f product (KCM.inl:434,444,480,490,579), here for
 * explicit (instance, output) pairs: J[p] = dense 6 x nv Jacobian of mapped output out_index[p] (< n_out) of
 * instance instance[p], column-major ([lin; ang] per column, pinocchio's Matrix6x layout), columns off the path
 * to the root zero.  Read from the Jacobian cache of the last apply (see rbd_workspace_set_rows_cache); this is
 * what a real getJ() (KinematicChainMapping.h:83 returns nullptr) would assemble for SOFA's direct solvers.
 *   _dev: instance, out_index, J are DEVICE pointers, J 16-byte aligned, [npairs][nv][6].   _host: host pointers.  */
RBD_API int rbd_getFrameJacobian_dev(rbd_workspace* ws, int64_t npairs, const int32_t* instance,
                                     const int32_t* out_index, double* J);
RBD_API int rbd_getFrameJacobian_host(rbd_workspace* ws, int64_t npairs, const int32_t* instance,
                                      const int32_t* out_index, double* J);

/* getJ — the ASSEMBLED mapping Jacobian.  The reference's getJ() returns nullptr (KinematicChainMapping.h:82-83), so
 * SOFA's direct solvers (the example scenes use SparseLDLSolver, examples/sofa/RigidBodyDynamics/LoadRobot/robot.py:48)
 * cannot see through the mapping.  The matrix has one 6 x nv block row per mapped output, and block row k is
 * structurally non-zero only in the dofs on the path from the output's parent joint to the root — exactly the columns
 * getFrameJacobian(LOCAL_WORLD_ALIGNED) fills at KinematicChainMapping.inl:434-490.
 *   rbd_model_getJ_pattern  the fixed pattern for a selection of n_sel mapped outputs (sel_out == NULL: all n_out, in
 *                           order): block row s has the dof columns cols[blk_ptr[s] .. blk_ptr[s+1]), ascending;
 *                           blk_ptr [n_sel + 1]; cols may be NULL (sizing pass).
 *   rbd_getJ_dev / _host    the values for instances 0 .. n-1 of the last apply: val [n][6 * blk_ptr[n_sel]], block row
 *                           s of instance i at val[i][6 blk_ptr[s] ...], 6 doubles [lin; ang] per column (pinocchio's
 *                           Matrix6x column layout).  sel_out is a HOST array in both variants (a per-scene constant);
 *                           _dev: val is a 16-byte aligned DEVICE pointer, _host: a host pointer.
 * With one UniformMass per body below the mapping (URDFModelLoader.cpp:369-391) the joint-space matrix SOFA would
 * assemble through this Jacobian, sum_b J_b^T M_b J_b, is what rbd_crba_soa_sparse returns directly (same support
 * pattern as rbd_model_mass_pattern; tests/test_gpu_configs.py::test_mass_matrix_through_the_mapping ties the two).  */
RBD_API int rbd_model_getJ_pattern(const rbd_model* model, int32_t n_sel, const int32_t* sel_out, int32_t* blk_ptr,
                                   int32_t* cols);
RBD_API int rbd_getJ_dev(rbd_workspace* ws, int64_t n, int32_t n_sel, const int32_t* sel_out, double* val);
RBD_API int rbd_getJ_host(rbd_workspace* ws, int64_t n, int32_t n_sel, const int32_t* sel_out, double* val);

/* ---------------------------------------------------------------------------------------------------- */
/* Mass / ForceField path, device SoA (north-star additions; the reference reaches them matrix-free
 * through UniformMass + applyJ/applyJT, URDFModelLoader.cpp:369-391).                                    */

/* Fused evaluation = the BASELINE metric unit: forward kinematics + world joint Jacobian
 * (pinocchio::computeJointJacobians, call site KCM.inl:373) + CRBA + RNEA.
 *   q: nq fields (full configuration incl. free-flyer), v, a: nv fields.
 *   oMi : 12*(njoints-1) fields — placement of joint i (1-based) at fields [12(i-1), 12i): R row-major, p
 *   J   : 6*nv fields — column c at fields [6c, 6c+6): [lin; ang], world frame (pinocchio data.J)
 *   M   : nv*(nv+1)/2 fields — upper triangle packed by columns: M(r,c), r <= c, at field c(c+1)/2 + r
 *   tau : nv fields — rnea(q, v, a)
 * Any output pointer may be NULL; the corresponding work is skipped.                                     */
RBD_API int rbd_eval_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                         const double* a, double* oMi, double* J, double* M, double* tau);
/* Optional FP32 mode of the same fused evaluation (BASELINE north-star: tolerance 1e-4 relative instead of 1e-10):
 * float inputs, float outputs, float sweep state; same layouts (fields are floats), same kernel structure.     */
RBD_API int rbd_eval_soa_f32(rbd_workspace* ws, int64_t n, int64_t tile, const float* q, const float* v,
                             const float* a, float* oMi, float* J, float* M, float* tau);
RBD_API int rbd_crba_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* M);
/* The same with M in the support-only format: Mval has mass_support_fields fields per instance, field e = the entry
 * (row_idx[e], column of e) of rbd_model_mass_pattern; the structural zeros of the packed triangle (Talos: 372 of 741)
 * are neither computed, stored nor — through rbd_eval_host_sparse — copied over PCIe.  Same kernel, same values.   */
RBD_API int rbd_eval_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                const double* a, double* oMi, double* J, double* Mval, double* tau);
RBD_API int rbd_crba_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* Mval);
RBD_API int rbd_rnea_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                         const double* a, double* tau);

/* Forward dynamics — pinocchio::aba(model, data, q, v, tau): the joint accelerations ddq [nv fields] of the articulated
 * body algorithm (not in the reference, which never integrates joint dynamics itself).  Same
 * conventions as rbd_rnea_soa, of which it is the inverse: rbd_aba(q, v, rbd_rnea(q, v, a)) == a.  A free-flyer is
 * supported as the root joint only (RBD_ERR_UNSUPPORTED otherwise).  The sweep keeps 45 doubles per joint in a scratch
 * record per resident warp (allocated in the workspace on first use).                                              */
RBD_API int rbd_aba_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v, const double* tau,
                        double* ddq);
/* ... on host vectors (instance-major q [n][nq]; v, tau, ddq [n][nv]), chunked and pipelined like rbd_eval_host.      */
RBD_API int rbd_aba_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* tau, double* ddq);
/* Inverse of the joint-space mass matrix, pinocchio::computeMinverse(model, data, q) (no call site in the reference;
 * SURVEY.md §8(f) rank 4): Minv = upper triangle of M(q)^-1 packed by columns exactly like M (field c (c + 1) / 2 + r,
 * r <= c; rbd_model_info::mass_fields fields per instance) — pinocchio fills the upper triangle of data.Minv too.  Computed
 * without forming M: the articulated-body quantities of the forward-dynamics sweep are shared by all columns (unit torque
 * per dof, v = 0, no gravity).  A free-flyer is supported as the root joint only; RBD_ERR_UNSUPPORTED for trees whose
 * level records do not fit an SM.                                                                                */
RBD_API int rbd_minv_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, double* Minv);
RBD_API int rbd_minv_host(rbd_workspace* ws, int64_t n, const double* q, double* Minv);
/* Analytical derivatives of inverse dynamics, pinocchio::computeRNEADerivatives(model, data, q, v, a, dtau_dq, dtau_dv,
 * dtau_da) (no call site in the reference): for tau = rnea(q, v, a),
 *   dtau_dq [nv * nv fields]  entry (r, c) at field r * nv + c (row-major, as pinocchio's RowMatrixXs) = d tau_r / d q_c,
 *                             taken along the tangent of pinocchio::integrate (free-flyer: velocity in its own frame);
 *   dtau_dv [nv * nv fields]  d tau_r / d v_c, same layout;
 *   M       [mass_fields]     optional (may be NULL): dtau_da = M(q), packed upper triangle exactly like rbd_crba_soa.
 * Both matrices are written completely, structural zeros (dofs on different branches of the tree) included.  One
 * depth-first sweep with all state on chip (shared + tensor memory); a free-flyer is supported as the root joint only;
 * RBD_ERR_UNSUPPORTED for trees whose level records do not fit an SM.                                              */
RBD_API int rbd_rnea_derivatives_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                     const double* a, double* dtau_dq, double* dtau_dv, double* M);
/* ... on host vectors (instance-major q [n][nq]; v, a [n][nv]; dtau_dq, dtau_dv [n][nv][nv]; M [n][mass_fields] or NULL) */
RBD_API int rbd_rnea_derivatives_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                      double* dtau_dq, double* dtau_dv, double* M);
/* The same matrices in a support-only format: entry (r, c) of dtau_dq / dtau_dv is structurally non-zero iff the joints
 * that own dofs r and c lie on one root-to-leaf path (fewer than half of the entries for a humanoid: Talos 700 of 1 444).
 *   rbd_model_drnea_pattern   the fixed pattern, compressed by rows: row_ptr [nv + 1], col_idx [nnz], columns ascending
 *                             within a row; either array may be NULL (sizing pass); *nnz (may be NULL) its size;
 *   *_sparse                  dq_val / dv_val [nnz fields]: entry e of the pattern is field e.  Same sweep, same values
 *                             as the dense entry points; the structural zeros are neither written nor copied.        */
RBD_API int rbd_model_drnea_pattern(const rbd_model* model, int32_t* row_ptr, int32_t* col_idx, int32_t* nnz);
RBD_API int rbd_rnea_derivatives_soa_sparse(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v,
                                            const double* a, double* dq_val, double* dv_val);
RBD_API int rbd_rnea_derivatives_host_sparse(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                             double* dq_val, double* dv_val);

/* Matrix-free Mass / ForceField products in joint space — what the reference's SOFA graph evaluates through the
 * mapping with one UniformMass<Rigid3> per body (URDFModelLoader.cpp:369-391): Mass::addMDx and ForceField::addForce
 * reach the joint dofs as applyJT(M_b * applyJ(dx)) and applyJT(m_b g).  Here they are one sweep each, with pinocchio's
 * exact rigid-body inertias (SURVEY.md §3.5 explains where SOFA's UniformMass deviates for non-isotropic bodies):
 *   rbd_addMDx_soa    res <- res + factor * M(q) dx                      (Mass::addMDx(res, dx, factor))
 *   rbd_addForce_soa  f   <- f - (C(q, v) v + g(q)) = f - rnea(q, v, 0)  (ForceField::addForce(f, x, v)); v == NULL
 *                     leaves the gravity term alone, f <- f + sum_b J_b^T m_b g, which is exactly what the reference
 *                     computes (applyDJT being a no-op, KinematicChainMapping.h:73-80, it has no velocity term).
 * q: nq fields; dx, v, res, f: nv fields; both ACCUMULATE as the SOFA methods do.                           */
RBD_API int rbd_addMDx_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* dx, double factor,
                           double* res);
RBD_API int rbd_addForce_soa(rbd_workspace* ws, int64_t n, int64_t tile, const double* q, const double* v, double* f);

/* ---------------------------------------------------------------------------------------------------- */
/* Layout adapters (Conversions.h:58-119 at batch scale): instance-major [n][K] <-> tiled SoA, on device. */
RBD_API int rbd_aos_to_soa_dev(rbd_workspace* ws, int64_t n, int32_t K, int64_t tile, const double* aos,
                               double* soa);
RBD_API int rbd_soa_to_aos_dev(rbd_workspace* ws, int64_t n, int32_t K, int64_t tile, const double* soa,
                               double* aos);

/* ---------------------------------------------------------------------------------------------------- */
/* Host (SOFA AoS) entry points — what the patched KinematicChainMapping calls.  Pointers are HOST
 * pointers, instance-major: [n][nq_joints], [n][7], [n][n_out][7], [n][n_out][6] ...  Pinned memory
 * makes the copies asynchronous; pageable memory works.  Each call returns after its outputs are ready. */
RBD_API int rbd_apply_host(rbd_workspace* ws, int64_t n, const double* q_joints, const double* root_pose,
                           double* out);
RBD_API int rbd_applyJ_host(rbd_workspace* ws, int64_t n, const double* dq_joints, const double* root_vel,
                            double* out);
RBD_API int rbd_applyJT_host(rbd_workspace* ws, int64_t n, const double* in, double* out_tau,
                             double* out_root);
RBD_API int rbd_applyJT_rows_host(rbd_workspace* ws, int64_t nrows, const int32_t* row_ptr,
                                  const int32_t* row_instance, const int32_t* col, const double* val,
                                  double* out_rows, int32_t* keep);
/* q [n][nq], v,a [n][nv] -> oMi [n][12(njoints-1)], J [n][6 nv], M [n][nv(nv+1)/2], tau [n][nv];
 * outputs may be NULL.  Processed in pipelined chunks (H2D, kernels and D2H overlap).                    */
RBD_API int rbd_eval_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                          double* oMi, double* J, double* M, double* tau);
/* ... with M in the support-only format, Mval [n][mass_support_fields] (see rbd_eval_soa_sparse).  Only the arrays the
 * requested outputs need cross PCIe: a NULL output is neither computed nor copied, v and a travel only when tau is
 * requested (a Mass / ForceField consumer asks for Mval + tau and gets 8 (mass_support_fields + nv) bytes back).    */
RBD_API int rbd_eval_host_sparse(rbd_workspace* ws, int64_t n, const double* q, const double* v, const double* a,
                                 double* oMi, double* J, double* Mval, double* tau);

/* Mass / ForceField products on host vectors (instance-major q [n][nq]; dx, v, res, f [n][nv]): res and f are
 * ACCUMULATED in place, as rbd_addMDx_soa / rbd_addForce_soa do on the device.  v may be NULL (gravity only).   */
RBD_API int rbd_addMDx_host(rbd_workspace* ws, int64_t n, const double* q, const double* dx, double factor, double* res);
RBD_API int rbd_addForce_host(rbd_workspace* ws, int64_t n, const double* q, const double* v, double* f);

/* ---------------------------------------------------------------------------------------------------- */
/* Multi-robot batching context.  N KinematicChainMapping components of one SOFA scene —
 * N robots built from the same URDF by N URDFModelLoader instances (URDFModelLoader.cpp:263-282) — share one
 * workspace of capacity >= N.  Each component binds its instance slot to its own SOFA vectors once per topology
 * change (the vectors live wherever SOFA allocated them: scattered host memory), and calls rbd_batch_apply /
 * _applyJ / _applyJT from its worker overload (KinematicChainMapping.h:103-117) with the scene's step counter
 * (`epoch`).  The FIRST call of an epoch gathers the inputs of ALL bound robots, runs ONE batched launch and
 * scatters every robot's outputs; the calls of the other components in the same epoch return immediately.
 * Buffers bound to a slot must stay valid until rebound; sizes are those of the single-robot entry points.    */
typedef struct rbd_batch rbd_batch;
RBD_API int rbd_batch_create(rbd_workspace* ws, int64_t n_slots, rbd_batch** out);
RBD_API void rbd_batch_destroy(rbd_batch* batch);
RBD_API int rbd_batch_bind_apply(rbd_batch* batch, int64_t slot, const double* q_joints, const double* root_pose,
                                 double* out);
RBD_API int rbd_batch_bind_applyJ(rbd_batch* batch, int64_t slot, const double* dq_joints, const double* root_vel,
                                  double* out);
RBD_API int rbd_batch_bind_applyJT(rbd_batch* batch, int64_t slot, const double* in, double* out_tau,
                                   double* out_root);
/* Return 1 when this call ran the batched launch, 0 when the epoch had already been served, < 0 on error.   */
RBD_API int rbd_batch_apply(rbd_batch* batch, int64_t slot, int64_t epoch);
RBD_API int rbd_batch_applyJ(rbd_batch* batch, int64_t slot, int64_t epoch);
RBD_API int rbd_batch_applyJT(rbd_batch* batch, int64_t slot, int64_t epoch);

/* ---------------------------------------------------------------------------------------------------- */
/* model ingest (host only): URDF -> descriptor, for callers without pinocchio.
 * Produces what the reference's loader hands to the mapping: the model pinocchio::urdf::buildModel builds — with a
 * free-flyer root joint when `free_flyer` != 0 (URDFModelLoader.cpp:126-139) — plus the loader's two frame tables: every
 * parsed frame as an "extra frame" (URDFModelLoader.cpp:159-165) and one "Body_<j>_CoM" OP_FRAME per joint, universe
 * included (URDFModelLoader.cpp:175-183); out_frames = CoM frames then extra frames (KCM.inl:382-395).  Joint types:
 * revolute, continuous (unbounded cos/sin), prismatic (axis-aligned or unaligned), floating, fixed (merged into the
 * parent joint, inertias aggregated); <mimic> / limits / geometry are ignored; planar -> RBD_ERR_UNSUPPORTED; malformed
 * XML or trees -> RBD_ERR_MODEL with the reason in rbd_last_error().  Conventions are pinocchio 3.3.1's restated (see
 * csrc/rbd_urdf.cpp).  The handle owns every array rbd_urdf_desc() points to; pass the descriptor to rbd_model_create
 * (which copies) and destroy the handle when the names are no longer needed.  Frame types are pinocchio::FrameType
 * bit values (OP_FRAME 1, JOINT 2, FIXED_JOINT 4, BODY 8).                                                        */
typedef struct rbd_urdf_model rbd_urdf_model;
RBD_API int rbd_urdf_load_file(const char* path, int free_flyer, rbd_urdf_model** out);
RBD_API int rbd_urdf_load_string(const char* xml, int free_flyer, rbd_urdf_model** out);
RBD_API void rbd_urdf_destroy(rbd_urdf_model* m);
RBD_API const rbd_model_desc* rbd_urdf_desc(const rbd_urdf_model* m);
RBD_API int32_t rbd_urdf_nframes0(const rbd_urdf_model* m); /* frames before the CoM frames were appended        */
RBD_API const char* rbd_urdf_robot_name(const rbd_urdf_model* m);
RBD_API const char* rbd_urdf_joint_name(const rbd_urdf_model* m, int32_t joint);
RBD_API const char* rbd_urdf_frame_name(const rbd_urdf_model* m, int32_t frame);
RBD_API int32_t rbd_urdf_frame_type(const rbd_urdf_model* m, int32_t frame);
/* first frame called `name` whose type is in `type_mask` (pinocchio Model::getFrameId), joint by name; -1 = none */
RBD_API int32_t rbd_urdf_find_frame(const rbd_urdf_model* m, const char* name, int32_t type_mask);
RBD_API int32_t rbd_urdf_find_joint(const rbd_urdf_model* m, const char* name);

#ifdef __cplusplus
}
#endif
#endif /* RBD_B200_H */
