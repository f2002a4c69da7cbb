// rbd_sweeps.h — per-instance bodies of the sm_100a kernels (one thread = one robot instance).
//
// Everything here is written in the WORLD frame: with oMi known, joint motion subspaces become the columns
// of pinocchio's data.J (A_i = oMi.act(S_i)), spatial velocities/accelerations/forces of all bodies live in
// one common frame, and the backward passes of RNEA and CRBA reduce to plain additions
// (f_parent += f_child, Ycrb_parent += Ycrb_child) instead of an SE3 action per joint.  Results equal
// pinocchio's LOCAL-convention crba/rnea (reference oracle: oracle/rbd_oracle.c) up to rounding.
//
// Joints are numbered in DFS pre-order (pinocchio's URDF order; checked by rbd_model_create), so one
// root-to-leaf sweep with a stack indexed by tree depth visits every joint once; a joint's subtree is
// complete the moment the sweep leaves it, and its backward step ("pop") runs right then.  Live state per
// instance is therefore O(depth), not O(njoints): that is what lets a Talos-sized robot keep its whole
// working set in shared memory ("state budget").
//
// The functions are templates over two accessor types so that the identical code is compiled
//   * by nvcc into the __global__ kernels (rbd_kernels.cu): Stack = shared memory column of this thread,
//     Field = strided global memory;
//   * by g++ into tests/emul (CPU unit tests of the kernel logic, never shipped, never a fallback).
//
// Reference semantics covered (paths relative to the reference root, KCM.inl =
// src/sofa/RigidBodyDynamics/KinematicChainMapping.inl):
//   eval_instance     pinocchio::computeJointJacobians (call site KCM.inl:373) + crba + rnea
//   apply_instance    KCM.inl:338-396   (computeJointJacobians, updateFramePlacements, se3ToSofaType)
//   applyJ_instance   KCM.inl:398-448   (getFrameJacobian LOCAL_WORLD_ALIGNED times dq, matrix-free)
//   applyJT_instance  KCM.inl:450-513 and :515-639 (sum_k J_k^T w_k, matrix-free)
#pragma once
#include <math.h>
#ifndef RBD_PTR_OMIJ
#define RBD_PTR_OMIJ 1
#endif
#ifndef RBD_COOP_ZERO
#define RBD_COOP_ZERO 0 /* structural zeros of M written cooperatively with 16-byte stores (tile 32): fewer
                           instructions, yet measured SLOWER (A/B r1h: 2.361 vs 2.333 ms) -> off */
#endif
#ifndef RBD_UNROLL_ANC
#define RBD_UNROLL_ANC 1 /* A/B r1e: 1 -> 2.45 ms, 2 -> 2.49 ms, 4 -> 2.67 ms (instruction-cache pressure) */
#endif
#ifndef RBD_UNROLL_FF
#define RBD_UNROLL_FF 6 /* rolling this loop indexes A dynamically -> local memory: measured 5 % slower */
#endif
#ifndef RBD_UNROLL_ZERO
#define RBD_UNROLL_ZERO 4
#endif
#define RBD_PRAGMA_(x) _Pragma(#x)
#define RBD_PRAGMA(x) RBD_PRAGMA_(x)
#ifndef RBD_EARLY_RESTORE
#define RBD_EARLY_RESTORE 1 /* state of the joint a new branch starts from is reloaded before the unwind of the previous
                               branch (r1f: -3 %); 0 = at the top of the next joint: fewer live registers */
#endif
#ifndef RBD_PTR_COL
#define RBD_PTR_COL 0 /* measured slower (A/B r1d: 2.98 vs 2.85 ms): pointer bumps serialise the zero-fill */
#endif

#include "rbd_dev_model.h"

#if defined(__CUDACC__)
#define RBD_HD __host__ __device__ __forceinline__
#else
#define RBD_HD inline
#endif

// The bodies live in rbd_sweeps_body.h, written against the scalar type `real` and the model types `DevModel` /
// `DevJoint`, and are compiled twice: in namespace rbd with real = double (the product), and in namespace rbd::f32
// with real = float (the optional FP32 mode of the fused evaluation, tolerance 1e-4).
namespace rbd {
typedef double real;
#include "rbd_sweeps_body.h"

// ... and a third time in namespace rbd::simple for models whose joints are all RX / RY / RZ / free-flyer: the fused
// evaluation without the code of the rarer joint types (rbd_dev_model.h: DevModelT::simple_joints picks it)
namespace simple {
typedef double real;
#define RBD_SIMPLE_JOINTS 1
#define RBD_BODY_EVAL_ONLY 1
#include "rbd_sweeps_body.h"
#undef RBD_BODY_EVAL_ONLY
#undef RBD_SIMPLE_JOINTS
}  // namespace simple

// ... for floating-base robots whose free-flyer root is the universe's only child and the model's only free-flyer
// (DevModelT::ff_root_only: Solo, Talos): the root is stepped in front of the joint loop and emitted behind it
namespace ffroot {
typedef double real;
#define RBD_SIMPLE_JOINTS 1
#define RBD_FF_ROOT_ONLY 1
#define RBD_BODY_EVAL_ONLY 1
#include "rbd_sweeps_body.h"
#undef RBD_BODY_EVAL_ONLY
#undef RBD_FF_ROOT_ONLY
#undef RBD_SIMPLE_JOINTS
}  // namespace ffroot

// ... the same for plans of <= 8 warps per SM (Talos: the sweep state, not the registers, caps the SM): compiled with
// the full register budget, so the early reload of the branch state stays in
namespace ffroot8 {
typedef double real;
#define RBD_SIMPLE_JOINTS 1
#define RBD_FF_ROOT_ONLY 1
#define RBD_FF_ROOT_WIDE_REGS 1
#define RBD_BODY_EVAL_ONLY 1
#include "rbd_sweeps_body.h"
#undef RBD_BODY_EVAL_ONLY
#undef RBD_FF_ROOT_WIDE_REGS
#undef RBD_FF_ROOT_ONLY
#undef RBD_SIMPLE_JOINTS
}  // namespace ffroot8

// ... and for fixed-base serial chains of RX / RY / RZ joints (DevModelT::chain_only): additionally without the
// free-flyer and branch code
namespace chain {
typedef double real;
#define RBD_SIMPLE_JOINTS 1
#define RBD_CHAIN_ONLY 1
#define RBD_BODY_EVAL_ONLY 1
#include "rbd_sweeps_body.h"
#include "rbd_arm_body.h"
#undef RBD_BODY_EVAL_ONLY
#undef RBD_CHAIN_ONLY
#undef RBD_SIMPLE_JOINTS
}  // namespace chain

namespace f32 {
typedef float real;
using DevModel = ::DevModelT<float>;
using DevJoint = ::DevJointT<float>;
#define RBD_BODY_F32 1
#include "rbd_sweeps_body.h"
#undef RBD_BODY_F32
}  // namespace f32
}  // namespace rbd
