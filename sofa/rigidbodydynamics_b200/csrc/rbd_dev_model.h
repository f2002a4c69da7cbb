// rbd_dev_model.h — kinematic-tree constants as the kernels see them.
//
// The joint table is uploaded once per workspace and staged once per block into shared memory by every kernel
// ("tree constants staged once into shared/constant memory", BASELINE north-star): all threads of a warp read the
// same joint record at the same time, which shared memory serves as one broadcast wavefront (DESIGN.md §4.1 explains
// why not the constant bank).  Replaces the pinocchio::Model the reference shares between loader and mapping
// (reference KinematicChainMapping.inl:304-311, URDFModelLoader.cpp:263-268).
// The structs are templates over the scalar type: double (the product's precision) and float (the optional FP32
// mode of the fused evaluation); the layout differs only in the width of the constants.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifndef RBD_MAX_JOINTS
#define RBD_MAX_JOINTS 64 /* joint records that fit the 32 KB kernel-parameter space with room to spare */
#endif

// joint type codes == include/rbd_b200.h RBD_JOINT_*
enum : int {
  JT_UNIVERSE = 0, JT_RX = 1, JT_RY = 2, JT_RZ = 3, JT_RU = 4, JT_PX = 5, JT_PY = 6, JT_PZ = 7, JT_PU = 8,
  JT_RUBX = 9, JT_RUBY = 10, JT_RUBZ = 11, JT_RUBU = 12, JT_FF = 13
};

// 16-byte aligned (and the DevModel header padded to a multiple of 16) so that the compiler may read the staged
// record with 128-bit shared-memory loads
template <class real>
struct alignas(16) DevJointT {
  real pl[12];      // jointPlacement: R row-major (9), p (3)
  real axis[3];     // unit axis in the joint frame
  real Y[10];       // m, c(3), Ic (xx,xy,yy,xz,yz,zz)
  real msub;        // mass of the whole subtree below (and including) this joint: the composite-inertia mass is a
                    // model constant, so the Y-stack carries 9 values per level, not 10
  int parent, type, idx_q, idx_v;
  int nv;           // dof of this joint (0 universe, 1, or 6)
  int nchild;       // number of child joints
  int slot[2];      // stack-slot offset (doubles) of a joint that has children: [0] full layout (A|f|Y), [1] lite (A|f);
                    // leaves never touch the stack (they are emitted from registers).  Mapping kernels only: in the
                    // staged blob of the FUSED EVAL the two words carry the joint's place in the M output instead
                    // (RBD_MBASE / RBD_MOWN below; the eval sweep never reads stack slots through this table)
  int bidx;         // branch-slot index (joints with >= 2 children), -1 otherwise
  int out_begin, out_end;  // range of this joint's mapped outputs in DevTables::sorted_*
  int pl_rot_identity;     // jointPlacement rotation is exactly I (skips one 3x3 product)
  int depth;               // number of joints on the path universe -> this joint
  // fused-eval storage plan (EvalPlan below; filled by plan_eval for one kernel variant):
  int offA;                // offset of this joint's A record (6 doubles, free-flyer 12) in the shared-memory A-stack
  int boff;                // branch joints: offset of the saved (R|p)|v|a record; bit 30 set = lives in tensor memory
  // column program of this joint's column(s) of M (ColProg table): depth-1 ancestor entries, then nzseg zero segments
  int prog_off, nzseg;
};
using DevJoint = DevJointT<double>;
// ColProg entries (int32).  Ancestor entry, root first: row (idx_v of the ancestor) | free-flyer flag | A-stack offset.
// Zero segment: first row | length.  Rows not covered by either belong to the joint itself.
// Place of joint j's column(s) in the M output (eval blob only, see DevJointT::slot).  Column cc (0 .. nv-1) of the
// joint starts at field  RBD_MBASE(j) + cc * RBD_MOWN(j) + cc (cc + 1) / 2 ; within it the rows of the ancestors' dofs
// sit at the offsets the column program gives, the joint's own rows 0 .. cc at offset RBD_MOWN(j) + row.
//   packed upper triangle (pinocchio data.M):  MBASE = c0 (c0 + 1) / 2, MOWN = c0 (c0 = idx_v), program row = idx_v of
//     the ancestor, zero segments for the dofs off the ancestor chain;
//   support-only ("sparse") M:  MBASE = entries before this joint, MOWN = number of ancestor dofs, program row = position
//     of the ancestor's first dof in the column, no zero segments — the structural zeros are not stored at all.
#define RBD_MBASE(j) ((j).slot[0])
#define RBD_MOWN(j) ((j).slot[1])
#define RBD_PROG_ROW(e) ((e) & 0xffff)
#define RBD_PROG_FF(e) (((e) >> 16) & 1)
#define RBD_PROG_OFFA(e) ((e) >> 17)
#define RBD_PROG_ANC(row, ff, offA) ((row) | ((ff) << 16) | (((offA) & 0x3fff) << 17))
#define RBD_PROG_ZSEG(row, len) ((row) | ((len) << 16))
#define RBD_PROG_ZLEN(e) ((e) >> 16)

// Where the fused eval kernel keeps its per-instance sweep state (doubles per instance, DESIGN.md "state budget").
// Shared memory holds the A-stack and the f-stack; the Y-stack and individual branch records go to TENSOR MEMORY
// (tcgen05.ld/st, 256 KB per SM that this FP64 path would otherwise leave unused) when that raises the number of
// resident warps.  All offsets are in doubles from the thread's base in the respective space.
struct EvalPlan {
  int fbase;         // shared: f-stack, 6 per internal depth level (RNEA), levels [0, fsplit)
  int fsplit;        // first f-stack level that lives in tensor memory (= number of levels when none does)
  int ftbase;        // tensor memory: f-stack levels [fsplit, D)
  int ybase;         // shared: Y-stack, 9 per internal depth level (CRBA; the mass is a model constant), levels [0, ysplit)
  int ysplit;        // first Y-stack level that lives in tensor memory (0 = all of it, D = none)
  int ytbase;        // tensor memory: Y-stack levels [ysplit, D)
  int y_in_tmem;     // 1 = at least one Y-stack level lives in tensor memory (introspection)
  int smem_doubles;  // shared-memory doubles per instance
  int tmem_doubles;  // tensor-memory doubles per instance (0 = tensor memory unused)
  int warps;         // warps per block
  int blocks_per_sm; // blocks the plan expects to be resident per SM
  int tmem_cols;     // columns the block allocates (power of two >= 32, or 0)
  int in_slots;      // slots of the asynchronous input ring (rbd_sweeps_body.h InRing; 3 elements each per instance), 0 = none
  int jt_fwd;        // applyJT ring kernel: 1 = forward-projection variant (applyJT_instance_fwd)
  int kvariant;      // fused eval: the instantiation this plan was made for (RBD_EVAL_*; the 12-warp plans need the
                     // 168-register chain / ffroot kernels, every other plan runs the generic or simple one)
  int pad_[1];
};
// Instantiations of the fused eval (rbd_sweeps.h compiles the sweep once per namespace):
#define RBD_EVAL_GENERIC 0 /* any supported joint type                                                              */
#define RBD_EVAL_SIMPLE 1  /* RX / RY / RZ / free-flyer joints only                                                 */
#define RBD_EVAL_CHAIN 2   /* fixed-base serial chain of RX / RY / RZ joints (Panda-class arms): no free-flyer, no
                              branch code -> 166 registers, 12 warps per SM                                         */
#define RBD_EVAL_FFROOT 3  /* RX / RY / RZ joints below a free-flyer root that is the universe's only child (Solo,
                              Talos): the root is stepped outside the joint loop -> 168 registers, 12 warps per SM   */
#define RBD_EVAL_FFROOT8 4 /* the ffroot sweep with the full register budget, for plans of <= 8 warps per SM (Talos)       */
#define RBD_BOFF_TMEM (1 << 30)
#define RBD_BOFF_NOVA (1 << 29) /* v|a not saved: recomputed when the sweep returns (joint attached to the universe) */
#define RBD_BOFF_RPOMI (1 << 28) /* R|p not saved: re-read from the oMi output this thread wrote (oMi requested) */
#define RBD_BOFF_MASK ((1 << 28) - 1)

template <class real>
struct alignas(16) DevModelT {
  int njoints, nq, nv, n_out;
  int slot_total[2];  // doubles of stack slots per instance for the two layouts
  int nbranch;
  int has_ff_root;    // joint 1 is a free-flyer attached to the universe
  int nq_joints, nv_joints;
  int max_depth, max_depth_internal;  // longest joint chain; deepest joint that has children
  int eval_variant;                   // specialised instantiation of the fused eval this model qualifies for (RBD_EVAL_*)
  int mfields;                        // eval blob: fields per instance of the M output (nv (nv + 1) / 2 packed, or the
                                      // support-only count); 0 in unplanned models
  int prog_words_off, prog_len;       // eval blob: 8-byte-word offset of the ColProg table behind the joint records; ints
  real gravity[3];
  EvalPlan plan;
  DevJointT<real> joint[RBD_MAX_JOINTS];
};
using DevModel = DevModelT<double>;

// the same model with float constants (tree indices and the storage plan are copied verbatim)
inline void dev_model_to_f32(const DevModel& d, DevModelT<float>* f) {
  f->njoints = d.njoints; f->nq = d.nq; f->nv = d.nv; f->n_out = d.n_out;
  f->slot_total[0] = d.slot_total[0]; f->slot_total[1] = d.slot_total[1];
  f->nbranch = d.nbranch; f->has_ff_root = d.has_ff_root; f->nq_joints = d.nq_joints; f->nv_joints = d.nv_joints;
  f->max_depth = d.max_depth; f->max_depth_internal = d.max_depth_internal; f->eval_variant = d.eval_variant; f->mfields = d.mfields;
  f->prog_words_off = d.prog_words_off; f->prog_len = d.prog_len;
  for (int k = 0; k < 3; ++k) f->gravity[k] = (float)d.gravity[k];
  f->plan = d.plan;
  for (int i = 0; i < RBD_MAX_JOINTS; ++i) {
    const DevJoint& a = d.joint[i];
    DevJointT<float>& b = f->joint[i];
    for (int k = 0; k < 12; ++k) b.pl[k] = (float)a.pl[k];
    for (int k = 0; k < 3; ++k) b.axis[k] = (float)a.axis[k];
    for (int k = 0; k < 10; ++k) b.Y[k] = (float)a.Y[k];
    b.msub = (float)a.msub;
    b.parent = a.parent; b.type = a.type; b.idx_q = a.idx_q; b.idx_v = a.idx_v; b.nv = a.nv; b.nchild = a.nchild;
    b.slot[0] = a.slot[0]; b.slot[1] = a.slot[1]; b.bidx = a.bidx; b.out_begin = a.out_begin; b.out_end = a.out_end;
    b.pl_rot_identity = a.pl_rot_identity; b.depth = a.depth; b.offA = a.offA; b.boff = a.boff;
    b.prog_off = a.prog_off; b.nzseg = a.nzseg;
  }
}

// bytes of a DevModel that are meaningful (header + the first njoints records): what the eval kernel stages
template <class real>
inline size_t dev_model_bytes(const DevModelT<real>& m) {
  return sizeof(DevModelT<real>) - sizeof(DevJointT<real>) * (size_t)(RBD_MAX_JOINTS - m.njoints);
}

// shared-memory bytes the eval kernel reserves in front of the stacks: the staged table rounded to 16 bytes plus two
// spare 8-byte words (tensor-memory base address)
template <class real>
inline size_t eval_model_bytes(const DevModelT<real>& m) {
  return (dev_model_bytes(m) + 15) / 16 * 16 + ((size_t)m.prog_len * 4 + 15) / 16 * 16 + 16;
}

#ifndef RBD_ABA2_V3
#define RBD_ABA2_V3 1 /* 1: pass 3 recomputes A and c (FK + velocity recursion once more) instead of re-reading them: the
                         per-joint record shrinks from 20 to 8 doubles (U, 1 / D, u), A | c of the joints on the current path
                         wait in a small per-LEVEL record for the unwind (A/B: DESIGN.md §4.5) */
#endif
#if RBD_ABA2_V3
#define RBD_ABA2_REC 8   /* doubles per joint of the global pass-3 record of the fused forward-dynamics sweep: U (6), 1 / D, u
                            (root free-flyer: its solved acceleration in [0, 6)) */
#define RBD_ABA2_LREC 12 /* ... and per depth level: A | c of the joint on the current path (root free-flyer: R | p) */
#else
#define RBD_ABA2_REC 20 /* doubles per joint of the global pass-3 record of the fused forward-dynamics sweep (rbd_dyn_body.h) */
#define RBD_ABA2_LREC 0
#endif
#ifndef RBD_MINV_CB
#define RBD_MINV_CB 6 /* columns of M^-1 one backward / forward sweep of minv_instance serves together (A/B r2: 8 spills) */
#endif
#define RBD_MINV_REC 13 /* ... of the M^-1 sweep (minv_instance): A (6), U (6), 1 / D; + RBD_MINV_ROOT behind the joints   */
#define RBD_MINV_ROOT_REC 33 /* root free-flyer: R | p (12), (X^T I^A X)^-1 (21)                                              */
// ... and the fused forward-dynamics kernel (aba2_kernel): the joint table rounded to 16 bytes plus the two spare words
inline size_t aba2_model_bytes(const DevModelT<double>& m) { return (dev_model_bytes(m) + 15) / 16 * 16 + 16; }
// doubles of global scratch per instance in flight of aba2_instance
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int aba2_scratch_doubles(const DevModelT<double>& m) { return RBD_ABA2_REC * (m.njoints - 1) + RBD_ABA2_LREC * (m.max_depth + 1); }

// 8-byte words the vector mapping kernels stage behind the joint table for the per-output tables of n_out outputs:
// placements (12 doubles each), output indices and identity flags (one int each, padded to even counts)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int table_words(int n_out) { return 12 * n_out + (n_out + 1) / 2 * 2; }

// applyJT ring kernel (rbd_kernels.cu: applyJT_ring_kernel; plan + ring program: rbd_model_build.h plan_jt_ring)
struct JtRing {
  static constexpr int kWarps = 8;       // block size when the sweep state needs the whole tensor-memory share ...
  static constexpr int kWarpsSmall = 12; // ... and when it fits a third less (small robots: three warps per scheduler)
  static constexpr int kSlotBytes = 6 * 32 * 8;           // one wrench of the 32 instances of a tile
  static constexpr int kStageBytes = 4 * kSlotBytes;      // one piece: <= RBD_JT_GMAX = 4 outputs of one joint
  static constexpr int kMaxStages = 8;
  static constexpr int kMinStages = 2;
};

// Per-output tables (global memory; only the mapping kernels read them).
struct DevTables {
  const int* sorted_k;       // [n_out] output index, grouped by parent joint (DevJoint::out_begin/out_end)
  const double* sorted_pl;   // [n_out][12] frame placement in the parent joint frame, same order
  const int* sorted_ident;   // [n_out] 1 when that placement's rotation is exactly I (the output shares the joint's
                             // rotation: one rotation->quaternion conversion per joint serves all such outputs)
  const int* outk_parent;    // [n_out] parent joint of output k
  const double* outk_pl;     // [n_out][12] frame placement of output k
};

// Scratch record of the forward-dynamics sweep (aba_instance, rbd_sweeps_body.h): doubles per joint / per branch joint
#define RBD_ABA_REC 45
#define RBD_ABA_BRANCH 18
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int aba_scratch_doubles(const DevModel& m) { return RBD_ABA_REC * (m.njoints - 1) + RBD_ABA_BRANCH * m.nbranch; }

// Stack-slot geometry (doubles): A is 6 wide (1-dof joints: [p x a ; a] world frame) or 12 (free-flyer: R|p).
#define RBD_SLOT_A(type) ((type) == JT_FF ? 12 : 6)
#define RBD_BRANCH_EVAL 24 /* R(9) p(3) v(6) a(6) */
#define RBD_BRANCH_FK 12   /* R(9) p(3) */
#define RBD_BRANCH_J 18    /* R(9) p(3) V(6) */

// Warps per block of ONE launch of a persistent one-block-per-SM kernel planned for W warps (<= W): the time of a tile
// grows almost linearly with the warps sharing the SM (t(w) ~ c + k w, c / k ~ 2.5: measured on the Panda eval,
// rbd_kernels_arm.cu), so a batch of only a few tiles per warp is fastest with the block size that leaves the fewest warps
// idle in the last round (N = 65 536 = 2 048 tiles: two full rounds of 7-warp blocks, but 1.15 rounds of 12-warp blocks).
inline int pick_warps_per_block(int W, long long ntiles, int num_sms) {
  int wpb = W;
  double best = 0;
  for (int w = W; w >= 4; --w) {
    const long long rounds = (ntiles + (long long)num_sms * w - 1) / ((long long)num_sms * w);
    const double cost = (double)rounds * (2.5 + w);
    if (w == W || cost < best - 1e-9) { best = cost; wpb = w; }
  }
  return wpb;
}
