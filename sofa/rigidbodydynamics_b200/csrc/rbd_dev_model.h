// rbd_dev_model.h — kinematic-tree constants as the kernels see them.
//
// The whole joint table travels as a __grid_constant__ kernel parameter, i.e. it is staged in the constant
// bank of every launch ("tree constants staged once into constant memory", BASELINE north-star): all threads
// of a warp read the same joint record at the same time, which is exactly the broadcast access pattern the
// constant cache serves, and it keeps L1/shared memory free for per-instance state.
// Replaces the pinocchio::Model the reference shares between loader and mapping
// (reference KinematicChainMapping.inl:304-311, URDFModelLoader.cpp:263-268).
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
struct alignas(16) DevJoint {
  double pl[12];    // jointPlacement: R row-major (9), p (3)
  double axis[3];   // unit axis in the joint frame
  double Y[10];     // m, c(3), Ic (xx,xy,yy,xz,yz,zz)
  double msub;      // mass of the whole subtree below (and including) this joint: the composite-inertia mass is a
                    // model constant, so the Y-stack carries 9 doubles per level, not 10
  int parent, type, idx_q, idx_v;
  int nv;           // dof of this joint (0 universe, 1, or 6)
  int nchild;       // number of child joints
  int slot[2];      // stack-slot offset (doubles) of a joint that has children: [0] full layout (A|f|Y), [1] lite (A|f);
                    // leaves never touch the stack (they are emitted from registers)
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
// ColProg entries (int32).  Ancestor entry, root first: row (idx_v of the ancestor) | free-flyer flag | A-stack offset.
// Zero segment: first row | length.  Rows not covered by either belong to the joint itself.
#define RBD_PROG_ROW(e) ((e) & 0xffff)
#define RBD_PROG_FF(e) (((e) >> 16) & 1)
#define RBD_PROG_OFFA(e) ((e) >> 17)
#define RBD_PROG_ANC(row, ff, offA) ((row) | ((ff) << 16) | ((offA) << 17))
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
};
#define RBD_BOFF_TMEM (1 << 30)
#define RBD_BOFF_NOVA (1 << 29) /* v|a not saved: recomputed when the sweep returns (joint attached to the universe) */
#define RBD_BOFF_RPOMI (1 << 28) /* R|p not saved: re-read from the oMi output this thread wrote (oMi requested) */
#define RBD_BOFF_MASK ((1 << 28) - 1)

struct alignas(16) DevModel {
  int njoints, nq, nv, n_out;
  int slot_total[2];  // doubles of stack slots per instance for the two layouts
  int nbranch;
  int has_ff_root;    // joint 1 is a free-flyer attached to the universe
  int nq_joints, nv_joints;
  int max_depth, max_depth_internal;  // longest joint chain; deepest joint that has children
  int prog_words_off, prog_len;       // eval blob: 8-byte-word offset of the ColProg table behind the joint records; ints
  double gravity[3];
  EvalPlan plan;
  DevJoint joint[RBD_MAX_JOINTS];
};

// bytes of a DevModel that are meaningful (header + the first njoints records): what the eval kernel stages
inline size_t dev_model_bytes(const DevModel& m) {
  return sizeof(DevModel) - sizeof(DevJoint) * (size_t)(RBD_MAX_JOINTS - m.njoints);
}

// shared-memory bytes the eval kernel reserves in front of the stacks: the staged table rounded to 16 bytes plus two
// spare 8-byte words (tensor-memory base address)
inline size_t eval_model_bytes(const DevModel& m) {
  return (dev_model_bytes(m) + 15) / 16 * 16 + ((size_t)m.prog_len * 4 + 15) / 16 * 16 + 16;
}

// 8-byte words the vector mapping kernels stage behind the joint table for the per-output tables of n_out outputs:
// placements (12 doubles each), output indices and identity flags (one int each, padded to even counts)
#if defined(__CUDACC__)
__host__ __device__
#endif
inline int table_words(int n_out) { return 12 * n_out + (n_out + 1) / 2 * 2; }

// Per-output tables (global memory; only the mapping kernels read them).
struct DevTables {
  const int* sorted_k;       // [n_out] output index, grouped by parent joint (DevJoint::out_begin/out_end)
  const double* sorted_pl;   // [n_out][12] frame placement in the parent joint frame, same order
  const int* sorted_ident;   // [n_out] 1 when that placement's rotation is exactly I (the output shares the joint's
                             // rotation: one rotation->quaternion conversion per joint serves all such outputs)
  const int* outk_parent;    // [n_out] parent joint of output k
  const double* outk_pl;     // [n_out][12] frame placement of output k
};

// Stack-slot geometry (doubles): A is 6 wide (1-dof joints: [p x a ; a] world frame) or 12 (free-flyer: R|p).
#define RBD_SLOT_A(type) ((type) == JT_FF ? 12 : 6)
#define RBD_BRANCH_EVAL 24 /* R(9) p(3) v(6) a(6) */
#define RBD_BRANCH_FK 12   /* R(9) p(3) */
#define RBD_BRANCH_J 18    /* R(9) p(3) V(6) */
