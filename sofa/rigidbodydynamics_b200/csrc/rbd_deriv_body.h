// rbd_deriv_body.h — analytical derivatives of inverse dynamics (drnea_instance): d tau / d q, d tau / d v [, d tau / d a =
// M] of tau = rnea(q, v, a), pinocchio::computeRNEADerivatives (Carpentier & Mansard, RSS 2018; SURVEY.md §8(f) rank 4 — no
// call site in the reference).  Included from rbd_sweeps_body.h inside namespace rbd (FP64 only).
//
// Formulation.  World frame like every sweep here.  With J_c the world subspace column of dof c, P its joint's parent,
//     dVdq_c = v_P x J_c          dAdq_c = a_P x J_c + v_P x dVdq_c          dAdv_c = v_joint(c) x J_c + dVdq_c
// (a includes the -g offset; a_universe = -g, v_universe = 0), and for joint j with the COMPOSITE quantities of its
// subtree  Y^c = sum Y_k,  B^c = sum (v_k x* Y_k - Y_k v_k x + (Y_k v_k) xbar*),  F^c = sum f_k :
//     dFdv_c = B^c J_c + Y^c dAdv_c        dFdq_c = B^c dVdq_c + Y^c dAdq_c        (c a dof of j)
//     row j, column c' of an ancestor:     dtau_dq = (Y^c J_c) . dAdq_c' + (B^c^T J_c) . dVdq_c'
//                                          dtau_dv = (Y^c J_c) . dAdv_c' + (B^c^T J_c) . J_c'
//     row of an ancestor-or-self r, column c:  dtau_dv = J_r . dFdv_c,  dtau_dq = J_r . (dFdq_c [+ J_c x* F^c for r != c's joint])
// which is pinocchio's ComputeRNEADerivativesBackwardStep regrouped per (joint, ancestor) PAIR: every entry that involves
// joint j and one of its ancestors is produced when the depth-first sweep leaves j, from j's composite quantities and
// the ancestor's record on the current root-to-leaf path.  Nothing of a finished subtree is kept: O(depth) state, as in
// the fused eval and aba2.
//
// Compact forms.  Y^c is a rigid-body inertia (mass = DevJoint::msub, h, I_O: 9 doubles).  B^c splits into a symmetric
// part, which is the TIME DERIVATIVE of a rigid-body inertia moving with v (v x* Y - Y v x = dY/dt) and therefore again of
// inertia form with zero mass rate: (dh, dI_O), 9 doubles — and an antisymmetric part m -> m x* H with H = sum Y_k v_k, 6
// doubles, linear in the momenta.  15 doubles instead of pinocchio's dense 6 x 6 doYcrb; B m = dY m + m x* H,
// B^T m = dY m - m x* H.
//
// Sweep state, one record per DEPTH LEVEL that holds a joint with children (plan_drnea), split by how often it is touched:
//   HOT, on chip (shared / tensor memory; DevJoint::offA, bit RBD_BOFF_TMEM = tensor memory) — read by the pair loop of
//   every descendant:        1-dof joint [v | a (12) | J (6)]       free-flyer root [v | a (12) | R | p (12)]
//   COLD, in the thread's global scratch record X (DevJoint::boff; L2-resident: written once, read once or twice):
//     one child   -> the joint's own Y (h, I_O: 9), re-read when the sweep leaves the joint;
//     >= 2        -> accumulator Y^c (9) | dY^c (9) | H (6) | F (6), initialised with the joint's own body, the finished
//                    child subtrees ADD into it without reading it back (red.global.add: nothing waits), + R | p (12) of a
//                    1-dof joint for the restart of the next branch.
//   18 doubles per level on chip instead of 27 .. 60: Talos (ten levels) runs 8 warps per SM instead of 4, and it is the
//   number of warps that hides the issue latency of this sweep (first version, everything on chip: 11.8 ms per 2^20).
// dVdq / dAdq of an ancestor are recomputed in the pair loop from its J and its parent's v | a (3 cross products) rather
// than stored.  Measured alternatives on Talos N = 2^20 (this version: 8.4 ms), all slower:
//   hot records [J | dVdq | dAdq] with v | a moved to the cold records (21 % fewer FP64 instructions, 12 more doubles of
//   scratch traffic per joint)                                                                             9.1 ms
//   row / column base addresses instead of an index multiplication per store (10 more live registers)      9.9 ms
//   the pair loop as two walks, column entries then row entries (fewer spills, less work in flight per walk) 9.2 ms
// — at 255 registers per thread every extra live value costs more than the instructions it saves.
//
// Outputs: DQ, DV: nv x nv, field r nv + c (row-major like pinocchio's RowMatrixXs), structural zeros WRITTEN (dofs on
// different branches) — or, SPARSE, the structurally non-zero entries only, compressed by rows (rbd_model_drnea_pattern:
// 700 of Talos' 1 444 entries per matrix; no zero fill); M (optional, dense format only): packed upper triangle like rbd_crba_soa.  A free-flyer only as joint 1 on the universe.
// All lanes of a warp run the sweep together (tensor-memory accesses are warp-collective).

#define RBD_DR_HOT 18       /* on-chip level record of a 1-dof joint: v | a | J                                       */
#define RBD_DR_HOT_FF 24    /* ... of the free-flyer root: v | a | R | p                                              */
#define RBD_DR_COLD_ONE 9   /* scratch record, one child: Y                                                          */
#define RBD_DR_COLD_BRANCH 42 /* ... >= 2 children: accumulator (30) | R | p (12; unused below a free-flyer)           */
#define RBD_DR_SLOTS 3      /* joints of lookahead of the asynchronous input ring (3 doubles each; == kDrneaSlots);
                               9 staging doubles for a parent's cold Y sit behind the ring                             */

// motion x motion: [w_a x v_b + v_a x w_b ; w_a x w_b]
RBD_HD void mxm(const real* a, const real* b, real* o) {
  real tx, ty, tz;
  RBD_CROSS(o[0], o[1], o[2], a[3], a[4], a[5], b[0], b[1], b[2]);
  RBD_CROSS(tx, ty, tz, a[0], a[1], a[2], b[3], b[4], b[5]);
  o[0] += tx; o[1] += ty; o[2] += tz;
  RBD_CROSS(o[3], o[4], o[5], a[3], a[4], a[5], b[3], b[4], b[5]);
}
// motion x* force: [w x f ; w x n + v x f]
RBD_HD void mxf(const real* a, const real* f, real* o) {
  real tx, ty, tz;
  RBD_CROSS(o[0], o[1], o[2], a[3], a[4], a[5], f[0], f[1], f[2]);
  RBD_CROSS(o[3], o[4], o[5], a[3], a[4], a[5], f[3], f[4], f[5]);
  RBD_CROSS(tx, ty, tz, a[0], a[1], a[2], f[0], f[1], f[2]);
  o[3] += tx; o[4] += ty; o[5] += tz;
}
// time derivative of the world inertia Y = (m, h, I_O) of a body moving with v: dY = v x* Y - Y v x, again of inertia form
// with zero mass: dY[0] = 0, dh = m v_l + w x h, dI = [w]x I - I [w]x - (h v_l^T + v_l h^T) + 2 (h . v_l) 1
RBD_HD void inertia_rate(const real* Y, const real* v, real* dY) {
  const real hx = Y[1], hy = Y[2], hz = Y[3];
  const real wx = v[3], wy = v[4], wz = v[5];
  dY[0] = 0;
  dY[1] = Y[0] * v[0] + (wy * hz - wz * hy);
  dY[2] = Y[0] * v[1] + (wz * hx - wx * hz);
  dY[3] = Y[0] * v[2] + (wx * hy - wy * hx);
  // K = [w]x I, I = [[Y4, Y5, Y7], [Y5, Y6, Y8], [Y7, Y8, Y9]]
  const real k00 = wy * Y[7] - wz * Y[5], k01 = wy * Y[8] - wz * Y[6], k02 = wy * Y[9] - wz * Y[8];
  const real k10 = wz * Y[4] - wx * Y[7], k11 = wz * Y[5] - wx * Y[8], k12 = wz * Y[7] - wx * Y[9];
  const real k20 = wx * Y[5] - wy * Y[4], k21 = wx * Y[6] - wy * Y[5], k22 = wx * Y[8] - wy * Y[7];
  const real hv = hx * v[0] + hy * v[1] + hz * v[2];
  dY[4] = 2 * (k00 - hx * v[0] + hv);
  dY[5] = k01 + k10 - (hx * v[1] + hy * v[0]);
  dY[6] = 2 * (k11 - hy * v[1] + hv);
  dY[7] = k02 + k20 - (hx * v[2] + hz * v[0]);
  dY[8] = k12 + k21 - (hy * v[2] + hz * v[1]);
  dY[9] = 2 * (k22 - hz * v[2] + hv);
}
// own contribution of a body to the composite quantities: Yc[1..9] | dYc[1..9] | H | F  (30 doubles, C[0..29])
RBD_HD void dr_own(const real* Y, const real* vel, const real* acc, real* C) {
  real dY[10], h[6], f[6], t[6];
  inertia_rate(Y, vel, dY);
  inertia_mul(Y, vel, h);
  inertia_mul(Y, acc, f);
  mxf(vel, h, t);
#pragma unroll
  for (int e = 0; e < 9; ++e) { C[e] = Y[1 + e]; C[9 + e] = dY[1 + e]; }
#pragma unroll
  for (int e = 0; e < 6; ++e) { C[18 + e] = h[e]; C[24 + e] = f[e] + t[e]; }
}

// Everything joint `cur` (1-dof) contributes when the sweep leaves it: its diagonal entry, the entries it shares with
// each ancestor (both triangles), the structural zeros it shares with the dofs of earlier branches.
// A: J of cur; va: v | a of cur's parent; C: composite of cur's subtree (dr_own layout).
// SPARSE: the support-only format (rbd_model_drnea_pattern: compressed by rows, structural zeros not stored).  Row of dof e
// of joint j starts at RBD_DR_ROWBASE(j) + e RBD_DR_ROWLEN(j) and holds the ancestors' dofs (root first), then the dofs
// [idx_v(j), idx_v(j) + nv_subtree(j)): entry (row of j, dof e' of ancestor k) sits at offset RBD_DR_NANC(k) + e', entry
// (row of j, dof c of its subtree) at RBD_DR_NANC(j) + c - idx_v(j).
#define RBD_DR_ROWBASE(j) ((j).slot[0])
#define RBD_DR_NANC(j) ((j).slot[1])
#define RBD_DR_ROWLEN(j) ((j).out_begin)
template <bool SPARSE, bool WITH_M, class FLD>
RBD_HD void dr_pop_1dof(const DevModel& m, int cur, const real* A, const real* va, const real* C, const FLD& DQ,
                        const FLD& DV, const FLD& M, const Stack& S, const TmStack& T) {
  const DevJoint& jc = m.joint[cur];
  const int nv = m.nv, iv = jc.idx_v;
  // field of entry (iv, column c of ancestor k) / (row r of ancestor k, iv)
  const int rowj = SPARSE ? RBD_DR_ROWBASE(jc) : iv * nv;
  real Yc[10], dYc[10];
  Yc[0] = jc.msub; dYc[0] = 0;
#pragma unroll
  for (int e = 0; e < 9; ++e) { Yc[1 + e] = C[e]; dYc[1 + e] = C[9 + e]; }
  const real* H = C + 18;
  const real* F = C + 24;
  real dV[6], dA[6], u[6], w[6], dFv[6], dFq[6];
  {
    real t[6], cb[6];
    mxm(va, A, dV);
    mxm(va + 6, A, dA);
    mxm(va, dV, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) dA[e] += t[e];
    inertia_mul(Yc, A, u);
    inertia_mul(dYc, A, w);   // dY J
    mxf(A, H, cb);            // J x* H
    inertia_mul(Yc, dV, t);   // Y dVdq  (dAdv = 2 dVdq for a 1-dof joint)
#pragma unroll
    for (int e = 0; e < 6; ++e) { dFv[e] = w[e] + cb[e] + 2 * t[e]; w[e] -= cb[e]; }
    inertia_mul(dYc, dV, dFq);
    mxf(dV, H, cb);
    inertia_mul(Yc, dA, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) dFq[e] += cb[e] + t[e];
  }
  DV.st(rowj + (SPARSE ? RBD_DR_NANC(jc) : iv), dot6(A, dFv));
  DQ.st(rowj + (SPARSE ? RBD_DR_NANC(jc) : iv), dot6(A, dFq));
  if (WITH_M) M.st(iv * (iv + 1) / 2 + iv, dot6(A, u));
  {
    real t[6];
    mxf(A, F, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) dFq[e] += t[e];
  }
  // ancestors, nearest first; the dofs between two consecutive ancestors belong to earlier branches: structural zeros
  int prev = iv;
  for (int k = jc.parent; k != 0;) {
    const DevJoint& jk = m.joint[k];
    const int ik = jk.idx_v;
    if (!SPARSE) {
      for (int c = ik + jk.nv; c < prev; ++c) {
        DQ.st(iv * nv + c, real(0)); DV.st(iv * nv + c, real(0));
        DQ.st(c * nv + iv, real(0)); DV.st(c * nv + iv, real(0));
        if (WITH_M) M.st(iv * (iv + 1) / 2 + c, real(0));
      }
    }
    prev = ik;
    const int rjk = SPARSE ? RBD_DR_NANC(jk) : ik;  // entry (iv, first dof of k) within row iv
    // entry (first dof of k, iv), and the distance to the same column in k's next row
    const int ckj = SPARSE ? RBD_DR_ROWBASE(jk) + RBD_DR_NANC(jk) + (iv - ik) : ik * nv + iv;
    const int ckstep = SPARSE ? RBD_DR_ROWLEN(jk) : nv;
    if (RBD_T_IS_FF(jk.type)) {  // the root: six columns from R | p; dVdq = 0, dAdq_c = a0 x J_c, dAdv_c = v_root x J_c
      real Rp[12], vr[6], o[6], t[6];
      lev_ld<12>(S, T, jk.offA + 12, Rp);
      lev_ld<6>(S, T, jk.offA, vr);
      ff_rows(Rp, Rp + 9, dFq, o);
#pragma unroll
      for (int e = 0; e < 6; ++e) DQ.st(ckj + e * ckstep, o[e]);
      ff_rows(Rp, Rp + 9, dFv, o);
#pragma unroll
      for (int e = 0; e < 6; ++e) DV.st(ckj + e * ckstep, o[e]);
      if (WITH_M) {
        ff_rows(Rp, Rp + 9, u, o);
#pragma unroll
        for (int e = 0; e < 6; ++e) M.st(iv * (iv + 1) / 2 + ik + e, o[e]);
      }
      // u . (a0 x J_c) = -(a0 x* u) . J_c,  a0 = (-g, 0):  -(a0 x* u) = [0 ; g x u_l]
      t[0] = t[1] = t[2] = 0;
      RBD_CROSS(t[3], t[4], t[5], m.gravity[0], m.gravity[1], m.gravity[2], u[0], u[1], u[2]);
      ff_rows(Rp, Rp + 9, t, o);
#pragma unroll
      for (int e = 0; e < 6; ++e) DQ.st(rowj + rjk + e, o[e]);
      // u . (v_root x J_c) + w . J_c = (w - v_root x* u) . J_c
      mxf(vr, u, t);
#pragma unroll
      for (int e = 0; e < 6; ++e) t[e] = w[e] - t[e];
      ff_rows(Rp, Rp + 9, t, o);
#pragma unroll
      for (int e = 0; e < 6; ++e) DV.st(rowj + rjk + e, o[e]);
      k = jk.parent;
      continue;
    }
    real Jk[6], vak[12], dVk[6], dAk[6], t[6];
    lev_ld<6>(S, T, jk.offA + 12, Jk);
    const int kp = jk.parent;
    if (kp == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) vak[e] = 0;
      vak[6] = -m.gravity[0]; vak[7] = -m.gravity[1]; vak[8] = -m.gravity[2];
    } else {
      lev_ld<12>(S, T, m.joint[kp].offA, vak);
    }
    mxm(vak, Jk, dVk);
    mxm(vak + 6, Jk, dAk);
    mxm(vak, dVk, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) dAk[e] += t[e];
    DQ.st(ckj, dot6(Jk, dFq));
    DV.st(ckj, dot6(Jk, dFv));
    if (WITH_M) M.st(iv * (iv + 1) / 2 + ik, dot6(Jk, u));
    DQ.st(rowj + rjk, dot6(u, dAk) + dot6(w, dVk));
    DV.st(rowj + rjk, 2 * dot6(u, dVk) + dot6(w, Jk));
    k = kp;
  }
  if (!SPARSE) {
    for (int c = 0; c < prev; ++c) {
      DQ.st(iv * nv + c, real(0)); DV.st(iv * nv + c, real(0));
      DQ.st(c * nv + iv, real(0)); DV.st(c * nv + iv, real(0));
      if (WITH_M) M.st(iv * (iv + 1) / 2 + c, real(0));
    }
  }
}

// ... and the root free-flyer's own 6 x 6 blocks: dVdq = 0, dAdq_c = a0 x J_c, dAdv_c = v x J_c (parent = universe)
template <bool SPARSE, bool WITH_M, class FLD>
RBD_HD void dr_pop_ff(const DevModel& m, int cur, const real* Rp, const real* vel, const real* C, const FLD& DQ,
                      const FLD& DV, const FLD& M) {
  const DevJoint& jc = m.joint[cur];
  const int nv = m.nv, iv = jc.idx_v;
  const int blk = SPARSE ? RBD_DR_ROWBASE(jc) : iv * nv + iv, rstep = SPARSE ? RBD_DR_ROWLEN(jc) : nv;  // entry (iv + r, iv + cc) at blk + r rstep + cc
  real Yc[10], dYc[10], a0[6];
  Yc[0] = jc.msub; dYc[0] = 0;
#pragma unroll
  for (int e = 0; e < 9; ++e) { Yc[1 + e] = C[e]; dYc[1 + e] = C[9 + e]; }
  a0[0] = -m.gravity[0]; a0[1] = -m.gravity[1]; a0[2] = -m.gravity[2]; a0[3] = a0[4] = a0[5] = 0;
  const real* H = C + 18;
#pragma unroll 1  // once per instance: rolled, the six columns are a sixth of the code (instruction-cache pressure)
  for (int cc = 0; cc < 6; ++cc) {
    real Jc[6], t[6], dF[6], o[6];
    RBD_FF_COL(Jc, Rp, (Rp + 9), cc);
    // dFdv = dY J + J x* H + Y (v x J)
    inertia_mul(dYc, Jc, dF);
    mxf(Jc, H, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) dF[e] += t[e];
    mxm(vel, Jc, t);
    inertia_mul(Yc, t, o);
#pragma unroll
    for (int e = 0; e < 6; ++e) dF[e] += o[e];
    ff_rows(Rp, Rp + 9, dF, o);
#pragma unroll
    for (int r = 0; r < 6; ++r) DV.st(blk + r * rstep + cc, o[r]);
    // dFdq = Y (a0 x J)
    mxm(a0, Jc, t);
    inertia_mul(Yc, t, dF);
    ff_rows(Rp, Rp + 9, dF, o);
#pragma unroll
    for (int r = 0; r < 6; ++r) DQ.st(blk + r * rstep + cc, o[r]);
    if (WITH_M) {
      inertia_mul(Yc, Jc, dF);
      ff_rows(Rp, Rp + 9, dF, o);
#pragma unroll
      for (int r = 0; r <= cc; ++r) M.st((iv + cc) * (iv + cc + 1) / 2 + iv + r, o[r]);
    }
  }
}

// WITH_M: the M output is requested (dense format only); compiled out otherwise — at 255 registers the dead stores and
// their addresses cost 10 % (A/B s4h)
template <bool SPARSE, bool WITH_M, class FLD, class SCR>
RBD_HD void drnea_instance(const DevModel& m, const FLD& q, const FLD& v, const FLD& a, const FLD& DQ, const FLD& DV,
                           const FLD& M, const SCR& X, const Stack& S, const TmStack& T, const InRing& inr) {
  const int nj = m.njoints;
  real R[9], p[3], vel[6], acc[6];
#define RBD_DR_ROOT()                                                                           \
  do {                                                                                          \
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;   \
    p[0] = p[1] = p[2] = 0;                                                                     \
    _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) { vel[e_] = 0; acc[e_] = 0; }              \
    acc[0] = -m.gravity[0]; acc[1] = -m.gravity[1]; acc[2] = -m.gravity[2];                     \
  } while (0)
  RBD_DR_ROOT();
  // Inputs of the 1-dof joints: through an asynchronous ring in shared memory (cp.async, no destination register)
  // RBD_DR_SLOTS joints ahead of the sweep when the plan has room for it (InRing, as in the fused eval: under the kernel's
  // own store stream a global read takes thousands of cycles, and a register prefetch is spilled by ptxas at 255
  // registers — the spill store then waits for the load: profile s3e); else (tests/emul) read at the point of use.
  const bool ring = inr.p != nullptr;
#define RBD_DR_ISSUE(J)                                                                         \
  do {                                                                                          \
    if ((J) < nj && !RBD_T_IS_FF(m.joint[J].type)) {                                            \
      const DevJoint& jx_ = m.joint[J];                                                         \
      const int sl_ = 3 * ((J) % RBD_DR_SLOTS);                                                 \
      cp_async_real(inr, sl_, q.at(jx_.idx_q));                                                 \
      cp_async_real(inr, sl_ + 1, v.at(jx_.idx_v));                                             \
      cp_async_real(inr, sl_ + 2, a.at(jx_.idx_v));                                             \
    }                                                                                           \
    cp_async_commit();                                                                          \
  } while (0)
  if (ring) {
#pragma unroll
    for (int s_ = 1; s_ <= RBD_DR_SLOTS; ++s_) RBD_DR_ISSUE(s_);
  }
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const bool ff = RBD_T_IS_FF(jn.type);
    real q0 = 0, q1 = 0, qd = 0, qdd = 0;
    if (ring) {
      cp_async_wait<RBD_DR_SLOTS - 1>();
      if (!ff) {
        const int sl = 3 * (i % RBD_DR_SLOTS);
        q0 = inr.p[sl * RBD_WARP]; qd = inr.p[(sl + 1) * RBD_WARP]; qdd = inr.p[(sl + 2) * RBD_WARP];
        if (is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
      }
      RBD_DR_ISSUE(i + RBD_DR_SLOTS);
    } else if (!ff) {
      q0 = q.ld(jn.idx_q); qd = v.ld(jn.idx_v); qdd = a.ld(jn.idx_v);
      if (is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
    }
    real aw[3], A[6], va[12];
    fk_joint(jn, R, p, q0, q1, q, aw);
#pragma unroll
    for (int e = 0; e < 6; ++e) { va[e] = vel[e]; va[6 + e] = acc[e]; }  // the parent's v | a
    if (!ff) {
      subspace_1dof(jn.type, p, aw, A);
      real vj[6], cr[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) { vj[e] = A[e] * qd; vel[e] += vj[e]; }
      mxm(vel, vj, cr);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] += A[e] * qdd + cr[e];
    } else {  // root free-flyer: v = vJ, hence v x vJ = 0
      real t[6];
      ff_act(R, p, v, jn.idx_v, vel);
      ff_act(R, p, a, jn.idx_v, t);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] += t[e];
    }
    real Y[10];
    inertia_world(jn, R, p, Y);
    if (jn.nchild > 0) {
      real t[12];
#pragma unroll
      for (int e = 0; e < 6; ++e) { t[e] = vel[e]; t[6 + e] = acc[e]; }
      lev_st<12>(S, T, jn.offA, t);
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) t[9 + e] = p[e];
      if (!ff) lev_st<6>(S, T, jn.offA + 12, A);
      else lev_st<12>(S, T, jn.offA + 12, t);
      const SCR Xc = X.at(jn.boff);
      if (jn.nchild == 1) {
#pragma unroll
        for (int e = 0; e < 9; ++e) Xc.st(e, Y[1 + e]);
      } else {
        real C[30];
        dr_own(Y, vel, acc, C);
#pragma unroll
        for (int e = 0; e < 30; ++e) Xc.st(e, C[e]);
        if (!ff) {
#pragma unroll
          for (int e = 0; e < 12; ++e) Xc.st(30 + e, t[e]);
        }
      }
      continue;
    }
    // ---- leaf: every joint this leaf completes is finished now, bottom-up --------------------------------------------
    const int stop_at = (i + 1 < nj) ? m.joint[i + 1].parent : 0;  // the joint that still has a child to come
    real C[30];
    dr_own(Y, vel, acc, C);
    int cur = i;
    for (;;) {
      const DevJoint& jc = m.joint[cur];
      if (RBD_T_IS_FF(jc.type)) {
        real Rp[12], vr[6];
        if (cur == i) {  // a free-flyer leaf: R | p | v are still in registers
#pragma unroll
          for (int e = 0; e < 9; ++e) Rp[e] = R[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) Rp[9 + e] = p[e];
#pragma unroll
          for (int e = 0; e < 6; ++e) vr[e] = vel[e];
        } else {
          lev_ld<12>(S, T, jc.offA + 12, Rp);
          lev_ld<6>(S, T, jc.offA, vr);
        }
        dr_pop_ff<SPARSE, WITH_M>(m, cur, Rp, vr, C, DQ, DV, M);
        break;  // (the host checks that a free-flyer is joint 1 on the universe)
      }
      const int pp = jc.parent;
      // the own inertia of a one-child parent that this pop completes comes back from the scratch record: start the copy
      // into the staging slots now, behind the pair loop it has arrived (a plain load here would hold 18 registers across
      // the loop — ptxas spills them, and the spill store waits for the load)
      const bool merge_own = pp != 0 && pp != stop_at && m.joint[pp].nchild == 1;
      if (ring && merge_own) {
        const SCR Xp = X.at(m.joint[pp].boff);
#pragma unroll
        for (int e = 0; e < 9; ++e) cp_async_real(inr, 3 * RBD_DR_SLOTS + e, Xp.ptr(e));
        cp_async_commit();
      }
      dr_pop_1dof<SPARSE, WITH_M>(m, cur, A, va, C, DQ, DV, M, S, T);
      if (pp == 0) break;
      const DevJoint& jp = m.joint[pp];
      const SCR Xp = X.at(jp.boff);
      if (pp == stop_at) {  // the parent has another child to come: add the contribution to its accumulator (no read back)
#pragma unroll
        for (int e = 0; e < 30; ++e) Xp.add(e, C[e]);
        break;
      }
      // the parent is complete: its accumulator (>= 2 children) or its own body (one child) joins in registers
      if (jp.nchild >= 2) {
#pragma unroll
        for (int e = 0; e < 30; ++e) C[e] += Xp.ld(e);
      } else {
        real Yp[10], t[30], vap[12];
        Yp[0] = jp.Y[0];
        if (ring) {
          cp_async_wait<0>();
#pragma unroll
          for (int e = 0; e < 9; ++e) Yp[1 + e] = inr.p[(3 * RBD_DR_SLOTS + e) * RBD_WARP];
        } else {
#pragma unroll
          for (int e = 0; e < 9; ++e) Yp[1 + e] = Xp.ld(e);
        }
        lev_ld<12>(S, T, jp.offA, vap);  // (re-read: keeping the parent's v | a in registers across the pair loop spills)
        dr_own(Yp, vap, vap + 6, t);
#pragma unroll
        for (int e = 0; e < 30; ++e) C[e] += t[e];
      }
      cur = pp;
      if (!RBD_T_IS_FF(jp.type)) {
        lev_ld<6>(S, T, jp.offA + 12, A);
        const int gp = jp.parent;
        if (gp == 0) {
#pragma unroll
          for (int e = 0; e < 12; ++e) va[e] = 0;
          va[6] = -m.gravity[0]; va[7] = -m.gravity[1]; va[8] = -m.gravity[2];
        } else {
          lev_ld<12>(S, T, m.joint[gp].offA, va);
        }
      }
    }
    // the joint after a leaf starts a new branch: restore R | p | v | a of its parent (or the universe)
    if (i + 1 < nj) {
      if (stop_at == 0) {
        RBD_DR_ROOT();
      } else {
        const DevJoint& js = m.joint[stop_at];
        real t[12];
        lev_ld<12>(S, T, js.offA, t);
#pragma unroll
        for (int e = 0; e < 6; ++e) { vel[e] = t[e]; acc[e] = t[6 + e]; }
        if (RBD_T_IS_FF(js.type)) {
          lev_ld<12>(S, T, js.offA + 12, t);
        } else {
          const SCR Xs = X.at(js.boff);
#pragma unroll
          for (int e = 0; e < 12; ++e) t[e] = Xs.ld(30 + e);
        }
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = t[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = t[9 + e];
      }
    }
  }
#undef RBD_DR_ISSUE
#undef RBD_DR_ROOT
}
