// rbd_dyn_body.h — forward dynamics, second version (aba2_instance).  Included from rbd_sweeps_body.h inside namespace
// rbd (FP64 only); the formulation (world frame, passes 1-3) is the one documented above aba_instance.
//
// What changed against aba_instance: passes 1 and 2 are FUSED on the depth-first order of the joints.  The sweep walks
// root -> leaf keeping, per DEPTH LEVEL that holds a joint with children, one on-chip record
//     one child :  [h (3) | I_O (6) | p^A (6)]                    15 doubles  (the joint's own inertia and bias force)
//     >= 2      :  [I^A (21) | p^A (6)] + restart [R | p | v (18)]  45 doubles  (accumulator of the children's I^a, p^a)
// and at a leaf runs pass 2 upwards at once: U, 1/D, u of the joint, I^a / p^a handed to the parent IN REGISTERS as long
// as that parent is complete (this was its last child), added to the parent's accumulator record otherwise.  Only one
// joint per depth level is open at any time, so records are indexed by level: O(depth) on-chip state (shared memory +
// tensor memory, planned by plan_aba2) instead of the O(njoints) global-memory record of aba_instance.  What pass 3
// needs of EVERY joint after the whole backward pass — A (6), c (6), U (6), 1/D, u = 20 doubles — is the only thing that
// still goes through global memory: written once (A, c in pass 1, U, 1/D, u in pass 2), A and c re-read by the unwind,
// all 20 read by pass 3.  Talos: 10.6 KB + 3.2 KB per instance instead of 29 KB.
// Pass 3 parks the acceleration of a joint with >= 2 children in that joint's (now free) level record.
//
// Level-record offsets: DevJoint::offA (record) / DevJoint::boff (restart record), bit RBD_BOFF_TMEM = tensor memory.
// All lanes of a warp execute the sweep together (tensor-memory accesses are warp-collective): the kernel lets lanes
// past the end of the batch recompute the last instance.

// (RBD_ABA2_REC = 20 doubles per joint of the global pass-3 record: rbd_dev_model.h)
#define RBD_ABA2_LEV1 15 /* level record of a joint with one child                                                */
#define RBD_ABA2_LEVB 27 /* accumulator record of a joint with >= 2 children ...                                  */
#define RBD_ABA2_RST 18  /* ... and its restart record                                                            */

template <int N>
RBD_HD void lev_ld(const Stack& S, const TmStack& T, int off, real* v) {
  if (off & RBD_BOFF_TMEM) {
    T.template ldv<N>(off & RBD_BOFF_MASK, v);
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = S.ld(off + e);
  }
}
template <int N>
RBD_HD void lev_st(const Stack& S, const TmStack& T, int off, const real* v) {
  if (off & RBD_BOFF_TMEM) {
    T.template stv<N>(off & RBD_BOFF_MASK, v);
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) S.st(off + e, v[e]);
  }
}
template <int N>
RBD_HD void lev_add(const Stack& S, const TmStack& T, int off, const real* v) {
  if (off & RBD_BOFF_TMEM) {
    T.template addv<N>(off & RBD_BOFF_MASK, v);
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) S.add(off + e, v[e]);
  }
}

// S += the rigid-body inertia (m, h, I_O) as a symmetric 6 x 6 matrix (sym6_from_inertia, accumulating)
RBD_HD void sym6_add_inertia(real m, const real* h, const real* I, real* S6) {
  S6[s6(0, 0)] += m; S6[s6(1, 1)] += m; S6[s6(2, 2)] += m;
  S6[s6(0, 4)] += h[2]; S6[s6(0, 5)] -= h[1];
  S6[s6(1, 3)] -= h[2]; S6[s6(1, 5)] += h[0];
  S6[s6(2, 3)] += h[1]; S6[s6(2, 4)] -= h[0];
  S6[s6(3, 3)] += I[0]; S6[s6(3, 4)] += I[1]; S6[s6(3, 5)] += I[3];
  S6[s6(4, 4)] += I[2]; S6[s6(4, 5)] += I[4];
  S6[s6(5, 5)] += I[5];
}

// Root free-flyer at the end of the backward pass: (X^T I^A X) qdd = tau - X^T (p^A + I^A a'), a' = -g (c = 0);
// Cholesky in registers.  Rp = R | p of the joint; returns qdd (6) and the joint's world acceleration a' + X qdd.
template <class FLD>
RBD_HD void aba_ff_root_solve(const DevModel& m, const real* Rp, const real* IA, const real* pA, const FLD& tau, int iv,
                              real* b, real* acc) {
  real a0[6], f[6];
  a0[0] = -m.gravity[0]; a0[1] = -m.gravity[1]; a0[2] = -m.gravity[2]; a0[3] = a0[4] = a0[5] = 0;
  sym6_mul(IA, a0, f);
#pragma unroll
  for (int e = 0; e < 6; ++e) f[e] += pA[e];
  ff_rows(Rp, Rp + 9, f, b);
#pragma unroll
  for (int e = 0; e < 6; ++e) b[e] = tau.ld(iv + e) - b[e];
  real D[21];  // upper triangle of X^T I^A X
#pragma unroll
  for (int cc = 0; cc < 6; ++cc) {
    real Xc[6], F[6], col[6];
    RBD_FF_COL(Xc, Rp, (Rp + 9), cc);
    sym6_mul(IA, Xc, F);
    ff_rows(Rp, Rp + 9, F, col);
#pragma unroll
    for (int r = 0; r <= cc; ++r) D[s6(r, cc)] = col[r];
  }
  // D = L L^T (D[s6(r, c)], r <= c, holds L(c, r)), then the two triangular solves
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    real d = D[s6(k, k)];
#pragma unroll
    for (int j = 0; j < k; ++j) d -= D[s6(j, k)] * D[s6(j, k)];
    d = sqrt(d);
    D[s6(k, k)] = d;
    const real di = real(1) / d;
#pragma unroll
    for (int r = k + 1; r < 6; ++r) {
      real t = D[s6(k, r)];
#pragma unroll
      for (int j = 0; j < k; ++j) t -= D[s6(j, r)] * D[s6(j, k)];
      D[s6(k, r)] = t * di;
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    real t = b[k];
#pragma unroll
    for (int j = 0; j < k; ++j) t -= D[s6(j, k)] * b[j];
    b[k] = t / D[s6(k, k)];
  }
#pragma unroll
  for (int k = 5; k >= 0; --k) {
    real t = b[k];
#pragma unroll
    for (int j = k + 1; j < 6; ++j) t -= D[s6(k, j)] * b[j];
    b[k] = t / D[s6(k, k)];
  }
  real al[3], aa[3], tx, ty, tz;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    al[e] = Rp[3 * e] * b[0] + Rp[3 * e + 1] * b[1] + Rp[3 * e + 2] * b[2];
    aa[e] = Rp[3 * e] * b[3] + Rp[3 * e + 1] * b[4] + Rp[3 * e + 2] * b[5];
  }
  RBD_CROSS(tx, ty, tz, Rp[9], Rp[10], Rp[11], aa[0], aa[1], aa[2]);
  acc[0] = a0[0] + al[0] + tx; acc[1] = a0[1] + al[1] + ty; acc[2] = a0[2] + al[2] + tz;
  acc[3] = aa[0]; acc[4] = aa[1]; acc[5] = aa[2];
}

// X: this thread's global pass-3 record (ld(e) / st(e, v)), RBD_ABA2_REC doubles per joint; S / T: on-chip level records.
template <class FLD, class SCR>
RBD_HD void aba2_instance(const DevModel& m, const FLD& q, const FLD& v, const FLD& tau, const FLD& ddq, const SCR& X,
                          const Stack& S, const TmStack& T) {
  const int nj = m.njoints;
  real R[9], p[3], vel[6];
#define RBD_ABA_ROOT()                                                                          \
  do {                                                                                          \
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;   \
    p[0] = p[1] = p[2] = 0;                                                                     \
    _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) vel[e_] = 0;                               \
  } while (0)
  RBD_ABA_ROOT();
  // ---- passes 1 + 2 ---------------------------------------------------------------------------------------------
  // (joint angle, velocity and torque of 1-dof joints are fetched one joint ahead: a global load under the kernel's own
  // traffic takes thousands of cycles; record accesses go through a per-joint base with compile-time offsets; R | p | v
  // of the joint a new branch starts from are restored right behind the unwind that precedes it, so that they are dead
  // — and their registers free — during the unwind)
  real nq0 = 0, nqd = 0, ntau = 0;
#define RBD_ABA_PREFETCH(J)                                                                                  \
  do {                                                                                                       \
    const DevJoint& jx_ = m.joint[J];                                                                        \
    if (jx_.type != JT_FF) { nq0 = q.ld(jx_.idx_q); nqd = v.ld(jx_.idx_v); if (jx_.nchild == 0) ntau = tau.ld(jx_.idx_v); } \
  } while (0)
  if (nj > 1) RBD_ABA_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
#if RBD_ABA2_V3
    const SCR Xi = X.at(RBD_ABA2_REC * (nj - 1) + RBD_ABA2_LREC * jn.depth);  // A | c (R | p) of the joint at this depth
#else
    const SCR Xi = X.at(RBD_ABA2_REC * (i - 1));
#endif
    const real q0c = nq0, qdc = nqd, tau0 = ntau;
    if (i + 1 < nj) RBD_ABA_PREFETCH(i + 1);
    const bool ff = jn.type == JT_FF;
    real q0 = q0c, q1 = 0, aw[3];
    if (!ff && is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    real A[6], cr[6], vj[6];
    if (!ff) {
      subspace_1dof(jn.type, p, aw, A);
      const real qd = qdc;
#pragma unroll
      for (int e = 0; e < 6; ++e) { Xi.st(e, A[e]); vj[e] = A[e] * qd; }
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) Xi.st(e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) Xi.st(9 + e, p[e]);
      ff_act(R, p, v, jn.idx_v, vj);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) vel[e] += vj[e];
    if (!ff) {  // c = v x vJ (a root free-flyer has v = vJ, hence c = 0)
      real tx, ty, tz;
      RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
      cr[0] += tx; cr[1] += ty; cr[2] += tz;
      RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
#pragma unroll
      for (int e = 0; e < 6; ++e) Xi.st(6 + e, cr[e]);
    }
    real Y[10], h[6], pA[6];
    inertia_world(jn, R, p, Y);
    inertia_mul(Y, vel, h);
    {
      real tx, ty, tz;
      RBD_CROSS(pA[0], pA[1], pA[2], vel[3], vel[4], vel[5], h[0], h[1], h[2]);
      RBD_CROSS(pA[3], pA[4], pA[5], vel[3], vel[4], vel[5], h[3], h[4], h[5]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], h[0], h[1], h[2]);
      pA[3] += tx; pA[4] += ty; pA[5] += tz;
    }
    if (jn.nchild == 1) {
      real t[RBD_ABA2_LEV1];
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = Y[1 + e];
#pragma unroll
      for (int e = 0; e < 6; ++e) t[9 + e] = pA[e];
      lev_st<RBD_ABA2_LEV1>(S, T, jn.offA, t);
      continue;
    }
    real IA[21 + 6];  // I^A | p^A
    sym6_from_inertia(Y, IA);
#pragma unroll
    for (int e = 0; e < 6; ++e) IA[21 + e] = pA[e];
    if (jn.nchild >= 2) {
      lev_st<RBD_ABA2_LEVB>(S, T, jn.offA, IA);
      real t[RBD_ABA2_RST];
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) t[9 + e] = p[e];
#pragma unroll
      for (int e = 0; e < 6; ++e) t[12 + e] = vel[e];
      lev_st<RBD_ABA2_RST>(S, T, jn.boff, t);
      continue;
    }
    // ---- leaf: pass 2 upwards for every joint this leaf completes -----------------------------------------------
    const int stop_at = (i + 1 < nj) ? m.joint[i + 1].parent : 0;  // the joint that still has a child to come
    int cur = i;
    real tauc = tau0;
    for (;;) {
      const DevJoint& jc = m.joint[cur];
      const SCR Xc = X.at(RBD_ABA2_REC * (cur - 1));
      if (jc.type == JT_FF) {  // the root (the host checks that a free-flyer is joint 1 on the universe)
        real Rp[12], b[6], acc[6];
#if RBD_ABA2_V3
        const SCR Xl = X.at(RBD_ABA2_REC * (nj - 1) + RBD_ABA2_LREC * jc.depth);
#pragma unroll
        for (int e = 0; e < 12; ++e) Rp[e] = Xl.ld(e);
#else
#pragma unroll
        for (int e = 0; e < 12; ++e) Rp[e] = Xc.ld(e);
#endif
        aba_ff_root_solve(m, Rp, IA, IA + 21, tau, jc.idx_v, b, acc);
#pragma unroll
        for (int e = 0; e < 6; ++e) { ddq.st(jc.idx_v + e, b[e]); Xc.st((RBD_ABA2_V3 ? 0 : 12) + e, acc[e]); }
        break;
      }
      const int pp = jc.parent;
      // A, c, tau of the parent are fetched now, a whole joint step ahead of their use (they are in L2 at best)
      real An[6], cn[6], taun = 0;
      const bool fetch_parent = pp != 0 && pp != stop_at && m.joint[pp].type != JT_FF;
#ifndef RBD_ABA2_UNW
#define RBD_ABA2_UNW 0
#endif
      if (RBD_ABA2_UNW && fetch_parent) {
#if RBD_ABA2_V3
        const SCR Xp = X.at(RBD_ABA2_REC * (nj - 1) + RBD_ABA2_LREC * m.joint[pp].depth);
#else
        const SCR Xp = X.at(RBD_ABA2_REC * (pp - 1));
#endif
#pragma unroll
        for (int e = 0; e < 6; ++e) { An[e] = Xp.ld(e); cn[e] = Xp.ld(6 + e); }
        taun = tau.ld(m.joint[pp].idx_v);
      }
      real U[6];
      sym6_mul(IA, A, U);
      const real dinv = real(1) / dot6(A, U);
      const real u = tauc - dot6(A, IA + 21);
#pragma unroll
      for (int e = 0; e < 6; ++e) Xc.st((RBD_ABA2_V3 ? 0 : 12) + e, U[e]);
      Xc.st(RBD_ABA2_V3 ? 6 : 18, dinv);
      Xc.st(RBD_ABA2_V3 ? 7 : 19, u);
      if (pp == 0) break;
      // I^a = I^A - U U^T / D ;  p^a = p^A + I^a c + U u / D
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = r; cc < 6; ++cc) IA[s6(r, cc)] -= U[r] * (U[cc] * dinv);
      {
        real Ic[6];
        sym6_mul(IA, cr, Ic);
        const real ud = u * dinv;
#pragma unroll
        for (int e = 0; e < 6; ++e) IA[21 + e] += Ic[e] + U[e] * ud;
      }
      const DevJoint& jp = m.joint[pp];
      if (pp == stop_at) {  // the parent has another child to come: park the contribution there
        lev_add<RBD_ABA2_LEVB>(S, T, jp.offA, IA);
        break;
      }
      // the parent is complete: its own part (one child) or its accumulator (>= 2 children) joins in registers
      if (jp.nchild >= 2) {
        real t[RBD_ABA2_LEVB];
        lev_ld<RBD_ABA2_LEVB>(S, T, jp.offA, t);
#pragma unroll
        for (int e = 0; e < RBD_ABA2_LEVB; ++e) IA[e] += t[e];
      } else {
        real t[RBD_ABA2_LEV1];
        lev_ld<RBD_ABA2_LEV1>(S, T, jp.offA, t);
        sym6_add_inertia(jp.Y[0], t, t + 3, IA);
#pragma unroll
        for (int e = 0; e < 6; ++e) IA[21 + e] += t[9 + e];
      }
      cur = pp;
      if (fetch_parent) {
        if (RBD_ABA2_UNW) {
#pragma unroll
          for (int e = 0; e < 6; ++e) { A[e] = An[e]; cr[e] = cn[e]; }
          tauc = taun;
        } else {
#if RBD_ABA2_V3
          const SCR Xp = X.at(RBD_ABA2_REC * (nj - 1) + RBD_ABA2_LREC * m.joint[pp].depth);
#else
          const SCR Xp = X.at(RBD_ABA2_REC * (pp - 1));
#endif
#pragma unroll
          for (int e = 0; e < 6; ++e) { A[e] = Xp.ld(e); cr[e] = Xp.ld(6 + e); }
          tauc = tau.ld(m.joint[pp].idx_v);
        }
      }
    }
    // the joint after a leaf starts a new branch: restore R | p | v of its parent (or the universe)
    if (i + 1 < nj) {
      if (stop_at == 0) {
        RBD_ABA_ROOT();
      } else {
        real t[RBD_ABA2_RST];
        lev_ld<RBD_ABA2_RST>(S, T, m.joint[stop_at].boff, t);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = t[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = t[9 + e];
#pragma unroll
        for (int e = 0; e < 6; ++e) vel[e] = t[12 + e];
      }
    }
  }
#undef RBD_ABA_PREFETCH
  // ---- pass 3 ----------------------------------------------------------------------------------------------------
#if RBD_ABA2_V3
  // The forward kinematics and the velocity recursion run once more (A, c recomputed: ~150 instructions per joint instead
  // of 12 doubles written and read twice through a record that does not fit L2), the 8-double record [U | 1 / D | u] is
  // all that comes back from global memory; joint angle and velocity are fetched one joint ahead as in pass 1.
  real acc[6];
  RBD_ABA_ROOT();
#define RBD_ABA_PREFETCH3(J)                                                                                 \
  do {                                                                                                       \
    const DevJoint& jx_ = m.joint[J];                                                                        \
    if (jx_.type != JT_FF) { nq0 = q.ld(jx_.idx_q); nqd = v.ld(jx_.idx_v); }                                 \
  } while (0)
#ifndef RBD_ABA2_P3V3
#define RBD_ABA2_P3V3 0 /* 1: the record of the next joint is fetched while this one is stepped — measured slower (A/B s4a,
                           Talos 2^20: 3.44 vs 3.31 ms), like every register prefetch in this sweep */
#endif
  real Bn[RBD_ABA2_REC];
#define RBD_ABA_FETCH3(J)                                                                       \
  do {                                                                                          \
    const SCR Xn_ = X.at(RBD_ABA2_REC * ((J) - 1));                                             \
    _Pragma("unroll") for (int e_ = 0; e_ < RBD_ABA2_REC; ++e_) Bn[e_] = Xn_.ld(e_);            \
  } while (0)
  if (nj > 1) { RBD_ABA_PREFETCH3(1); if (RBD_ABA2_P3V3) RBD_ABA_FETCH3(1); }
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    real B[RBD_ABA2_REC];
    if (!RBD_ABA2_P3V3) RBD_ABA_FETCH3(i);
#pragma unroll
    for (int e = 0; e < RBD_ABA2_REC; ++e) B[e] = Bn[e];
    if (RBD_ABA2_P3V3 && i + 1 < nj) RBD_ABA_FETCH3(i + 1);
    if (par != i - 1) {  // a new branch: R | p | v of the joint it starts from (or the universe)
      if (par == 0) {
        RBD_ABA_ROOT();
      } else {
        real t[RBD_ABA2_RST];
        lev_ld<RBD_ABA2_RST>(S, T, m.joint[par].boff, t);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = t[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = t[9 + e];
#pragma unroll
        for (int e = 0; e < 6; ++e) vel[e] = t[12 + e];
      }
    }
    const real q0c = nq0, qdc = nqd;
    if (i + 1 < nj) RBD_ABA_PREFETCH3(i + 1);
    const bool ff = jn.type == JT_FF;
    real q0 = q0c, q1 = 0, aw[3];
    if (!ff && is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    if (ff) {
      real vj[6];
      ff_act(R, p, v, jn.idx_v, vj);
#pragma unroll
      for (int e = 0; e < 6; ++e) { vel[e] += vj[e]; acc[e] = B[e]; }
    } else {
      real A[6], vj[6], cr[6], tx, ty, tz;
      subspace_1dof(jn.type, p, aw, A);
#pragma unroll
      for (int e = 0; e < 6; ++e) { vj[e] = A[e] * qdc; vel[e] += vj[e]; }
      RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
      cr[0] += tx; cr[1] += ty; cr[2] += tz;
      RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
      if (par == 0) {
        acc[0] = -m.gravity[0]; acc[1] = -m.gravity[1]; acc[2] = -m.gravity[2]; acc[3] = acc[4] = acc[5] = 0;
      } else if (par != i - 1) {  // (else: acc still holds the parent's acceleration)
        lev_ld<6>(S, T, m.joint[par].offA, acc);
      }
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] += cr[e];
      const real qdd = (B[7] - dot6(B, acc)) * B[6];
      ddq.st(jn.idx_v, qdd);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] += A[e] * qdd;
    }
    if (jn.nchild >= 2) {
      // (the restart record is written again: level records are shared by all joints of a depth, and a later joint of
      // this depth has overwritten what passes 1 + 2 left there)
      real t[RBD_ABA2_RST];
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) t[9 + e] = p[e];
#pragma unroll
      for (int e = 0; e < 6; ++e) t[12 + e] = vel[e];
      lev_st<RBD_ABA2_RST>(S, T, jn.boff, t);
      lev_st<6>(S, T, jn.offA, acc);
    }
  }
#undef RBD_ABA_PREFETCH3
#undef RBD_ABA_FETCH3
#else
  // A chain of short dependent steps behind 20 loads each: the record of the next joint is fetched while this one is
  // stepped (two buffers in ping-pong: the loop is unrolled by two so that no buffer is ever copied).
  real acc[6];
  real b0[RBD_ABA2_REC], b1[RBD_ABA2_REC];
#define RBD_ABA_FETCH(J, B)                                                                     \
  do {                                                                                          \
    const SCR Xn_ = X.at(RBD_ABA2_REC * ((J) - 1));                                             \
    if (m.joint[J].type != JT_FF) {                                                             \
      _Pragma("unroll") for (int e_ = 0; e_ < RBD_ABA2_REC; ++e_) B[e_] = Xn_.ld(e_);           \
    } else { /* the acceleration solved at the end of the backward pass */                      \
      _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) B[e_] = Xn_.ld(12 + e_);                 \
    }                                                                                           \
  } while (0)
#define RBD_ABA_STEP3(I, B)                                                                     \
  do {                                                                                          \
    const DevJoint& jn = m.joint[I];                                                            \
    const int par = jn.parent;                                                                  \
    if (jn.type == JT_FF) {                                                                     \
      _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) acc[e_] = B[e_];                         \
    } else {                                                                                    \
      if (par == 0) {                                                                           \
        acc[0] = -m.gravity[0]; acc[1] = -m.gravity[1]; acc[2] = -m.gravity[2]; acc[3] = acc[4] = acc[5] = 0; \
      } else if (par != (I) - 1) { /* (else: acc still holds the parent's acceleration) */      \
        lev_ld<6>(S, T, m.joint[par].offA, acc);                                                \
      }                                                                                         \
      _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) acc[e_] += B[6 + e_];                    \
      const real qdd_ = (B[19] - dot6(B + 12, acc)) * B[18];                                    \
      ddq.st(jn.idx_v, qdd_);                                                                   \
      _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) acc[e_] += B[e_] * qdd_;                 \
    }                                                                                           \
    if (jn.nchild >= 2) lev_st<6>(S, T, jn.offA, acc);                                          \
  } while (0)
#ifndef RBD_ABA2_P3
#define RBD_ABA2_P3 0
#endif
#if RBD_ABA2_P3 == 2
  if (nj > 1) RBD_ABA_FETCH(1, b0);
  for (int i = 1; i < nj; i += 2) {
    if (i + 1 < nj) RBD_ABA_FETCH(i + 1, b1);
    RBD_ABA_STEP3(i, b0);
    if (i + 1 < nj) {
      if (i + 2 < nj) RBD_ABA_FETCH(i + 2, b0);
      RBD_ABA_STEP3(i + 1, b1);
    }
  }
#elif RBD_ABA2_P3 == 1
  if (nj > 1) RBD_ABA_FETCH(1, b1);
  for (int i = 1; i < nj; ++i) {
#pragma unroll
    for (int e = 0; e < RBD_ABA2_REC; ++e) b0[e] = b1[e];
    if (i + 1 < nj) RBD_ABA_FETCH(i + 1, b1);
    RBD_ABA_STEP3(i, b0);
  }
#else
  for (int i = 1; i < nj; ++i) {
    RBD_ABA_FETCH(i, b0);
    RBD_ABA_STEP3(i, b0);
  }
#endif
#undef RBD_ABA_STEP3
#undef RBD_ABA_FETCH
#endif  /* RBD_ABA2_V3 */
#undef RBD_ABA_ROOT
}

// =========================================================================================================
// Inverse of the joint-space mass matrix (pinocchio::computeMinverse; SURVEY.md §8(f) rank 4 — not in the reference)
// =========================================================================================================
// M^-1 column c is the forward dynamics of a unit torque on dof c with v = 0 and no gravity, so the articulated-body
// quantities are shared by all columns:
//   phase 1  passes 1 + 2 of the sweep above restricted to the matrix part (I^A, U = I^A A, 1/D; no bias forces), on the
//            same depth-level records (one child: the joint's own inertia h | I_O (9); >= 2 children: I^A accumulator (21)
//            + restart R | p (12)).  Per joint A (6), U (6), 1/D go to the thread's global record (RBD_MINV_REC = 13); a
//            root free-flyer leaves R | p (12) and (X^T I^A X)^-1 (21, Cholesky) in a 33-double record behind them.
//   phase 2  columns in blocks of RBD_MINV_CB consecutive dofs.  A unit torque on joint j loads only the path j -> root:
//            the backward walk p <- U_j / D_j, then u_a = -A_a . p, p += U_a u_a / D_a for every ancestor a, leaves the
//            partial entries u_a / D_a in the OUTPUT (rows of the ancestors, column c), the root free-flyer closes it
//            with qdd_root = -(X^T I^A X)^-1 X^T p.  One forward sweep per block then completes the rows:
//            qdd_i = partial(i, c) - (U_i . a_parent) / D_i,  a_i = a_parent + A_i qdd_i, every joint record (13 doubles)
//            read once per block; the accelerations of a block at a joint with >= 2 children (6 RBD_MINV_CB doubles) wait
//            in that joint's level record.  Only the upper triangle is produced (rows <= column, packed by columns like M):
//            a column is finished as soon as the sweep passes its joint, and the sweep of a block stops at its last joint.
// Joints other than a root free-flyer have one dof, so dof c belongs to joint c + 1 (fixed base) or c - 4 (c >= 6 below a
// free-flyer root).  DevJoint::slot[0] of the planned copy holds the end of the joint's subtree (exclusive).
// (RBD_MINV_CB, RBD_MINV_REC, RBD_MINV_ROOT_REC: rbd_dev_model.h)
static_assert(RBD_MINV_CB >= 6, "the six columns of a root free-flyer must sit in the first block");
#define RBD_MINV_LEV1 9
#define RBD_MINV_LEVB 21
#define RBD_MINV_RST 12

// Mv: packed upper triangle of the output (field c (c + 1) / 2 + r, r <= c); live = 0 for the lanes past the end of the batch
template <class FLD, class SCR>
RBD_HD void minv_instance(const DevModel& m, const FLD& q, const FLD& Mv, const SCR& X, const Stack& S, const TmStack& T,
                          const int live) {
  const int nj = m.njoints, nv = m.nv;
  const bool ffroot = nj > 1 && m.joint[1].type == JT_FF;
  const SCR XR = X.at(RBD_MINV_REC * (nj - 1));
  real R[9], p[3];
#define RBD_MINV_ROOT()                                                                         \
  do {                                                                                          \
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;   \
    p[0] = p[1] = p[2] = 0;                                                                     \
  } while (0)
  RBD_MINV_ROOT();
  // ---- phase 1 ---------------------------------------------------------------------------------------------------
  real nq0 = 0;
  if (nj > 1 && m.joint[1].type != JT_FF) nq0 = q.ld(m.joint[1].idx_q);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const SCR Xi = X.at(RBD_MINV_REC * (i - 1));
    const real q0 = nq0;
    if (i + 1 < nj && m.joint[i + 1].type != JT_FF) nq0 = q.ld(m.joint[i + 1].idx_q);
    const bool ff = jn.type == JT_FF;
    real q1 = 0, aw[3];
    if (!ff && is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    real A[6];
    if (!ff) {
      subspace_1dof(jn.type, p, aw, A);
#pragma unroll
      for (int e = 0; e < 6; ++e) Xi.st(e, A[e]);
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) XR.st(e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) XR.st(9 + e, p[e]);
    }
    real Y[10];
    inertia_world(jn, R, p, Y);
    if (jn.nchild == 1) {
      lev_st<RBD_MINV_LEV1>(S, T, jn.offA, Y + 1);
      continue;
    }
    real IA[21];
    sym6_from_inertia(Y, IA);
    if (jn.nchild >= 2) {
      lev_st<RBD_MINV_LEVB>(S, T, jn.offA, IA);
      real t[RBD_MINV_RST];
#pragma unroll
      for (int e = 0; e < 9; ++e) t[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) t[9 + e] = p[e];
      lev_st<RBD_MINV_RST>(S, T, jn.boff, t);
      continue;
    }
    // leaf: backward pass for every joint this leaf completes
    const int stop_at = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    int cur = i;
    for (;;) {
      const DevJoint& jc = m.joint[cur];
      if (jc.type == JT_FF) {  // root: (X^T I^A X)^-1 by Cholesky, upper triangle
        real Rp[12], D[21], Di[21];
#pragma unroll
        for (int e = 0; e < 12; ++e) Rp[e] = XR.ld(e);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          real Xc[6], F[6], col[6];
          RBD_FF_COL(Xc, Rp, (Rp + 9), cc);
          sym6_mul(IA, Xc, F);
          ff_rows(Rp, Rp + 9, F, col);
#pragma unroll
          for (int r = 0; r <= cc; ++r) D[s6(r, cc)] = col[r];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {  // D = L L^T, D[s6(r, c)] (r <= c) <- L(c, r)
          real d = D[s6(k, k)];
#pragma unroll
          for (int j = 0; j < k; ++j) d -= D[s6(j, k)] * D[s6(j, k)];
          d = sqrt(d);
          D[s6(k, k)] = d;
          const real di = real(1) / d;
#pragma unroll
          for (int r = k + 1; r < 6; ++r) {
            real t = D[s6(k, r)];
#pragma unroll
            for (int j = 0; j < k; ++j) t -= D[s6(j, r)] * D[s6(j, k)];
            D[s6(k, r)] = t * di;
          }
        }
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {  // column cc of the inverse: L y = e_cc (y_k = 0 for k < cc), L^T x = y
          real b[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            real t = k == cc ? real(1) : real(0);
#pragma unroll
            for (int j = cc; j < k; ++j) t -= D[s6(j, k)] * b[j];
            b[k] = k < cc ? real(0) : t / D[s6(k, k)];
          }
#pragma unroll
          for (int k = 5; k >= 0; --k) {
            real t = b[k];
#pragma unroll
            for (int j = k + 1; j < 6; ++j) t -= D[s6(k, j)] * b[j];
            b[k] = t / D[s6(k, k)];
          }
#pragma unroll
          for (int r = 0; r <= cc; ++r) Di[s6(r, cc)] = b[r];
        }
#pragma unroll
        for (int e = 0; e < 21; ++e) XR.st(12 + e, Di[e]);
        break;
      }
      const SCR Xc = X.at(RBD_MINV_REC * (cur - 1));
      real U[6];
      sym6_mul(IA, A, U);
      const real dinv = real(1) / dot6(A, U);
#pragma unroll
      for (int e = 0; e < 6; ++e) Xc.st(6 + e, U[e]);
      Xc.st(12, dinv);
      const int pp = jc.parent;
      if (pp == 0) break;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = r; cc < 6; ++cc) IA[s6(r, cc)] -= U[r] * (U[cc] * dinv);
      const DevJoint& jp = m.joint[pp];
      if (pp == stop_at) {
        lev_add<RBD_MINV_LEVB>(S, T, jp.offA, IA);
        break;
      }
      if (jp.nchild >= 2) {
        real t[RBD_MINV_LEVB];
        lev_ld<RBD_MINV_LEVB>(S, T, jp.offA, t);
#pragma unroll
        for (int e = 0; e < RBD_MINV_LEVB; ++e) IA[e] += t[e];
      } else {
        real t[RBD_MINV_LEV1];
        lev_ld<RBD_MINV_LEV1>(S, T, jp.offA, t);
        sym6_add_inertia(jp.Y[0], t, t + 3, IA);
      }
      cur = pp;
      if (jp.type != JT_FF) {
        const SCR Xp = X.at(RBD_MINV_REC * (pp - 1));
#pragma unroll
        for (int e = 0; e < 6; ++e) A[e] = Xp.ld(e);
      }
    }
    if (i + 1 < nj) {  // the joint after a leaf starts a new branch
      if (stop_at == 0) {
        RBD_MINV_ROOT();
      } else {
        real t[RBD_MINV_RST];
        lev_ld<RBD_MINV_RST>(S, T, m.joint[stop_at].boff, t);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = t[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = t[9 + e];
      }
    }
  }
#undef RBD_MINV_ROOT
  // ---- phase 2 ---------------------------------------------------------------------------------------------------
  // Per block of columns: ONE backward sweep over the joints (descending) serves the walks of all its columns — the
  // columns of a block are neighbouring dofs and share most of their ancestors, each joint record is read once, and the
  // steps of different columns are independent —, then the root, then ONE forward sweep.  Records (and, in the forward
  // sweep, the partial entries) are fetched one joint ahead.
  const int jlo = ffroot ? 2 : 1;
#define RBD_MINV_JOINT_OF(c) (ffroot ? ((c) < 6 ? 1 : (c) - 4) : (c) + 1)
#define RBD_MINV_FLD(r, c) ((c) * ((c) + 1) / 2 + (r))
// the block's accelerations <-> level record, two columns at a time (a 48-double tensor-memory access would need 96
// staging registers next to the 96 of acc itself)
#define RBD_MINV_ACC_LD(OFF)                                                                      \
  do {                                                                                            \
    _Pragma("unroll") for (int h_ = 0; h_ < RBD_MINV_CB / 2; ++h_)                                \
      lev_ld<12>(S, T, (OFF) + 12 * h_, &acc[2 * h_][0]);                                         \
  } while (0)
#define RBD_MINV_ACC_ST(OFF)                                                                      \
  do {                                                                                            \
    _Pragma("unroll") for (int h_ = 0; h_ < RBD_MINV_CB / 2; ++h_)                                \
      lev_st<12>(S, T, (OFF) + 12 * h_, &acc[2 * h_][0]);                                         \
  } while (0)
#define RBD_MINV_FETCH(J)                                                                         \
  do {                                                                                            \
    const SCR Xn_ = X.at(RBD_MINV_REC * ((J) - 1));                                               \
    _Pragma("unroll") for (int e_ = 0; e_ < RBD_MINV_REC; ++e_) nrec[e_] = Xn_.ld(e_);            \
  } while (0)
  for (int c0 = 0; c0 < nv; c0 += RBD_MINV_CB) {
    const int ncb = nv - c0 < RBD_MINV_CB ? nv - c0 : RBD_MINV_CB;
    const int jlast = RBD_MINV_JOINT_OF(c0 + ncb - 1);
    real acc[RBD_MINV_CB][6];  // p of the backward sweep, then the accelerations of the forward sweep
#pragma unroll
    for (int k = 0; k < RBD_MINV_CB; ++k)
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[k][e] = 0;
#ifndef RBD_MINV_PREFETCH
#define RBD_MINV_PREFETCH 0 /* A/B r2: fetching records one joint ahead costs registers the sweep does not have (spills) */
#endif
    real nrec[RBD_MINV_REC];
    // backward sweep
    if (RBD_MINV_PREFETCH && jlast >= jlo) RBD_MINV_FETCH(jlast);
    for (int a = jlast; a >= jlo; --a) {
#if RBD_MINV_PREFETCH
      real rec[RBD_MINV_REC];
#pragma unroll
      for (int e = 0; e < RBD_MINV_REC; ++e) rec[e] = nrec[e];
      if (a - 1 >= jlo) RBD_MINV_FETCH(a - 1);
#else
      RBD_MINV_FETCH(a);
      real* const rec = nrec;
#endif
      const int r = m.joint[a].idx_v, send = m.joint[a].slot[0];
      const real dinv = rec[12];
#pragma unroll
      for (int k = 0; k < RBD_MINV_CB; ++k) {
        const int c = c0 + k;
        const int jc = RBD_MINV_JOINT_OF(c);
        if (k < ncb) {
          if (jc == a) {  // the column's own joint: unit torque, p = U / D
#pragma unroll
            for (int e = 0; e < 6; ++e) acc[k][e] = rec[6 + e] * dinv;
            if (live) Mv.st(RBD_MINV_FLD(c, c), dinv);
          } else if (a < jc && jc < send) {  // an ancestor of the column's joint
            const real ua = -dot6(rec, acc[k]) * dinv;
            if (live) Mv.st(RBD_MINV_FLD(r, c), ua);
#pragma unroll
            for (int e = 0; e < 6; ++e) acc[k][e] += rec[6 + e] * ua;
          }
        }
      }
    }
    // the root: free-flyer dofs of every column of the block, and the accelerations the forward sweep starts from
    real Rp[12], Di[21];
    if (ffroot) {
#pragma unroll
      for (int e = 0; e < 12; ++e) Rp[e] = XR.ld(e);
#pragma unroll
      for (int e = 0; e < 21; ++e) Di[e] = XR.ld(12 + e);
    }
#pragma unroll
    for (int k = 0; k < RBD_MINV_CB; ++k) {
      const int c = c0 + k;
      if (k < ncb && ffroot) {
        real qr[6];
        if (c < 6) {
#pragma unroll
          for (int r = 0; r < 6; ++r) qr[r] = Di[s6(r, k)];  // (c0 == 0 for these columns: c == k) column c of (X^T I^A X)^-1
        } else {  // qdd_root = -(X^T I^A X)^-1 X^T p
          real b[6];
          ff_rows(Rp, Rp + 9, acc[k], b);
#pragma unroll
          for (int e = 0; e < 6; ++e) b[e] = -b[e];
          sym6_mul(Di, b, qr);
        }
        const int rmax = c < 5 ? c : 5;
#pragma unroll
        for (int r = 0; r < 6; ++r)
          if (r <= rmax && live) Mv.st(RBD_MINV_FLD(r, c), qr[r]);
        real aa[3], tx, ty, tz;  // a_root = X qdd_root
#pragma unroll
        for (int e = 0; e < 3; ++e) {
          acc[k][e] = Rp[3 * e] * qr[0] + Rp[3 * e + 1] * qr[1] + Rp[3 * e + 2] * qr[2];
          aa[e] = Rp[3 * e] * qr[3] + Rp[3 * e + 1] * qr[4] + Rp[3 * e + 2] * qr[5];
        }
        RBD_CROSS(tx, ty, tz, Rp[9], Rp[10], Rp[11], aa[0], aa[1], aa[2]);
        acc[k][0] += tx; acc[k][1] += ty; acc[k][2] += tz;
        acc[k][3] = aa[0]; acc[k][4] = aa[1]; acc[k][5] = aa[2];
      } else {
#pragma unroll
        for (int e = 0; e < 6; ++e) acc[k][e] = 0;
      }
    }
    // forward sweep of the block up to the joint of its last column
    if (ffroot && m.joint[1].nchild >= 2 && jlast >= jlo) RBD_MINV_ACC_ST(m.joint[1].offA);
    real npart[RBD_MINV_CB];
#define RBD_MINV_FETCH_PART(J)                                                                    \
  do {                                                                                            \
    const int rn_ = m.joint[J].idx_v, sn_ = m.joint[J].slot[0];                                   \
    _Pragma("unroll") for (int k_ = 0; k_ < RBD_MINV_CB; ++k_) {                                  \
      const int cn_ = c0 + k_;                                                                    \
      const int jn_ = RBD_MINV_JOINT_OF(cn_);                                                     \
      npart[k_] = 0;                                                                              \
      if (k_ < ncb && jn_ >= (J) && jn_ < sn_ && live) npart[k_] = Mv.ld(RBD_MINV_FLD(rn_, cn_)); \
    }                                                                                             \
  } while (0)
    if (RBD_MINV_PREFETCH && jlast >= jlo) { RBD_MINV_FETCH(jlo); RBD_MINV_FETCH_PART(jlo); }
    for (int i = jlo; i <= jlast; ++i) {
      const DevJoint& jn = m.joint[i];
      const int par = jn.parent;
#if RBD_MINV_PREFETCH
      real rec[RBD_MINV_REC], part[RBD_MINV_CB];
#pragma unroll
      for (int e = 0; e < RBD_MINV_REC; ++e) rec[e] = nrec[e];
#pragma unroll
      for (int k = 0; k < RBD_MINV_CB; ++k) part[k] = npart[k];
      if (i + 1 <= jlast) { RBD_MINV_FETCH(i + 1); RBD_MINV_FETCH_PART(i + 1); }
#else
      RBD_MINV_FETCH(i);
      RBD_MINV_FETCH_PART(i);
      real* const rec = nrec;
      real* const part = npart;
#endif
      if (par == 0) {
#pragma unroll
        for (int k = 0; k < RBD_MINV_CB; ++k)
#pragma unroll
          for (int e = 0; e < 6; ++e) acc[k][e] = 0;
      } else if (par != i - 1) {
        RBD_MINV_ACC_LD(m.joint[par].offA);
      }
      const real dinv = rec[12];
      const int r = jn.idx_v;
#pragma unroll
      for (int k = 0; k < RBD_MINV_CB; ++k) {
        const int c = c0 + k;
        const int jc = RBD_MINV_JOINT_OF(c);
        if (k < ncb && jc >= i) {  // (a column is finished once the sweep has passed its joint: rows > column)
          const real qdd = part[k] - dot6(rec + 6, acc[k]) * dinv;
          if (live) Mv.st(RBD_MINV_FLD(r, c), qdd);
#pragma unroll
          for (int e = 0; e < 6; ++e) acc[k][e] += rec[e] * qdd;
        }
      }
      if (jn.nchild >= 2) RBD_MINV_ACC_ST(jn.offA);
    }
#undef RBD_MINV_FETCH_PART
  }
#undef RBD_MINV_FETCH
#undef RBD_MINV_ACC_LD
#undef RBD_MINV_ACC_ST
#undef RBD_MINV_JOINT_OF
#undef RBD_MINV_FLD
}
