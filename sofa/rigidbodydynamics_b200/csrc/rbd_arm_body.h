// rbd_arm_body.h — fused FK + world Jacobian + CRBA + RNEA for fixed-base SERIAL CHAINS of RX / RY / RZ joints with the
// joint count known at compile time (arms: Panda, iiwa, UR: BASELINE configs[0] / [1]).  Included from rbd_sweeps.h inside
// namespace rbd::chain (same helpers, same formulas as eval_instance: results agree to rounding).
//
// Why a second sweep for such models: N = 65 536 Panda instances are bound by the per-SM instruction throughput of the
// generic sweep, not by occupancy — 5 120 warp-instructions per tile, of which only 47 % are FP64: the
// rest is what a run-time tree costs (joint-table reads from shared memory, stack addressing, column programs, loop and
// type dispatch).  With NJ a template parameter both passes unroll completely, and then
//   * the joint table travels as a __grid_constant__ kernel parameter: every constant is an immediate-offset operand from
//     the constant bank (c[0x0][imm]) inside the FP64 instruction that uses it — no load at all;
//   * stack slots, output fields and the rows / columns of M are compile-time offsets folded into the instructions;
//   * the world subspace columns A_k of ALL joints are in registers during the backward pass (6 NJ doubles), so the M
//     columns are dot products of registers; during the forward pass, where the registers are needed for the kinematics,
//     they wait in TENSOR MEMORY (one tcgen05.st of 6 doubles per joint, one vector load of all 6 NJ at the turn); the
//     shared-memory stack holds only f_k and Y_k of the joints with a child (15 doubles per level).
// A chain of NJ joints has nq = nv = NJ, joint k owns dof k, M's packed upper triangle has no structural zeros (the
// support-only format is the same array).
template <int NJ>
struct ArmParams {
  DevJoint joint[NJ];  // the moving joints 1 .. NJ of the model, universe dropped
  real gravity[3];
};

#define RBD_ARM_LEVEL 15 /* shared-memory doubles per joint with a child: f (6), h | I_O (9) */

template <int NJ, int TILE>
RBD_HD void arm_instance(const ArmParams<NJ>& m, const FieldT<TILE>& q, const FieldT<TILE>& v, const FieldT<TILE>& a,
                         const FieldT<TILE>& oMi, const FieldT<TILE>& J, const FieldT<TILE>& M, const FieldT<TILE>& tau,
                         const Stack& S, const TmStack& T) {
  real R[9], p[3], vel[6], acc[6];
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  p[0] = p[1] = p[2] = 0;
#pragma unroll
  for (int e = 0; e < 6; ++e) { vel[e] = 0; acc[e] = 0; }
  acc[0] = -m.gravity[0]; acc[1] = -m.gravity[1]; acc[2] = -m.gravity[2];
  real F[6], Yc[10];  // subtree force and composite inertia of the backward pass (start at the last joint)
  real qn = q.ld(0), vn = v.ld(0), an = a.ld(0);
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const DevJoint& jn = m.joint[i];
    const real q0 = qn, qd = vn, qdd = an;
    if (i + 1 < NJ) { qn = q.ld(i + 1); vn = v.ld(i + 1); an = a.ld(i + 1); }  // one joint ahead
    real aw[3];
    fk_joint(jn, R, p, q0, real(0), q, aw);
    if (oMi.ok()) {
#pragma unroll
      for (int e = 0; e < 9; ++e) oMi.st(12 * i + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) oMi.st(12 * i + 9 + e, p[e]);
    }
    real Ai[6];
    subspace_1dof(jn.type, p, aw, Ai);
    T.template stv<6>(6 * i, Ai);
    if (J.ok()) {
#pragma unroll
      for (int e = 0; e < 6; ++e) J.st(6 * i + e, Ai[e]);
    }
    {
      real vj[6], cr[6], tx, ty, tz;
#pragma unroll
      for (int e = 0; e < 6; ++e) { vj[e] = Ai[e] * qd; vel[e] += vj[e]; }
      // a_i = a_parent + A qdd + v_i x vJ
      RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
      cr[0] += tx; cr[1] += ty; cr[2] += tz;
      RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] += Ai[e] * qdd + cr[e];
    }
    real Yk[10], fk[6];
    inertia_world(jn, R, p, Yk);
    {
      // f = Y a + v x* (Y v)
      real h[6], tx, ty, tz;
      inertia_mul(Yk, acc, fk);
      inertia_mul(Yk, vel, h);
      RBD_CROSS(tx, ty, tz, vel[3], vel[4], vel[5], h[0], h[1], h[2]);
      fk[0] += tx; fk[1] += ty; fk[2] += tz;
      RBD_CROSS(tx, ty, tz, vel[3], vel[4], vel[5], h[3], h[4], h[5]);
      fk[3] += tx; fk[4] += ty; fk[5] += tz;
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], h[0], h[1], h[2]);
      fk[3] += tx; fk[4] += ty; fk[5] += tz;
    }
    if (i + 1 < NJ) {
#pragma unroll
      for (int e = 0; e < 6; ++e) S.st(RBD_ARM_LEVEL * i + e, fk[e]);
#pragma unroll
      for (int e = 0; e < 9; ++e) S.st(RBD_ARM_LEVEL * i + 6 + e, Yk[1 + e]);
    } else {
#pragma unroll
      for (int e = 0; e < 6; ++e) F[e] = fk[e];
#pragma unroll
      for (int e = 0; e < 10; ++e) Yc[e] = Yk[e];
    }
  }
  real A[NJ][6];
  T.template ldv<6 * NJ>(0, &A[0][0]);
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    if (i + 1 < NJ) {
#pragma unroll
      for (int e = 0; e < 6; ++e) F[e] += S.ld(RBD_ARM_LEVEL * i + e);
#pragma unroll
      for (int e = 0; e < 9; ++e) Yc[1 + e] += S.ld(RBD_ARM_LEVEL * i + 6 + e);
      Yc[0] += m.joint[i].Y[0];
    }
    tau.st(i, dot6(A[i], F));
    real Fc[6];
    inertia_mul(Yc, A[i], Fc);
#pragma unroll
    for (int k = 0; k <= i; ++k) M.st(i * (i + 1) / 2 + k, dot6(A[k], Fc));
  }
}

// The vector applyJT of the same chains (reference KinematicChainMapping.inl:450-513; the ring kernel's sweep,
// applyJT_instance_ring, with the tree taken out): tau_i = A_i . sum of the origin-shifted wrenches on joints >= i.  A chain
// needs no stack for that — with Pre_i the sum over the joints < i, tau_i = A_i . Total - A_i . Pre_i, so the forward pass
// keeps the NJ columns A_i in registers, one running sum and NJ partial dot products; nothing goes to shared or tensor
// memory.  Joint constants from the kernel-parameter constant bank, wrenches from the ring (Ring: see rbd_sweeps_body.h).
template <int NJ, class QC, class Ring, class Sink>
RBD_HD void applyJT_arm_instance(const ArmParams<NJ>& m, const real* opl, const QC& q, Ring& ring, const Sink& sink) {
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real A[NJ][6], part[NJ], Pre[6] = {0, 0, 0, 0, 0, 0};
  real qv[NJ];  // all joint angles at once: NJ independent loads in flight instead of one round trip per joint
#pragma unroll
  for (int i = 0; i < NJ; ++i) qv[i] = q.ld(i);
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const DevJoint& jn = m.joint[i];
    const real q0 = qv[i];
    real aw[3];
    ::rbd::chain::fk_joint(jn, R, p, q0, real(0), q, aw);  // (qualified: `q` may be a type of namespace rbd, whose fk_joint ADL would find too)
    subspace_1dof(jn.type, p, aw, A[i]);
    part[i] = dot6(A[i], Pre);
    // (one rolled loop over the wrenches of a piece, not the four straight-line cases of the generic ring sweep: with the
    // joint loop unrolled those made 95 KB of code that every warp runs through once per tile — `no_instruction` was the
    // largest stall of the first version, 2.0 cycles per issue, profile s3q)
    for (int o = jn.out_begin; o < jn.out_end;) {
      const int cnt = ring.acquire();
#pragma unroll 1
      for (int j = 0; j < cnt; ++j) {
        real w[6];
        ring.ldr(j, w);
        const real* pl = opl + 3 * (o + j);
        const real px = R[0] * pl[0] + R[1] * pl[1] + R[2] * pl[2] + p[0];
        const real py = R[3] * pl[0] + R[4] * pl[1] + R[5] * pl[2] + p[1];
        const real pz = R[6] * pl[0] + R[7] * pl[1] + R[8] * pl[2] + p[2];
        real tx, ty, tz;
        RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);
        Pre[0] += w[0]; Pre[1] += w[1]; Pre[2] += w[2];
        Pre[3] += w[3] + tx; Pre[4] += w[4] + ty; Pre[5] += w[5] + tz;
      }
      ring.release(cnt);
      o += cnt;
    }
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i) sink.put(i, dot6(A[i], Pre) - part[i]);
}
