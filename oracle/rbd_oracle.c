/*
 * rbd_oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
 * load or call this file.  The product (librbd_b200.so) never links it and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference delegates its arithmetic to pinocchio 3.3.1 (pixi.lock:223; Eigen 3.4.0,
 * pixi.lock:35), which is neither vendored under the reference tree nor installable here, and the
 * reference has no tests or golden vectors.  This file restates pinocchio's published algorithms
 * (algorithm/{kinematics,jacobian,frames,crba,rnea}.hxx, spatial/{se3,motion,force,inertia,symmetric3},
 * multibody/joint/joint-{revolute,revolute-unaligned,revolute-unbounded,prismatic,free-flyer}.hpp) in the
 * order the reference calls them, and is cross-checked against two independent implementations
 * (oracle/rbd_numpy.py in double and in 50-digit mpmath) and analytic cases, not against pinocchio.
 *
 * Reference call sites restated here (paths relative to the reference root):
 *   src/sofa/RigidBodyDynamics/KinematicChainMapping.inl  (KCM.inl)
 *     :338-396 apply      -> orc_apply            (computeJointJacobians :373, updateFramePlacements :375)
 *     :398-448 applyJ     -> orc_applyJ           (getFrameJacobian LWA :434,:444; dense J*dq :435,:445)
 *     :450-513 applyJT    -> orc_applyJT          (getFrameJacobian :480,:490; J^T f :481,:491; += :498-511)
 *     :515-639 applyJT(M) -> orc_applyJT_row      (kMinTorqueSqrd filter :43-45,:587,:605)
 *   src/sofa/RigidBodyDynamics/Conversions.h :58-61 se3ToSofaType (Eigen::Quaternion{R}) -> orc_rot_to_quat
 *   north-star additions with no reference call site: pinocchio::crba (LOCAL convention), pinocchio::rnea.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#ifdef ORC_COUNT_FLOPS
/* Flop-counting build (g++ -x c++ -DORC_COUNT_FLOPS, oracle/Makefile target librbd_oracle_flops.so): from here on
 * `double` is a counting scalar of the same size and layout, so every arithmetic operation of the restatement below
 * is tallied (orc_flop_counts) while the C entry points keep their binary interface.  SURVEY.md §8(d). */
#include "orc_flopcount.hpp"
thread_local orc_counters orc_cnt = {0, 0, 0, 0, 0, 0};
#define double orc_counted
#endif
#include "rbd_oracle.h"

/* ------------------------------------------------------------------------------------------------ */
/* small spatial algebra, [linear; angular], SE3 = R row-major + p                                    */
typedef struct { double R[9], p[3]; } se3;
typedef struct { double m, c[3], I[6]; } inertia; /* I: xx,xy,yy,xz,yz,zz about the CoM */

static void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matTvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matmul3(const double* A, const double* B, double* O) {
  double T[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(O, T, sizeof T);
}
static void se3_identity(se3* M) {
  memset(M, 0, sizeof *M);
  M->R[0] = M->R[4] = M->R[8] = 1.0;
}
/* SE3 composition (R1,p1)*(R2,p2) = (R1 R2, R1 p2 + p1) */
static void se3_mul(const se3* a, const se3* b, se3* o) {
  se3 t; double rp[3];
  matmul3(a->R, b->R, t.R);
  matvec3(a->R, b->p, rp);
  for (int k = 0; k < 3; ++k) t.p[k] = rp[k] + a->p[k];
  *o = t;
}
/* aMb.act(motion_b) = [R v + p x (R w); R w] */
static void se3_act_motion(const se3* M, const double* m, double* o) {
  double Rv[3], Rw[3], pxw[3];
  matvec3(M->R, m, Rv); matvec3(M->R, m + 3, Rw); cross3(M->p, Rw, pxw);
  for (int k = 0; k < 3; ++k) { o[k] = Rv[k] + pxw[k]; o[3 + k] = Rw[k]; }
}
/* aMb.actInv(motion_a) = [R^T (v - p x w); R^T w] */
static void se3_actinv_motion(const se3* M, const double* m, double* o) {
  double pxw[3], t[3];
  cross3(M->p, m + 3, pxw);
  for (int k = 0; k < 3; ++k) t[k] = m[k] - pxw[k];
  double a[3], b[3];
  matTvec3(M->R, t, a); matTvec3(M->R, m + 3, b);
  for (int k = 0; k < 3; ++k) { o[k] = a[k]; o[3 + k] = b[k]; }
}
/* aMb.act(force_b) = [R f; R n + p x (R f)] */
static void se3_act_force(const se3* M, const double* f, double* o) {
  double Rf[3], Rn[3], pxf[3];
  matvec3(M->R, f, Rf); matvec3(M->R, f + 3, Rn); cross3(M->p, Rf, pxf);
  for (int k = 0; k < 3; ++k) { o[k] = Rf[k]; o[3 + k] = Rn[k] + pxf[k]; }
}
/* motion x motion: [w_a x v_b + v_a x w_b; w_a x w_b] */
static void motion_cross(const double* a, const double* b, double* o) {
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1); cross3(a, b + 3, t2); cross3(a + 3, b + 3, t3);
  for (int k = 0; k < 3; ++k) { o[k] = t1[k] + t2[k]; o[3 + k] = t3[k]; }
}
/* motion x* force: [w x f; w x n + v x f] */
static void motion_cross_force(const double* a, const double* f, double* o) {
  double t1[3], t2[3], t3[3];
  cross3(a + 3, f, t1); cross3(a + 3, f + 3, t2); cross3(a, f, t3);
  for (int k = 0; k < 3; ++k) { o[k] = t1[k]; o[3 + k] = t2[k] + t3[k]; }
}
static void sym3_mulvec(const double* I, const double* w, double* o) {
  double x = I[0] * w[0] + I[1] * w[1] + I[3] * w[2];
  double y = I[1] * w[0] + I[2] * w[1] + I[4] * w[2];
  double z = I[3] * w[0] + I[4] * w[1] + I[5] * w[2];
  o[0] = x; o[1] = y; o[2] = z;
}
/* InertiaTpl::__mult__: f.lin = m (v - c x w); f.ang = I w + c x f.lin */
static void inertia_mul_motion(const inertia* Y, const double* v, double* f) {
  double cxw[3], lin[3], Iw[3], cxf[3];
  cross3(Y->c, v + 3, cxw);
  for (int k = 0; k < 3; ++k) lin[k] = Y->m * (v[k] - cxw[k]);
  sym3_mulvec(Y->I, v + 3, Iw);
  cross3(Y->c, lin, cxf);
  for (int k = 0; k < 3; ++k) { f[k] = lin[k]; f[3 + k] = Iw[k] + cxf[k]; }
}
/* InertiaTpl::se3Action_impl: (m, p + R c, R I R^T) */
static void inertia_se3_act(const se3* M, const inertia* Y, inertia* O) {
  inertia t;
  double Rc[3];
  t.m = Y->m;
  matvec3(M->R, Y->c, Rc);
  for (int k = 0; k < 3; ++k) t.c[k] = M->p[k] + Rc[k];
  double If[9] = {Y->I[0], Y->I[1], Y->I[3], Y->I[1], Y->I[2], Y->I[4], Y->I[3], Y->I[4], Y->I[5]};
  double RI[9], RIRt[9], Rt[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[3 * i + j] = M->R[3 * j + i];
  matmul3(M->R, If, RI); matmul3(RI, Rt, RIRt);
  t.I[0] = RIRt[0]; t.I[1] = RIRt[1]; t.I[2] = RIRt[4]; t.I[3] = RIRt[2]; t.I[4] = RIRt[5]; t.I[5] = RIRt[8];
  *O = t;
}
/* InertiaTpl::__plus__ with the mass floored at machine epsilon; SkewSquare(AB) = AB AB^T - |AB|^2 I */
static void inertia_add(const inertia* A, const inertia* B, inertia* O) {
  inertia t;
  double mab = A->m + B->m;
  double inv = 1.0 / (mab > DBL_EPSILON ? mab : DBL_EPSILON);
  double ab[3];
  for (int k = 0; k < 3; ++k) { ab[k] = A->c[k] - B->c[k]; t.c[k] = (A->m * A->c[k] + B->m * B->c[k]) * inv; }
  t.m = mab;
  double s = A->m * B->m * inv, n2 = ab[0] * ab[0] + ab[1] * ab[1] + ab[2] * ab[2];
  double SS[6] = {ab[0] * ab[0] - n2, ab[0] * ab[1], ab[1] * ab[1] - n2, ab[0] * ab[2], ab[1] * ab[2], ab[2] * ab[2] - n2};
  for (int k = 0; k < 6; ++k) t.I[k] = A->I[k] + B->I[k] - s * SS[k];
  *O = t;
}

/* ------------------------------------------------------------------------------------------------ */
struct orc_workspace {
  /* model (deep copy of the descriptor) */
  int njoints, nq, nv, nframes, n_out;
  int *parents, *jtype, *idx_q, *idx_v, *nvj, *nv_subtree, *frame_parent, *out_frames, *parent_col;
  double *axis;
  se3 *jplace, *fplace;
  inertia* Y;
  double gravity[3];
  /* data */
  se3 *liMi, *oMi, *oMf;
  double* S;     /* [njoints][6*6] joint motion subspace columns (col-major, up to 6 cols) */
  double* vj;    /* [njoints][6] joint velocity */
  double* J;     /* 6 x nv col-major: world joint Jacobian (data.J) */
  double* Jf;    /* 6 x nv scratch frame Jacobian */
  double *v, *a_gf, *f, *h; /* [njoints][6] */
  inertia* Ycrb;
  double* Fcrb;  /* [njoints] x (6 x nv col-major) */
  double* M;     /* nv x nv row-major */
  double *qfull, *dqfull, *tauacc;
  int applied;
};

static inline int joint_nq(int t) { return t == RBD_JOINT_FF ? 7 : (t >= RBD_JOINT_RUBX && t <= RBD_JOINT_RUBU) ? 2 : (t == RBD_JOINT_UNIVERSE ? 0 : 1); }
static int joint_nv(int t) { return t == RBD_JOINT_FF ? 6 : (t == RBD_JOINT_UNIVERSE ? 0 : 1); }

orc_workspace* orc_create(const rbd_model_desc* d) {
  orc_workspace* w = (orc_workspace*)calloc(1, sizeof *w);
  int nj = d->njoints, nv = d->nv, nf = d->nframes;
  w->njoints = nj; w->nq = d->nq; w->nv = nv; w->nframes = nf; w->n_out = d->n_out;
#define DUPI(dst, src, n) do { w->dst = (int*)malloc(sizeof(int) * (size_t)((n) > 0 ? (n) : 1)); for (int _i = 0; _i < (n); ++_i) w->dst[_i] = (src)[_i]; } while (0)
  DUPI(parents, d->parents, nj); DUPI(jtype, d->jtype, nj); DUPI(idx_q, d->idx_q, nj); DUPI(idx_v, d->idx_v, nj);
  DUPI(frame_parent, d->frame_parent, nf); DUPI(out_frames, d->out_frames, d->n_out);
  w->nvj = (int*)calloc((size_t)nj, sizeof(int)); w->nv_subtree = (int*)calloc((size_t)nj, sizeof(int));
  w->parent_col = (int*)malloc(sizeof(int) * (size_t)(nv > 0 ? nv : 1));
  w->axis = (double*)malloc(sizeof(double) * 3 * (size_t)nj); memcpy(w->axis, d->axis, sizeof(double) * 3 * (size_t)nj);
  w->jplace = (se3*)malloc(sizeof(se3) * (size_t)nj); w->fplace = (se3*)malloc(sizeof(se3) * (size_t)(nf > 0 ? nf : 1));
  w->Y = (inertia*)malloc(sizeof(inertia) * (size_t)nj);
  for (int i = 0; i < nj; ++i) {
    memcpy(w->jplace[i].R, d->placement + 12 * i, 9 * sizeof(double));
    memcpy(w->jplace[i].p, d->placement + 12 * i + 9, 3 * sizeof(double));
    const double* y = d->inertia + 10 * i;
    w->Y[i].m = y[0]; memcpy(w->Y[i].c, y + 1, 3 * sizeof(double)); memcpy(w->Y[i].I, y + 4, 6 * sizeof(double));
    w->nvj[i] = joint_nv(d->jtype[i]);
  }
  for (int f = 0; f < nf; ++f) {
    memcpy(w->fplace[f].R, d->frame_placement + 12 * f, 9 * sizeof(double));
    memcpy(w->fplace[f].p, d->frame_placement + 12 * f + 9, 3 * sizeof(double));
  }
  memcpy(w->gravity, d->gravity, sizeof w->gravity);
  /* nvSubtree (pinocchio Data::computeLastChild / nvSubtree) */
  for (int i = nj - 1; i >= 1; --i) { w->nv_subtree[i] += w->nvj[i]; w->nv_subtree[w->parents[i]] += w->nv_subtree[i]; }
  /* parents_fromRow: for column c of joint i: previous column of the same joint, or last column of the parent */
  for (int i = 1; i < nj; ++i)
    for (int k = 0; k < w->nvj[i]; ++k) {
      int c = w->idx_v[i] + k;
      if (k > 0) w->parent_col[c] = c - 1;
      else { int p = w->parents[i]; w->parent_col[c] = p > 0 ? w->idx_v[p] + w->nvj[p] - 1 : -1; }
    }
  size_t snj = (size_t)nj, snv = (size_t)(nv > 0 ? nv : 1);
  w->liMi = (se3*)calloc(snj, sizeof(se3)); w->oMi = (se3*)calloc(snj, sizeof(se3));
  w->oMf = (se3*)calloc((size_t)(nf > 0 ? nf : 1), sizeof(se3));
  w->S = (double*)calloc(snj * 36, sizeof(double)); w->vj = (double*)calloc(snj * 6, sizeof(double));
  w->J = (double*)calloc(6 * snv, sizeof(double)); w->Jf = (double*)calloc(6 * snv, sizeof(double));
  w->v = (double*)calloc(snj * 6, sizeof(double)); w->a_gf = (double*)calloc(snj * 6, sizeof(double));
  w->f = (double*)calloc(snj * 6, sizeof(double)); w->h = (double*)calloc(snj * 6, sizeof(double));
  w->Ycrb = (inertia*)calloc(snj, sizeof(inertia)); w->Fcrb = (double*)calloc(snj * 6 * snv, sizeof(double));
  w->M = (double*)calloc(snv * snv, sizeof(double));
  w->qfull = (double*)calloc((size_t)(d->nq > 0 ? d->nq : 1), sizeof(double));
  w->dqfull = (double*)calloc(snv, sizeof(double)); w->tauacc = (double*)calloc(snv, sizeof(double));
  for (int i = 0; i < nj; ++i) { se3_identity(&w->liMi[i]); se3_identity(&w->oMi[i]); }
  for (int f = 0; f < nf; ++f) se3_identity(&w->oMf[f]);
  return w;
}

void orc_destroy(orc_workspace* w) {
  if (!w) return;
  free(w->parents); free(w->jtype); free(w->idx_q); free(w->idx_v); free(w->nvj); free(w->nv_subtree);
  free(w->frame_parent); free(w->out_frames); free(w->parent_col); free(w->axis); free(w->jplace); free(w->fplace);
  free(w->Y); free(w->liMi); free(w->oMi); free(w->oMf); free(w->S); free(w->vj); free(w->J); free(w->Jf);
  free(w->v); free(w->a_gf); free(w->f); free(w->h); free(w->Ycrb); free(w->Fcrb); free(w->M);
  free(w->qfull); free(w->dqfull); free(w->tauacc);
  free(w);
}

/* JointModel::calc(q [, v]): joint transform M_j, motion subspace S (joint frame), joint velocity S*v.   */
static void joint_calc(const orc_workspace* w, int i, const double* q, const double* v, se3* Mj, double* S, double* vj) {
  int t = w->jtype[i];
  const double* qi = q + w->idx_q[i];
  const double* ax = w->axis + 3 * i;
  se3_identity(Mj);
  memset(S, 0, 36 * sizeof(double));
  double ca = 1.0, sa = 0.0;
  if (t >= RBD_JOINT_RX && t <= RBD_JOINT_RU) { ca = cos(qi[0]); sa = sin(qi[0]); }
  if (t >= RBD_JOINT_RUBX && t <= RBD_JOINT_RUBU) { ca = qi[0]; sa = qi[1]; }
  switch (t) {
    case RBD_JOINT_RX: case RBD_JOINT_RUBX:
      Mj->R[4] = ca; Mj->R[5] = -sa; Mj->R[7] = sa; Mj->R[8] = ca; S[3] = 1.0; break;
    case RBD_JOINT_RY: case RBD_JOINT_RUBY:
      Mj->R[0] = ca; Mj->R[2] = sa; Mj->R[6] = -sa; Mj->R[8] = ca; S[4] = 1.0; break;
    case RBD_JOINT_RZ: case RBD_JOINT_RUBZ:
      Mj->R[0] = ca; Mj->R[1] = -sa; Mj->R[3] = sa; Mj->R[4] = ca; S[5] = 1.0; break;
    case RBD_JOINT_RU: case RBD_JOINT_RUBU: {
      /* pinocchio toRotationMatrix(axis, cos, sin): cos I + (1-cos) a a^T + sin [a]x */
      double c1 = 1.0 - ca;
      double sx = sa * ax[0], sy = sa * ax[1], sz = sa * ax[2];
      Mj->R[0] = c1 * ax[0] * ax[0] + ca; Mj->R[1] = c1 * ax[0] * ax[1] - sz; Mj->R[2] = c1 * ax[0] * ax[2] + sy;
      Mj->R[3] = c1 * ax[1] * ax[0] + sz; Mj->R[4] = c1 * ax[1] * ax[1] + ca; Mj->R[5] = c1 * ax[1] * ax[2] - sx;
      Mj->R[6] = c1 * ax[2] * ax[0] - sy; Mj->R[7] = c1 * ax[2] * ax[1] + sx; Mj->R[8] = c1 * ax[2] * ax[2] + ca;
      S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2];
    } break;
    case RBD_JOINT_PX: Mj->p[0] = qi[0]; S[0] = 1.0; break;
    case RBD_JOINT_PY: Mj->p[1] = qi[0]; S[1] = 1.0; break;
    case RBD_JOINT_PZ: Mj->p[2] = qi[0]; S[2] = 1.0; break;
    case RBD_JOINT_PU:
      for (int k = 0; k < 3; ++k) { Mj->p[k] = qi[0] * ax[k]; S[k] = ax[k]; }
      break;
    case RBD_JOINT_FF: {
      /* Eigen QuaternionBase::toRotationMatrix on the raw (x,y,z,w) coefficients — no normalisation */
      double x = qi[3], y = qi[4], z = qi[5], s = qi[6];
      double tx = 2 * x, ty = 2 * y, tz = 2 * z;
      double twx = tx * s, twy = ty * s, twz = tz * s, txx = tx * x, txy = ty * x, txz = tz * x;
      double tyy = ty * y, tyz = tz * y, tzz = tz * z;
      Mj->R[0] = 1 - (tyy + tzz); Mj->R[1] = txy - twz; Mj->R[2] = txz + twy;
      Mj->R[3] = txy + twz; Mj->R[4] = 1 - (txx + tzz); Mj->R[5] = tyz - twx;
      Mj->R[6] = txz - twy; Mj->R[7] = tyz + twx; Mj->R[8] = 1 - (txx + tyy);
      Mj->p[0] = qi[0]; Mj->p[1] = qi[1]; Mj->p[2] = qi[2];
      for (int k = 0; k < 6; ++k) S[6 * k + k] = 1.0;
    } break;
    default: break;
  }
  if (vj) {
    memset(vj, 0, 6 * sizeof(double));
    if (v) {
      const double* vi = v + w->idx_v[i];
      for (int c = 0; c < w->nvj[i]; ++c)
        for (int k = 0; k < 6; ++k) vj[k] += S[6 * c + k] * vi[c];
    }
  }
}

/* pinocchio::computeJointJacobians(model, data, q)  (call site KCM.inl:373) */
void orc_compute_joint_jacobians(orc_workspace* w, const double* q) {
  for (int i = 1; i < w->njoints; ++i) {
    se3 Mj;
    double* S = w->S + 36 * i;
    joint_calc(w, i, q, NULL, &Mj, S, NULL);
    se3_mul(&w->jplace[i], &Mj, &w->liMi[i]);
    int p = w->parents[i];
    if (p > 0) se3_mul(&w->oMi[p], &w->liMi[i], &w->oMi[i]); else w->oMi[i] = w->liMi[i];
    for (int c = 0; c < w->nvj[i]; ++c) se3_act_motion(&w->oMi[i], S + 6 * c, w->J + 6 * (w->idx_v[i] + c));
  }
}

/* pinocchio::updateFramePlacements(model, data)  (call site KCM.inl:375) */
void orc_update_frame_placements(orc_workspace* w) {
  for (int f = 1; f < w->nframes; ++f) {
    int p = w->frame_parent[f];
    if (p > 0) se3_mul(&w->oMi[p], &w->fplace[f], &w->oMf[f]); else w->oMf[f] = w->fplace[f];
  }
}

/* pinocchio::getFrameJacobian(model, data, frame, LOCAL_WORLD_ALIGNED, J) on a caller-zeroed J
 * (call sites KCM.inl:434,444,480,490,579): only the support columns of the parent joint are written. */
void orc_get_frame_jacobian_lwa(orc_workspace* w, int frame, double* Jout) {
  int j = w->frame_parent[frame];
  se3 oMf;
  if (j > 0) se3_mul(&w->oMi[j], &w->fplace[frame], &oMf); else oMf = w->fplace[frame];
  w->oMf[frame] = oMf;
  if (j <= 0) return; /* universe: no column in its support */
  for (int c = w->idx_v[j] + w->nvj[j] - 1; c >= 0; c = w->parent_col[c]) {
    const double* in = w->J + 6 * c;
    double* out = Jout + 6 * c;
    double pxw[3];
    cross3(oMf.p, in + 3, pxw);
    for (int k = 0; k < 3; ++k) { out[k] = in[k] - pxw[k]; out[3 + k] = in[3 + k]; }
  }
}

/* Eigen 3.4 Quaternion(Matrix3) (QuaternionBase::operator=(MatrixBase), quat_product-free branch order),
 * reached through Conversions.h:60; output in SOFA order [x y z w] (Conversions.h:36). */
void orc_rot_to_quat(const double* R, double* xyzw) {
  double t = R[0] + R[4] + R[8];
  double qw, qv[3];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    qw = 0.5 * t;
    t = 0.5 / t;
    qv[0] = (R[7] - R[5]) * t; qv[1] = (R[2] - R[6]) * t; qv[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    qv[i] = 0.5 * t;
    t = 0.5 / t;
    qw = (R[3 * k + j] - R[3 * j + k]) * t;
    qv[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    qv[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  xyzw[0] = qv[0]; xyzw[1] = qv[1]; xyzw[2] = qv[2]; xyzw[3] = qw;
}

static int has_ff_root(const orc_workspace* w) { return w->njoints > 1 && w->jtype[1] == RBD_JOINT_FF && w->parents[1] == 0; }

/* KinematicChainMapping::apply worker, KCM.inl:338-396 */
void orc_apply(orc_workspace* w, const double* q_joints, const double* root_pose, double* out) {
  if (has_ff_root(w) && root_pose) {
    memcpy(w->qfull, root_pose, 7 * sizeof(double));                          /* KCM.inl:356-357 */
    memcpy(w->qfull + 7, q_joints, sizeof(double) * (size_t)(w->nq - 7));      /* KCM.inl:358 */
  } else {
    memcpy(w->qfull, q_joints, sizeof(double) * (size_t)w->nq);               /* KCM.inl:362 */
  }
  orc_compute_joint_jacobians(w, w->qfull);
  orc_update_frame_placements(w);
  for (int k = 0; k < w->n_out; ++k) {                                         /* KCM.inl:382-395 */
    const se3* M = &w->oMf[w->out_frames[k]];
    out[7 * k + 0] = M->p[0]; out[7 * k + 1] = M->p[1]; out[7 * k + 2] = M->p[2];
    orc_rot_to_quat(M->R, out + 7 * k + 3);
  }
  w->applied = 1;
}

static void assemble_dq(orc_workspace* w, const double* dq_joints, const double* root_vel) {
  if (has_ff_root(w) && root_vel) {
    memcpy(w->dqfull, root_vel, 6 * sizeof(double));                           /* KCM.inl:417-418 (verbatim) */
    memcpy(w->dqfull + 6, dq_joints, sizeof(double) * (size_t)(w->nv - 6));
  } else {
    memcpy(w->dqfull, dq_joints, sizeof(double) * (size_t)w->nv);
  }
}

/* applyJ worker, KCM.inl:398-448: dense frame Jacobian times dq per output */
void orc_applyJ(orc_workspace* w, const double* dq_joints, const double* root_vel, double* out) {
  assemble_dq(w, dq_joints, root_vel);
  for (int k = 0; k < w->n_out; ++k) {
    memset(w->Jf, 0, sizeof(double) * 6 * (size_t)w->nv);
    orc_get_frame_jacobian_lwa(w, w->out_frames[k], w->Jf);
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
      for (int c = 0; c < w->nv; ++c) s += w->Jf[6 * c + r] * w->dqfull[c];
      out[6 * k + r] = s;
    }
  }
}

/* applyJT (vector) worker, KCM.inl:450-513: tau = sum_k J_k^T in[k]; += into out / out_root */
void orc_applyJT(orc_workspace* w, const double* in, double* out_tau, double* out_root) {
  memset(w->tauacc, 0, sizeof(double) * (size_t)w->nv);
  for (int k = 0; k < w->n_out; ++k) {
    memset(w->Jf, 0, sizeof(double) * 6 * (size_t)w->nv);
    orc_get_frame_jacobian_lwa(w, w->out_frames[k], w->Jf);
    for (int c = 0; c < w->nv; ++c) {
      double s = 0.0;
      for (int r = 0; r < 6; ++r) s += w->Jf[6 * c + r] * in[6 * k + r];
      w->tauacc[c] += s;
    }
  }
  if (has_ff_root(w) && out_root) {
    for (int k = 0; k < 6; ++k) out_root[k] += w->tauacc[k];                   /* KCM.inl:498 */
    for (int i = 0; i < w->nv - 6; ++i) out_tau[i] += w->tauacc[6 + i];        /* KCM.inl:500-503 */
  } else {
    for (int i = 0; i < w->nv; ++i) out_tau[i] += w->tauacc[i];                /* KCM.inl:508-511 */
  }
}

/* one constraint row of applyJT(MatrixDeriv), KCM.inl:557-616.  Returns 1 when the row is kept. */
int orc_applyJT_row(orc_workspace* w, int nnz, const int* col, const double* val, double* row_out) {
  memset(w->tauacc, 0, sizeof(double) * (size_t)w->nv);
  for (int j = 0; j < nnz; ++j) {
    memset(w->Jf, 0, sizeof(double) * 6 * (size_t)w->nv);
    orc_get_frame_jacobian_lwa(w, w->out_frames[col[j]], w->Jf);              /* KCM.inl:566-579 */
    for (int c = 0; c < w->nv; ++c) {
      double s = 0.0;
      for (int r = 0; r < 6; ++r) s += w->Jf[6 * c + r] * val[6 * j + r];
      w->tauacc[c] += s;
    }
  }
  double n2 = 0.0;
  for (int c = 0; c < w->nv; ++c) { row_out[c] = w->tauacc[c]; n2 += w->tauacc[c] * w->tauacc[c]; }
  return n2 > 1e-12;                                                            /* KCM.inl:43-45,587,605 */
}

/* pinocchio::crba, Convention::LOCAL.  M: nv x nv row-major, upper triangle written, rest zero. */
void orc_crba(orc_workspace* w, const double* q, double* M) {
  int nj = w->njoints, nv = w->nv;
  memset(M, 0, sizeof(double) * (size_t)nv * (size_t)nv);
  for (int i = 1; i < nj; ++i) {
    se3 Mj;
    joint_calc(w, i, q, NULL, &Mj, w->S + 36 * i, NULL);
    se3_mul(&w->jplace[i], &Mj, &w->liMi[i]);
    w->Ycrb[i] = w->Y[i];
  }
  for (int i = nj - 1; i >= 1; --i) {
    double* Fi = w->Fcrb + (size_t)i * 6 * (size_t)nv;
    const double* S = w->S + 36 * i;
    int iv = w->idx_v[i], nvi = w->nvj[i], nsub = w->nv_subtree[i];
    for (int c = 0; c < nvi; ++c) inertia_mul_motion(&w->Ycrb[i], S + 6 * c, Fi + 6 * (iv + c));
    for (int r = 0; r < nvi; ++r)
      for (int c = 0; c < nsub; ++c) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += S[6 * r + k] * Fi[6 * (iv + c) + k];
        M[(size_t)(iv + r) * nv + (iv + c)] = s;
      }
    int p = w->parents[i];
    if (p > 0) {
      inertia t;
      inertia_se3_act(&w->liMi[i], &w->Ycrb[i], &t);
      inertia_add(&w->Ycrb[p], &t, &w->Ycrb[p]);
      double* Fp = w->Fcrb + (size_t)p * 6 * (size_t)nv;
      for (int c = 0; c < nsub; ++c) se3_act_force(&w->liMi[i], Fi + 6 * (iv + c), Fp + 6 * (iv + c));
    }
  }
}

/* pinocchio::rnea(model, data, q, v, a) */
void orc_rnea(orc_workspace* w, const double* q, const double* v, const double* a, double* tau) {
  int nj = w->njoints;
  memset(w->v, 0, 6 * sizeof(double));
  for (int k = 0; k < 3; ++k) { w->a_gf[k] = -w->gravity[k]; w->a_gf[3 + k] = 0.0; }
  for (int i = 1; i < nj; ++i) {
    se3 Mj;
    double* S = w->S + 36 * i;
    double* vj = w->vj + 6 * i;
    joint_calc(w, i, q, v, &Mj, S, vj);
    se3_mul(&w->jplace[i], &Mj, &w->liMi[i]);
    int p = w->parents[i];
    double* vi = w->v + 6 * i;
    double* ai = w->a_gf + 6 * i;
    memcpy(vi, vj, 6 * sizeof(double));
    if (p > 0) {
      double t[6];
      se3_actinv_motion(&w->liMi[i], w->v + 6 * p, t);
      for (int k = 0; k < 6; ++k) vi[k] += t[k];
    }
    double cr[6], t2[6];
    motion_cross(vi, vj, cr); /* joint bias c = 0 for every supported joint type */
    memcpy(ai, cr, sizeof cr);
    const double* aj = a + w->idx_v[i];
    for (int c = 0; c < w->nvj[i]; ++c)
      for (int k = 0; k < 6; ++k) ai[k] += S[6 * c + k] * aj[c];
    se3_actinv_motion(&w->liMi[i], w->a_gf + 6 * p, t2);
    for (int k = 0; k < 6; ++k) ai[k] += t2[k];
    inertia_mul_motion(&w->Y[i], vi, w->h + 6 * i);
    inertia_mul_motion(&w->Y[i], ai, w->f + 6 * i);
    double vxh[6];
    motion_cross_force(vi, w->h + 6 * i, vxh);
    for (int k = 0; k < 6; ++k) w->f[6 * i + k] += vxh[k];
  }
  for (int i = nj - 1; i >= 1; --i) {
    const double* S = w->S + 36 * i;
    for (int c = 0; c < w->nvj[i]; ++c) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += S[6 * c + k] * w->f[6 * i + k];
      tau[w->idx_v[i] + c] = s;
    }
    int p = w->parents[i];
    if (p > 0) {
      double t[6];
      se3_act_force(&w->liMi[i], w->f + 6 * i, t);
      for (int k = 0; k < 6; ++k) w->f[6 * p + k] += t[k];
    }
  }
}

/* ---- forward dynamics ------------------------------------------------------------------------------------------- */
/* dense 6x6 helpers (row-major) for the articulated-body inertias */
static void inertia_to_mat6(const inertia* Y, double* M) {
  /* InertiaTpl::matrix(): [[m I, -m [c]x], [m [c]x, I_c - m [c]x [c]x]] */
  const double m = Y->m, *c = Y->c;
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  double Ic[9] = {Y->I[0], Y->I[1], Y->I[3], Y->I[1], Y->I[2], Y->I[4], Y->I[3], Y->I[4], Y->I[5]};
  double cc[9];
  matmul3(cx, cx, cc);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      M[6 * r + k] = r == k ? m : 0.0;
      M[6 * r + 3 + k] = -m * cx[3 * r + k];
      M[6 * (3 + r) + k] = m * cx[3 * r + k];
      M[6 * (3 + r) + 3 + k] = Ic[3 * r + k] - m * cc[3 * r + k];
    }
}
/* X_f of aMb (force transform b -> a): [[R, 0], [[p]x R, R]] */
static void se3_force_matrix(const se3* M, double* X) {
  double px[9] = {0, -M->p[2], M->p[1], M->p[2], 0, -M->p[0], -M->p[1], M->p[0], 0}, pR[9];
  matmul3(px, M->R, pR);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      X[6 * r + k] = M->R[3 * r + k]; X[6 * r + 3 + k] = 0.0;
      X[6 * (3 + r) + k] = pR[3 * r + k]; X[6 * (3 + r) + 3 + k] = M->R[3 * r + k];
    }
}
/* solve the n x n system A x = b (A symmetric positive definite, row-major, n <= 6) by Gaussian elimination with
 * partial pivoting; A and b are overwritten */
static void solve_small(int n, double* A, double* b) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int r = k + 1; r < n; ++r) if (fabs(A[6 * r + k]) > fabs(A[6 * piv + k])) piv = r;
    if (piv != k) {
      for (int c = 0; c < n; ++c) { double t = A[6 * k + c]; A[6 * k + c] = A[6 * piv + c]; A[6 * piv + c] = t; }
      double t = b[k]; b[k] = b[piv]; b[piv] = t;
    }
    for (int r = k + 1; r < n; ++r) {
      double f = A[6 * r + k] / A[6 * k + k];
      for (int c = k; c < n; ++c) A[6 * r + c] -= f * A[6 * k + c];
      b[r] -= f * b[k];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int c = k + 1; c < n; ++c) s -= A[6 * k + c] * b[c];
    b[k] = s / A[6 * k + k];
  }
}

/* pinocchio::aba(model, data, q, v, tau), Convention::LOCAL (algorithm/aba.hxx: AbaLocalConventionForwardStep1,
 * AbaLocalConventionBackwardStep, AbaLocalConventionForwardStep2; JointModel::calc_aba: U = Ia S, Dinv = (S^T U)^-1,
 * UDinv = U Dinv, Ia -= UDinv U^T).  North-star extension: no reference call site.
 * Joint frames, dense 6x6 articulated inertias, SE3 actions between parent and child — a formulation that shares no
 * code path with the kernels' world-frame sweep.  ddq [nv]. */
void orc_aba(orc_workspace* w, const double* q, const double* v, const double* tau, double* ddq) {
  int nj = w->njoints, nv = w->nv;
  double* Yaba = (double*)calloc((size_t)nj * 36, sizeof(double));
  double* U = (double*)calloc((size_t)nj * 36, sizeof(double));     /* [joint][6 x nvj] col-major */
  double* UDinv = (double*)calloc((size_t)nj * 36, sizeof(double));
  double* Dinv = (double*)calloc((size_t)nj * 36, sizeof(double));  /* [joint][nvj x nvj] row-major, stride 6 */
  double* u = (double*)calloc((size_t)(nv > 0 ? nv : 1), sizeof(double));
  memcpy(u, tau, sizeof(double) * (size_t)nv);
  memset(w->v, 0, 6 * sizeof(double));
  /* pass 1 */
  for (int i = 1; i < nj; ++i) {
    se3 Mj;
    double* S = w->S + 36 * i;
    double* vj = w->vj + 6 * i;
    joint_calc(w, i, q, v, &Mj, S, vj);
    se3_mul(&w->jplace[i], &Mj, &w->liMi[i]);
    int p = w->parents[i];
    double* vi = w->v + 6 * i;
    memcpy(vi, vj, 6 * sizeof(double));
    if (p > 0) {
      double t[6];
      se3_actinv_motion(&w->liMi[i], w->v + 6 * p, t);
      for (int k = 0; k < 6; ++k) vi[k] += t[k];
    }
    motion_cross(vi, vj, w->a_gf + 6 * i);                       /* joint bias c = 0 */
    inertia_to_mat6(&w->Y[i], Yaba + 36 * i);
    inertia_mul_motion(&w->Y[i], vi, w->h + 6 * i);
    motion_cross_force(vi, w->h + 6 * i, w->f + 6 * i);
  }
  /* pass 2 */
  for (int i = nj - 1; i >= 1; --i) {
    const double* S = w->S + 36 * i;
    double* Ia = Yaba + 36 * i;
    int iv = w->idx_v[i], nvi = w->nvj[i], p = w->parents[i];
    for (int c = 0; c < nvi; ++c) {
      double s = 0.0;
      for (int k = 0; k < 6; ++k) s += S[6 * c + k] * w->f[6 * i + k];
      u[iv + c] -= s;
    }
    double* Ui = U + 36 * i; double* UDi = UDinv + 36 * i; double* Di = Dinv + 36 * i;
    for (int c = 0; c < nvi; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < 6; ++k) s += Ia[6 * r + k] * S[6 * c + k];
        Ui[6 * c + r] = s;
      }
    /* Dinv = (S^T U)^-1, column by column */
    for (int c = 0; c < nvi; ++c) {
      double D[36], e[6] = {0, 0, 0, 0, 0, 0};
      for (int r = 0; r < nvi; ++r)
        for (int cc = 0; cc < nvi; ++cc) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += S[6 * r + k] * Ui[6 * cc + k];
          D[6 * r + cc] = s;
        }
      e[c] = 1.0;
      solve_small(nvi, D, e);
      for (int r = 0; r < nvi; ++r) Di[6 * r + c] = e[r];
    }
    for (int c = 0; c < nvi; ++c)
      for (int r = 0; r < 6; ++r) {
        double s = 0.0;
        for (int k = 0; k < nvi; ++k) s += Ui[6 * k + r] * Di[6 * k + c];
        UDi[6 * c + r] = s;
      }
    if (p > 0) {
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) {
          double s = 0.0;
          for (int k = 0; k < nvi; ++k) s += UDi[6 * k + r] * Ui[6 * k + cc];
          Ia[6 * r + cc] -= s;
        }
      double pa[6];
      for (int r = 0; r < 6; ++r) {
        double s = w->f[6 * i + r];
        for (int k = 0; k < 6; ++k) s += Ia[6 * r + k] * w->a_gf[6 * i + k];
        for (int k = 0; k < nvi; ++k) s += UDi[6 * k + r] * u[iv + k];
        pa[r] = s;
      }
      double X[36], XI[36];
      se3_force_matrix(&w->liMi[i], X);
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += X[6 * r + k] * Ia[6 * k + cc];
          XI[6 * r + cc] = s;
        }
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc < 6; ++cc) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += XI[6 * r + k] * X[6 * cc + k];
          Yaba[36 * p + 6 * r + cc] += s;
        }
      double t[6];
      se3_act_force(&w->liMi[i], pa, t);
      for (int k = 0; k < 6; ++k) w->f[6 * p + k] += t[k];
    }
  }
  /* pass 3 */
  for (int k = 0; k < 3; ++k) { w->a_gf[k] = -w->gravity[k]; w->a_gf[3 + k] = 0.0; }
  for (int i = 1; i < nj; ++i) {
    const double* S = w->S + 36 * i;
    int iv = w->idx_v[i], nvi = w->nvj[i], p = w->parents[i];
    double t[6];
    se3_actinv_motion(&w->liMi[i], w->a_gf + 6 * p, t);
    for (int k = 0; k < 6; ++k) w->a_gf[6 * i + k] += t[k];
    const double* UDi = UDinv + 36 * i; const double* Di = Dinv + 36 * i;
    for (int c = 0; c < nvi; ++c) {
      double s = 0.0;
      for (int k = 0; k < nvi; ++k) s += Di[6 * c + k] * u[iv + k];
      for (int k = 0; k < 6; ++k) s -= UDi[6 * c + k] * w->a_gf[6 * i + k];
      ddq[iv + c] = s;
    }
    for (int c = 0; c < nvi; ++c)
      for (int k = 0; k < 6; ++k) w->a_gf[6 * i + k] += S[6 * c + k] * ddq[iv + c];
  }
  free(Yaba); free(U); free(UDinv); free(Dinv); free(u);
}

/* ---- analytical derivatives of inverse dynamics ------------------------------------------------------------------ */
/* y = A x, A 6x6 row-major */
static void mat6_mulvec(const double* A, const double* x, double* y) {
  for (int r = 0; r < 6; ++r) {
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += A[6 * r + k] * x[k];
    y[r] = s;
  }
}
/* motion cross-product matrix: crm(v) m = v x m, [[w x, v x], [0, w x]] on [lin; ang] */
static void motion_cross_matrix(const double* v, double* X) {
  const double vx[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
  const double wx[9] = {0, -v[5], v[4], v[5], 0, -v[3], -v[4], v[3], 0};
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) {
      X[6 * r + k] = wx[3 * r + k]; X[6 * r + 3 + k] = vx[3 * r + k];
      X[6 * (3 + r) + k] = 0.0; X[6 * (3 + r) + 3 + k] = wx[3 * r + k];
    }
}

/* pinocchio::computeRNEADerivatives(model, data, q, v, a, dtau_dq, dtau_dv, dtau_da) (algorithm/rnea-derivatives.hxx:
 * ComputeRNEADerivativesForwardStep / BackwardStep; Carpentier & Mansard, RSS 2018).  North-star extension (SURVEY.md
 * §8(f) rank 4): no reference call site.  Restated with pinocchio's data layout — every joint's world-frame quantities
 * kept in arrays, dense 6x6 oYcrb / doYcrb, dense 6 x nv dVdq / dAdq / dAdv / dFdq / dFdv / dFda, a full forward pass
 * followed by a full backward pass — which shares no code path with the kernels' depth-first sweep on compact forms.
 * dtau_dq, dtau_dv: nv x nv row-major (pinocchio's RowMatrixXs), entry (r, c) = d tau_r / d q_c (tangent of integrate)
 * resp. d tau_r / d v_c; structural zeros written.  M (optional): dtau_da, upper triangle filled as by pinocchio, rest 0. */
void orc_rnea_derivatives(orc_workspace* w, const double* q, const double* v, const double* a, double* dtau_dq,
                          double* dtau_dv, double* M) {
  const int nj = w->njoints, nv = w->nv;
  const size_t snj = (size_t)nj, snv = (size_t)(nv > 0 ? nv : 1);
  double* ov = (double*)calloc(snj * 6, sizeof(double));
  double* oa_gf = (double*)calloc(snj * 6, sizeof(double));
  double* of = (double*)calloc(snj * 6, sizeof(double));
  double* oY = (double*)calloc(snj * 36, sizeof(double));   /* oYcrb */
  double* doY = (double*)calloc(snj * 36, sizeof(double));  /* doYcrb */
  double* dVdq = (double*)calloc(6 * snv, sizeof(double));  /* 6 x nv, column-major like data.J */
  double* dAdq = (double*)calloc(6 * snv, sizeof(double));
  double* dAdv = (double*)calloc(6 * snv, sizeof(double));
  double* dFdq = (double*)calloc(6 * snv, sizeof(double));
  double* dFdv = (double*)calloc(6 * snv, sizeof(double));
  double* dFda = (double*)calloc(6 * snv, sizeof(double));
  memset(dtau_dq, 0, sizeof(double) * snv * snv);
  memset(dtau_dv, 0, sizeof(double) * snv * snv);
  if (M) memset(M, 0, sizeof(double) * snv * snv);
  for (int k = 0; k < 3; ++k) oa_gf[k] = -w->gravity[k];  /* data.oa_gf[0] = -gravity */
  /* forward step */
  for (int i = 1; i < nj; ++i) {
    se3 Mj;
    double* S = w->S + 36 * i;
    double* vj = w->vj + 6 * i;
    joint_calc(w, i, q, v, &Mj, S, vj);
    se3_mul(&w->jplace[i], &Mj, &w->liMi[i]);
    const int p = w->parents[i], iv = w->idx_v[i], nvi = w->nvj[i];
    double* vi = w->v + 6 * i;       /* data.v[i], joint frame */
    double* ai = w->a_gf + 6 * i;    /* data.a[i], joint frame, WITHOUT gravity */
    memcpy(vi, vj, 6 * sizeof(double));
    if (p > 0) {
      double t[6];
      se3_mul(&w->oMi[p], &w->liMi[i], &w->oMi[i]);
      se3_actinv_motion(&w->liMi[i], w->v + 6 * p, t);
      for (int k = 0; k < 6; ++k) vi[k] += t[k];
    } else {
      w->oMi[i] = w->liMi[i];
    }
    motion_cross(vi, vj, ai); /* joint bias c = 0 */
    for (int c = 0; c < nvi; ++c)
      for (int k = 0; k < 6; ++k) ai[k] += S[6 * c + k] * a[iv + c];
    if (p > 0) {
      double t[6];
      se3_actinv_motion(&w->liMi[i], w->a_gf + 6 * p, t);
      for (int k = 0; k < 6; ++k) ai[k] += t[k];
    }
    inertia Yw;
    inertia_se3_act(&w->oMi[i], &w->Y[i], &Yw);
    inertia_to_mat6(&Yw, oY + 36 * i);
    double oa[6], oh[6], t6[6];
    se3_act_motion(&w->oMi[i], vi, ov + 6 * i);
    se3_act_motion(&w->oMi[i], ai, oa);
    for (int k = 0; k < 6; ++k) oa_gf[6 * i + k] = oa[k];
    for (int k = 0; k < 3; ++k) oa_gf[6 * i + k] -= w->gravity[k];
    mat6_mulvec(oY + 36 * i, ov + 6 * i, oh);
    mat6_mulvec(oY + 36 * i, oa_gf + 6 * i, of + 6 * i);
    motion_cross_force(ov + 6 * i, oh, t6);
    for (int k = 0; k < 6; ++k) of[6 * i + k] += t6[k];
    for (int c = 0; c < nvi; ++c) {
      double* Jc = w->J + 6 * (iv + c);
      double dJ[6];
      se3_act_motion(&w->oMi[i], S + 6 * c, Jc);
      motion_cross(ov + 6 * i, Jc, dJ);
      motion_cross(oa_gf + 6 * p, Jc, dAdq + 6 * (iv + c));
      memcpy(dAdv + 6 * (iv + c), dJ, sizeof dJ);
      if (p > 0) {
        double t[6];
        motion_cross(ov + 6 * p, Jc, dVdq + 6 * (iv + c));
        motion_cross(ov + 6 * p, dVdq + 6 * (iv + c), t);
        for (int k = 0; k < 6; ++k) { dAdq[6 * (iv + c) + k] += t[k]; dAdv[6 * (iv + c) + k] += dVdq[6 * (iv + c) + k]; }
      } else {
        memset(dVdq + 6 * (iv + c), 0, 6 * sizeof(double));
      }
    }
    /* doYcrb = oYcrb.variation(ov) + forceCrossMatrix(oh):  crf(v) Y - Y crm(v), crf(v) = -crm(v)^T */
    {
      double X[36];
      double* D = doY + 36 * i;
      const double* Y6 = oY + 36 * i;
      motion_cross_matrix(ov + 6 * i, X);
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += -X[6 * k + r] * Y6[6 * k + c] - Y6[6 * r + k] * X[6 * k + c];
          D[6 * r + c] = s;
        }
      /* addForceCrossMatrix: -[f]x on (lin, ang) and (ang, lin), -[n]x on (ang, ang) */
      const double fx[9] = {0, -oh[2], oh[1], oh[2], 0, -oh[0], -oh[1], oh[0], 0};
      const double nx[9] = {0, -oh[5], oh[4], oh[5], 0, -oh[3], -oh[4], oh[3], 0};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          D[6 * r + 3 + c] -= fx[3 * r + c];
          D[6 * (3 + r) + c] -= fx[3 * r + c];
          D[6 * (3 + r) + 3 + c] -= nx[3 * r + c];
        }
    }
  }
  /* backward step */
  for (int i = nj - 1; i >= 1; --i) {
    const int p = w->parents[i], iv = w->idx_v[i], nvi = w->nvj[i], nsub = w->nv_subtree[i];
    const double* Y6 = oY + 36 * i;
    const double* D = doY + 36 * i;
    for (int c = 0; c < nvi; ++c) {
      const double* Jc = w->J + 6 * (iv + c);
      double t[6];
      /* dtau/dv */
      mat6_mulvec(D, Jc, dFdv + 6 * (iv + c));
      mat6_mulvec(Y6, dAdv + 6 * (iv + c), t);
      for (int k = 0; k < 6; ++k) dFdv[6 * (iv + c) + k] += t[k];
      /* dtau/dq */
      if (p > 0) {
        mat6_mulvec(D, dVdq + 6 * (iv + c), dFdq + 6 * (iv + c));
        mat6_mulvec(Y6, dAdq + 6 * (iv + c), t);
        for (int k = 0; k < 6; ++k) dFdq[6 * (iv + c) + k] += t[k];
      } else {
        mat6_mulvec(Y6, dAdq + 6 * (iv + c), dFdq + 6 * (iv + c));
      }
      mat6_mulvec(Y6, Jc, dFda + 6 * (iv + c));
    }
    for (int r = 0; r < nvi; ++r)
      for (int c = 0; c < nsub; ++c) {
        double sv = 0.0, sq = 0.0, sa = 0.0;
        for (int k = 0; k < 6; ++k) {
          sv += w->J[6 * (iv + r) + k] * dFdv[6 * (iv + c) + k];
          sq += w->J[6 * (iv + r) + k] * dFdq[6 * (iv + c) + k];
          sa += w->J[6 * (iv + r) + k] * dFda[6 * (iv + c) + k];
        }
        dtau_dv[(size_t)(iv + r) * nv + (iv + c)] = sv;
        dtau_dq[(size_t)(iv + r) * nv + (iv + c)] = sq;
        if (M) M[(size_t)(iv + r) * nv + (iv + c)] = sa;
      }
    for (int c = 0; c < nvi; ++c) {  /* motionSet::act<ADDTO>(J_cols, of[i], dFdq_cols) */
      double t[6];
      motion_cross_force(w->J + 6 * (iv + c), of + 6 * i, t);
      for (int k = 0; k < 6; ++k) dFdq[6 * (iv + c) + k] += t[k];
    }
    /* columns of the ancestors' dofs (parents_fromRow walk) */
    for (int r = 0; r < nvi; ++r) {
      double YJ[6], DJ[6];  /* rows of J^T oYcrb and J^T doYcrb */
      for (int c = 0; c < 6; ++c) {
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < 6; ++k) { s1 += w->J[6 * (iv + r) + k] * Y6[6 * k + c]; s2 += w->J[6 * (iv + r) + k] * D[6 * k + c]; }
        YJ[c] = s1; DJ[c] = s2;
      }
      for (int j = w->parent_col[iv]; j >= 0; j = w->parent_col[j]) {
        double sq = 0.0, sv = 0.0;
        for (int k = 0; k < 6; ++k) {
          sq += YJ[k] * dAdq[6 * j + k] + DJ[k] * dVdq[6 * j + k];
          sv += YJ[k] * dAdv[6 * j + k] + DJ[k] * w->J[6 * j + k];
        }
        dtau_dq[(size_t)(iv + r) * nv + j] = sq;
        dtau_dv[(size_t)(iv + r) * nv + j] = sv;
      }
    }
    if (p > 0) {
      for (int k = 0; k < 36; ++k) { oY[36 * p + k] += oY[36 * i + k]; doY[36 * p + k] += doY[36 * i + k]; }
      for (int k = 0; k < 6; ++k) of[6 * p + k] += of[6 * i + k];
    }
  }
  free(ov); free(oa_gf); free(of); free(oY); free(doY); free(dVdq); free(dAdq); free(dAdv); free(dFdq); free(dFdv); free(dFda);
}

/* One full evaluation in the algorithmic output format of SURVEY.md §8(d):
 * oMi [12(nj-1)], J [6 nv] col-major, M packed upper by columns, tau [nv].  Any output may be NULL. */
void orc_eval(orc_workspace* w, const double* q, const double* v, const double* a, double* oMi, double* J,
              double* Mp, double* tau) {
  int nv = w->nv;
  if (oMi || J) {
    orc_compute_joint_jacobians(w, q);
    if (oMi)
      for (int i = 1; i < w->njoints; ++i) {
        memcpy(oMi + 12 * (i - 1), w->oMi[i].R, 9 * sizeof(double));
        memcpy(oMi + 12 * (i - 1) + 9, w->oMi[i].p, 3 * sizeof(double));
      }
    if (J) memcpy(J, w->J, sizeof(double) * 6 * (size_t)nv);
  }
  if (Mp) {
    orc_crba(w, q, w->M);
    for (int c = 0; c < nv; ++c)
      for (int r = 0; r <= c; ++r) Mp[c * (c + 1) / 2 + r] = w->M[(size_t)r * nv + c];
  }
  if (tau) orc_rnea(w, q, v, a, tau);
}

/* Batched evaluation over instance-major arrays on `nthreads` POSIX threads (one workspace per thread,
 * contiguous instance ranges): the CPU baseline timed by bench.py.  nthreads <= 0 -> all online cores.
 * Returns the thread count actually used. */
typedef struct {
  const rbd_model_desc* d; long lo, hi; const double *q, *v, *a; double *oMi, *J, *Mp, *tau;
} batch_job;

static void* batch_worker(void* arg) {
  batch_job* b = (batch_job*)arg;
  const rbd_model_desc* d = b->d;
  long nq = d->nq, nv = d->nv, no = 12L * (d->njoints - 1), njac = 6L * nv, nm = nv * (nv + 1) / 2;
  orc_workspace* w = orc_create(d);
  for (long i = b->lo; i < b->hi; ++i)
    orc_eval(w, b->q + i * nq, b->v + i * nv, b->a + i * nv, b->oMi ? b->oMi + i * no : NULL,
             b->J ? b->J + i * njac : NULL, b->Mp ? b->Mp + i * nm : NULL, b->tau ? b->tau + i * nv : NULL);
  orc_destroy(w);
  return NULL;
}

int orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int orc_eval_batch(const rbd_model_desc* d, long n, const double* q, const double* v, const double* a,
                   double* oMi, double* J, double* Mp, double* tau, int nthreads) {
  if (nthreads <= 0) nthreads = orc_max_threads();
  if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    batch_job b = {d, n * t / nthreads, n * (t + 1) / nthreads, q, v, a, oMi, J, Mp, tau};
    jobs[t] = b;
    if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  batch_worker(&jobs[0]);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th); free(jobs);
  return nthreads;
}

#ifdef ORC_COUNT_FLOPS
#undef double
/* counts of the calling thread since the last reset: add/sub, mul, div, sqrt, sin+cos evaluations, comparisons */
extern "C" void orc_flop_reset(void) { orc_counters z = {0, 0, 0, 0, 0, 0}; orc_cnt = z; }
extern "C" void orc_flop_counts(long long out[6]) {
  out[0] = orc_cnt.add; out[1] = orc_cnt.mul; out[2] = orc_cnt.div; out[3] = orc_cnt.sqrt_; out[4] = orc_cnt.sincos;
  out[5] = orc_cnt.cmp;
}
#define double orc_counted
#endif

/* The mapping's apply for n robots one after the other on the calling thread (what n KinematicChainMapping components
 * of one scene cost the reference per propagation pass): experiments/latency_table.py times it next to the batched
 * back end.  q_joints [n][nq_joints], root_pose [n][7] or NULL, out [n][n_out][7]. */
void orc_apply_loop(orc_workspace* w, long n, const double* q_joints, const double* root_pose, double* out) {
  const long kq = has_ff_root(w) && root_pose ? w->nq - 7 : w->nq;
  for (long i = 0; i < n; ++i)
    orc_apply(w, q_joints + i * kq, root_pose ? root_pose + i * 7 : NULL, out + i * 7L * w->n_out);
}

/* accessors used by the tests */
const double* orc_data_J(const orc_workspace* w) { return w->J; }
void orc_get_oMi(const orc_workspace* w, int i, double* out12) {
  memcpy(out12, w->oMi[i].R, 9 * sizeof(double)); memcpy(out12 + 9, w->oMi[i].p, 3 * sizeof(double));
}
void orc_get_oMf(const orc_workspace* w, int f, double* out12) {
  memcpy(out12, w->oMf[f].R, 9 * sizeof(double)); memcpy(out12 + 9, w->oMf[f].p, 3 * sizeof(double));
}
