
// ---------------------------------------------------------------------------------------------------------
// accessors
//
// FieldT<TILE>: strided view of one instance's fields in a tiled-SoA array.  TILE = 0 reads the tile (= stride
// between consecutive fields) at run time; TILE = 32 is the AoSoA layout the hot path is tuned for: the 32
// instances of a warp form one tile, field k of the warp sits at byte offset k*256 from the warp's base, so the
// stride folds into the immediate offset of every load/store and a warp's whole output is one contiguous block.
// rbd::chain instantiation (RBD_CHAIN_ONLY): fixed-base serial chains of RX / RY / RZ joints (Panda-class arms) — no
// free-flyer code, no branch records, no return-to-branch restores: fewer registers, so more warps fit an SM.
#undef RBD_T_IS_FF
#undef RBD_CAN_BRANCH
#undef RBD_EARLY_RESTORE_NS
#if defined(RBD_FF_ROOT_ONLY) && !defined(RBD_FF_ROOT_WIDE_REGS)
#define RBD_EARLY_RESTORE_NS 0 /* 12-warp blocks at 168 registers: the early reload would spill (ptxas: 88 bytes) */
#else
#define RBD_EARLY_RESTORE_NS RBD_EARLY_RESTORE
#endif
#undef RBD_CUR_FF
#undef RBD_ANC_FF
#undef RBD_EMIT_CUR
#undef RBD_EMIT_ANC
struct FfYes { static constexpr bool value = true; };
struct FfNo { static constexpr bool value = false; };
#if defined(RBD_FF_ROOT_ONLY)
// rbd::ffroot instantiation: joint 1 is a free-flyer attached to the universe, it is the universe's only child and the
// only free-flyer of the model (Solo, Talos: every floating-base robot).  The root is stepped in front of the joint
// loop and emitted behind it; inside the loop "the current joint" and "an ancestor being emitted" are never
// free-flyers, at compile time.
#define RBD_CUR_FF(ffj, t) (decltype(ffj)::value)
#define RBD_ANC_FF(t) false
#define RBD_EMIT_CUR(ffj) (decltype(ffj)::value ? 1 : 2)
#define RBD_EMIT_ANC 2
#else
#define RBD_CUR_FF(ffj, t) RBD_T_IS_FF(t)
#define RBD_ANC_FF(t) RBD_T_IS_FF(t)
#define RBD_EMIT_CUR(ffj) 0
#define RBD_EMIT_ANC 0
#endif
#if defined(RBD_CHAIN_ONLY)
#define RBD_T_IS_FF(t) false
#define RBD_CAN_BRANCH false
#else
#define RBD_T_IS_FF(t) ((t) == JT_FF)
#define RBD_CAN_BRANCH true
#endif
template <int TILE>
struct FieldT {
  real* p;
  long long stride;
  RBD_HD long long off(int k) const { return TILE ? (long long)k * TILE : (long long)k * stride; }
  RBD_HD real ld(int k) const { return p[off(k)]; }
  RBD_HD real ldz(int k) const { return p ? p[off(k)] : real(0); }  // an absent input array reads as zeros
  template <int N>
  RBD_HD void ldn(int k, real* v) const {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] = ld(k + e);
  }
  RBD_HD void st(int k, real v) const { p[off(k)] = v; }
  RBD_HD bool ok() const { return p != nullptr; }
  // address of field k, and a store at a (small, usually compile-time) field offset e from such an address: with
  // TILE fixed the offset folds into the instruction's immediate instead of costing one IMAD.WIDE per store
  RBD_HD real* at(int k) const { return p + off(k); }
  RBD_HD void st_at(real* a, int e, real v) const { a[TILE ? (long long)e * TILE : (long long)e * stride] = v; }
  RBD_HD long long step(int e) const { return TILE ? (long long)e * TILE : (long long)e * stride; }
  // address `elems` elements behind this lane's field 0 (32-bit offset: one IMAD.WIDE)
  RBD_HD const real* at_elems(int elems) const { return p + elems; }
  // The same store / load with the L2 eviction priority "evict_last": for the few lines this thread re-reads much
  // later (R|p of a branch joint out of its own oMi output, the root free-flyer's v / a) — without the hint the
  // kernel's 5 TB/s write stream evicts them from L2 within ~25 us and the re-read goes to HBM behind that stream
  // (profile r1g: 10 % of all stall samples on the first consumer of the re-read rotation).
  RBD_HD void st_at_keep(real* a, int e, real v) const {
#if defined(__CUDA_ARCH__) && !defined(RBD_NO_L2_KEEP)
    real* addr = a + step(e);
    if constexpr (sizeof(real) == 8)
      asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"((double)v), "l"(l2_keep_policy()) : "memory");
    else
      asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(addr), "f"((float)v), "l"(l2_keep_policy()) : "memory");
#else
    st_at(a, e, v);
#endif
  }
  RBD_HD real ld_keep(int k) const {
#if defined(__CUDA_ARCH__) && !defined(RBD_NO_L2_KEEP)
    const real* addr = p + off(k);
    if constexpr (sizeof(real) == 8) {
      double v;
      asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(addr), "l"(l2_keep_policy()));
      return (real)v;
    } else {
      float v;
      asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(addr), "l"(l2_keep_policy()));
      return (real)v;
    }
#else
    return ld(k);
#endif
  }
#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
  }
#endif
};
using Field = FieldT<0>;
template <int TILE>
RBD_HD FieldT<TILE> make_field_t(real* base, long long inst, int K, long long tile) {
  FieldT<TILE> f;
  if constexpr (TILE != 0) {
    f.stride = TILE;
    f.p = base ? base + (inst / TILE) * (long long)K * TILE + (inst % TILE) : nullptr;
  } else {
    f.stride = tile;
    f.p = base ? base + (inst / tile) * (long long)K * tile + (inst % tile) : nullptr;
  }
  return f;
}
RBD_HD Field make_field(real* base, long long inst, int K, long long tile) { return make_field_t<0>(base, inst, K, tile); }

// Stack: this thread's column of its warp's shared-memory state.  Every warp owns a contiguous [slots][32]
// region, so the stride between slots is the compile-time constant 32 (bank-conflict free: the 32 lanes of a
// warp touch 32 consecutive doubles), whatever the block size.
#define RBD_WARP 32
struct Stack {
  real* p;
  RBD_HD real ld(int k) const { return p[k * RBD_WARP]; }
  RBD_HD void st(int k, real v) const { p[k * RBD_WARP] = v; }
  RBD_HD void add(int k, real v) const { p[k * RBD_WARP] += v; }
};
// base of thread t's stack column inside a block-wide region of `slots` doubles per thread
RBD_HD Stack make_stack(real* smem, int t, int slots) {
  Stack s;
  s.p = smem + (long long)(t / RBD_WARP) * slots * RBD_WARP + (t % RBD_WARP);
  return s;
}

// TmStack: this thread's private slice of TENSOR MEMORY used as scratch (tcgen05.ld / tcgen05.st, shape 32x32b:
// lane i of the warp <-> TMEM lane 32*(warp%4)+i, one 32-bit column per register; a real takes two columns).
// `base` already contains the warp's lane quarter (bits 31:16) and the thread's first column (bits 15:0).
// All lanes of a warp must execute these together (.sync.aligned): the eval sweep is warp-uniform by construction.
// On the host (tests/emul) the same interface is backed by a plain array.
#if defined(__CUDACC__)
struct TmStack {
  unsigned base;
  // NC consecutive 32-bit columns -> registers, in power-of-two pieces
  template <int NC>
  __device__ __forceinline__ void issue_ld(unsigned addr, unsigned* r) const {
#if defined(__CUDA_ARCH__)
    if constexpr (NC >= 16) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(addr) : "memory");
      issue_ld<NC - 16>(addr + 16, r + 16);
    } else if constexpr (NC >= 8) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(addr) : "memory");
      issue_ld<NC - 8>(addr + 8, r + 8);
    } else if constexpr (NC >= 4) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
      issue_ld<NC - 4>(addr + 4, r + 4);
    } else if constexpr (NC >= 2) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
      issue_ld<NC - 2>(addr + 2, r + 2);
    } else if constexpr (NC == 1) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(addr) : "memory");
    }
#endif
  }
  template <int NC>
  __device__ __forceinline__ void issue_st(unsigned addr, const unsigned* r) const {
#if defined(__CUDA_ARCH__)
    if constexpr (NC >= 16) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
          :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
             "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
      issue_st<NC - 16>(addr + 16, r + 16);
    } else if constexpr (NC >= 8) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
      issue_st<NC - 8>(addr + 8, r + 8);
    } else if constexpr (NC >= 4) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                   :: "r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
      issue_st<NC - 4>(addr + 4, r + 4);
    } else if constexpr (NC >= 2) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(addr), "r"(r[0]), "r"(r[1]) : "memory");
      issue_st<NC - 2>(addr + 2, r + 2);
    } else if constexpr (NC == 1) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(addr), "r"(r[0]) : "memory");
    }
#endif
  }
  // v[0..N) <- values [k, k+N)   (a double takes two columns, a float one)
  template <int N>
  __host__ __device__ __forceinline__ void ldv(int k, real* v) const {
#if defined(__CUDA_ARCH__)
    constexpr int C = (int)(sizeof(real) / 4);
    unsigned r[C * N];
    issue_ld<C * N>(base + (unsigned)(C * k), r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if constexpr (C == 2) v[i] = (real)__hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
      else v[i] = (real)__uint_as_float(r[i]);
    }
#endif
  }
  // values [k, k+N) += v[0..N)
  template <int N>
  __host__ __device__ __forceinline__ void addv(int k, const real* v) const {
    real t[N];
    ldv<N>(k, t);
#pragma unroll
    for (int i = 0; i < N; ++i) t[i] += v[i];
    stv<N>(k, t);
  }
  // values [k, k+N) <- v[0..N); complete before returning (a later ldv of the same columns must see it)
  template <int N>
  __host__ __device__ __forceinline__ void stv(int k, const real* v) const {
#if defined(__CUDA_ARCH__)
    constexpr int C = (int)(sizeof(real) / 4);
    unsigned r[C * N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if constexpr (C == 2) {
        r[2 * i] = (unsigned)__double2loint((double)v[i]); r[2 * i + 1] = (unsigned)__double2hiint((double)v[i]);
      } else {
        r[i] = __float_as_uint((float)v[i]);
      }
    }
    issue_st<C * N>(base + (unsigned)(C * k), r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
  }
};
#else
struct TmStack {
  real* p;
  template <int N> void ldv(int k, real* v) const { for (int i = 0; i < N; ++i) v[i] = p[k + i]; }
  template <int N> void stv(int k, const real* v) const { for (int i = 0; i < N; ++i) p[k + i] = v[i]; }
  template <int N> void addv(int k, const real* v) const { for (int i = 0; i < N; ++i) p[k + i] += v[i]; }
};
#endif

// ---------------------------------------------------------------------------------------------------------
// 3-vector helpers on scalars held in registers
#define RBD_CROSS(ox, oy, oz, ax, ay, az, bx, by, bz) \
  do {                                                \
    ox = (ay) * (bz) - (az) * (by);                   \
    oy = (az) * (bx) - (ax) * (bz);                   \
    oz = (ax) * (by) - (ay) * (bx);                   \
  } while (0)

// sin and cos of a joint angle.  Branch-free on the common path so that the compiler can interleave it with the
// surrounding spatial algebra (the library sincos() carries a Payne-Hanek slow path behind a branch and cost ~120
// instructions per call in the r1b profile): three-term Cody-Waite reduction by pi/2 (exact products through FMA,
// good to |x| ~ 1e5), then the fdlibm minimax kernels on [-pi/4, pi/4]; error <= ~1.5 ulp.  Larger or non-finite
// arguments take the library path.
RBD_HD void sincos_f64(double x, double* s, double* c) {
  if (!(fabs(x) < 1.0e5)) {
#if defined(__CUDA_ARCH__)
    sincos(x, s, c);
#else
    *s = sin(x);
    *c = cos(x);
#endif
    return;
  }
  const double q = rint(x * 0.6366197723675814);
  double r = fma(-q, 1.5707963267948966, x);
  r = fma(-q, 6.123233995736766e-17, r);
  r = fma(-q, -1.4973849048591698e-33, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double sr = fma(ps * z, r, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double cr = fma(pc * z, z, fma(-0.5, z, 1.0));
  const int n = (int)q;
  const double a = (n & 1) ? cr : sr;   // quadrant: sin takes cos(r) in odd quadrants
  const double b = (n & 1) ? sr : cr;
  *s = (n & 2) ? -a : a;
  *c = ((n + 1) & 2) ? -b : b;
}

RBD_HD void sincos_(real x, real* s, real* c) {
  if constexpr (sizeof(real) == 4) {
    // optional FP32 mode (1e-4 tolerance): the single-precision library routine
#if defined(__CUDA_ARCH__)
    float fs, fc;
    sincosf((float)x, &fs, &fc);
    *s = (real)fs; *c = (real)fc;
#else
    *s = (real)sinf((float)x);
    *c = (real)cosf((float)x);
#endif
  } else {
    sincos_f64((double)x, (double*)s, (double*)c);
  }
}

// O = A * B (3x3 row-major, O distinct from A and B)
RBD_HD void mat3_mul(const real* A, const real* B, real* O) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) O[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  }
}

// rotate columns (ci, cj) of B by (c, s): col_i' = c col_i + s col_j ; col_j' = -s col_i + c col_j
#define RBD_ROTCOLS(R, B, ck, ci, cj, c, s)        \
  do {                                             \
    _Pragma("unroll") for (int r_ = 0; r_ < 3; ++r_) { \
      real bi_ = B[3 * r_ + ci], bj_ = B[3 * r_ + cj]; \
      R[3 * r_ + ck] = B[3 * r_ + ck];             \
      R[3 * r_ + ci] = c * bi_ + s * bj_;          \
      R[3 * r_ + cj] = c * bj_ - s * bi_;          \
    }                                              \
  } while (0)

// the same rotation of columns (ci, cj) in place (column ck, the joint axis, is untouched)
#define RBD_ROTCOLS_IP(R, ci, cj, c, s)             \
  do {                                              \
    _Pragma("unroll") for (int r_ = 0; r_ < 3; ++r_) {  \
      const real bi_ = R[3 * r_ + ci], bj_ = R[3 * r_ + cj]; \
      R[3 * r_ + ci] = c * bi_ + s * bj_;           \
      R[3 * r_ + cj] = c * bj_ - s * bi_;           \
    }                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// One forward-kinematics step: (R,p) of the parent joint in, (R,p) of joint `jn` out (world frame), plus the
// world joint axis aw (1-dof joints).  q0/q1 are the joint's configuration values (angle | displacement |
// cos,sin); the free-flyer reads its 7 values through `qf`.
// pinocchio: liMi = jointPlacement * M_joint(q); oMi = oMi[parent] * liMi (JointJacobiansForwardStep).
// FFK: 0 = joint type tested at run time, 1 = the joint is a free-flyer, 2 = it is not (free-flyer-root instantiation)
template <int FFK = 0, class QF>
RBD_HD void fk_joint(const DevJoint& jn, real* R, real* p, real q0, real q1, const QF& qf, real* aw) {
  // p <- R pl_p + p with the parent's R, then R <- R pl_R IN PLACE: most URDF joints have rpy = 0, and for those the
  // step is nothing at all (no copy of R into a scratch matrix: profile r1h counted 18 register moves per joint for it)
  {
    const real pn0 = R[0] * jn.pl[9] + R[1] * jn.pl[10] + R[2] * jn.pl[11] + p[0];
    const real pn1 = R[3] * jn.pl[9] + R[4] * jn.pl[10] + R[5] * jn.pl[11] + p[1];
    const real pn2 = R[6] * jn.pl[9] + R[7] * jn.pl[10] + R[8] * jn.pl[11] + p[2];
    p[0] = pn0; p[1] = pn1; p[2] = pn2;
  }
  if (!jn.pl_rot_identity) {  // uniform branch
    real B[9];
    mat3_mul(R, jn.pl, B);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = B[k];
  }
  const int t = jn.type;
#if !defined(RBD_CHAIN_ONLY)
  if (FFK != 2 && (FFK == 1 || t == JT_FF)) {
    const int o = jn.idx_q;
    const real px = qf.ld(o), py = qf.ld(o + 1), pz = qf.ld(o + 2);
    const real x = qf.ld(o + 3), y = qf.ld(o + 4), z = qf.ld(o + 5), w = qf.ld(o + 6);
    // Eigen toRotationMatrix on raw coefficients (no normalisation), as pinocchio's JointModelFreeFlyer::calc
    const real tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const real twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    const real tyy = ty * y, tyz = tz * y, tzz = tz * z;
    real Rq[9], B[9];
    Rq[0] = 1 - (tyy + tzz); Rq[1] = txy - twz;       Rq[2] = txz + twy;
    Rq[3] = txy + twz;       Rq[4] = 1 - (txx + tzz); Rq[5] = tyz - twx;
    Rq[6] = txz - twy;       Rq[7] = tyz + twx;       Rq[8] = 1 - (txx + tyy);
    p[0] += R[0] * px + R[1] * py + R[2] * pz;
    p[1] += R[3] * px + R[4] * py + R[5] * pz;
    p[2] += R[6] * px + R[7] * py + R[8] * pz;
    mat3_mul(R, Rq, B);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = B[k];
    aw[0] = aw[1] = aw[2] = 0.0;
    return;
  }
#endif
  real c = q0, s = q1;
#if defined(RBD_SIMPLE_JOINTS)
  // instantiation for models whose joints are all RX / RY / RZ (+ free-flyers) — every URDF robot of the BASELINE
  // configs: the unaligned / prismatic / unbounded cases are not compiled in (A/B r1i: 2.344 -> 2.308 ms; the kernel
  // is instruction-cache sensitive, see RBD_UNROLL_ZERO)
  sincos_(q0, &s, &c);  // (a free-flyer returned above)
#else
  if (t >= JT_RX && t <= JT_RU) sincos_(q0, &s, &c);
#endif
  switch (t) {
    case JT_RX: case JT_RUBX:
      aw[0] = R[0]; aw[1] = R[3]; aw[2] = R[6];
      RBD_ROTCOLS_IP(R, 1, 2, c, s);
      break;
    case JT_RY: case JT_RUBY:
      aw[0] = R[1]; aw[1] = R[4]; aw[2] = R[7];
      RBD_ROTCOLS_IP(R, 2, 0, c, s);
      break;
    case JT_RZ: case JT_RUBZ:
      aw[0] = R[2]; aw[1] = R[5]; aw[2] = R[8];
      RBD_ROTCOLS_IP(R, 0, 1, c, s);
      break;
#if !defined(RBD_SIMPLE_JOINTS)
    case JT_RU: case JT_RUBU: {
      const real ax = jn.axis[0], ay = jn.axis[1], az = jn.axis[2];
      const real c1 = real(1) - c, sx = s * ax, sy = s * ay, sz = s * az;
      real Rj[9], B[9];
      Rj[0] = c1 * ax * ax + c;  Rj[1] = c1 * ax * ay - sz; Rj[2] = c1 * ax * az + sy;
      Rj[3] = c1 * ay * ax + sz; Rj[4] = c1 * ay * ay + c;  Rj[5] = c1 * ay * az - sx;
      Rj[6] = c1 * az * ax - sy; Rj[7] = c1 * az * ay + sx; Rj[8] = c1 * az * az + c;
      aw[0] = R[0] * ax + R[1] * ay + R[2] * az;
      aw[1] = R[3] * ax + R[4] * ay + R[5] * az;
      aw[2] = R[6] * ax + R[7] * ay + R[8] * az;
      mat3_mul(R, Rj, B);
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = B[k];
    } break;
    case JT_PX: case JT_PY: case JT_PZ: case JT_PU: {
      const real ax = jn.axis[0], ay = jn.axis[1], az = jn.axis[2];
      aw[0] = R[0] * ax + R[1] * ay + R[2] * az;
      aw[1] = R[3] * ax + R[4] * ay + R[5] * az;
      aw[2] = R[6] * ax + R[7] * ay + R[8] * az;
      p[0] += q0 * aw[0]; p[1] += q0 * aw[1]; p[2] += q0 * aw[2];
    } break;
#endif
    default:
      aw[0] = aw[1] = aw[2] = 0.0;
      break;
  }
}

#if defined(RBD_SIMPLE_JOINTS)
RBD_HD bool is_prismatic(int) { return false; }
RBD_HD bool is_unbounded(int) { return false; }
#else
RBD_HD bool is_prismatic(int t) { return t >= JT_PX && t <= JT_PU; }
RBD_HD bool is_unbounded(int t) { return t >= JT_RUBX && t <= JT_RUBU; }
#endif

// world-frame motion subspace column of a 1-dof joint: revolute [p x a ; a], prismatic [a ; 0]
RBD_HD void subspace_1dof(int type, const real* p, const real* aw, real* A) {
  // revolute first, unconditionally; the (rare, uniform) prismatic case overwrites: no merge of two register sets
  RBD_CROSS(A[0], A[1], A[2], p[0], p[1], p[2], aw[0], aw[1], aw[2]);
  A[3] = aw[0]; A[4] = aw[1]; A[5] = aw[2];
  if (is_prismatic(type)) {
    A[0] = aw[0]; A[1] = aw[1]; A[2] = aw[2];
    A[3] = A[4] = A[5] = 0.0;
  }
}
// column e (0..5) of the free-flyer subspace oMi.act(I6): e<3 -> [R e ; 0], e>=3 -> [p x R e ; R e]
#define RBD_FF_COL(A, R, p, e)                                                        \
  do {                                                                                \
    const real cx_ = R[(e) % 3], cy_ = R[3 + (e) % 3], cz_ = R[6 + (e) % 3];        \
    if ((e) < 3) {                                                                    \
      A[0] = cx_; A[1] = cy_; A[2] = cz_; A[3] = A[4] = A[5] = 0.0;                   \
    } else {                                                                          \
      RBD_CROSS(A[0], A[1], A[2], p[0], p[1], p[2], cx_, cy_, cz_);                   \
      A[3] = cx_; A[4] = cy_; A[5] = cz_;                                             \
    }                                                                                 \
  } while (0)

// free-flyer rows of (column F): A_ff^T F = [R^T Fl ; R^T (Fw - p x Fl)]
RBD_HD void ff_rows(const real* R, const real* p, const real* F, real* out6) {
  real tx, ty, tz;
  RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], F[0], F[1], F[2]);
  tx = F[3] - tx; ty = F[4] - ty; tz = F[5] - tz;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    out6[e] = R[e] * F[0] + R[3 + e] * F[1] + R[6 + e] * F[2];
    out6[3 + e] = R[e] * tx + R[3 + e] * ty + R[6 + e] * tz;
  }
}

// free-flyer joint motion given in the joint (body) frame, 6 fields of `fld` from `o`: world = oMi.act(.)
// (the six values are read with the L2 "keep" hint: the fused eval re-reads them when the sweep returns to the joint)
template <class FLD>
RBD_HD void ff_act(const real* R, const real* p, const FLD& fld, int o, real* out, bool may_be_absent = false) {
  if (may_be_absent && !fld.ok()) {  // absent input array (Mass / ForceField products): zero motion
#pragma unroll
    for (int e = 0; e < 6; ++e) out[e] = real(0);
    return;
  }
  const real l0 = fld.ld_keep(o), l1 = fld.ld_keep(o + 1), l2 = fld.ld_keep(o + 2);
  const real w0 = fld.ld_keep(o + 3), w1 = fld.ld_keep(o + 4), w2 = fld.ld_keep(o + 5);
  out[3] = R[0] * w0 + R[1] * w1 + R[2] * w2;
  out[4] = R[3] * w0 + R[4] * w1 + R[5] * w2;
  out[5] = R[6] * w0 + R[7] * w1 + R[8] * w2;
  real tx, ty, tz;
  RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], out[3], out[4], out[5]);
  out[0] = R[0] * l0 + R[1] * l1 + R[2] * l2 + tx;
  out[1] = R[3] * l0 + R[4] * l1 + R[5] * l2 + ty;
  out[2] = R[6] * l0 + R[7] * l1 + R[8] * l2 + tz;
}

// world spatial inertia about the world origin: Y = (m, h = m c_w, I_O (xx,xy,yy,xz,yz,zz))
RBD_HD void inertia_world(const DevJoint& jn, const real* R, const real* p, real* Y) {
  const real m = jn.Y[0];
  const real cx = R[0] * jn.Y[1] + R[1] * jn.Y[2] + R[2] * jn.Y[3] + p[0];
  const real cy = R[3] * jn.Y[1] + R[4] * jn.Y[2] + R[5] * jn.Y[3] + p[1];
  const real cz = R[6] * jn.Y[1] + R[7] * jn.Y[2] + R[8] * jn.Y[3] + p[2];
  // T = R * Ic
  const real ixx = jn.Y[4], ixy = jn.Y[5], iyy = jn.Y[6], ixz = jn.Y[7], iyz = jn.Y[8], izz = jn.Y[9];
  real T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T[3 * i + 0] = R[3 * i] * ixx + R[3 * i + 1] * ixy + R[3 * i + 2] * ixz;
    T[3 * i + 1] = R[3 * i] * ixy + R[3 * i + 1] * iyy + R[3 * i + 2] * iyz;
    T[3 * i + 2] = R[3 * i] * ixz + R[3 * i + 1] * iyz + R[3 * i + 2] * izz;
  }
  const real mcx = m * cx, mcy = m * cy, mcz = m * cz;
  Y[0] = m; Y[1] = mcx; Y[2] = mcy; Y[3] = mcz;
  // I_O = R Ic R^T + m (|c|^2 I - c c^T)
  Y[4] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + (mcy * cy + mcz * cz);      // xx
  Y[5] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - mcx * cy;                   // xy
  Y[6] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + (mcx * cx + mcz * cz);      // yy
  Y[7] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - mcx * cz;                   // xz
  Y[8] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - mcy * cz;                   // yz
  Y[9] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + (mcx * cx + mcy * cy);      // zz
}

// F = Y * motion (both about the world origin): lin = m v + w x h ; ang = I_O w + h x v
RBD_HD void inertia_mul(const real* Y, const real* mo, real* F) {
  real ax, ay, az, bx, by, bz;
  RBD_CROSS(ax, ay, az, mo[3], mo[4], mo[5], Y[1], Y[2], Y[3]);
  RBD_CROSS(bx, by, bz, Y[1], Y[2], Y[3], mo[0], mo[1], mo[2]);
  F[0] = Y[0] * mo[0] + ax;
  F[1] = Y[0] * mo[1] + ay;
  F[2] = Y[0] * mo[2] + az;
  F[3] = Y[4] * mo[3] + Y[5] * mo[4] + Y[7] * mo[5] + bx;
  F[4] = Y[5] * mo[3] + Y[6] * mo[4] + Y[8] * mo[5] + by;
  F[5] = Y[7] * mo[3] + Y[8] * mo[4] + Y[9] * mo[5] + bz;
}

RBD_HD real dot6(const real* a, const real* b) {
  return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) + (a[3] * b[3] + a[4] * b[4] + a[5] * b[5]);
}

// Eigen 3.4 Quaternion(Matrix3) branch order (reference Conversions.h:60), output [x y z w] (Conversions.h:36).
// Eigen picks one of four formulas (trace > 0, else the largest diagonal entry with its tie order); which one is
// data dependent, so neighbouring instances of a warp diverge.  Written with selects instead: the case index is
// computed first, then ONE square root and ONE division serve whichever case was picked, and every value is the very
// expression Eigen evaluates in that case (same operations, same order: bit-identical to the branched form, which the
// profile r1h of the apply kernel showed as 3 600 instructions per tile of mostly register moves behind divergent
// branches).
RBD_HD void rot_to_quat(const real* R, real* q) {
  const real tr = R[0] + R[4] + R[8];
  const bool cw = tr > real(0);
  const bool c0 = !cw && !(R[4] > R[0]) && !(R[8] > R[0]);  // i = 0
  const bool c1 = !cw && !c0 && (R[4] > R[0]) && !(R[8] > R[4]);  // i = 1; otherwise i = 2
  // argument of the square root in the four cases
  const real aw_ = tr + real(1);
  const real a0_ = R[0] - R[4] - R[8] + real(1);
  const real a1_ = R[4] - R[8] - R[0] + real(1);
  const real a2_ = R[8] - R[0] - R[4] + real(1);
  const real arg = cw ? aw_ : (c0 ? a0_ : (c1 ? a1_ : a2_));
  const real t = sqrt(arg);
  const real big = real(0.5) * t;
  const real f = real(0.5) / t;
  // antisymmetric and symmetric off-diagonal combinations
  const real d0 = R[7] - R[5], d1 = R[2] - R[6], d2 = R[3] - R[1];
  const real s0 = R[7] + R[5], s1 = R[6] + R[2], s2 = R[3] + R[1];  // (R21+R12), (R20+R02), (R10+R01)
  // case w: x = d0 f, y = d1 f, z = d2 f, w = big
  // case 0: x = big,  y = s2 f, z = s1 f, w = d0 f
  // case 1: x = (R[1]+R[3]) f, y = big, z = s0 f, w = d1 f
  // case 2: x = (R[2]+R[6]) f, y = (R[5]+R[7]) f, z = big, w = d2 f
  const real s2b = R[1] + R[3], s1b = R[2] + R[6], s0b = R[5] + R[7];  // operand order of Eigen's cases 1 and 2
  const real nx = cw ? d0 : (c1 ? s2b : s1b);
  const real ny = cw ? d1 : (c0 ? s2 : s0b);
  const real nz = cw ? d2 : (c0 ? s1 : s0);
  const real nw = c0 ? d0 : (c1 ? d1 : d2);
  const real xv = nx * f, yv = ny * f, zv = nz * f, wv = nw * f;
  q[0] = c0 ? big : xv;
  q[1] = c1 ? big : yv;
  q[2] = (!cw && !c0 && !c1) ? big : zv;
  q[3] = cw ? big : wv;
}

// =========================================================================================================
// Fused evaluation: FK + world joint Jacobian (+ CRBA) (+ RNEA)                      [kernel K7 / K5 / K6]
// =========================================================================================================
template <int TILE>
struct EvalIOT {
  FieldT<TILE> q, v, a;          // inputs: nq, nv, nv fields
  FieldT<TILE> oMi, J, M, tau;   // outputs (p == nullptr -> not written)
  real* Mtile = nullptr;         // field 0, lane 0 of this warp's tile of M (tile-32, 16-byte aligned M), else nullptr
  long long next = 0;            // tiles from this tile to the next one the same warp will process (0 = none): the
                                 // asynchronous input ring runs on into that tile
  // Mass / ForceField products (rbd_addMDx_soa, rbd_addForce_soa): v or a may be absent (zeros), gravity may be
  // switched off, and tau may be accumulated, tau <- tau + tau_scale * rnea(...), as SOFA's addMDx / addForce do
  real tau_scale = real(1);
  int tau_acc = 0;
  int gravity_on = 1;
  int tau_live = 1;  // 0 on the lanes past the end of the batch (they recompute the last instance)
  // MF: the instantiation serves the Mass / ForceField products (RNEA without CRBA); the fused kernel compiles the
  // plain store (the run-time switches cost it 3 % when they were unconditional)
  template <bool MF>
  RBD_HD void put_tau(int k, real v) const {
    if (MF && tau_acc) {
      // a reduction that returns nothing (RED.ADD.F64: one IEEE add per address and launch = load + add + store),
      // so the sweep does not wait for the old value (r1i: addMDx 1.87 ms with load + add + store)
#if defined(__CUDA_ARCH__)
      if (tau_live) atomicAdd(tau.at(k), tau_scale * v);  // lanes past the end duplicate the last instance: no add
#else
      tau.st(k, tau.ld(k) + tau_scale * v);
#endif
    } else {
      tau.st(k, v);
    }
  }
};
// Asynchronous input ring of the fused eval (device, tile-32 batches; the plan says whether it fits): the q | v | a
// values of the 1-dof joints travel global -> shared memory with cp.async (no destination register, so nothing
// waits on them until they are read) RBD_INR_SLOTS joints ahead of the sweep, running on into the warp's next tile.
// Why: under the kernel's own write stream (5.3 TB/s) a global read takes 2 500 cycles on average and 6 000 at the
// 99th percentile (experiments/load_under_stores.cu) — more than a whole joint's work, which is what a
// one-joint-ahead register prefetch gives: its consumer held 10 % of all stall samples (profile r1h).  A register
// ring cannot go further ahead without unrolling the joint loop (shifting the ring reads the in-flight registers).
#define RBD_INR_SLOTS 2
struct InRing {
  real* p = nullptr;  // this thread's column of its warp's [RBD_INR_SLOTS * 3][32] region; nullptr = register prefetch
  unsigned ps = 0;    // the same address in the shared window (device)
  int slot = 0;       // slot of the joint the sweep is about to visit
  int primed = 0;     // the previous tile already started the copies for this tile's first joints
};
// Input ring of the mapping kernels (apply, applyJ): NV values per joint (q | q, dq) travel global -> shared memory
// with cp.async RBD_MAP_SLOTS joints ahead of the sweep.  p: this thread's column of its warp's
// [RBD_MAP_SLOTS * NV][32] region, nullptr (host, tests/emul) = one joint ahead through registers / direct loads.
#define RBD_MAP_SLOTS 4
struct MapRing {
  real* p = nullptr;
  unsigned ps = 0;  // the same address in the shared window (device)
};
RBD_HD void cp_async_map(const MapRing& r, int e, const real* src) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(r.ps + (unsigned)e * (unsigned)(RBD_WARP * 8)), "l"(src) : "memory");
#else
  r.p[e * RBD_WARP] = *src;
#endif
}
// element `e` of the ring <- *src, asynchronously
RBD_HD void cp_async_real(const InRing& r, int e, const real* src) {
#if defined(__CUDA_ARCH__)
  const unsigned d = r.ps + (unsigned)e * (unsigned)(RBD_WARP * sizeof(real));
  if constexpr (sizeof(real) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
#else
  r.p[e * RBD_WARP] = *src;
#endif
}
RBD_HD void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
RBD_HD void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
using EvalIO = EvalIOT<0>;

// ---- sweep-state storage (EvalPlan): vector accessors over the two spaces ----------------------------------------
template <int N>
RBD_HD void s_ldv(const Stack& S, int k, real* v) {
#pragma unroll
  for (int e = 0; e < N; ++e) v[e] = S.ld(k + e);
}
template <int N>
RBD_HD void s_stv(const Stack& S, int k, const real* v) {
#pragma unroll
  for (int e = 0; e < N; ++e) S.st(k + e, v[e]);
}
// load / store / accumulate N doubles at offset k of shared (tm == false) or tensor memory (tm == true)
template <int N>
RBD_HD void x_ldv(bool tm, const Stack& S, const TmStack& T, int k, real* v) {
  if (tm) T.ldv<N>(k, v); else s_ldv<N>(S, k, v);
}
template <int N>
RBD_HD void x_stv(bool tm, const Stack& S, const TmStack& T, int k, const real* v) {
  if (tm) T.stv<N>(k, v); else s_stv<N>(S, k, v);
}
// v[0..N) += the N values at offset k: each space adds into v inside its own branch, so that the two load paths do not
// merge through a temporary (the merge cost ~2 register moves per value: 7 % of the eval kernel's instructions, r1i)
template <int N>
RBD_HD void x_accv(bool tm, const Stack& S, const TmStack& T, int k, real* v) {
  if (tm) {
    real t[N];
    T.template ldv<N>(k, t);
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] += t[e];
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) v[e] += S.ld(k + e);
  }
}
template <int N>
RBD_HD void x_addv(bool tm, const Stack& S, const TmStack& T, int k, const real* v) {
  if (tm) {
    real t[N];
    T.ldv<N>(k, t);
#pragma unroll
    for (int e = 0; e < N; ++e) t[e] += v[e];
    T.stv<N>(k, t);
  } else {
#pragma unroll
    for (int e = 0; e < N; ++e) S.add(k + e, v[e]);
  }
}

// ColProg table of the staged eval blob (rbd_model_build.h: build_eval_blob): `m` IS the start of the blob.
RBD_HD const int* col_prog(const DevModel& m) {
  return reinterpret_cast<const int*>(reinterpret_cast<const unsigned long long*>(&m) + m.prog_words_off);
}

// Rows of packed column `base` above the joint that owns it, driven by the joint's column program: one entry per
// ancestor dof (A_anc . F) — independent iterations, so the loads of several ancestors are in flight together —
// then the zero segments (dofs of joints off the ancestor chain).
template <int TILE>
RBD_HD void column_rows(const DevModel& m, const DevJoint& jn, int base, const real* F, const Stack& S,
                        const FieldT<TILE>& M, real* Mtile) {
  const int* pr = col_prog(m) + jn.prog_off;
  const int na = jn.depth - 1;
#if RBD_PTR_COL
  real* const pc = M.at(base);
#pragma unroll 2
  for (int d = 0; d < na; ++d) {
    const int e = pr[d];
    const int sa = RBD_PROG_OFFA(e);
    real* const pa = pc + M.step(RBD_PROG_ROW(e));
    if (RBD_PROG_FF(e)) {
      real Rq[12], p6[6];
      s_ldv<12>(S, sa, Rq);
      ff_rows(Rq, Rq + 9, F, p6);
#pragma unroll
      for (int k = 0; k < 6; ++k) M.st_at(pa, k, p6[k]);
    } else {
      real Aa[6];
      s_ldv<6>(S, sa, Aa);
      *pa = dot6(Aa, F);
    }
  }
  for (int z = 0; z < jn.nzseg; ++z) {
    const int e = pr[na + z];
    real* pz = pc + M.step(RBD_PROG_ROW(e));
    int len = RBD_PROG_ZLEN(e);
    for (; len >= 4; len -= 4, pz += M.step(4)) {
      M.st_at(pz, 0, 0.0); M.st_at(pz, 1, 0.0); M.st_at(pz, 2, 0.0); M.st_at(pz, 3, 0.0);
    }
    for (; len > 0; --len, pz += M.step(1)) *pz = 0.0;
  }
#else
  RBD_PRAGMA(unroll RBD_UNROLL_ANC)
  for (int d = 0; d < na; ++d) {
    const int e = pr[d];
    const int sa = RBD_PROG_OFFA(e);
    if (RBD_PROG_FF(e)) {
      real Rq[12], p6[6];
      s_ldv<12>(S, sa, Rq);
      ff_rows(Rq, Rq + 9, F, p6);
#pragma unroll
      for (int k = 0; k < 6; ++k) M.st(base + RBD_PROG_ROW(e) + k, p6[k]);
    } else {
      real Aa[6];
      s_ldv<6>(S, sa, Aa);
      M.st(base + RBD_PROG_ROW(e), dot6(Aa, F));
    }
  }
#if defined(__CUDA_ARCH__) && RBD_COOP_ZERO
  if constexpr (TILE == 32 && sizeof(real) == 8) {
    // Structural zeros do not depend on the instance, so the warp fills a zero segment (len consecutive 256-byte
    // fields of its tile) cooperatively with 16-byte stores: one instruction covers two fields instead of one
    // (r1h: 372 zero stores + 650 loop instructions per tile before).
    if (Mtile != nullptr) {
      const int lane = threadIdx.x & 31;
      for (int z = 0; z < jn.nzseg; ++z) {
        const int e = pr[na + z];
        const int r0 = base + RBD_PROG_ROW(e), len = RBD_PROG_ZLEN(e);
        double2* const pz = reinterpret_cast<double2*>(Mtile + (long long)r0 * 32) + lane;
        const int pairs = len >> 1;
        for (int i = 0; i < pairs; ++i) pz[i * 32] = make_double2(0.0, 0.0);
        if ((len & 1) && lane < 16) pz[pairs * 32] = make_double2(0.0, 0.0);
      }
      return;
    }
  }
#endif
  for (int z = 0; z < jn.nzseg; ++z) {
    const int e = pr[na + z];
    const int r0 = base + RBD_PROG_ROW(e), len = RBD_PROG_ZLEN(e);
    RBD_PRAGMA(unroll RBD_UNROLL_ZERO)
    for (int r = 0; r < len; ++r) M.st(r0 + r, 0.0);
  }
#endif
}

// Everything joint k owes once its subtree is complete: tau_k = A_k^T f_k and its column(s) of M.
// A = world subspace column (6) of a 1-dof joint, or R|p (12) of a free-flyer; f = total subtree force;
// Y = composite inertia of the subtree — all three in registers.
// FFK: 0 = the joint type is tested at run time, 1 = joint k is a free-flyer, 2 = it is not (compile-time knowledge of
// the free-flyer-root instantiation).
template <bool DO_M, bool DO_TAU, int TILE, int FFK = 0>
RBD_HD void emit_joint(const DevModel& m, int k, const real* A, const real* f, const real* Y, const Stack& S,
                       const EvalIOT<TILE>& io) {
  const DevJoint& jn = m.joint[k];
  const bool is_ff = FFK == 1 ? true : (FFK == 2 ? false : RBD_T_IS_FF(jn.type));
  if (!is_ff) {
    if (DO_TAU) io.template put_tau<(DO_TAU && !DO_M)>(jn.idx_v, dot6(A, f));
    if (DO_M) {
      real F[6];
      inertia_mul(Y, A, F);
      const int base = RBD_MBASE(jn);  // packed: col (col + 1) / 2, own row at + col; support-only: see rbd_dev_model.h
      *io.M.at(base + RBD_MOWN(jn)) = dot6(A, F);
      column_rows(m, jn, base, F, S, io.M, io.Mtile);
    }
  } else {
    // a real loop (not unrolled): runs once per free-flyer, and the kernel is instruction-cache sensitive
    RBD_PRAGMA(unroll RBD_UNROLL_FF)
    for (int cc = 0; cc < 6; ++cc) {
      real Ac[6];
      RBD_FF_COL(Ac, A, (A + 9), cc);
      if (DO_TAU) io.template put_tau<(DO_TAU && !DO_M)>(jn.idx_v + cc, dot6(Ac, f));
      if (DO_M) {
        real F[6], o6[6];
        inertia_mul(Y, Ac, F);
        const int base = RBD_MBASE(jn) + cc * RBD_MOWN(jn) + cc * (cc + 1) / 2;
        ff_rows(A, A + 9, F, o6);
        real* const pd = io.M.at(base + RBD_MOWN(jn));
#pragma unroll
        for (int e = 0; e < 6; ++e)
          if (e <= cc) io.M.st_at(pd, e, o6[e]);
        column_rows(m, jn, base, F, S, io.M, io.Mtile);
      }
    }
  }
}

// One instance, one sweep.  S = this thread's shared-memory stack, T = its tensor-memory slice (m.plan says what
// lives where).  Control flow depends on the model only, so all lanes of a warp stay converged.
template <bool DO_M, bool DO_TAU, int TILE>
RBD_HD void eval_instance(const DevModel& m, const EvalIOT<TILE>& io, const Stack& S, const TmStack& T, InRing& inr) {
  constexpr bool NEED_STACK = DO_M || DO_TAU;
  const int nj = m.njoints;
  const int fbase = m.plan.fbase, fsplit = m.plan.fsplit, ftbase = m.plan.ftbase, ybase = m.plan.ybase;
  const int ysplit = m.plan.ysplit, ytbase = m.plan.ytbase;
  // Y- and f-stack level -> (space, offset): levels below the split in shared memory, the rest in tensor memory
#define RBD_Y_TM(lvl) ((lvl) >= ysplit)
#define RBD_Y_OFF(lvl) ((lvl) >= ysplit ? ytbase + 9 * ((lvl)-ysplit) : ybase + 9 * (lvl))
#define RBD_F_TM(lvl) ((lvl) >= fsplit)
#define RBD_F_OFF(lvl) ((lvl) >= fsplit ? ftbase + 6 * ((lvl)-fsplit) : fbase + 6 * (lvl))
  real R[9], p[3], vel[6], acc[6];
  constexpr bool MF = DO_TAU && !DO_M;  // Mass / ForceField switches live in the RNEA-only instantiation
  // The single-algorithm instantiations (RNEA only, CRBA only) never write oMi / J (the C API serves "tau or M alone + oMi /
  // J" with two launches): their stores and addresses are compiled out instead of predicated off — at 168 registers that
  // is worth 4-5 % (A/B s4j: Talos 2^20 rnea 1.056 -> 1.009 ms, addMDx 0.927 -> 0.879 ms)
#ifndef RBD_MF_NO_OMIJ
#define RBD_MF_NO_OMIJ 1
#endif
#define RBD_OMIJ(F) ((!RBD_MF_NO_OMIJ || (DO_M == DO_TAU)) && (F).ok())
  const real gs_ = (MF && !io.gravity_on) ? real(0) : real(1);
  const bool has_v = !MF || io.v.ok();  // (uniform) the Mass / ForceField products without a velocity skip its terms
#define RBD_RESET_ROOT()                                                   \
  do {                                                                     \
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1; \
    p[0] = p[1] = p[2] = 0;                                                \
    _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) { vel[e_] = 0; acc[e_] = 0; } \
    acc[0] = -gs_ * m.gravity[0]; acc[1] = -gs_ * m.gravity[1]; acc[2] = -gs_ * m.gravity[2]; \
  } while (0)
  RBD_RESET_ROOT();

  // Prefetch of the 1-dof inputs: through the asynchronous ring (see InRing) when the plan has room for it, else one
  // joint ahead through registers.
  real qn = 0, q1n = 0, vn = 0, an = 0;
#define RBD_IN_FETCH(jidx)                                                                         \
  do {                                                                                             \
    if ((jidx) < nj && !RBD_T_IS_FF(m.joint[jidx].type)) {                                              \
      const DevJoint& jx_ = m.joint[jidx];                                                         \
      qn = io.q.ld(jx_.idx_q);                                                                     \
      if (is_unbounded(jx_.type)) q1n = io.q.ld(jx_.idx_q + 1);                                    \
      if (DO_TAU) {                                                                                \
        vn = MF ? io.v.ldz(jx_.idx_v) : io.v.ld(jx_.idx_v);                                        \
        an = MF ? io.a.ldz(jx_.idx_v) : io.a.ld(jx_.idx_v);                                        \
      }                                                                                            \
    }                                                                                              \
  } while (0)
  // start the copies of joint `jidx` (`ahead` tiles further on) into ring slot `slot_`; always closes one group
#define RBD_INR_ISSUE(jidx, ahead, slot_)                                                          \
  do {                                                                                             \
    if ((jidx) < nj && !RBD_T_IS_FF(m.joint[jidx].type)) {                                              \
      const DevJoint& jx_ = m.joint[jidx];                                                         \
      const int oq_ = (jx_.idx_q + (ahead) * m.nq) * TILE, ov_ = (jx_.idx_v + (ahead) * m.nv) * TILE; \
      cp_async_real(inr, 3 * (slot_), io.q.at_elems(oq_));                                         \
      if (DO_TAU) {                                                                                \
        if (!MF || io.v.ok()) cp_async_real(inr, 3 * (slot_) + 1, io.v.at_elems(ov_));             \
        if (!MF || io.a.ok()) cp_async_real(inr, 3 * (slot_) + 2, io.a.at_elems(ov_));             \
      }                                                                                            \
    }                                                                                              \
    cp_async_commit();                                                                             \
  } while (0)
  // (FP64 only: the FP32 mode is capped at 128 registers for two blocks per SM and has no room for the ring's code)
  const bool ring = sizeof(real) == 8 && inr.p != nullptr;
  if (!ring) {
    RBD_IN_FETCH(1);
  } else if (!inr.primed) {
#pragma unroll
    for (int s_ = 0; s_ < RBD_INR_SLOTS; ++s_) {
      const int sl_ = inr.slot + s_ >= RBD_INR_SLOTS ? inr.slot + s_ - RBD_INR_SLOTS : inr.slot + s_;
      RBD_INR_ISSUE(1 + s_, 0, sl_);
    }
  }

  // State of the joint the sweep returns to when a new branch starts below it: R|p (+ v|a).  With a stack
  // (CRBA / RNEA) this runs at the LEAF that ends the previous branch, BEFORE its unwind — R, p, vel, acc are dead
  // during the unwind, so the loads (tensor memory, the oMi re-read, the input re-reads) overlap it instead of
  // stalling the next joint (r1f: the return block was the top stall site, 7.7 % of samples).
  auto restore = [&](int par) {
    if (par == 0) {
      RBD_RESET_ROOT();
    } else {
      const DevJoint& jb = m.joint[par];
      const bool btm = (jb.boff & RBD_BOFF_TMEM) != 0;
      int b = jb.boff & RBD_BOFF_MASK;
      // every source assigns R | p (v | a) inside its own branch: no merge through a temporary
      if (NEED_STACK && RBD_T_IS_FF(jb.type)) {
        // a free-flyer's A record is its R|p
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S.ld(jb.offA + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S.ld(jb.offA + 9 + e);
      } else if (jb.boff & RBD_BOFF_RPOMI) {
        const int o = 12 * (par - 1);  // this thread wrote oMi[par] earlier in the sweep
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = io.oMi.ld(o + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = io.oMi.ld(o + 9 + e);
      } else {
        if (btm) {
          real Rp[12];
          T.template ldv<12>(b, Rp);
#pragma unroll
          for (int e = 0; e < 9; ++e) R[e] = Rp[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) p[e] = Rp[9 + e];
        } else {
#pragma unroll
          for (int e = 0; e < 9; ++e) R[e] = S.ld(b + e);
#pragma unroll
          for (int e = 0; e < 3; ++e) p[e] = S.ld(b + 9 + e);
        }
        b += 12;
      }
      if (DO_TAU) {
        if (!(jb.boff & RBD_BOFF_NOVA)) {
          if (btm) {
            real va[12];
            T.template ldv<12>(b, va);
#pragma unroll
            for (int e = 0; e < 6; ++e) { vel[e] = va[e]; acc[e] = va[6 + e]; }
          } else {
#pragma unroll
            for (int e = 0; e < 6; ++e) { vel[e] = S.ld(b + e); acc[e] = S.ld(b + 6 + e); }
          }
        } else {
          // branch joint attached to the universe: v = A qd, a = -g + A qdd (v x vJ = 0), nothing was saved
          real vj[6], aj[6];
          if (RBD_T_IS_FF(jb.type)) {
            ff_act(R, p, io.v, jb.idx_v, vj, MF);
            ff_act(R, p, io.a, jb.idx_v, aj, MF);
          } else {
            real Ab[6];
            s_ldv<6>(S, jb.offA, Ab);
            const real qd_b = MF ? io.v.ldz(jb.idx_v) : io.v.ld(jb.idx_v);
            const real qdd_b = MF ? io.a.ldz(jb.idx_v) : io.a.ld(jb.idx_v);
#pragma unroll
            for (int e = 0; e < 6; ++e) { vj[e] = Ab[e] * qd_b; aj[e] = Ab[e] * qdd_b; }
          }
#pragma unroll
          for (int e = 0; e < 6; ++e) { vel[e] = vj[e]; acc[e] = aj[e]; }
          acc[0] -= gs_ * m.gravity[0]; acc[1] -= gs_ * m.gravity[1]; acc[2] -= gs_ * m.gravity[2];
        }
      }
    }
  };

  // Ak: world subspace column (1-dof) or R|p (free-flyer); fk / Yk: this joint's force / inertia, then the running
  // subtree sums of the unwind.  Declared outside the joint loop: the free-flyer-root instantiation carries fk / Yk of
  // the root's last subtree out of the loop into its epilogue.
#if defined(RBD_FF_ROOT_ONLY)
  real Ak[12], fk[6], Yk[10];
#define RBD_STEP_STATE
#else
#define RBD_STEP_STATE real Ak[12], fk[6], Yk[10];
#endif
  // One joint of the sweep.  `ffj` is a compile-time tag in the free-flyer-root instantiation (RBD_FF_ROOT_ONLY: joint 1
  // is visited by joint_step(FfYes) in front of the loop, every other joint by joint_step(FfNo), so the loop body
  // carries no free-flyer code and fits 12-warp blocks); elsewhere the joint type is tested at run time.
  auto joint_step = [&](auto ffj, const int i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (RBD_CAN_BRANCH && (!NEED_STACK || !RBD_EARLY_RESTORE_NS) && par != i - 1) restore(par);  // FK-only variant: no unwind to hide behind
    real q0 = qn, q1 = q1n, qd = vn, qdd = an;
    const int nxt_par = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    if (!ring) {
      RBD_IN_FETCH(i + 1);
    } else {
      cp_async_wait<RBD_INR_SLOTS - 1>();  // everything but the newest groups has landed: this joint's values are in
      if (!RBD_CUR_FF(ffj, jn.type)) {
        const real* const s_ = inr.p + 3 * inr.slot * RBD_WARP;
        q0 = s_[0];
        if (is_unbounded(jn.type)) q1 = io.q.ld(jn.idx_q + 1);
        if (DO_TAU) {
          qd = (!MF || io.v.ok()) ? s_[RBD_WARP] : real(0);
          qdd = (!MF || io.a.ok()) ? s_[2 * RBD_WARP] : real(0);
        }
      }
      // the slot just read receives the joint RBD_INR_SLOTS positions further on (possibly in the warp's next tile)
      int pj = i + RBD_INR_SLOTS;
      int ahead = 0;
      if (pj >= nj) { pj -= nj - 1; ahead = (int)io.next; if (ahead == 0) pj = nj; }
      RBD_INR_ISSUE(pj, ahead, inr.slot);
      inr.slot = inr.slot + 1 == RBD_INR_SLOTS ? 0 : inr.slot + 1;
    }

    real aw[3];
    fk_joint<RBD_EMIT_CUR(ffj)>(jn, R, p, q0, q1, io.q, aw);
    if (RBD_OMIJ(io.oMi)) {
#if RBD_PTR_OMIJ
      real* const po = io.oMi.at(12 * (i - 1));
      if (jn.boff & RBD_BOFF_RPOMI) {  // this record is re-read when the sweep returns to joint i: keep it in L2
#pragma unroll
        for (int e = 0; e < 9; ++e) io.oMi.st_at_keep(po, e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) io.oMi.st_at_keep(po, 9 + e, p[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 9; ++e) io.oMi.st_at(po, e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) io.oMi.st_at(po, 9 + e, p[e]);
      }
#else
      const int o = 12 * (i - 1);
#pragma unroll
      for (int e = 0; e < 9; ++e) io.oMi.st(o + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) io.oMi.st(o + 9 + e, p[e]);
#endif
    }
    // Ak: world subspace column (1-dof) or R|p (free-flyer) — what the A-stack keeps for the CRBA/RNEA projections
    RBD_STEP_STATE
    if (!RBD_CUR_FF(ffj, jn.type)) {
      subspace_1dof(jn.type, p, aw, Ak);
      if (RBD_OMIJ(io.J)) {
#if RBD_PTR_OMIJ
        real* const pj = io.J.at(6 * jn.idx_v);
#pragma unroll
        for (int e = 0; e < 6; ++e) io.J.st_at(pj, e, Ak[e]);
#else
#pragma unroll
        for (int e = 0; e < 6; ++e) io.J.st(6 * jn.idx_v + e, Ak[e]);
#endif
      }
      if (DO_TAU) {
        if (has_v) {
          real vj[6], cr[6];
#pragma unroll
          for (int e = 0; e < 6; ++e) { vj[e] = Ak[e] * qd; vel[e] += vj[e]; }
          // a_i = a_parent + A qdd + v_i x vJ   (joint bias c = 0 for all supported joints)
          RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
          real tx, ty, tz;
          RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
          cr[0] += tx; cr[1] += ty; cr[2] += tz;
          RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
#pragma unroll
          for (int e = 0; e < 6; ++e) acc[e] += Ak[e] * qdd + cr[e];
        } else {
          // no velocity (M dx, gravity forces): the velocity-product terms vanish identically
#pragma unroll
          for (int e = 0; e < 6; ++e) acc[e] += Ak[e] * qdd;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) Ak[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) Ak[9 + e] = p[e];
      if (RBD_OMIJ(io.J)) {
        real* const pj = io.J.at(6 * jn.idx_v);
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) {
          real A[6];
          RBD_FF_COL(A, R, p, cc);
#pragma unroll
          for (int e = 0; e < 6; ++e) io.J.st_at(pj, 6 * cc + e, A[e]);
        }
      }
      if (DO_TAU) {
        // joint velocity / acceleration given in the joint (body) frame: world = oMi.act(.)
        real vj[6], aj[6], cr[6];
        ff_act(R, p, io.v, jn.idx_v, vj, MF);
        ff_act(R, p, io.a, jn.idx_v, aj, MF);
#pragma unroll
        for (int e = 0; e < 6; ++e) vel[e] += vj[e];
        RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
        real tx, ty, tz;
        RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
        cr[0] += tx; cr[1] += ty; cr[2] += tz;
        RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
#pragma unroll
        for (int e = 0; e < 6; ++e) acc[e] += aj[e] + cr[e];
      }
    }
    if (NEED_STACK) {
      inertia_world(jn, R, p, Yk);
      if (DO_TAU) {
        // f = Y a + v x* (Y v)
        inertia_mul(Yk, acc, fk);
        if (has_v) {
          real h[6];
          inertia_mul(Yk, vel, h);
          real tx, ty, tz;
          RBD_CROSS(tx, ty, tz, vel[3], vel[4], vel[5], h[0], h[1], h[2]);
          fk[0] += tx; fk[1] += ty; fk[2] += tz;
          RBD_CROSS(tx, ty, tz, vel[3], vel[4], vel[5], h[3], h[4], h[5]);
          fk[3] += tx; fk[4] += ty; fk[5] += tz;
          RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], h[0], h[1], h[2]);
          fk[3] += tx; fk[4] += ty; fk[5] += tz;
        }
      }
    }
    if (nxt_par == i) {
      // joint i has children: park A | f | Y at its depth level; its subtree sums accumulate there
      if (NEED_STACK) {
        const int lvl = jn.depth - 1;
        if (!RBD_CUR_FF(ffj, jn.type)) s_stv<6>(S, jn.offA, Ak); else s_stv<12>(S, jn.offA, Ak);
        if (DO_TAU) x_stv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fk);
        if (DO_M) x_stv<9>(RBD_Y_TM(lvl), S, T, RBD_Y_OFF(lvl), Yk + 1);
      }
      if (RBD_CAN_BRANCH && jn.nchild >= 2) {
        const bool btm = (jn.boff & RBD_BOFF_TMEM) != 0;
        int b = jn.boff & RBD_BOFF_MASK;
        if (!(NEED_STACK && RBD_CUR_FF(ffj, jn.type)) && !(jn.boff & RBD_BOFF_RPOMI)) {
          real Rp[12];
#pragma unroll
          for (int e = 0; e < 9; ++e) Rp[e] = R[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) Rp[9 + e] = p[e];
          x_stv<12>(btm, S, T, b, Rp);
          b += 12;
        }
        if (DO_TAU && !(jn.boff & RBD_BOFF_NOVA)) {
          real va[12];
#pragma unroll
          for (int e = 0; e < 6; ++e) { va[e] = vel[e]; va[6 + e] = acc[e]; }
          x_stv<12>(btm, S, T, b, va);
        }
      }
    } else if (NEED_STACK) {
      // joint i is a leaf: its subtree is complete.  Emit it straight from registers, then unwind: every
      // ancestor whose last child this was is complete too; its sums continue in registers up the chain and
      // are written back only at the joint that still has children to come (nxt_par).
      if (RBD_CAN_BRANCH && RBD_EARLY_RESTORE_NS && i + 1 < nj) restore(nxt_par);  // early: see `restore`
      int k = i;
#if defined(RBD_FF_ROOT_ONLY)
      // (peeled: joint i is of the step's own kind, the ancestors emitted on the way up are never the root; this form
      // also needs fewer registers than the loop below — 168 without spills)
      emit_joint<DO_M, DO_TAU, TILE, RBD_EMIT_CUR(ffj)>(m, k, Ak, fk, Yk, S, io);
      while (true) {
        const int pk = m.joint[k].parent;
        if (pk == 0) break;
        const DevJoint& jp = m.joint[pk];
        const int lvl = jp.depth - 1;
        if (pk == nxt_par) {
          if (DO_TAU) x_addv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fk);
          if (DO_M) x_addv<9>(RBD_Y_TM(lvl), S, T, RBD_Y_OFF(lvl), Yk + 1);
          break;
        }
        // the free-flyer root is complete (this was the last joint): fk / Yk carry its last subtree to the epilogue
        if (pk == 1) break;
        if (DO_TAU) x_accv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fk);
        if (DO_M) {
          x_accv<9>(RBD_Y_TM(lvl), S, T, RBD_Y_OFF(lvl), Yk + 1);
          Yk[0] = jp.msub;  // composite mass of the subtree: model constant
        }
        s_ldv<6>(S, jp.offA, Ak);
        k = pk;
        emit_joint<DO_M, DO_TAU, TILE, 2>(m, k, Ak, fk, Yk, S, io);
      }
#else
      while (true) {
        emit_joint<DO_M, DO_TAU, TILE>(m, k, Ak, fk, Yk, S, io);
        const int pk = m.joint[k].parent;
        if (pk == 0) break;
        const DevJoint& jp = m.joint[pk];
        const int lvl = jp.depth - 1;
        if (pk == nxt_par) {
          if (DO_TAU) x_addv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fk);
          if (DO_M) x_addv<9>(RBD_Y_TM(lvl), S, T, RBD_Y_OFF(lvl), Yk + 1);
          break;
        }
        if (DO_TAU) x_accv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fk);
        if (DO_M) {
          x_accv<9>(RBD_Y_TM(lvl), S, T, RBD_Y_OFF(lvl), Yk + 1);
          Yk[0] = jp.msub;  // composite mass of the subtree: model constant
        }
        if (!RBD_T_IS_FF(jp.type)) s_ldv<6>(S, jp.offA, Ak); else s_ldv<12>(S, jp.offA, Ak);
        k = pk;
      }
#endif
    }
  };
#if defined(RBD_FF_ROOT_ONLY)
  joint_step(FfYes{}, 1);
  for (int i = 2; i < nj; ++i) joint_step(FfNo{}, i);
  if (NEED_STACK && m.joint[1].nchild > 0) {
    // epilogue: the root's subtree sums = its own record + what the last unwind carried here in registers
    const DevJoint& jr = m.joint[1];
    if (DO_TAU) x_accv<6>(RBD_F_TM(0), S, T, RBD_F_OFF(0), fk);
    if (DO_M) {
      x_accv<9>(RBD_Y_TM(0), S, T, RBD_Y_OFF(0), Yk + 1);
      Yk[0] = jr.msub;
    }
    s_ldv<12>(S, jr.offA, Ak);
    emit_joint<DO_M, DO_TAU, TILE, 1>(m, 1, Ak, fk, Yk, S, io);
  }
#else
  for (int i = 1; i < nj; ++i) joint_step(FfNo{}, i);
#endif
  inr.primed = io.next != 0 ? 1 : 0;
#undef RBD_STEP_STATE
#undef RBD_RESET_ROOT
#undef RBD_IN_FETCH
#undef RBD_INR_ISSUE
#undef RBD_F_TM
#undef RBD_F_OFF
#undef RBD_Y_TM
#undef RBD_Y_OFF
}

#if !defined(RBD_BODY_F32) && !defined(RBD_BODY_EVAL_ONLY)  /* the mapping path exists in FP64 only */
// =========================================================================================================
// Mapping path
// =========================================================================================================
// one-joint-ahead prefetch of a 1-dof joint's configuration value(s) in the mapping sweeps (`q`, `m`, `nj` in scope)
#define RBD_Q_PREFETCH(idx)                                              \
  do {                                                                   \
    if ((idx) < nj) {                                                    \
      const DevJoint& jx_ = m.joint[idx];                                \
      if (jx_.type != JT_FF) {                                           \
        q0n = q.ld(jx_.idx_q);                                           \
        if (is_unbounded(jx_.type)) q1n = q.ld(jx_.idx_q + 1);           \
      }                                                                  \
    }                                                                    \
  } while (0)

// Full configuration seen through the SOFA split (root pose | joint dofs): KCM.inl:353-363.
template <class FLD>
struct SplitQT {
  FLD root, joints;
  int off;  // 7 (or 6 for velocities) when a root is supplied, else 0
  RBD_HD real ld(int k) const { return k < off ? root.ld(k) : joints.ld(k - off); }
  RBD_HD const real* at(int k) const { return k < off ? root.at(k) : joints.at(k - off); }
};
using SplitQ = SplitQT<Field>;

// apply: q -> [p, quat] of every mapped output; also copies the assembled q into the workspace cache.
template <class QF, class QC, class OUT>
RBD_HD void apply_instance(const DevModel& m, const DevTables& tb, const QF& q, const QC& qcache,
                           const OUT& out, const Stack& S, const MapRing& ring = MapRing()) {
  const int nj = m.njoints;
  // The assembled configuration goes to the workspace cache joint by joint, as the sweep consumes it (a copy loop up
  // front was 39 dependent load -> store round trips before the first joint: profile r1h, long scoreboard 3.1 / issue)
  // ring: slot s holds the angle of the joint RBD_MAP_SLOTS positions ahead
  const bool use_ring = ring.p != nullptr;
#define RBD_MAP_ISSUE(jidx, slot_)                                                    \
  do {                                                                                \
    if ((jidx) < nj && m.joint[jidx].type != JT_FF) cp_async_map(ring, (slot_), q.at(m.joint[jidx].idx_q)); \
    cp_async_commit();                                                                \
  } while (0)
  int slot = 0;
  if (use_ring) {
#pragma unroll
    for (int s_ = 0; s_ < RBD_MAP_SLOTS; ++s_) RBD_MAP_ISSUE(1 + s_, s_);
  }
  // outputs attached to the universe (kSkipUniverse = 0, reference Types.h:32): constant poses
  {
    const DevJoint& j0 = m.joint[0];
    for (int o = j0.out_begin; o < j0.out_end; ++o) {
      const real* pl = tb.sorted_pl + 12 * o;
      const int k = tb.sorted_k[o];
      real qt[4];
      rot_to_quat(pl, qt);
      real* const po = out.at(7 * k);
      out.st_at(po, 0, pl[9]); out.st_at(po, 1, pl[10]); out.st_at(po, 2, pl[11]);
      out.st_at(po, 3, qt[0]); out.st_at(po, 4, qt[1]); out.st_at(po, 5, qt[2]); out.st_at(po, 6, qt[3]);
    }
  }
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real q0n = 0.0, q1n = 0.0;
  if (!use_ring) RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
      } else {
        const int b = m.joint[par].bidx * RBD_BRANCH_FK;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S.ld(b + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S.ld(b + 9 + e);
      }
    }
    real aw[3];
    real q0 = q0n, q1 = q1n;
    if (!use_ring) {
      RBD_Q_PREFETCH(i + 1);  // next joint's angle: its HBM latency hides behind this joint's work
    } else {
      cp_async_wait<RBD_MAP_SLOTS - 1>();
      if (jn.type != JT_FF) {
        q0 = ring.p[slot * RBD_WARP];
        if (is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
      }
      RBD_MAP_ISSUE(i + RBD_MAP_SLOTS, slot);
      slot = slot + 1 == RBD_MAP_SLOTS ? 0 : slot + 1;
    }
    if (qcache.ok()) {
      if (jn.type == JT_FF) {
        real t7[7];
#pragma unroll
        for (int e = 0; e < 7; ++e) t7[e] = q.ld(jn.idx_q + e);
#pragma unroll
        for (int e = 0; e < 7; ++e) qcache.st(jn.idx_q + e, t7[e]);
      } else {
        qcache.st(jn.idx_q, q0);
        if (is_unbounded(jn.type)) qcache.st(jn.idx_q + 1, q1);
      }
    }
    fk_joint(jn, R, p, q0, q1, q, aw);
    // outputs whose placement has an identity rotation share the joint's rotation: convert it once
    real qj[4];
    bool have_qj = false;
    for (int o = jn.out_begin; o < jn.out_end; ++o) {
      // updateFramePlacements: oMf = oMi * placement (KCM.inl:375), then se3ToSofaType (KCM.inl:385,394)
      const real* pl = tb.sorted_pl + 12 * o;
      const int k = tb.sorted_k[o];
      real qt[4];
      if (tb.sorted_ident[o]) {
        if (!have_qj) { rot_to_quat(R, qj); have_qj = true; }
        qt[0] = qj[0]; qt[1] = qj[1]; qt[2] = qj[2]; qt[3] = qj[3];
      } else {
        real Rf[9];
        mat3_mul(R, pl, Rf);
        rot_to_quat(Rf, qt);
      }
      const real px = R[0] * pl[9] + R[1] * pl[10] + R[2] * pl[11] + p[0];
      const real py = R[3] * pl[9] + R[4] * pl[10] + R[5] * pl[11] + p[1];
      const real pz = R[6] * pl[9] + R[7] * pl[10] + R[8] * pl[11] + p[2];
      real* const po = out.at(7 * k);
      out.st_at(po, 0, px); out.st_at(po, 1, py); out.st_at(po, 2, pz);
      out.st_at(po, 3, qt[0]); out.st_at(po, 4, qt[1]); out.st_at(po, 5, qt[2]); out.st_at(po, 6, qt[3]);
    }
    if (jn.nchild >= 2) {
      const int b = jn.bidx * RBD_BRANCH_FK;
#pragma unroll
      for (int e = 0; e < 9; ++e) S.st(b + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) S.st(b + 9 + e, p[e]);
    }
  }
#undef RBD_MAP_ISSUE
}

// applyJ: out_k = J_k(LWA) dq, matrix-free: world twists V_i = V_parent + A_i dq_i, then
// out_k = [V_lin + V_ang x p_k ; V_ang]  (== [J_lin - p_k x J_ang ; J_ang] dq, getFrameJacobian LWA)
template <class QC, class DQ, class OUT>
RBD_HD void applyJ_instance(const DevModel& m, const DevTables& tb, const QC& q, const DQ& dq, const OUT& out,
                            const Stack& S, const MapRing& ring = MapRing()) {
  const int nj = m.njoints;
  // ring: slot s holds (angle, joint velocity) of the joint RBD_MAP_SLOTS positions ahead (the velocity used to be
  // loaded at its point of use: profile r1h, long scoreboard 6.7 / issue)
  const bool use_ring = ring.p != nullptr;
#define RBD_MAP_ISSUE(jidx, slot_)                                                    \
  do {                                                                                \
    if ((jidx) < nj && m.joint[jidx].type != JT_FF) {                                 \
      cp_async_map(ring, 2 * (slot_), q.at(m.joint[jidx].idx_q));                     \
      cp_async_map(ring, 2 * (slot_) + 1, dq.at(m.joint[jidx].idx_v));                \
    }                                                                                 \
    cp_async_commit();                                                                \
  } while (0)
  int slot = 0;
  if (use_ring) {
#pragma unroll
    for (int s_ = 0; s_ < RBD_MAP_SLOTS; ++s_) RBD_MAP_ISSUE(1 + s_, s_);
  }
  {
    const DevJoint& j0 = m.joint[0];
    for (int o = j0.out_begin; o < j0.out_end; ++o) {
      const int k = tb.sorted_k[o];
#pragma unroll
      for (int e = 0; e < 6; ++e) out.st(6 * k + e, 0.0);  // universe: empty support -> zero Jacobian
    }
  }
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}, V[6] = {0, 0, 0, 0, 0, 0};
  real q0n = 0.0, q1n = 0.0;
  if (!use_ring) RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
#pragma unroll
        for (int e = 0; e < 6; ++e) V[e] = 0;
      } else {
        const int b = m.joint[par].bidx * RBD_BRANCH_J;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S.ld(b + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S.ld(b + 9 + e);
#pragma unroll
        for (int e = 0; e < 6; ++e) V[e] = S.ld(b + 12 + e);
      }
    }
    real aw[3];
    real q0 = q0n, q1 = q1n, d = 0.0;
    if (!use_ring) {
      RBD_Q_PREFETCH(i + 1);  // next joint's angle: its HBM latency hides behind this joint's work
      if (jn.type != JT_FF) d = dq.ld(jn.idx_v);
    } else {
      cp_async_wait<RBD_MAP_SLOTS - 1>();
      if (jn.type != JT_FF) {
        q0 = ring.p[2 * slot * RBD_WARP];
        d = ring.p[(2 * slot + 1) * RBD_WARP];
        if (is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
      }
      RBD_MAP_ISSUE(i + RBD_MAP_SLOTS, slot);
      slot = slot + 1 == RBD_MAP_SLOTS ? 0 : slot + 1;
    }
    fk_joint(jn, R, p, q0, q1, q, aw);
    if (jn.type != JT_FF) {
      real A[6];
      subspace_1dof(jn.type, p, aw, A);
#pragma unroll
      for (int e = 0; e < 6; ++e) V[e] += A[e] * d;
    } else {
      const int o = jn.idx_v;
      const real l0 = dq.ld(o), l1 = dq.ld(o + 1), l2 = dq.ld(o + 2);
      const real w0 = dq.ld(o + 3), w1 = dq.ld(o + 4), w2 = dq.ld(o + 5);
      const real wx = R[0] * w0 + R[1] * w1 + R[2] * w2;
      const real wy = R[3] * w0 + R[4] * w1 + R[5] * w2;
      const real wz = R[6] * w0 + R[7] * w1 + R[8] * w2;
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], wx, wy, wz);
      V[0] += R[0] * l0 + R[1] * l1 + R[2] * l2 + tx;
      V[1] += R[3] * l0 + R[4] * l1 + R[5] * l2 + ty;
      V[2] += R[6] * l0 + R[7] * l1 + R[8] * l2 + tz;
      V[3] += wx; V[4] += wy; V[5] += wz;
    }
    for (int o = jn.out_begin; o < jn.out_end; ++o) {
      const real* pl = tb.sorted_pl + 12 * o;
      const int k = tb.sorted_k[o];
      const real px = R[0] * pl[9] + R[1] * pl[10] + R[2] * pl[11] + p[0];
      const real py = R[3] * pl[9] + R[4] * pl[10] + R[5] * pl[11] + p[1];
      const real pz = R[6] * pl[9] + R[7] * pl[10] + R[8] * pl[11] + p[2];
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, V[3], V[4], V[5], px, py, pz);
      real* const po = out.at(6 * k);
      out.st_at(po, 0, V[0] + tx); out.st_at(po, 1, V[1] + ty); out.st_at(po, 2, V[2] + tz);
      out.st_at(po, 3, V[3]); out.st_at(po, 4, V[4]); out.st_at(po, 5, V[5]);
    }
    if (jn.nchild >= 2) {
      const int b = jn.bidx * RBD_BRANCH_J;
#pragma unroll
      for (int e = 0; e < 9; ++e) S.st(b + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) S.st(b + 9 + e, p[e]);
#pragma unroll
      for (int e = 0; e < 6; ++e) S.st(b + 12 + e, V[e]);
    }
  }
#undef RBD_MAP_ISSUE
}

// applyJT, matrix-free: every wrench w_k = [f; n] given at output k (world-aligned axes) is moved to the
// world origin, [f ; n + p_k x f], summed per joint, accumulated leaf->root, and projected: tau_i = A_i^T F_i.
// `Src::gather(i, begin, end, R, p, F)` adds the wrenches attached to joint i; `Sink::put(c, value)` receives
// the generalized force of velocity index c.
template <class QC, class Src, class Sink>
RBD_HD void applyJT_instance(const DevModel& m, const DevTables& tb, const QC& q, const Src& src, const Sink& sink,
                             const Stack& S) {
  const int nj = m.njoints;
  const int bbase = m.slot_total[1];
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real q0n = 0.0, q1n = 0.0;
  RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (par != i - 1) {
      // joint i starts a new branch below `par` (the subtree that ended at i-1 was unwound already)
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
      } else {
        const int b = bbase + m.joint[par].bidx * RBD_BRANCH_FK;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S.ld(b + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S.ld(b + 9 + e);
      }
    }
    const int nxt_par = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    real aw[3];
    const real q0 = q0n, q1 = q1n;
    RBD_Q_PREFETCH(i + 1);  // next joint's angle: its HBM latency hides behind this joint's work
    fk_joint(jn, R, p, q0, q1, q, aw);
    // Ak: world subspace column (1-dof) or R|p (free-flyer); Fk: wrenches attached to this joint, about the origin
    real Ak[12], Fk[6] = {0, 0, 0, 0, 0, 0};
    if (jn.type != JT_FF) {
      subspace_1dof(jn.type, p, aw, Ak);
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) Ak[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) Ak[9 + e] = p[e];
    }
    src.gather(i, jn.out_begin, jn.out_end, R, p, Fk);
    if (nxt_par == i) {
      // joint i has children: park A | F in its depth slot; the subtree's wrenches accumulate there
      const int s = jn.slot[1];
      const int aw_ = RBD_SLOT_A(jn.type);
      if (jn.type != JT_FF) s_stv<6>(S, s, Ak); else s_stv<12>(S, s, Ak);
      s_stv<6>(S, s + aw_, Fk);
      if (jn.nchild >= 2) {
        const int b = bbase + jn.bidx * RBD_BRANCH_FK;
#pragma unroll
        for (int e = 0; e < 9; ++e) S.st(b + e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) S.st(b + 9 + e, p[e]);
      }
    } else {
      // leaf: project from registers, then unwind — sums travel up the chain in registers and are written back
      // only at the joint that still has children to come
      int k = i;
      while (true) {
        const DevJoint& jk = m.joint[k];
        if (jk.type != JT_FF) {
          sink.put(jk.idx_v, dot6(Ak, Fk));
        } else {
          real o6[6];
          ff_rows(Ak, Ak + 9, Fk, o6);
#pragma unroll
          for (int e = 0; e < 6; ++e) sink.put(jk.idx_v + e, o6[e]);
        }
        const int pk = jk.parent;
        if (pk == 0) break;
        const DevJoint& jp = m.joint[pk];
        const int sp = jp.slot[1];
        const int awp = RBD_SLOT_A(jp.type);
        if (pk == nxt_par) {
#pragma unroll
          for (int e = 0; e < 6; ++e) S.add(sp + awp + e, Fk[e]);
          break;
        }
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += S.ld(sp + awp + e);
        if (jp.type != JT_FF) s_ldv<6>(S, sp, Ak); else s_ldv<12>(S, sp, Ak);
        k = pk;
      }
    }
  }
}

// applyDJT — the geometric stiffness of the mapping: d/de [ J(q (+) e dq)^T w ] at e = 0, i.e. how the generalized
// forces of FIXED world-aligned wrenches w_k change when the configuration moves along dq.  The reference leaves it a
// no-op (KinematicChainMapping.h:73-80: "geometric stiffness absent"); this is the
// optional real thing (rbd_applyDJT_soa), evaluated matrix-free in one sweep:
//   tau_i = A_i . F_i,   F_i = sum over the subtree of [f_k ; n_k + p_k x f_k]
//   dA_i  = dV_parent(i) x A_i            (motion cross product; dV = body twists of applyJ for this dq)
//   dF_i  = sum over the subtree of [0 ; dp_k x f_k],   dp_k = dV_lin + dV_ang x p_k of the body carrying output k
//   dtau_i = dA_i . F_i + A_i . dF_i
// and for a free-flyer ROOT (its velocity is given in the joint frame, so its columns move with its own pose):
//   dtau_root = ff_rows(R, p, dF) + [ -w_b x tau_l ; -w_b x tau_a - v_b x tau_l ].
// Stack: per depth level [A (12) | dA (6) | F (6) | dF (6)], per branch joint [R|p (12) | dV (6)], shared memory.
// sink.put(c, value) receives kfac * dtau_c (accumulated by the caller's sink, like SOFA's parentForce +=).
#define RBD_DJT_LEVEL 30
#define RBD_DJT_BRANCH 18
template <class QC, class DQ, class FLD, class Sink>
RBD_HD void applyDJT_instance(const DevModel& m, const DevTables& tb, const QC& q, const DQ& dq, const FLD& in,
                              const Sink& sink, const Stack& S, real kfac) {
  const int nj = m.njoints;
  const int bbase = RBD_DJT_LEVEL * m.max_depth;
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}, dV[6] = {0, 0, 0, 0, 0, 0};
  real q0n = 0.0, q1n = 0.0;
  RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
#pragma unroll
        for (int e = 0; e < 6; ++e) dV[e] = 0;
      } else {
        const int b = bbase + m.joint[par].bidx * RBD_DJT_BRANCH;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = S.ld(b + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = S.ld(b + 9 + e);
#pragma unroll
        for (int e = 0; e < 6; ++e) dV[e] = S.ld(b + 12 + e);
      }
    }
    const int nxt_par = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    real aw[3];
    const real q0 = q0n, q1 = q1n;
    RBD_Q_PREFETCH(i + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    // rec = [A (12) | dA (6) | F (6) | dF (6)] of this joint
    real rec[RBD_DJT_LEVEL];
#pragma unroll
    for (int e = 0; e < RBD_DJT_LEVEL; ++e) rec[e] = 0;
    real vb[6] = {0, 0, 0, 0, 0, 0};  // free-flyer: its joint-frame velocity
    if (jn.type != JT_FF) {
      subspace_1dof(jn.type, p, aw, rec);
      // dA = dV_parent x A
      RBD_CROSS(rec[12], rec[13], rec[14], dV[3], dV[4], dV[5], rec[0], rec[1], rec[2]);
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, dV[0], dV[1], dV[2], rec[3], rec[4], rec[5]);
      rec[12] += tx; rec[13] += ty; rec[14] += tz;
      RBD_CROSS(rec[15], rec[16], rec[17], dV[3], dV[4], dV[5], rec[3], rec[4], rec[5]);
      const real d = dq.ld(jn.idx_v);
#pragma unroll
      for (int e = 0; e < 6; ++e) dV[e] += rec[e] * d;
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) rec[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) rec[9 + e] = p[e];
#pragma unroll
      for (int e = 0; e < 6; ++e) vb[e] = dq.ld(jn.idx_v + e);
      // world twist of the free-flyer body: oMi.act(v_b)
      const real wx = R[0] * vb[3] + R[1] * vb[4] + R[2] * vb[5];
      const real wy = R[3] * vb[3] + R[4] * vb[4] + R[5] * vb[5];
      const real wz = R[6] * vb[3] + R[7] * vb[4] + R[8] * vb[5];
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], wx, wy, wz);
      dV[0] += R[0] * vb[0] + R[1] * vb[1] + R[2] * vb[2] + tx;
      dV[1] += R[3] * vb[0] + R[4] * vb[1] + R[5] * vb[2] + ty;
      dV[2] += R[6] * vb[0] + R[7] * vb[1] + R[8] * vb[2] + tz;
      dV[3] += wx; dV[4] += wy; dV[5] += wz;
      // the joint-frame velocity is needed again when the root is projected: keep it in the dA slot
#pragma unroll
      for (int e = 0; e < 6; ++e) rec[12 + e] = vb[e];
    }
    // wrenches of this joint's outputs and their first-order change
    real* const F = rec + 18;
    real* const dF = rec + 24;
    for (int o = jn.out_begin; o < jn.out_end; ++o) {
      const int k = tb.sorted_k[o];
      const real* pl = tb.sorted_pl + 12 * o;
      real w[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) w[e] = in.ld(6 * k + e);
      const real px = R[0] * pl[9] + R[1] * pl[10] + R[2] * pl[11] + p[0];
      const real py = R[3] * pl[9] + R[4] * pl[10] + R[5] * pl[11] + p[1];
      const real pz = R[6] * pl[9] + R[7] * pl[10] + R[8] * pl[11] + p[2];
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);
      F[0] += w[0]; F[1] += w[1]; F[2] += w[2];
      F[3] += w[3] + tx; F[4] += w[4] + ty; F[5] += w[5] + tz;
      // dp = dV_lin + dV_ang x p_k
      real dx, dy, dz;
      RBD_CROSS(dx, dy, dz, dV[3], dV[4], dV[5], px, py, pz);
      dx += dV[0]; dy += dV[1]; dz += dV[2];
      RBD_CROSS(tx, ty, tz, dx, dy, dz, w[0], w[1], w[2]);
      dF[3] += tx; dF[4] += ty; dF[5] += tz;
    }
    if (nxt_par == i) {
      const int s = RBD_DJT_LEVEL * (jn.depth - 1);
      s_stv<RBD_DJT_LEVEL>(S, s, rec);
      if (jn.nchild >= 2) {
        const int b = bbase + jn.bidx * RBD_DJT_BRANCH;
#pragma unroll
        for (int e = 0; e < 9; ++e) S.st(b + e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) S.st(b + 9 + e, p[e]);
#pragma unroll
        for (int e = 0; e < 6; ++e) S.st(b + 12 + e, dV[e]);
      }
    } else {
      int k = i;
      while (true) {
        const DevJoint& jk = m.joint[k];
        if (jk.type != JT_FF) {
          sink.put(jk.idx_v, kfac * (dot6(rec + 12, F) + dot6(rec, dF)));
        } else {
          // free-flyer root: tau = ff_rows(R, p, F); dtau = ff_rows(R, p, dF) + [-w x tau_l ; -w x tau_a - v x tau_l]
          real t6[6], d6[6];
          ff_rows(rec, rec + 9, F, t6);
          ff_rows(rec, rec + 9, dF, d6);
          const real* v_b = rec + 12;
          real ax, ay, az, bx, by, bz, cx, cy, cz;
          RBD_CROSS(ax, ay, az, v_b[3], v_b[4], v_b[5], t6[0], t6[1], t6[2]);
          RBD_CROSS(bx, by, bz, v_b[3], v_b[4], v_b[5], t6[3], t6[4], t6[5]);
          RBD_CROSS(cx, cy, cz, v_b[0], v_b[1], v_b[2], t6[0], t6[1], t6[2]);
          sink.put(jk.idx_v + 0, kfac * (d6[0] - ax)); sink.put(jk.idx_v + 1, kfac * (d6[1] - ay));
          sink.put(jk.idx_v + 2, kfac * (d6[2] - az));
          sink.put(jk.idx_v + 3, kfac * (d6[3] - bx - cx)); sink.put(jk.idx_v + 4, kfac * (d6[4] - by - cy));
          sink.put(jk.idx_v + 5, kfac * (d6[5] - bz - cz));
        }
        const int pk = jk.parent;
        if (pk == 0) break;
        const int sp = RBD_DJT_LEVEL * (m.joint[pk].depth - 1);
        if (pk == nxt_par) {
#pragma unroll
          for (int e = 0; e < 12; ++e) S.add(sp + 18 + e, rec[18 + e]);
          break;
        }
        real Fs[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) Fs[e] = rec[18 + e];
        s_ldv<RBD_DJT_LEVEL>(S, sp, rec);
#pragma unroll
        for (int e = 0; e < 12; ++e) rec[18 + e] += Fs[e];
        k = pk;
      }
    }
  }
}

#ifndef RBD_JT_DIST
#define RBD_JT_DIST 1 /* joints of lookahead of the wrench prefetch (1 or 2) */
#endif
#ifndef RBD_JT_PREF
#define RBD_JT_PREF 4 /* outputs per joint whose wrenches are prefetched one joint ahead (A/B: 3..6 within 3 %) */
#endif
RBD_HD void add_shifted_wrench(const real* R, const real* p, const real* pl, const real* w, real* F);

// applyJT on the fused-eval storage plan (plan_eval(..., jt = true)): A-stack in shared memory, F-stack levels and
// branch R|p records in shared or tensor memory — the batched vector path (K3).  Same sweep as applyJT_instance.
template <class QC, class Src, class Sink>
RBD_HD void applyJT_instance_plan(const DevModel& m, const DevTables& tb, const QC& q, const Src& src, const Sink& sink,
                                  const Stack& S, const TmStack& T) {
  const int nj = m.njoints;
  const int fbase = m.plan.fbase, fsplit = m.plan.fsplit, ftbase = m.plan.ftbase;
#define RBD_F_TM(lvl) ((lvl) >= fsplit)
#define RBD_F_OFF(lvl) ((lvl) >= fsplit ? ftbase + 6 * ((lvl)-fsplit) : fbase + 6 * (lvl))
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real wn[6 * RBD_JT_PREF];
#pragma unroll
  for (int e = 0; e < 6 * RBD_JT_PREF; ++e) wn[e] = 0.0;
  if (nj > 1) src.prefetch(m.joint[1].out_begin, m.joint[1].out_end, wn);
#if RBD_JT_DIST == 2
  real wn2[6 * RBD_JT_PREF];
#pragma unroll
  for (int e = 0; e < 6 * RBD_JT_PREF; ++e) wn2[e] = 0.0;
  if (nj > 2) src.prefetch(m.joint[2].out_begin, m.joint[2].out_end, wn2);
#endif
  real q0n = 0.0, q1n = 0.0;
  RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    // this joint's wrenches arrived while the previous joint(s) were processed; start the loads of the joint
    // RBD_JT_DIST ahead now (A/B: issuing them only after this joint's gather, without the copy, is 16 % slower)
    real wc[6 * RBD_JT_PREF];
#pragma unroll
    for (int e = 0; e < 6 * RBD_JT_PREF; ++e) wc[e] = wn[e];
#if RBD_JT_DIST == 2
#pragma unroll
    for (int e = 0; e < 6 * RBD_JT_PREF; ++e) wn[e] = wn2[e];
    if (i + 2 < nj) src.prefetch(m.joint[i + 2].out_begin, m.joint[i + 2].out_end, wn2);
#else
    if (i + 1 < nj) src.prefetch(m.joint[i + 1].out_begin, m.joint[i + 1].out_end, wn);
#endif
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
      } else {
        const DevJoint& jb = m.joint[par];
        real Rp[12];
        if (jb.type == JT_FF) s_ldv<12>(S, jb.offA, Rp);
        else x_ldv<12>((jb.boff & RBD_BOFF_TMEM) != 0, S, T, jb.boff & RBD_BOFF_MASK, Rp);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = Rp[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = Rp[9 + e];
      }
    }
    const int nxt_par = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    real aw[3];
    const real q0 = q0n, q1 = q1n;
    RBD_Q_PREFETCH(i + 1);  // next joint's angle: its HBM latency hides behind this joint's work
    fk_joint(jn, R, p, q0, q1, q, aw);
    real Ak[12], Fk[6] = {0, 0, 0, 0, 0, 0};
    if (jn.type != JT_FF) {
      subspace_1dof(jn.type, p, aw, Ak);
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) Ak[e] = R[e];
#pragma unroll
      for (int e = 0; e < 3; ++e) Ak[9 + e] = p[e];
    }
    src.gather_pref(jn.out_begin, jn.out_end, R, p, wc, Fk);
    if (nxt_par == i) {
      const int lvl = jn.depth - 1;
      if (jn.type != JT_FF) s_stv<6>(S, jn.offA, Ak); else s_stv<12>(S, jn.offA, Ak);
      x_stv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), Fk);
      if (jn.nchild >= 2 && jn.type != JT_FF) {
        real Rp[12];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rp[e] = R[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) Rp[9 + e] = p[e];
        x_stv<12>((jn.boff & RBD_BOFF_TMEM) != 0, S, T, jn.boff & RBD_BOFF_MASK, Rp);
      }
    } else {
      int k = i;
      while (true) {
        const DevJoint& jk = m.joint[k];
        if (jk.type != JT_FF) {
          sink.put(jk.idx_v, dot6(Ak, Fk));
        } else {
          real o6[6];
          ff_rows(Ak, Ak + 9, Fk, o6);
#pragma unroll
          for (int e = 0; e < 6; ++e) sink.put(jk.idx_v + e, o6[e]);
        }
        const int pk = jk.parent;
        if (pk == 0) break;
        const DevJoint& jp = m.joint[pk];
        const int lvl = jp.depth - 1;
        if (pk == nxt_par) {
          x_addv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), Fk);
          break;
        }
        real fp[6];
        x_ldv<6>(RBD_F_TM(lvl), S, T, RBD_F_OFF(lvl), fp);
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += fp[e];
        if (jp.type != JT_FF) s_ldv<6>(S, jp.offA, Ak); else s_ldv<12>(S, jp.offA, Ak);
        k = pk;
      }
    }
  }
#undef RBD_F_TM
#undef RBD_F_OFF
}

// add wrench w (at point pf = R pl_p + p, world-aligned axes) moved to the world origin
RBD_HD void add_shifted_wrench(const real* R, const real* p, const real* pl, const real* w, real* F) {
  const real px = R[0] * pl[9] + R[1] * pl[10] + R[2] * pl[11] + p[0];
  const real py = R[3] * pl[9] + R[4] * pl[10] + R[5] * pl[11] + p[1];
  const real pz = R[6] * pl[9] + R[7] * pl[10] + R[8] * pl[11] + p[2];
  real tx, ty, tz;
  RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);
  F[0] += w[0]; F[1] += w[1]; F[2] += w[2];
  F[3] += w[3] + tx; F[4] += w[4] + ty; F[5] += w[5] + tz;
}

// dense wrench source: one wrench per mapped output (KCM.inl:472-492)
template <class FLD>
struct DenseWrenchSrcT {
  FLD in;  // n_out*6 fields
  const DevTables* tb;
  // Issue the loads of the first RBD_JT_PREF wrenches of outputs [begin, end) (indices past the end are clamped, so
  // every load is valid; the duplicates are ignored by gather_pref).  The sweep calls this one joint ahead: the HBM
  // latency of the wrench stream hides behind the previous joint's kinematics.
  RBD_HD void prefetch(int begin, int end, real* w) const {
    if (end <= begin) return;
#pragma unroll
    for (int j = 0; j < RBD_JT_PREF; ++j) {
      const int o = begin + j < end ? begin + j : end - 1;
      const int k = tb->sorted_k[o];
#pragma unroll
      for (int e = 0; e < 6; ++e) w[6 * j + e] = in.ld(6 * k + e);
    }
  }
  RBD_HD void gather_pref(int begin, int end, const real* R, const real* p, const real* w, real* F) const {
#pragma unroll
    for (int j = 0; j < RBD_JT_PREF; ++j)
      if (begin + j < end) add_shifted_wrench(R, p, tb->sorted_pl + 12 * (begin + j), w + 6 * j, F);
    if (end > begin + RBD_JT_PREF) gather(0, begin + RBD_JT_PREF, end, R, p, F);
  }
  RBD_HD void gather(int, int begin, int end, const real* R, const real* p, real* F) const {
    // independent iterations: unrolled so that the wrench loads of several outputs are in flight together
#pragma unroll 4
    for (int o = begin; o < end; ++o) {
      const int k = tb->sorted_k[o];
      real w[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) w[e] = in.ld(6 * k + e);
      add_shifted_wrench(R, p, tb->sorted_pl + 12 * o, w, F);
    }
  }
};
using DenseWrenchSrc = DenseWrenchSrcT<Field>;
// sparse wrench source: the non-zeros of one constraint row (KCM.inl:558-583)
struct RowWrenchSrc {
  const int* col;     // [nnz] mapped-output index
  const real* val;  // [nnz][6]
  int nnz;
  const DevTables* tb;
  RBD_HD void gather(int joint, int, int, const real* R, const real* p, real* F) const {
    for (int j = 0; j < nnz; ++j) {
      const int k = col[j];
      if (tb->outk_parent[k] == joint) add_shifted_wrench(R, p, tb->outk_pl + 12 * k, val + 6 * j, F);
    }
  }
};
// += into SOFA's split outputs (root wrench | joint torques): KCM.inl:494-512
template <class FLD>
struct SplitAccumSinkT {
  FLD root, joints;
  int off;
  RBD_HD void put(int c, real v) const {
    if (c < off) {
      if (root.ok()) root.st(c, root.ld(c) + v);
    } else if (joints.ok()) {
      joints.st(c - off, joints.ld(c - off) + v);
    }
  }
};
using SplitAccumSink = SplitAccumSinkT<Field>;
// the same += as a reduction that returns nothing (device: RED.ADD.F64 — one IEEE add per address and launch, so the
// result is that of load + add + store, but the sweep never waits for the old value)
template <class FLD>
struct SplitRedSinkT {
  FLD root, joints;
  int off;
  static RBD_HD void red(real* a, real v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(a, v);
#else
    *a += v;
#endif
  }
  RBD_HD void put(int c, real v) const {
    if (c < off) {
      if (root.ok()) red(root.at(c), v);
    } else if (joints.ok()) {
      red(joints.at(c - off), v);
    }
  }
};
// dense row + squared norm (KCM.inl:585-615)
struct RowSink {
  real* row;
  real* n2;
  RBD_HD void put(int c, real v) const { row[c] = v; *n2 += v * v; }
};

// ---------------------------------------------------------------------------------------------------------
// applyJT, ring variant (K3 on full tile-32 batches): the wrenches do not travel through registers.  A warp's 32
// instances form one tile, so the wrench of mapped output k is ONE contiguous 1 536-byte block of the input
// (6 fields x 32 lanes x 8 bytes): the kernel streams those blocks with bulk asynchronous copies
// (cp.async.bulk, completion on an mbarrier) into a per-warp ring in shared memory, many outputs — and across
// tile boundaries — ahead of the sweep, and the sweep reads them from there.  `Ring` is that ring on the device
// (rbd_kernels.cu) and a plain reader of the input array in tests/emul:
//   int  acquire()         wait for the next piece (<= gmax outputs of one joint, in sweep order); returns its length
//   void ld(j, w)          wrench of the j-th output of the acquired piece
//   void release(cnt)      the piece has been consumed: its slots are refilled with the outputs furthest ahead
// Sweep state (plan_jt_ring), placed at COMPILE time so that the register allocator never has to merge a
// tensor-memory tuple with a shared-memory load (profile r1h: 22 % of the executed instructions of the first ring
// kernel were register moves around exactly those merges):
//   tensor memory   one record [F (6) | A (6)] per depth level of 1-dof joints with children; R|p (12) per 1-dof
//                   branch joint
//   shared memory   free-flyer level record [R|p (12) | F (6)]
// Accumulated outputs go through Sink::put (a fire-and-forget reduction on the device: the += of the reference,
// KCM.inl:498-511, without a load the sweep would have to wait for).
template <class QC, class Ring, class Sink>
RBD_HD void applyJT_instance_ring(const DevModel& m, const real* opl, const QC& q, Ring& ring, const Sink& sink,
                                  const Stack& S, const TmStack& T) {
  const int nj = m.njoints;
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real q0n = 0.0, q1n = 0.0;
  RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
      } else {
        const DevJoint& jb = m.joint[par];
        real Rp[12];
        if (jb.type == JT_FF) s_ldv<12>(S, jb.offA, Rp);  // a free-flyer's level record starts with its R|p
        else T.template ldv<12>(jb.boff, Rp);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = Rp[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = Rp[9 + e];
      }
    }
    const int nxt_par = (i + 1 < nj) ? m.joint[i + 1].parent : 0;
    real aw[3];
    const real q0 = q0n, q1 = q1n;
    RBD_Q_PREFETCH(i + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    // wrenches attached to this joint, moved to the world origin.  Pieces hold 1..4 outputs; each size is
    // straight-line code (four independent shift chains the scheduler can interleave, summed pairwise).
    real Fk[6] = {0, 0, 0, 0, 0, 0};
#define RBD_JT_W(j, c)                                                                  \
  real c[6];                                                                            \
  {                                                                                     \
    real w[6];                                                                          \
    ring.template ld<j>(w);                                                             \
    const real* pl = opl + 3 * (o + j);                                                 \
    const real px = R[0] * pl[0] + R[1] * pl[1] + R[2] * pl[2] + p[0];                  \
    const real py = R[3] * pl[0] + R[4] * pl[1] + R[5] * pl[2] + p[1];                  \
    const real pz = R[6] * pl[0] + R[7] * pl[1] + R[8] * pl[2] + p[2];                  \
    real tx, ty, tz;                                                                    \
    RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);                                \
    c[0] = w[0]; c[1] = w[1]; c[2] = w[2];                                              \
    c[3] = w[3] + tx; c[4] = w[4] + ty; c[5] = w[5] + tz;                               \
  }
    for (int o = jn.out_begin; o < jn.out_end;) {
      const int cnt = ring.acquire();
      if (cnt == 1) {
        RBD_JT_W(0, c0)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += c0[e];
      } else if (cnt == 2) {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += c0[e] + c1[e];
      } else if (cnt == 3) {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1) RBD_JT_W(2, c2)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += (c0[e] + c1[e]) + c2[e];
      } else {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1) RBD_JT_W(2, c2) RBD_JT_W(3, c3)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += (c0[e] + c1[e]) + (c2[e] + c3[e]);
      }
      ring.release(cnt);
      o += cnt;
    }
#undef RBD_JT_W
    if (nxt_par == i) {
      // joint i has children: park F | A at its depth level; the subtree's wrenches accumulate there
      if (jn.type != JT_FF) {
        real rec[12];
#pragma unroll
        for (int e = 0; e < 6; ++e) rec[e] = Fk[e];
        subspace_1dof(jn.type, p, aw, rec + 6);
        T.template stv<12>(jn.offA, rec);
        if (jn.nchild >= 2) {
          real Rp[12];
#pragma unroll
          for (int e = 0; e < 9; ++e) Rp[e] = R[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) Rp[9 + e] = p[e];
          T.template stv<12>(jn.boff, Rp);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 9; ++e) S.st(jn.offA + e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) S.st(jn.offA + 9 + e, p[e]);
#pragma unroll
        for (int e = 0; e < 6; ++e) S.st(jn.offA + 12 + e, Fk[e]);
      }
    } else {
      // leaf: project from registers, then unwind — only the running sum Fk is carried from level to level
      if (jn.type != JT_FF) {
        real A[6];
        subspace_1dof(jn.type, p, aw, A);
        sink.put(jn.idx_v, dot6(A, Fk));
      } else {
        real o6[6];
        ff_rows(R, p, Fk, o6);
#pragma unroll
        for (int e = 0; e < 6; ++e) sink.put(jn.idx_v + e, o6[e]);
      }
      int pk = par;
      while (pk != 0) {
        const DevJoint& jp = m.joint[pk];
        if (jp.type != JT_FF) {
          if (pk == nxt_par) { T.template addv<6>(jp.offA, Fk); break; }
          real rec[12];
          T.template ldv<12>(jp.offA, rec);
#pragma unroll
          for (int e = 0; e < 6; ++e) Fk[e] += rec[e];
          sink.put(jp.idx_v, dot6(rec + 6, Fk));
        } else {
          if (pk == nxt_par) {
#pragma unroll
            for (int e = 0; e < 6; ++e) S.add(jp.offA + 12 + e, Fk[e]);
            break;
          }
          real Rp[12], o6[6];
          s_ldv<12>(S, jp.offA, Rp);
#pragma unroll
          for (int e = 0; e < 6; ++e) Fk[e] += S.ld(jp.offA + 12 + e);
          ff_rows(Rp, Rp + 9, Fk, o6);
#pragma unroll
          for (int e = 0; e < 6; ++e) sink.put(jp.idx_v + e, o6[e]);
        }
        pk = jp.parent;
      }
    }
  }
}
// ---------------------------------------------------------------------------------------------------------
// applyJT (MatrixDeriv rows) on CACHED Jacobians — the reference's own structure: apply leaves data.J / data.oMi
// behind (computeJointJacobians, KinematicChainMapping.inl:365-373) and every constraint row only reads them
// (getFrameJacobian + J^T w per non-zero, inl:558-583).  Here the cache is the world joint Jacobian J (6 nv fields)
// and oMi (12 (njoints-1) fields) of the applied instances, filled once per apply by the FK + Jacobian instantiation of
// the eval kernel; a row then costs, per non-zero, one wrench shift and one 6-flop projection per dof on the path
// to the root — no kinematics at all.  `acc(c, t)` receives the contribution t to entry c of the dense row.
template <class JF, class OF, class ACC>
RBD_HD void rows_cached_row(const DevModel& m, const DevTables& tb, const JF& Jc, const OF& Oc, const int* col,
                            const real* val, int nnz, const ACC& acc) {
  for (int j = 0; j < nnz; ++j) {
    const int k = col[j];
    const int pj = tb.outk_parent[k];
    if (pj == 0) continue;  // attached to the universe: empty support, zero Jacobian
    const real* pl = tb.outk_pl + 12 * k;
    const real* w = val + 6 * j;
    real Rp[12];
    Oc.template ldn<12>(12 * (pj - 1), Rp);
    const real px = Rp[0] * pl[9] + Rp[1] * pl[10] + Rp[2] * pl[11] + Rp[9];
    const real py = Rp[3] * pl[9] + Rp[4] * pl[10] + Rp[5] * pl[11] + Rp[10];
    const real pz = Rp[6] * pl[9] + Rp[7] * pl[10] + Rp[8] * pl[11] + Rp[11];
    real F[6], tx, ty, tz;
    RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);
    F[0] = w[0]; F[1] = w[1]; F[2] = w[2];
    F[3] = w[3] + tx; F[4] = w[4] + ty; F[5] = w[5] + tz;
    // walk to the root four dof columns at a time: the column indices come from the staged joint table, then the
    // four Jacobian columns are loaded together (the row kernel is bound by the latency of these scattered reads:
    // one column in flight per lane left it at 24 long-scoreboard stall cycles per issue, r1i)
    int a = pj, cc = 0;  // next dof to visit: dof cc of joint a
    while (a != 0) {
      int cidx[4];
      int ncol = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        cidx[u] = 0;
        if (a != 0) {
          const DevJoint& ja = m.joint[a];
          cidx[u] = ja.idx_v + cc;
          ncol = u + 1;
          if (++cc >= ja.nv) { a = ja.parent; cc = 0; }
        }
      }
      real A[4][6];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < ncol) Jc.template ldn<6>(6 * cidx[u], A[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < ncol) acc(cidx[u], dot6(A[u], F));
    }
  }
}

// getFrameJacobian(LOCAL_WORLD_ALIGNED) of one mapped output from the same cache — what the reference evaluates per
// output before every J dq / J^T f product (KinematicChainMapping.inl:434,444,480,490,579):
//   J_k[:, c] = [J_lin[:, c] - p_k x J_ang[:, c] ; J_ang[:, c]]   for every dof c on the path to the root, else 0,
// p_k = world position of the output.  `put(c, col6)` receives the non-zero columns (the caller zeroes the rest).
template <class JF, class OF, class PUT>
RBD_HD void frame_jacobian_cached(const DevModel& m, const DevTables& tb, const JF& Jc, const OF& Oc, int k,
                                  const PUT& put) {
  const int pj = tb.outk_parent[k];
  if (pj == 0) return;  // attached to the universe: empty support
  const real* pl = tb.outk_pl + 12 * k;
  real Rp[12];
  Oc.template ldn<12>(12 * (pj - 1), Rp);
  const real px = Rp[0] * pl[9] + Rp[1] * pl[10] + Rp[2] * pl[11] + Rp[9];
  const real py = Rp[3] * pl[9] + Rp[4] * pl[10] + Rp[5] * pl[11] + Rp[10];
  const real pz = Rp[6] * pl[9] + Rp[7] * pl[10] + Rp[8] * pl[11] + Rp[11];
  for (int a = pj; a != 0; a = m.joint[a].parent) {
    const DevJoint& ja = m.joint[a];
    for (int cc = 0; cc < ja.nv; ++cc) {
      const int c = ja.idx_v + cc;
      real A[6], o[6], tx, ty, tz;
      Jc.template ldn<6>(6 * c, A);
      RBD_CROSS(tx, ty, tz, px, py, pz, A[3], A[4], A[5]);
      o[0] = A[0] - tx; o[1] = A[1] - ty; o[2] = A[2] - tz;
      o[3] = A[3]; o[4] = A[4]; o[5] = A[5];
      put(c, o);
    }
  }
}

// One block row of the ASSEMBLED mapping Jacobian (what a real getJ() hands SOFA's direct solvers): the non-zero
// columns of output k's LOCAL_WORLD_ALIGNED frame Jacobian only — the dofs on the path to the root, in ascending dof
// order — 6 values [lin; ang] each, written back to back at `dst` (16-byte aligned).  Same arithmetic as
// frame_jacobian_cached; the dof at (joint a, component cc) lands at position (dofs of a's strict ancestors) + cc.
template <class JF, class OF>
RBD_HD void getJ_block_cached(const DevModel& m, const DevTables& tb, const JF& Jc, const OF& Oc, int k, real* dst) {
  const int pj = tb.outk_parent[k];
  if (pj == 0) return;  // attached to the universe: empty block
  int pos = 0;
  for (int a = pj; a != 0; a = m.joint[a].parent) pos += m.joint[a].nv;
  const real* pl = tb.outk_pl + 12 * k;
  real Rp[12];
  Oc.template ldn<12>(12 * (pj - 1), Rp);
  const real px = Rp[0] * pl[9] + Rp[1] * pl[10] + Rp[2] * pl[11] + Rp[9];
  const real py = Rp[3] * pl[9] + Rp[4] * pl[10] + Rp[5] * pl[11] + Rp[10];
  const real pz = Rp[6] * pl[9] + Rp[7] * pl[10] + Rp[8] * pl[11] + Rp[11];
  for (int a = pj; a != 0; a = m.joint[a].parent) {
    const DevJoint& ja = m.joint[a];
    pos -= ja.nv;
    for (int cc = 0; cc < ja.nv; ++cc) {
      real A[6], tx, ty, tz;
      Jc.template ldn<6>(6 * (ja.idx_v + cc), A);
      RBD_CROSS(tx, ty, tz, px, py, pz, A[3], A[4], A[5]);
      real* o = dst + 6 * (pos + cc);
      o[0] = A[0] - tx; o[1] = A[1] - ty; o[2] = A[2] - tz;
      o[3] = A[3]; o[4] = A[4]; o[5] = A[5];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// applyJT, ring variant with FORWARD projection (shallow trees: depth <= RBD_JT_MAXD, a free-flyer only as the
// root).  Instead of parking the subtree wrench F per depth level and projecting on the way back
// (tau_i = A_i . sum_k F_k), every joint k projects its own wrench on all its ancestors as it is visited,
//   tau_i += A_i . F_k   for every ancestor i of k (and k itself),
// with the running tau of the joint currently occupying depth level L in a REGISTER (the level loop is unrolled, so
// the index is a compile-time constant).  The sweep state shrinks to the ancestors' A columns (6 per level, tensor
// memory at compile-time offsets) + R|p of the branch joints: Talos 66 doubles instead of 138, which is what lets 12
// warps (three per scheduler) share an SM — the kernel is issue-latency-bound, so that is worth more than the extra
// 6-flop projections (sum of depths = 156 for Talos instead of 33).  Below a free-flyer root the total wrench is
// summed in registers and projected once.  `anc` (ring program): joint at level L of the chain that ends at joint i.
#define RBD_JT_MAXD 12
template <class QC, class Ring, class Sink>
RBD_HD void applyJT_instance_fwd(const DevModel& m, const real* opl, const int* anc, const QC& q, Ring& ring,
                                 const Sink& sink, const Stack& S, const TmStack& T) {
  const int nj = m.njoints;
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  real tacc[RBD_JT_MAXD];
#pragma unroll
  for (int L = 0; L < RBD_JT_MAXD; ++L) tacc[L] = 0.0;
  real Ft[6] = {0, 0, 0, 0, 0, 0};  // total wrench below the free-flyer root
  int ffroot = 0;                   // the free-flyer occupying level 0 (0 = none)
  int dprev = 0;                    // depth of the previous joint
  // the joints at levels [lo, hi) of the chain that ends at joint `last` are complete: hand their tau over
#define RBD_JT_FLUSH(lo, hi, last)                                                              \
  _Pragma("unroll") for (int L = 0; L < RBD_JT_MAXD; ++L)                                       \
    if (L >= (lo) && L < (hi)) {                                                                \
      if (L == 0 && ffroot != 0) {                                                              \
        const DevJoint& jf_ = m.joint[ffroot];                                                  \
        real Rp_[12], o6_[6];                                                                   \
        s_ldv<12>(S, jf_.offA, Rp_);                                                            \
        ff_rows(Rp_, Rp_ + 9, Ft, o6_);                                                         \
        _Pragma("unroll") for (int e = 0; e < 6; ++e) { sink.put(jf_.idx_v + e, o6_[e]); Ft[e] = 0.0; } \
        ffroot = 0;                                                                             \
      } else {                                                                                  \
        sink.put(m.joint[anc[(last) * RBD_JT_MAXD + L]].idx_v, tacc[L]);                        \
      }                                                                                         \
    }
  real q0n = 0.0, q1n = 0.0;
  RBD_Q_PREFETCH(1);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    const int d = jn.depth;
    if (dprev >= d) { RBD_JT_FLUSH(d - 1, dprev, i - 1) }
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
      } else {
        const DevJoint& jb = m.joint[par];
        real Rp[12];
        if (jb.type == JT_FF) s_ldv<12>(S, jb.offA, Rp);
        else T.template ldv<12>(jb.boff, Rp);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = Rp[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = Rp[9 + e];
      }
    }
    real aw[3];
    const real q0 = q0n, q1 = q1n;
    RBD_Q_PREFETCH(i + 1);
    fk_joint(jn, R, p, q0, q1, q, aw);
    real Fk[6] = {0, 0, 0, 0, 0, 0};
#define RBD_JT_W(j, c)                                                                  \
  real c[6];                                                                            \
  {                                                                                     \
    real w[6];                                                                          \
    ring.template ld<j>(w);                                                             \
    const real* pl = opl + 3 * (o + j);                                                 \
    const real px = R[0] * pl[0] + R[1] * pl[1] + R[2] * pl[2] + p[0];                  \
    const real py = R[3] * pl[0] + R[4] * pl[1] + R[5] * pl[2] + p[1];                  \
    const real pz = R[6] * pl[0] + R[7] * pl[1] + R[8] * pl[2] + p[2];                  \
    real tx, ty, tz;                                                                    \
    RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);                                \
    c[0] = w[0]; c[1] = w[1]; c[2] = w[2];                                              \
    c[3] = w[3] + tx; c[4] = w[4] + ty; c[5] = w[5] + tz;                               \
  }
    for (int o = jn.out_begin; o < jn.out_end;) {
      const int cnt = ring.acquire();
      if (cnt == 1) {
        RBD_JT_W(0, c0)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += c0[e];
      } else if (cnt == 2) {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += c0[e] + c1[e];
      } else if (cnt == 3) {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1) RBD_JT_W(2, c2)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += (c0[e] + c1[e]) + c2[e];
      } else {
        RBD_JT_W(0, c0) RBD_JT_W(1, c1) RBD_JT_W(2, c2) RBD_JT_W(3, c3)
#pragma unroll
        for (int e = 0; e < 6; ++e) Fk[e] += (c0[e] + c1[e]) + (c2[e] + c3[e]);
      }
      ring.release(cnt);
      o += cnt;
    }
#undef RBD_JT_W
    if (jn.type == JT_FF) {
      // root free-flyer (level 0): its R|p serves the children's branch returns and the final projection
#pragma unroll
      for (int e = 0; e < 9; ++e) S.st(jn.offA + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) S.st(jn.offA + 9 + e, p[e]);
      ffroot = i;
#pragma unroll
      for (int e = 0; e < 6; ++e) Ft[e] = Fk[e];
    } else {
      real A[6];
      subspace_1dof(jn.type, p, aw, A);
      const real own = dot6(A, Fk);
      // own level, then every ancestor level below it: two jump tables on the depth instead of a compare per level
      // (the second one falls through from level d-2 down to level 0)
#define RBD_JT_OWN(D_) case D_: tacc[D_ - 1] = own; break;
      switch (d) {
        RBD_JT_OWN(1) RBD_JT_OWN(2) RBD_JT_OWN(3) RBD_JT_OWN(4) RBD_JT_OWN(5) RBD_JT_OWN(6)
        RBD_JT_OWN(7) RBD_JT_OWN(8) RBD_JT_OWN(9) RBD_JT_OWN(10) RBD_JT_OWN(11) RBD_JT_OWN(12)
        default: break;
      }
#undef RBD_JT_OWN
#define RBD_JT_ANC(L_)                              \
  {                                                 \
    real Aa[6];                                     \
    T.template ldv<6>(6 * (L_), Aa);                \
    tacc[L_] += dot6(Aa, Fk);                       \
  }
      switch (d) {
        case 12: RBD_JT_ANC(10) [[fallthrough]];
        case 11: RBD_JT_ANC(9) [[fallthrough]];
        case 10: RBD_JT_ANC(8) [[fallthrough]];
        case 9: RBD_JT_ANC(7) [[fallthrough]];
        case 8: RBD_JT_ANC(6) [[fallthrough]];
        case 7: RBD_JT_ANC(5) [[fallthrough]];
        case 6: RBD_JT_ANC(4) [[fallthrough]];
        case 5: RBD_JT_ANC(3) [[fallthrough]];
        case 4: RBD_JT_ANC(2) [[fallthrough]];
        case 3: RBD_JT_ANC(1) [[fallthrough]];
        case 2:
          if (ffroot != 0) {
#pragma unroll
            for (int e = 0; e < 6; ++e) Ft[e] += Fk[e];
          } else RBD_JT_ANC(0)
          break;
        default: break;
      }
#undef RBD_JT_ANC
      if (jn.nchild > 0) T.template stv<6>(6 * (d - 1), A);
      if (jn.nchild >= 2) {
        real Rp[12];
#pragma unroll
        for (int e = 0; e < 9; ++e) Rp[e] = R[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) Rp[9 + e] = p[e];
        T.template stv<12>(jn.boff, Rp);
      }
    }
    dprev = d;
  }
  if (nj > 1) { RBD_JT_FLUSH(0, dprev, nj - 1) }
#undef RBD_JT_FLUSH
}

// applyDJT on the ring kernel: the forward-projection sweep of applyJT_instance_fwd carrying the body twists dV of the
// direction dq.  Level records [A | dA] (12 doubles) at compile-time tensor-memory offsets, branch records
// [R|p | dV] (18) and the free-flyer root's [R|p | v_b] (18) in shared memory; per joint k the own wrench F_k and its
// first-order change dF_k = [0 ; sum dp x f] are projected on every ancestor level:
//   dtau_L += dA_L . F_k + A_L . dF_k.
template <class QC, class DQ, class Ring, class Sink>
RBD_HD void applyDJT_instance_fwd(const DevModel& m, const real* opl, const int* anc, const QC& q, const DQ& dq,
                                  Ring& ring, const Sink& sink, const Stack& S, const TmStack& T, real kfac) {
  const int nj = m.njoints;
  real R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0}, dV[6] = {0, 0, 0, 0, 0, 0};
  real tacc[RBD_JT_MAXD];
#pragma unroll
  for (int L = 0; L < RBD_JT_MAXD; ++L) tacc[L] = 0.0;
  real Ft[6] = {0, 0, 0, 0, 0, 0}, dFt[3] = {0, 0, 0};  // totals below a free-flyer root
  int ffroot = 0, dprev = 0;
#define RBD_DJT_FLUSH(lo, hi, last)                                                             \
  _Pragma("unroll") for (int L = 0; L < RBD_JT_MAXD; ++L)                                       \
    if (L >= (lo) && L < (hi)) {                                                                \
      if (L == 0 && ffroot != 0) {                                                              \
        const DevJoint& jf_ = m.joint[ffroot];                                                  \
        real Rp_[18], t6_[6], d6_[6], dF6_[6] = {0, 0, 0, dFt[0], dFt[1], dFt[2]};              \
        s_ldv<18>(S, jf_.offA, Rp_);                                                            \
        ff_rows(Rp_, Rp_ + 9, Ft, t6_);                                                         \
        ff_rows(Rp_, Rp_ + 9, dF6_, d6_);                                                       \
        const real* vb_ = Rp_ + 12;                                                             \
        real ax_, ay_, az_, bx_, by_, bz_, cx_, cy_, cz_;                                       \
        RBD_CROSS(ax_, ay_, az_, vb_[3], vb_[4], vb_[5], t6_[0], t6_[1], t6_[2]);               \
        RBD_CROSS(bx_, by_, bz_, vb_[3], vb_[4], vb_[5], t6_[3], t6_[4], t6_[5]);               \
        RBD_CROSS(cx_, cy_, cz_, vb_[0], vb_[1], vb_[2], t6_[0], t6_[1], t6_[2]);               \
        sink.put(jf_.idx_v + 0, kfac * (d6_[0] - ax_)); sink.put(jf_.idx_v + 1, kfac * (d6_[1] - ay_)); \
        sink.put(jf_.idx_v + 2, kfac * (d6_[2] - az_));                                         \
        sink.put(jf_.idx_v + 3, kfac * (d6_[3] - bx_ - cx_)); sink.put(jf_.idx_v + 4, kfac * (d6_[4] - by_ - cy_)); \
        sink.put(jf_.idx_v + 5, kfac * (d6_[5] - bz_ - cz_));                                   \
        _Pragma("unroll") for (int e = 0; e < 6; ++e) Ft[e] = 0.0;                              \
        dFt[0] = dFt[1] = dFt[2] = 0.0;                                                         \
        ffroot = 0;                                                                             \
      } else {                                                                                  \
        sink.put(m.joint[anc[(last) * RBD_JT_MAXD + L]].idx_v, kfac * tacc[L]);                 \
      }                                                                                         \
    }
  real q0n = 0.0, q1n = 0.0, dn = 0.0;
  RBD_Q_PREFETCH(1);
  if (nj > 1 && m.joint[1].type != JT_FF) dn = dq.ld(m.joint[1].idx_v);
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    const int d = jn.depth;
    if (dprev >= d) { RBD_DJT_FLUSH(d - 1, dprev, i - 1) }
    if (par != i - 1) {
      if (par == 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        p[0] = p[1] = p[2] = 0;
#pragma unroll
        for (int e = 0; e < 6; ++e) dV[e] = 0;
      } else {
        const DevJoint& jb = m.joint[par];
        if (jb.type == JT_FF) {
          // free-flyer root: R|p from its record, its world twist recomputed from v_b
          real Rp[18];
          s_ldv<18>(S, jb.offA, Rp);
#pragma unroll
          for (int e = 0; e < 9; ++e) R[e] = Rp[e];
#pragma unroll
          for (int e = 0; e < 3; ++e) p[e] = Rp[9 + e];
          const real* vb = Rp + 12;
          const real wx = R[0] * vb[3] + R[1] * vb[4] + R[2] * vb[5];
          const real wy = R[3] * vb[3] + R[4] * vb[4] + R[5] * vb[5];
          const real wz = R[6] * vb[3] + R[7] * vb[4] + R[8] * vb[5];
          real tx, ty, tz;
          RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], wx, wy, wz);
          dV[0] = R[0] * vb[0] + R[1] * vb[1] + R[2] * vb[2] + tx;
          dV[1] = R[3] * vb[0] + R[4] * vb[1] + R[5] * vb[2] + ty;
          dV[2] = R[6] * vb[0] + R[7] * vb[1] + R[8] * vb[2] + tz;
          dV[3] = wx; dV[4] = wy; dV[5] = wz;
        } else {
#pragma unroll
          for (int e = 0; e < 9; ++e) R[e] = S.ld(jb.boff + e);
#pragma unroll
          for (int e = 0; e < 3; ++e) p[e] = S.ld(jb.boff + 9 + e);
#pragma unroll
          for (int e = 0; e < 6; ++e) dV[e] = S.ld(jb.boff + 12 + e);
        }
      }
    }
    real aw[3];
    const real q0 = q0n, q1 = q1n, dqi = dn;
    RBD_Q_PREFETCH(i + 1);
    if (i + 1 < nj && m.joint[i + 1].type != JT_FF) dn = dq.ld(m.joint[i + 1].idx_v);
    fk_joint(jn, R, p, q0, q1, q, aw);
    // A, dA of this joint and the twist of its body
    real A[6] = {0, 0, 0, 0, 0, 0}, dA[6] = {0, 0, 0, 0, 0, 0};
    if (jn.type != JT_FF) {
      subspace_1dof(jn.type, p, aw, A);
      RBD_CROSS(dA[0], dA[1], dA[2], dV[3], dV[4], dV[5], A[0], A[1], A[2]);
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, dV[0], dV[1], dV[2], A[3], A[4], A[5]);
      dA[0] += tx; dA[1] += ty; dA[2] += tz;
      RBD_CROSS(dA[3], dA[4], dA[5], dV[3], dV[4], dV[5], A[3], A[4], A[5]);
#pragma unroll
      for (int e = 0; e < 6; ++e) dV[e] += A[e] * dqi;
    } else {
      real vb[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) vb[e] = dq.ld(jn.idx_v + e);
      const real wx = R[0] * vb[3] + R[1] * vb[4] + R[2] * vb[5];
      const real wy = R[3] * vb[3] + R[4] * vb[4] + R[5] * vb[5];
      const real wz = R[6] * vb[3] + R[7] * vb[4] + R[8] * vb[5];
      real tx, ty, tz;
      RBD_CROSS(tx, ty, tz, p[0], p[1], p[2], wx, wy, wz);
      dV[0] += R[0] * vb[0] + R[1] * vb[1] + R[2] * vb[2] + tx;
      dV[1] += R[3] * vb[0] + R[4] * vb[1] + R[5] * vb[2] + ty;
      dV[2] += R[6] * vb[0] + R[7] * vb[1] + R[8] * vb[2] + tz;
      dV[3] += wx; dV[4] += wy; dV[5] += wz;
#pragma unroll
      for (int e = 0; e < 9; ++e) S.st(jn.offA + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) S.st(jn.offA + 9 + e, p[e]);
#pragma unroll
      for (int e = 0; e < 6; ++e) S.st(jn.offA + 12 + e, vb[e]);
    }
    // own wrench about the origin and its first-order change (angular part only)
    real Fk[6] = {0, 0, 0, 0, 0, 0}, dFk[3] = {0, 0, 0};
#define RBD_DJT_W(j)                                                                    \
  {                                                                                     \
    real w[6];                                                                          \
    ring.template ld<j>(w);                                                             \
    const real* pl = opl + 3 * (o + j);                                                 \
    const real px = R[0] * pl[0] + R[1] * pl[1] + R[2] * pl[2] + p[0];                  \
    const real py = R[3] * pl[0] + R[4] * pl[1] + R[5] * pl[2] + p[1];                  \
    const real pz = R[6] * pl[0] + R[7] * pl[1] + R[8] * pl[2] + p[2];                  \
    real tx, ty, tz, dx, dy, dz;                                                        \
    RBD_CROSS(tx, ty, tz, px, py, pz, w[0], w[1], w[2]);                                \
    Fk[0] += w[0]; Fk[1] += w[1]; Fk[2] += w[2];                                        \
    Fk[3] += w[3] + tx; Fk[4] += w[4] + ty; Fk[5] += w[5] + tz;                         \
    RBD_CROSS(dx, dy, dz, dV[3], dV[4], dV[5], px, py, pz);                             \
    dx += dV[0]; dy += dV[1]; dz += dV[2];                                              \
    RBD_CROSS(tx, ty, tz, dx, dy, dz, w[0], w[1], w[2]);                                \
    dFk[0] += tx; dFk[1] += ty; dFk[2] += tz;                                           \
  }
    for (int o = jn.out_begin; o < jn.out_end;) {
      const int cnt = ring.acquire();
      RBD_DJT_W(0)
      if (cnt > 1) RBD_DJT_W(1)
      if (cnt > 2) RBD_DJT_W(2)
      if (cnt > 3) RBD_DJT_W(3)
      ring.release(cnt);
      o += cnt;
    }
#undef RBD_DJT_W
    if (jn.type == JT_FF) {
      ffroot = i;
#pragma unroll
      for (int e = 0; e < 6; ++e) Ft[e] = Fk[e];
      dFt[0] = dFk[0]; dFt[1] = dFk[1]; dFt[2] = dFk[2];
    } else {
      const real own = dot6(dA, Fk) + (A[3] * dFk[0] + A[4] * dFk[1] + A[5] * dFk[2]);
#define RBD_DJT_OWN(D_) case D_: tacc[D_ - 1] = own; break;
      switch (d) {
        RBD_DJT_OWN(1) RBD_DJT_OWN(2) RBD_DJT_OWN(3) RBD_DJT_OWN(4) RBD_DJT_OWN(5) RBD_DJT_OWN(6)
        RBD_DJT_OWN(7) RBD_DJT_OWN(8) RBD_DJT_OWN(9) RBD_DJT_OWN(10) RBD_DJT_OWN(11) RBD_DJT_OWN(12)
        default: break;
      }
#undef RBD_DJT_OWN
#define RBD_DJT_ANC(L_)                                                                  \
  {                                                                                      \
    real rc_[12];                                                                        \
    T.template ldv<12>(12 * (L_), rc_);                                                  \
    tacc[L_] += dot6(rc_ + 6, Fk) + (rc_[3] * dFk[0] + rc_[4] * dFk[1] + rc_[5] * dFk[2]); \
  }
      switch (d) {
        case 12: RBD_DJT_ANC(10) [[fallthrough]];
        case 11: RBD_DJT_ANC(9) [[fallthrough]];
        case 10: RBD_DJT_ANC(8) [[fallthrough]];
        case 9: RBD_DJT_ANC(7) [[fallthrough]];
        case 8: RBD_DJT_ANC(6) [[fallthrough]];
        case 7: RBD_DJT_ANC(5) [[fallthrough]];
        case 6: RBD_DJT_ANC(4) [[fallthrough]];
        case 5: RBD_DJT_ANC(3) [[fallthrough]];
        case 4: RBD_DJT_ANC(2) [[fallthrough]];
        case 3: RBD_DJT_ANC(1) [[fallthrough]];
        case 2:
          if (ffroot != 0) {
#pragma unroll
            for (int e = 0; e < 6; ++e) Ft[e] += Fk[e];
            dFt[0] += dFk[0]; dFt[1] += dFk[1]; dFt[2] += dFk[2];
          } else RBD_DJT_ANC(0)
          break;
        default: break;
      }
#undef RBD_DJT_ANC
      if (jn.nchild > 0) {
        real rc[12];
#pragma unroll
        for (int e = 0; e < 6; ++e) { rc[e] = A[e]; rc[6 + e] = dA[e]; }
        T.template stv<12>(12 * (d - 1), rc);
      }
      if (jn.nchild >= 2) {
#pragma unroll
        for (int e = 0; e < 9; ++e) S.st(jn.boff + e, R[e]);
#pragma unroll
        for (int e = 0; e < 3; ++e) S.st(jn.boff + 9 + e, p[e]);
#pragma unroll
        for (int e = 0; e < 6; ++e) S.st(jn.boff + 12 + e, dV[e]);
      }
    }
    dprev = d;
  }
  if (nj > 1) { RBD_DJT_FLUSH(0, dprev, nj - 1) }
#undef RBD_DJT_FLUSH
}

// Ring program (host-built by plan_jt_ring, staged with the blob in place of the ColProg table; ints):
//   [0] pieces  [1] copies  [2] offset (ints, even) of the translation table  [3] offset of the ancestor table
//   (forward-projection variant: joint at level L of the chain ending at joint i, RBD_JT_MAXD ints per joint; 0 = none)
//   pieces  int4 each: first sorted index | outputs << 16, bytes to expect, first copy, copies
//   copies  int2 each: source offset within the tile's wrench block (bytes), offset within the stage | bytes << 16
//   opl     3 doubles per mapped output (sorted order): translation of its placement in the parent joint frame
// A piece is <= RBD_JT_GMAX outputs of one joint, fetched into one ring stage by one bulk copy per run of consecutive
// output indices (the frames of one joint usually are consecutive).
#define RBD_JT_GMAX 4
struct JtProg {
  const int* base;
  RBD_HD int npieces() const { return base[0]; }
  RBD_HD const int* piece(int i) const { return base + 4 + 4 * i; }
  RBD_HD const int* copy(int c) const { return base + 4 + 4 * base[0] + 2 * c; }
  RBD_HD const real* opl() const { return reinterpret_cast<const real*>(base + base[2]); }
  RBD_HD const int* anc() const { return base + base[3]; }
};
// reader of the input array with the ring's interface (tests/emul; the device ring lives in rbd_kernels.cu): follows
// the COPY LIST of every piece, i.e. reads exactly the bytes the device's copy engine is told to fetch
template <class FLD>
struct DirectRingT {
  FLD in;
  JtProg prog;
  int c;
  const int* cur;
  RBD_HD int acquire() {
    cur = prog.piece(c);
    c = c + 1 == prog.npieces() ? 0 : c + 1;
    return cur[0] >> 16;
  }
  template <int J>
  RBD_HD void ld(real* w) const {
    // locate slot J of the stage in the copy list
    for (int k = 0; k < cur[3]; ++k) {
      const int* ce = prog.copy(cur[2] + k);
      const int dst = ce[1] & 0xffff, bytes = ce[1] >> 16;
      if (J * 1536 >= dst && J * 1536 < dst + bytes) {
        const int field0 = (ce[0] + (J * 1536 - dst)) / 256;
#pragma unroll
        for (int e = 0; e < 6; ++e) w[e] = in.ld(field0 + e);
        return;
      }
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) w[e] = real(0) / real(0);  // not covered by any copy: poison
  }
  // the same with the slot index at run time
  RBD_HD void ldr(int j, real* w) const {
    for (int k = 0; k < cur[3]; ++k) {
      const int* ce = prog.copy(cur[2] + k);
      const int dst = ce[1] & 0xffff, bytes = ce[1] >> 16;
      if (j * 1536 >= dst && j * 1536 < dst + bytes) {
        const int field0 = (ce[0] + (j * 1536 - dst)) / 256;
#pragma unroll
        for (int e = 0; e < 6; ++e) w[e] = in.ld(field0 + e);
        return;
      }
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) w[e] = real(0) / real(0);  // not covered by any copy: poison
  }
  RBD_HD void release(int) {}
};
// =========================================================================================================
// Forward dynamics: the articulated-body algorithm (pinocchio::aba; SURVEY.md §8(f) rank 4 — not in the reference)
// =========================================================================================================
// World-frame formulation, like everything else here: with all spatial quantities expressed about the world origin the
// backward pass needs no transform between child and parent,
//   pass 1 (root -> leaves)  FK, A_i = oMi.act(S_i), v_i, c_i = v_i x vJ_i, I^A_i = Y_i, p^A_i = v_i x* (Y_i v_i)
//   pass 2 (leaves -> root)  U_i = I^A_i A_i, D_i = A_i . U_i, u_i = tau_i - A_i . p^A_i;
//                            I^a = I^A_i - U_i U_i^T / D_i,  p^a = p^A_i + I^a c_i + U_i u_i / D_i,
//                            I^A_parent += I^a,  p^A_parent += p^a                      (plain additions)
//   pass 3 (root -> leaves)  a' = a_parent + c_i (a_universe = -g),  qdd_i = (u_i - U_i . a') / D_i,  a_i = a' + A_i qdd_i
// A free-flyer is supported as the root joint (joint 1 on the universe; the host checks): its 6 x 6 system
// (X^T I^A X) qdd = tau - X^T (p^A + I^A a') is solved by a Cholesky factorisation in registers at the end of pass 2.
// Unlike the fused eval, pass 3 needs A, c, U, D, u of EVERY joint after the whole backward pass, so the state is
// O(njoints), not O(depth): it lives in a per-thread scratch record in global memory (tile-32 interleaved, one record
// per RESIDENT thread, reused from instance to instance so that it stays in L2), 45 doubles per joint:
//   [0,6) A (free-flyer root: [0,12) R|p)   [6,12) c   [12,33) I^A (upper triangle, 21) -> after pass 2: U [12,18),
//   1/D [18], u [19]   [33,39) p^A   [39,45) a_i (pass 3, joints with >= 2 children)
// plus 18 doubles R|p|v per joint with >= 2 children (the state pass 1 returns to when a new branch starts).
// (RBD_ABA_REC, RBD_ABA_BRANCH, aba_scratch_doubles: rbd_dev_model.h)

// index of (r, c) in the row-major upper triangle of a symmetric 6 x 6 matrix
RBD_HD constexpr int s6(int r, int c) { return r <= c ? 6 * r - r * (r - 1) / 2 + (c - r) : 6 * c - c * (c - 1) / 2 + (r - c); }
// y = S x
RBD_HD void sym6_mul(const real* S, const real* x, real* y) {
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    real t = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) t += S[s6(r, c)] * x[c];
    y[r] = t;
  }
}
// the rigid-body inertia Y = (m, h, I_O) as a symmetric 6 x 6 matrix [[m 1, -[h]x], [[h]x, I_O]] ([lin; ang])
RBD_HD void sym6_from_inertia(const real* Y, real* S) {
  const real m = Y[0], hx = Y[1], hy = Y[2], hz = Y[3];
  S[s6(0, 0)] = m; S[s6(0, 1)] = 0; S[s6(0, 2)] = 0; S[s6(0, 3)] = 0;   S[s6(0, 4)] = hz;  S[s6(0, 5)] = -hy;
  S[s6(1, 1)] = m; S[s6(1, 2)] = 0; S[s6(1, 3)] = -hz; S[s6(1, 4)] = 0; S[s6(1, 5)] = hx;
  S[s6(2, 2)] = m; S[s6(2, 3)] = hy; S[s6(2, 4)] = -hx; S[s6(2, 5)] = 0;
  S[s6(3, 3)] = Y[4]; S[s6(3, 4)] = Y[5]; S[s6(3, 5)] = Y[7];
  S[s6(4, 4)] = Y[6]; S[s6(4, 5)] = Y[8];
  S[s6(5, 5)] = Y[9];
}

// X: this thread's scratch record (ld(e) / st(e, v)); q, v, tau, ddq: strided fields of the instance.
template <class FLD, class SCR>
RBD_HD void aba_instance(const DevModel& m, const FLD& q, const FLD& v, const FLD& tau, const FLD& ddq, const SCR& X) {
  const int nj = m.njoints;
  const int bbase = RBD_ABA_REC * (nj - 1);
  real R[9], p[3], vel[6];
#define RBD_ABA_ROOT()                                                                          \
  do {                                                                                          \
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;   \
    p[0] = p[1] = p[2] = 0;                                                                     \
    _Pragma("unroll") for (int e_ = 0; e_ < 6; ++e_) vel[e_] = 0;                               \
  } while (0)
  RBD_ABA_ROOT();
  // ---- pass 1 ----------------------------------------------------------------------------------------------------
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    const int rec = RBD_ABA_REC * (i - 1);
    if (par != i - 1) {  // a new branch starts below `par`
      if (par == 0) {
        RBD_ABA_ROOT();
      } else {
        const int b = bbase + RBD_ABA_BRANCH * m.joint[par].bidx;
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = X.ld(b + e);
#pragma unroll
        for (int e = 0; e < 3; ++e) p[e] = X.ld(b + 9 + e);
#pragma unroll
        for (int e = 0; e < 6; ++e) vel[e] = X.ld(b + 12 + e);
      }
    }
    real q0 = 0, q1 = 0, aw[3];
    if (jn.type != JT_FF) {
      q0 = q.ld(jn.idx_q);
      if (is_unbounded(jn.type)) q1 = q.ld(jn.idx_q + 1);
    }
    fk_joint(jn, R, p, q0, q1, q, aw);
    real vj[6], cr[6];
    if (jn.type != JT_FF) {
      real A[6];
      subspace_1dof(jn.type, p, aw, A);
      const real qd = v.ld(jn.idx_v);
#pragma unroll
      for (int e = 0; e < 6; ++e) { X.st(rec + e, A[e]); vj[e] = A[e] * qd; }
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) X.st(rec + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) X.st(rec + 9 + e, p[e]);
      ff_act(R, p, v, jn.idx_v, vj);
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) vel[e] += vj[e];
    // c = v x vJ
    {
      real tx, ty, tz;
      RBD_CROSS(cr[0], cr[1], cr[2], vel[3], vel[4], vel[5], vj[0], vj[1], vj[2]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], vj[3], vj[4], vj[5]);
      cr[0] += tx; cr[1] += ty; cr[2] += tz;
      RBD_CROSS(cr[3], cr[4], cr[5], vel[3], vel[4], vel[5], vj[3], vj[4], vj[5]);
    }
    if (jn.type != JT_FF) {
#pragma unroll
      for (int e = 0; e < 6; ++e) X.st(rec + 6 + e, cr[e]);
    }  // (a root free-flyer has v = vJ, hence c = 0; its record keeps R|p in [0, 12))
    real Y[10], IA[21], h[6], pA[6];
    inertia_world(jn, R, p, Y);
    sym6_from_inertia(Y, IA);
#pragma unroll
    for (int e = 0; e < 21; ++e) X.st(rec + 12 + e, IA[e]);
    inertia_mul(Y, vel, h);
    {
      real tx, ty, tz;
      RBD_CROSS(pA[0], pA[1], pA[2], vel[3], vel[4], vel[5], h[0], h[1], h[2]);
      RBD_CROSS(pA[3], pA[4], pA[5], vel[3], vel[4], vel[5], h[3], h[4], h[5]);
      RBD_CROSS(tx, ty, tz, vel[0], vel[1], vel[2], h[0], h[1], h[2]);
      pA[3] += tx; pA[4] += ty; pA[5] += tz;
    }
#pragma unroll
    for (int e = 0; e < 6; ++e) X.st(rec + 33 + e, pA[e]);
    if (jn.nchild >= 2) {
      const int b = bbase + RBD_ABA_BRANCH * jn.bidx;
#pragma unroll
      for (int e = 0; e < 9; ++e) X.st(b + e, R[e]);
#pragma unroll
      for (int e = 0; e < 3; ++e) X.st(b + 9 + e, p[e]);
#pragma unroll
      for (int e = 0; e < 6; ++e) X.st(b + 12 + e, vel[e]);
    }
  }
  // ---- pass 2 ----------------------------------------------------------------------------------------------------
  real IA[21], pA[6];
  bool have = false;  // IA | pA hold what joint i + 1 hands to joint i (its parent)
  for (int i = nj - 1; i >= 1; --i) {
    const DevJoint& jn = m.joint[i];
    const int rec = RBD_ABA_REC * (i - 1);
    if (have) {
#pragma unroll
      for (int e = 0; e < 21; ++e) IA[e] += X.ld(rec + 12 + e);
#pragma unroll
      for (int e = 0; e < 6; ++e) pA[e] += X.ld(rec + 33 + e);
    } else {
#pragma unroll
      for (int e = 0; e < 21; ++e) IA[e] = X.ld(rec + 12 + e);
#pragma unroll
      for (int e = 0; e < 6; ++e) pA[e] = X.ld(rec + 33 + e);
    }
    have = false;
    if (jn.type != JT_FF) {
      real A[6], U[6], c[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) { A[e] = X.ld(rec + e); c[e] = X.ld(rec + 6 + e); }
      sym6_mul(IA, A, U);
      const real dinv = real(1) / dot6(A, U);
      const real u = tau.ld(jn.idx_v) - dot6(A, pA);
#pragma unroll
      for (int e = 0; e < 6; ++e) X.st(rec + 12 + e, U[e]);
      X.st(rec + 18, dinv);
      X.st(rec + 19, u);
      const int par = jn.parent;
      if (par != 0) {
        // I^a = I^A - U U^T / D ;  p^a = p^A + I^a c + U u / D
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int cc = r; cc < 6; ++cc) IA[s6(r, cc)] -= U[r] * (U[cc] * dinv);
        real Ic[6];
        sym6_mul(IA, c, Ic);
        const real ud = u * dinv;
#pragma unroll
        for (int e = 0; e < 6; ++e) pA[e] += Ic[e] + U[e] * ud;
        if (par == i - 1) {
          have = true;  // the parent is visited next: keep the contribution in registers
        } else {
          const int pr = RBD_ABA_REC * (par - 1);
#pragma unroll
          for (int e = 0; e < 21; ++e) X.st(pr + 12 + e, X.ld(pr + 12 + e) + IA[e]);
#pragma unroll
          for (int e = 0; e < 6; ++e) X.st(pr + 33 + e, X.ld(pr + 33 + e) + pA[e]);
        }
      }
    } else {
      // root free-flyer: (X^T I^A X) qdd = tau - X^T (p^A + I^A a'),  a' = -g  (c = 0)
      real Rp[12], a0[6], f[6], b[6];
#pragma unroll
      for (int e = 0; e < 12; ++e) Rp[e] = X.ld(rec + e);
      a0[0] = -m.gravity[0]; a0[1] = -m.gravity[1]; a0[2] = -m.gravity[2]; a0[3] = a0[4] = a0[5] = 0;
      sym6_mul(IA, a0, f);
#pragma unroll
      for (int e = 0; e < 6; ++e) f[e] += pA[e];
      ff_rows(Rp, Rp + 9, f, b);
#pragma unroll
      for (int e = 0; e < 6; ++e) b[e] = tau.ld(jn.idx_v + e) - b[e];
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
      // Cholesky D = L L^T (L stored in D's triangle as its transpose: D[s6(r, c)], r <= c, holds L(c, r)), then two
      // triangular solves
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
      for (int k = 0; k < 6; ++k) {  // L y = b
        real t = b[k];
#pragma unroll
        for (int j = 0; j < k; ++j) t -= D[s6(j, k)] * b[j];
        b[k] = t / D[s6(k, k)];
      }
#pragma unroll
      for (int k = 5; k >= 0; --k) {  // L^T x = y
        real t = b[k];
#pragma unroll
        for (int j = k + 1; j < 6; ++j) t -= D[s6(k, j)] * b[j];
        b[k] = t / D[s6(k, k)];
      }
      // a_1 = a' + X qdd
      real al[3], aa[3], tx, ty, tz;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        al[e] = Rp[3 * e] * b[0] + Rp[3 * e + 1] * b[1] + Rp[3 * e + 2] * b[2];
        aa[e] = Rp[3 * e] * b[3] + Rp[3 * e + 1] * b[4] + Rp[3 * e + 2] * b[5];
      }
      RBD_CROSS(tx, ty, tz, Rp[9], Rp[10], Rp[11], aa[0], aa[1], aa[2]);
      X.st(rec + 39, a0[0] + al[0] + tx); X.st(rec + 40, a0[1] + al[1] + ty); X.st(rec + 41, a0[2] + al[2] + tz);
      X.st(rec + 42, aa[0]); X.st(rec + 43, aa[1]); X.st(rec + 44, aa[2]);
#pragma unroll
      for (int e = 0; e < 6; ++e) ddq.st(jn.idx_v + e, b[e]);
    }
  }
  // ---- pass 3 ----------------------------------------------------------------------------------------------------
  real acc[6];
  for (int i = 1; i < nj; ++i) {
    const DevJoint& jn = m.joint[i];
    const int par = jn.parent;
    const int rec = RBD_ABA_REC * (i - 1);
    if (jn.type == JT_FF) {  // solved at the end of pass 2
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] = X.ld(rec + 39 + e);
      continue;
    }
    if (par == 0) {
      acc[0] = -m.gravity[0]; acc[1] = -m.gravity[1]; acc[2] = -m.gravity[2]; acc[3] = acc[4] = acc[5] = 0;
    } else if (par != i - 1) {  // (else: acc still holds the parent's acceleration)
      const int pr = RBD_ABA_REC * (par - 1);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[e] = X.ld(pr + 39 + e);
    }
    real A[6], U[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) { A[e] = X.ld(rec + e); U[e] = X.ld(rec + 12 + e); acc[e] += X.ld(rec + 6 + e); }
    const real qdd = (X.ld(rec + 19) - dot6(U, acc)) * X.ld(rec + 18);
    ddq.st(jn.idx_v, qdd);
#pragma unroll
    for (int e = 0; e < 6; ++e) acc[e] += A[e] * qdd;
    if (jn.nchild >= 2) {
#pragma unroll
      for (int e = 0; e < 6; ++e) X.st(rec + 39 + e, acc[e]);
    }
  }
#undef RBD_ABA_ROOT
}

#include "rbd_dyn_body.h"
#include "rbd_deriv_body.h"
#endif  /* !RBD_BODY_F32 && !RBD_BODY_EVAL_ONLY */
