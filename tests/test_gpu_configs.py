"""BASELINE.json configurations at their full sizes, checked through size-independent properties of the domain
(every instance, on the device) plus a strided sample against the oracle:
  configs[1]  Panda 7-dof, N = 65,536: FK + joint Jacobians + CRBA + RNEA
  configs[2]  Solo-12 free-flyer, N = 262,144, with the applyJT force mapping
(configs[3], Talos N = 1,048,576, is tests/test_gpu_parity.py::test_full_size_identities.)"""
import numpy as np
import pytest

from helpers import TOL_FP64, rel_err

pytestmark = pytest.mark.gpu


def _rand_state(torch, d, n, seed):
    from sofa.rigidbodydynamics_b200 import model as mdl
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    nt = n // 32
    q = (torch.rand((nt, d.nq, 32), generator=g, dtype=torch.float64, device="cuda") * 2 - 1) * np.pi
    if d.has_freeflyer_root:
        q[:, 0:3] = q[:, 0:3] / np.pi
        quat = torch.randn((nt, 4, 32), generator=g, dtype=torch.float64, device="cuda")
        q[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    v = torch.rand((nt, d.nv, 32), generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    a = torch.rand((nt, d.nv, 32), generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    return q, v, a


def _Ma(torch, Mp, a, nv):
    """M a from the packed upper triangle; Mp [nt, nv(nv+1)/2, 32], a [nt, nv, 32]."""
    out = torch.zeros_like(a)
    for c in range(nv):
        base = c * (c + 1) // 2
        out[:, c] += Mp[:, base + c] * a[:, c]
        if c:
            col = Mp[:, base:base + c]
            out[:, :c] += col * a[:, c:c + 1]
            out[:, c] += (col * a[:, :c]).sum(dim=1)
    return out


def test_config1_panda_65536():
    import torch
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200.dynamics import BatchedDynamics
    d = mdl.panda7(); n = 65536
    dyn = BatchedDynamics(d, n, tile=32)
    q, v, a = _rand_state(torch, d, n, 11)
    out = dyn.alloc_outputs()
    dyn.eval(q, v, a, out)
    tau0 = dyn.empty("tau"); dyn.rnea(q, v, torch.zeros_like(a), tau0)
    torch.cuda.synchronize()
    lhs = out["tau"] - tau0
    assert float((lhs - _Ma(torch, out["M"], a, d.nv)).abs().max()) <= 1e-10 * max(1.0, float(lhs.abs().max()))
    # oMi rotations are orthonormal, J columns of revolute joints have unit angular part
    R = out["oMi"].reshape(n // 32, d.njoints - 1, 12, 32)[:, :, :9].reshape(n // 32, d.njoints - 1, 3, 3, 32)
    RtR = torch.einsum("tjabi,tjaci->tjbci", R, R)
    eye = torch.eye(3, dtype=torch.float64, device="cuda")[None, None, :, :, None]
    assert float((RtR - eye).abs().max()) <= 1e-12
    Jang = out["J"].reshape(n // 32, d.nv, 6, 32)[:, :, 3:]
    assert float((Jang.norm(dim=2) - 1).abs().max()) <= 1e-12
    idx = np.arange(0, n, 4099)
    im = lambda x: dyn.to_instance_major(x)[idx].cpu().numpy()
    ref, _ = oracle.eval_batch(d, im(q), im(v), im(a), 0)
    for k in ("oMi", "J", "M", "tau"):
        assert rel_err(im(out[k]), ref[k]) <= TOL_FP64, k


def test_config2_solo_262144_applyJT():
    """apply -> applyJ -> applyJT on device tile-32 buffers for 262,144 Solo instances:
       adjointness  <applyJ(dq), w> == <dq, applyJT(w)>  for every instance (ties K2 to K3),
       gravity      applyJT(m_b g on the body-CoM outputs) == -rnea(q, 0, 0)  (ties the mapping to the dynamics path),
       accumulate   two applyJT calls add up (KCM.inl:498-511)."""
    import torch
    import oracle
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.solo12(); n = 262144; nt = n // 32
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    q, v, _ = _rand_state(torch, d, n, 12)
    g = torch.Generator(device="cuda"); g.manual_seed(13)
    w = torch.rand((nt, 6 * d.n_out, 32), generator=g, dtype=torch.float64, device="cuda") * 2 - 1
    qj, root = q[:, 7:].contiguous(), q[:, :7].contiguous()
    dq, rootv = v[:, 6:].contiguous(), v[:, :6].contiguous()
    pos = torch.empty((nt, 7 * d.n_out, 32), dtype=torch.float64, device="cuda")
    vel = torch.empty((nt, 6 * d.n_out, 32), dtype=torch.float64, device="cuda")
    tau = torch.zeros((nt, d.nv_joints, 32), dtype=torch.float64, device="cuda")
    rt = torch.zeros((nt, 6, 32), dtype=torch.float64, device="cuda")
    ws.apply_soa(n, 32, qj, root, pos)
    ws.applyJ_soa(n, 32, dq, rootv, vel)
    ws.applyJT_soa(n, 32, w, tau, rt)
    torch.cuda.synchronize()
    lhs = (vel * w).sum(dim=1)
    rhs = (dq * tau).sum(dim=1) + (rootv * rt).sum(dim=1)
    assert float((lhs - rhs).abs().max()) <= 1e-10 * max(1.0, float(lhs.abs().max()))
    # accumulate: a second identical call doubles the result
    tau1, rt1 = tau.clone(), rt.clone()
    ws.applyJT_soa(n, 32, w, tau, rt)
    torch.cuda.synchronize()
    assert float((tau - 2 * tau1).abs().max()) <= 1e-12 * max(1.0, float(tau1.abs().max()))
    assert float((rt - 2 * rt1).abs().max()) <= 1e-12 * max(1.0, float(rt1.abs().max()))
    # gravity through the mapping == - generalized gravity of RNEA
    wg = torch.zeros_like(w).reshape(nt, d.n_out, 6, 32)
    for b in range(d.njoints):
        for k in range(3):
            wg[:, b, k] = float(d.inertia[b, 0] * d.gravity[k])
    wg = wg.reshape(nt, 6 * d.n_out, 32).contiguous()
    tg = torch.zeros_like(tau); rg = torch.zeros_like(rt)
    ws.applyJT_soa(n, 32, wg, tg, rg)
    gq = torch.empty((nt, d.nv, 32), dtype=torch.float64, device="cuda")
    ws.rnea_soa(n, 32, q, torch.zeros_like(v), torch.zeros_like(v), gq)
    torch.cuda.synchronize()
    full = torch.cat([rg, tg], dim=1)
    assert float((full + gq).abs().max()) <= 1e-10 * max(1.0, float(gq.abs().max()))
    # strided sample of the poses against the oracle
    orc = oracle.Oracle(d)
    P = pos.permute(0, 2, 1).reshape(n, d.n_out, 7)
    Q = q.permute(0, 2, 1).reshape(n, d.nq)
    for i in range(0, n, 32771):
        qi = Q[i].cpu().numpy()
        ref = orc.apply(qi[7:], qi[:7])
        assert rel_err(P[i, :, :3].cpu().numpy(), ref[:, :3]) <= TOL_FP64


@pytest.mark.parametrize("name,n,cols", [("solo12", 262144, None), ("talos_reduced", 1 << 20, (0, 3, 5, 6, 11, 12, 19, 27, 37)),
                                         ("panda7", 65536, None)])
def test_mass_matrix_through_the_mapping(name, n, cols):
    """SURVEY.md §0 fact 3 — the reference's own route to the joint-space mass: one UniformMass<Rigid3> per body hung
    under the mapping (URDFModelLoader.cpp:369-391), so SOFA evaluates  M e_j = sum_b J_b^T Y_b J_b e_j  matrix-free
    through applyJ and applyJT.  Here, for EVERY instance of the BASELINE batch: column j of the CRBA kernel's M equals
    applyJT( diag(m_b I_3, R_b I_c,b R_b^T) applyJ(e_j) ) built from the mapping kernels on the body-CoM outputs, with
    the rotated body inertias taken from the eval kernel's oMi.  Ties K1/K2/K3 (reference call sites) to K5 (no
    reference call site) without any second implementation."""
    import torch
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.ROBOTS[name](); nt = n // 32
    nj, nv, no = d.njoints, d.nv, d.n_out
    ff = d.has_freeflyer_root
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    q, _v, _a = _rand_state(torch, d, n, 17)
    E = lambda K: torch.empty((nt, K, 32), dtype=torch.float64, device="cuda")
    Mp, oMi = E(m.info.mass_fields), E(m.info.omi_fields)
    ws.eval_soa(n, 32, q, None, None, oMi, None, Mp, None)
    qj, root = (q[:, 7:].contiguous(), q[:, :7].contiguous()) if ff else (q, None)
    pos = E(7 * no)
    ws.apply_soa(n, 32, qj, root, pos)
    # world inertia about the CoM of body b (joint b, b >= 1): R_b Ic_b R_b^T, R_b from the eval kernel's oMi
    R = oMi.reshape(nt, nj - 1, 12, 32)[:, :, :9].reshape(nt, nj - 1, 3, 3, 32)
    Ic = torch.zeros((nj - 1, 3, 3), dtype=torch.float64, device="cuda")
    for b in range(1, nj):
        xx, xy, yy, xz, yz, zz = (float(x) for x in d.inertia[b, 4:10])
        Ic[b - 1] = torch.tensor([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]], dtype=torch.float64)
    Iw = torch.einsum("tbijl,bjk,tbmkl->tbiml", R, Ic, R)
    mass = torch.tensor([float(d.inertia[b, 0]) for b in range(1, nj)], dtype=torch.float64, device="cuda")
    del R, oMi
    scale = max(1.0, float(Mp.abs().max()))
    worst = 0.0
    vel = E(6 * no)
    for j in (range(nv) if cols is None else cols):
        dq = torch.zeros((nt, nv, 32), dtype=torch.float64, device="cuda"); dq[:, j] = 1.0
        dqj, rootv = (dq[:, 6:].contiguous(), dq[:, :6].contiguous()) if ff else (dq, None)
        ws.applyJ_soa(n, 32, dqj, rootv, vel)
        V = vel.reshape(nt, no, 6, 32)
        w = torch.zeros((nt, no, 6, 32), dtype=torch.float64, device="cuda")
        w[:, 1:nj, :3] = V[:, 1:nj, :3] * mass[None, :, None, None]           # bodies = mapped outputs 1 .. nj-1 (their CoM frames)
        w[:, 1:nj, 3:] = torch.einsum("tbijl,tbjl->tbil", Iw, V[:, 1:nj, 3:])
        tau = torch.zeros((nt, d.nv_joints, 32), dtype=torch.float64, device="cuda")
        rt = torch.zeros((nt, 6, 32), dtype=torch.float64, device="cuda") if ff else None
        ws.applyJT_soa(n, 32, w.reshape(nt, 6 * no, 32), tau, rt)
        col = torch.cat([rt, tau], dim=1) if ff else tau                        # [nt, nv, 32]
        ref = torch.stack([Mp[:, j * (j + 1) // 2 + i] if i <= j else Mp[:, i * (i + 1) // 2 + j] for i in range(nv)], dim=1)
        worst = max(worst, float((col - ref).abs().max()) / scale)
        del w, V
    assert worst <= 1e-10, worst


@pytest.mark.parametrize("name", ["solo12", "talos_reduced", "random_2"])
def test_world_jacobian_by_finite_differences_gpu(name):
    """The free-flyer velocity frame / [lin; ang] ordering / column ownership of the GPU kernel's world joint Jacobian,
    pinned by central differences of its OWN joint placements along pinocchio's integrate (helpers.integrate)."""
    import torch
    import helpers
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = [x for x in helpers.all_models() if x.name == name][0]
    q, _v, _a = mdl.random_state(d, 5, seed=33)
    m = api.Model(d)

    def eval_fn(qb):
        N = qb.shape[0]
        ws = api.Workspace(m, N)
        qs = torch.from_numpy(helpers.to_tiled(qb, 32)).cuda()
        nt = qs.shape[0]
        oMi = torch.zeros((nt, m.info.omi_fields, 32), dtype=torch.float64, device="cuda")
        J = torch.zeros((nt, m.info.jac_fields, 32), dtype=torch.float64, device="cuda")
        ws.eval_soa(N, 32, qs, None, None, oMi, J, None, None)
        ws.synchronize()
        return helpers.from_tiled(oMi.cpu().numpy(), N), helpers.from_tiled(J.cpu().numpy(), N)

    assert helpers.fd_world_jacobian_error(eval_fn, d, q) <= 1e-7


def test_aba_round_trip_talos_full_size():
    """Talos at N = 1,048,576: aba(q, v, rnea(q, v, a)) == a for every instance (forward dynamics inverts the inverse
    dynamics; both kernels on the device), and a strided sample against the oracle's ABA."""
    import torch
    import oracle
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.talos_reduced(); n = 1 << 20; nt = n // 32
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    q, v, a = _rand_state(torch, d, n, 19)
    tau = torch.empty((nt, d.nv, 32), dtype=torch.float64, device="cuda")
    ddq = torch.empty((nt, d.nv, 32), dtype=torch.float64, device="cuda")
    ws.rnea_soa(n, 32, q, v, a, tau)
    ws.aba_soa(n, 32, q, v, tau, ddq)
    torch.cuda.synchronize()
    assert float((ddq - a).abs().max()) <= 1e-8                # |a| <= 1; conditioning of M(q) enters
    orc = oracle.Oracle(d)
    im = lambda x, i: x.permute(0, 2, 1).reshape(n, -1)[i].cpu().numpy()
    for i in range(0, n, 131071):
        ref = orc.aba(im(q, i), im(v, i), im(tau, i))
        assert rel_err(im(ddq, i), ref) <= 1e-9


def _integrate_ff_tree(torch, q, dq, eps):
    """q (+) eps dq on the device for a free-flyer-rooted tree of revolute / prismatic joints (tiled [nt, nq, 32] /
    [nt, nv, 32]): p += R (eps v_b), quat <- quat * exp(eps w_b), the rest q += eps dq (pinocchio::integrate)."""
    out = q.clone()
    out[:, 7:] = q[:, 7:] + eps * dq[:, 6:]
    x, y, z, w = q[:, 3], q[:, 4], q[:, 5], q[:, 6]
    vb = eps * dq[:, 0:3]
    out[:, 0] = q[:, 0] + (1 - 2 * (y * y + z * z)) * vb[:, 0] + 2 * (x * y - z * w) * vb[:, 1] + 2 * (x * z + y * w) * vb[:, 2]
    out[:, 1] = q[:, 1] + 2 * (x * y + z * w) * vb[:, 0] + (1 - 2 * (x * x + z * z)) * vb[:, 1] + 2 * (y * z - x * w) * vb[:, 2]
    out[:, 2] = q[:, 2] + 2 * (x * z - y * w) * vb[:, 0] + 2 * (y * z + x * w) * vb[:, 1] + (1 - 2 * (x * x + y * y)) * vb[:, 2]
    th = eps * dq[:, 3:6]
    ang = th.norm(dim=1)
    s = torch.sin(ang / 2) / ang
    dx, dy, dz, dw = s * th[:, 0], s * th[:, 1], s * th[:, 2], torch.cos(ang / 2)
    out[:, 3] = w * dx + x * dw + y * dz - z * dy
    out[:, 4] = w * dy - x * dz + y * dw + z * dx
    out[:, 5] = w * dz + x * dy - y * dx + z * dw
    out[:, 6] = w * dw - x * dx - y * dy - z * dz
    return out


def test_rnea_derivatives_talos_full_size():
    """Talos at N = 1,048,576 (24 GB of derivative matrices), every instance on the device: tau is QUADRATIC in v, so
    rnea(q, v + dv, a) - rnea(q, v - dv, a) == 2 dtau_dv dv holds exactly for a finite dv; tau is linear in a with
    dtau_da = M; and a central difference along q (+) eps dq reproduces dtau_dq dq.  Plus a strided sample against the
    oracle's restatement of pinocchio::computeRNEADerivatives."""
    import torch
    import oracle
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.talos_reduced(); n = 1 << 20; nt = n // 32; nv = d.nv
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    q, v, a = _rand_state(torch, d, n, 23)
    _q2, dq, dv = _rand_state(torch, d, n, 29)
    new = lambda K: torch.empty((nt, K, 32), dtype=torch.float64, device="cuda")
    DQ, DV, Mp = new(nv * nv), new(nv * nv), new(nv * (nv + 1) // 2)
    ws.rnea_derivatives_soa(n, 32, q, v, a, DQ, DV, Mp)
    torch.cuda.synchronize()
    matvec = lambda D, x: torch.einsum("trcl,tcl->trl", D.view(nt, nv, nv, 32), x)
    tp, tm = new(nv), new(nv)
    ws.rnea_soa(n, 32, q, v + dv, a, tp); ws.rnea_soa(n, 32, q, v - dv, a, tm)
    torch.cuda.synchronize()
    lin = 2 * matvec(DV, dv)
    assert float((tp - tm - lin).abs().max()) <= 1e-10 * max(1.0, float(lin.abs().max()))
    ws.rnea_soa(n, 32, q, v, a + dv, tp); ws.rnea_soa(n, 32, q, v, a, tm)
    torch.cuda.synchronize()
    lin = _Ma(torch, Mp, dv, nv)
    assert float((tp - tm - lin).abs().max()) <= 1e-10 * max(1.0, float(lin.abs().max()))
    eps = 1e-6
    ws.rnea_soa(n, 32, _integrate_ff_tree(torch, q, dq, eps), v, a, tp)
    ws.rnea_soa(n, 32, _integrate_ff_tree(torch, q, dq, -eps), v, a, tm)
    torch.cuda.synchronize()
    lin = matvec(DQ, dq)
    assert float(((tp - tm) / (2 * eps) - lin).abs().max()) <= 2e-6 * max(1.0, float(lin.abs().max()))
    orc = oracle.Oracle(d)
    im = lambda x, i: x.permute(0, 2, 1).reshape(n, -1)[i].cpu().numpy()
    for i in range(0, n, 131071):
        rq, rv, _ = orc.rnea_derivatives(im(q, i), im(v, i), im(a, i))
        assert rel_err(im(DQ, i).reshape(nv, nv), rq) <= TOL_FP64
        assert rel_err(im(DV, i).reshape(nv, nv), rv) <= TOL_FP64
