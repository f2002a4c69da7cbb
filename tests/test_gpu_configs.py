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
