"""Derivatives of inverse dynamics (rbd_rnea_derivatives_soa) timing on one GPU for the three BASELINE robots at their
BASELINE batch sizes, L2 flushed between launches, with and without the optional M output, next to the HBM roofline of the
algorithmic bytes (q, v, a in; two nv x nv matrices [+ packed M] out).  python experiments/time_drnea.py [robot]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sofa.rigidbodydynamics_b200 import api, model as mdl  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    peak = 6544.3
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {}
    for name, n in (("panda7", 65536), ("solo12", 262144), ("talos_reduced", 1 << 20)):
        if only and name != only.split(":")[0]:
            continue
        d = mdl.ROBOTS[name]()
        if only and ":" in only:
            n = int(only.split(":")[1])
        m = api.Model(d); ws = api.Workspace(m, n)
        ws.set_stream(torch.cuda.current_stream().cuda_stream)
        nt = n // 32; nv = d.nv
        g = torch.Generator(device=dev); g.manual_seed(5)
        q = torch.rand((nt, d.nq, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        if d.has_freeflyer_root:
            qq = q[:, 3:7, :]; q[:, 3:7, :] = qq / qq.norm(dim=1, keepdim=True)
        v = torch.rand((nt, nv, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        a = torch.rand((nt, nv, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        mf = nv * (nv + 1) // 2
        DQ = torch.empty((nt, nv * nv, 32), dtype=torch.float64, device=dev)
        DV = torch.empty((nt, nv * nv, 32), dtype=torch.float64, device=dev)
        Mp = torch.empty((nt, mf, 32), dtype=torch.float64, device=dev)
        nnz = len(m.drnea_pattern()[1])
        res = {"support_entries": nnz, "dense_entries": nv * nv}
        for what, fn, fields in (("drnea", lambda: ws.rnea_derivatives_soa(n, 32, q, v, a, DQ, DV, None), d.nq + 2 * nv + 2 * nv * nv),
                                 ("drnea_sparse", lambda: ws.rnea_derivatives_soa_sparse(n, 32, q, v, a, DQ, DV), d.nq + 2 * nv + 2 * nnz),
                                 ("drnea+M", lambda: ws.rnea_derivatives_soa(n, 32, q, v, a, DQ, DV, Mp), d.nq + 2 * nv + 2 * nv * nv + mf)):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            res[what] = {"ms": ms, "kernel": api.last_kernel_name().split("(")[0], "bytes_per_instance": 8 * fields,
                         "GBps_algorithmic": 8 * fields * n / ms / 1e6, "frac_hbm": 8 * fields * n / ms / 1e6 / peak,
                         "M_per_s": n / ms / 1e3}
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
