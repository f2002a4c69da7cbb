"""Forward dynamics A/B on one GPU: the fused sweep (aba2_kernel) against the first version (aba_kernel) for the three
BASELINE robots at their BASELINE batch sizes; L2 flushed between launches.  python experiments/time_aba.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sofa.rigidbodydynamics_b200 import api, model as mdl  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    only = sys.argv[1] if len(sys.argv) > 1 else None     # optional: one robot (ncu captures)
    for name, n in (("panda7", 65536), ("solo12", 262144), ("talos_reduced", 1 << 20)):
        if only and name != only:
            continue
        d = mdl.ROBOTS[name]()
        m = api.Model(d); ws = api.Workspace(m, n)
        ws.set_stream(torch.cuda.current_stream().cuda_stream)
        nt = n // 32
        g = torch.Generator(device=dev); g.manual_seed(5)
        q = torch.rand((nt, d.nq, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        if d.has_freeflyer_root:
            qq = q[:, 3:7, :]; q[:, 3:7, :] = qq / qq.norm(dim=1, keepdim=True)
        v = torch.rand((nt, d.nv, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        tau = torch.rand((nt, d.nv, 32), dtype=torch.float64, device=dev, generator=g) * 2 - 1
        res = {}
        ref = None
        for fused in (1, 0):
            ws.set_aba_fused(bool(fused))
            ddq = torch.empty((nt, d.nv, 32), dtype=torch.float64, device=dev)
            for _ in range(3):
                ws.aba_soa(n, 32, q, v, tau, ddq)
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ws.aba_soa(n, 32, q, v, tau, ddq); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res["fused" if fused else "v1"] = {"ms": ts[len(ts) // 2], "kernel": api.last_kernel_name().split("(")[0]}
            if ref is None:
                ref = ddq.clone()
            else:
                res["max_abs_diff"] = float((ddq - ref).abs().max())
        res["M_per_s_fused"] = n / res["fused"]["ms"] / 1e3
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
