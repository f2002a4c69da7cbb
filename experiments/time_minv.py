"""M^-1 (rbd_minv_soa) timing on one GPU for the three BASELINE robots at their BASELINE batch sizes, L2 flushed between
launches, next to CRBA (the matrix it inverts).  python experiments/time_minv.py [robot]"""
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
    out = {}
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
        mf = d.nv * (d.nv + 1) // 2
        Mi = torch.empty((nt, mf, 32), dtype=torch.float64, device=dev)
        res = {}
        for what, fn in (("minv", lambda: ws.minv_soa(n, 32, q, Mi)), ("crba", lambda: ws.crba_soa(n, 32, q, Mi))):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            res[what] = {"ms": ts[len(ts) // 2], "kernel": api.last_kernel_name().split("(")[0]}
        bytes_ = 8 * (d.nq + mf) * n
        res["minv_GBps_algorithmic"] = bytes_ / res["minv"]["ms"] / 1e6
        res["minv_M_per_s"] = n / res["minv"]["ms"] / 1e3
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
