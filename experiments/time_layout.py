"""Layout adapters (instance-major <-> tile-32) on one GPU, L2 flushed between launches: GB/s moved and fraction of the
measured HBM copy roofline for the field counts of the Talos outputs.  python experiments/time_layout.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sofa.rigidbodydynamics_b200 import api, model as mdl  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6544.3
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    m = api.Model(mdl.panda7()); ws = api.Workspace(m, 8)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 1 << 20
    out = {}
    for K in (741, 396, 369, 228, 128, 38):
        aos = torch.randn((n, K), dtype=torch.float64, device=dev)
        soa = torch.empty((n // 32, K, 32), dtype=torch.float64, device=dev)
        res = {}
        for name, fn in (("aos_to_soa", lambda: ws.aos_to_soa(n, K, 32, aos, soa)), ("soa_to_aos", lambda: ws.soa_to_aos(n, K, 32, soa, aos))):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            res[name] = {"ms": round(ms, 4), "GBps": round(16 * K * n / ms / 1e6, 1), "frac": round(16 * K * n / ms / 1e6 / peak, 3)}
        out[f"K={K}"] = res
        del aos, soa
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
