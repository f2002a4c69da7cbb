#!/usr/bin/env python
"""bench_mapping.py — secondary measurement (NOT the driver contract; bench.py is): the Mapping path of BASELINE
configs[2], Solo-12 free-flyer N = 262,144: apply -> applyJ -> applyJT on device tiled-SoA buffers, one wrench per
mapped output; the constraint-row applyJT (two rows per instance, two non-zeros each); and the Mass / ForceField
products and the separate CRBA / RNEA entry points.  Prints one JSON line with the time and the algorithmic HBM
bandwidth of each kernel."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="solo12")
    ap.add_argument("--n", type=int, default=1 << 18)
    ap.add_argument("--tile", type=int, default=32)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import numpy as np
    import torch
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.ROBOTS[args.robot]()
    n, tile = args.n, args.tile
    nt = (n + tile - 1) // tile
    m = api.Model(d); ws = api.Workspace(m, n)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    I = m.info
    dev = "cuda"
    g = torch.Generator(device=dev); g.manual_seed(1)
    R = lambda K: torch.rand((nt, K, tile), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    ff = bool(I.has_freeflyer_root)
    qj = R(I.nq_joints) * np.pi
    root = None
    if ff:
        root = R(7); root[:, 3:7] /= root[:, 3:7].norm(dim=1, keepdim=True)
    dq = R(I.nv_joints); rootv = R(6) if ff else None
    w = R(6 * I.n_out)
    pos = torch.empty((nt, 7 * I.n_out, tile), dtype=torch.float64, device=dev)
    vel = torch.empty((nt, 6 * I.n_out, tile), dtype=torch.float64, device=dev)
    tau = torch.zeros((nt, I.nv_joints, tile), dtype=torch.float64, device=dev)
    rt = torch.zeros((nt, 6, tile), dtype=torch.float64, device=dev) if ff else None
    flush = torch.empty(3 * (126 << 20) // 8, dtype=torch.float64, device=dev)
    # constraint rows (KCM.inl:515-639): two rows per instance, two non-zeros per row on random outputs
    nrows, nnz = 2 * n, 4 * n
    row_ptr = torch.arange(0, nnz + 1, 2, dtype=torch.int32, device=dev)
    row_inst = (torch.arange(0, nrows, dtype=torch.int32, device=dev) // 2).contiguous()
    col = torch.randint(0, int(I.n_out), (nnz,), generator=g, dtype=torch.int32, device=dev)
    val = torch.rand((nnz, 6), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    rows = torch.empty((nrows, I.nv), dtype=torch.float64, device=dev)
    keep = torch.empty((nrows,), dtype=torch.int32, device=dev)
    # Mass / ForceField path: full configuration, velocity, direction
    q = R(I.nq) * np.pi
    if ff:
        q[:, 3:7] /= q[:, 3:7].norm(dim=1, keepdim=True)
    v = R(I.nv); dx = R(I.nv)
    res = torch.zeros((nt, I.nv, tile), dtype=torch.float64, device=dev)
    Mp = torch.empty((nt, I.mass_fields, tile), dtype=torch.float64, device=dev)
    calls = {"apply": lambda: ws.apply_soa(n, tile, qj, root, pos),
             "applyJ": lambda: ws.applyJ_soa(n, tile, dq, rootv, vel),
             "applyJT": lambda: ws.applyJT_soa(n, tile, w, tau, rt),
             "applyJT_rows": lambda: ws.applyJT_rows_dev(nrows, row_ptr, row_inst, col, val, rows, keep),
             "applyDJT": lambda: ws.applyDJT_soa(n, tile, dq, rootv, w, 1.0, tau, rt),
             "addMDx": lambda: ws.addMDx_soa(n, tile, q, dx, 1.0, res),
             "addForce": lambda: ws.addForce_soa(n, tile, q, v, res),
             "addForce_gravity": lambda: ws.addForce_soa(n, tile, q, None, res),
             "rnea": lambda: ws.rnea_soa(n, tile, q, v, dx, res),
             "crba": lambda: ws.crba_soa(n, tile, q, Mp)}
    # algorithmic bytes per instance: inputs read + outputs written (+ cached q for J / JT, += read-modify-write)
    B = {"apply": 8 * (I.nq + 7 * I.n_out + I.nq),            # q in, poses out, q cache out
         "applyJ": 8 * (I.nq + I.nv + 6 * I.n_out),
         "applyJT": 8 * (I.nq + 6 * I.n_out + 2 * I.nv),
         "applyJT_rows": 2 * (8 * I.nq + 2 * (4 + 48) + 8 * I.nv + 4 + 8),   # per instance: 2 rows x (q, 2 nnz, row, keep, ptr)
         "applyDJT": 8 * (I.nq + I.nv + 6 * I.n_out + 2 * I.nv),
         "addMDx": 8 * (I.nq + I.nv + 2 * I.nv),
         "addForce": 8 * (I.nq + I.nv + 2 * I.nv),
         "addForce_gravity": 8 * (I.nq + 2 * I.nv),
         "rnea": 8 * (I.nq + 2 * I.nv + I.nv),
         "crba": 8 * (I.nq + I.mass_fields)}
    out = {}
    for name, fn in calls.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.steps):
            flush.fill_(1.0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        out[name] = {"ms": ms, "instances_per_s": n / (ms * 1e-3), "bytes_per_instance": B[name],
                     "GBps": n * B[name] / (ms * 1e-3) / 1e9, "frac_of_6544": n * B[name] / (ms * 1e-3) / 1e9 / 6544.3}
    print(json.dumps({"robot": d.name, "n": n, "tile": tile, "n_out": int(I.n_out), "kernels": out,
                      "note": "L2 flushed between timed launches; " + d.note}))


if __name__ == "__main__":
    main()
