"""latency_table.py — what one propagation pass costs through the reference-facing HOST entry points as a function of
the number of robots N, next to N single-robot CPU passes.

    python experiments/latency_table.py [--robots panda7,talos_reduced] [--out gpurun_out/latency.json]

Per robot model and N in {1, 32, 1024, 32768}:
  host_contiguous   rbd_apply_host(ws, N, ...) on N robots stored back to back (the patched component's call; pageable
                    memory, as SOFA allocates its vectors) — and the same with pinned memory
  batch_scattered   rbd_batch_apply with N separately allocated per-robot vectors (gather + one launch + scatter)
  cpu_per_robot     the oracle's apply (pinocchio-equivalent restatement: computeJointJacobians +
                    updateFramePlacements + pose conversion) run N times on ONE thread, as N KinematicChainMapping
                    components do today — the stand-in for "one pinocchio call per robot"
and the crossover N above which the back end beats the CPU loop.  BASELINE configs[0] (single instance) is N = 1."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med_us(fn, reps):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robots", default="panda7,solo12,talos_reduced")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import oracle
    from sofa.rigidbodydynamics_b200 import _capi, api, model as mdl
    from sofa.rigidbodydynamics_b200.mapping import MappingBatch
    lib = _capi.load()
    OL = oracle.lib()
    res = {}
    for name in args.robots.split(","):
        d = mdl.ROBOTS[name]()
        ff = d.has_freeflyer_root
        m = api.Model(d)
        rows = {}
        for N in (1, 32, 1024, 32768):
            q, _v, _a = mdl.random_state(d, N, seed=3)
            qj = np.ascontiguousarray(q[:, 7:] if ff else q); root = np.ascontiguousarray(q[:, :7]) if ff else None
            out = np.zeros((N, d.n_out, 7))
            reps = 200 if N <= 1024 else 20
            ws = api.Workspace(m, N)
            p = lambda a: None if a is None else a.ctypes.data
            t_page = med_us(lambda: lib.rbd_apply_host(ws._h, N, p(qj), p(root), p(out)), reps)
            tq, tr, to = torch.from_numpy(qj).pin_memory(), (torch.from_numpy(root).pin_memory() if ff else None), torch.from_numpy(out).pin_memory()
            dp = lambda t: None if t is None else t.data_ptr()
            t_pin = med_us(lambda: lib.rbd_apply_host(ws._h, N, dp(tq), dp(tr), dp(to)), reps)
            # scattered robots: separate allocations per robot, bound once; one pass = the call that runs the launch
            batch = MappingBatch(d, N)
            in1 = [np.ascontiguousarray(qj[i]) for i in range(N)]
            rts = [np.ascontiguousarray(root[i]) if ff else None for i in range(N)]
            pos = [np.zeros((d.n_out, 7)) for _ in range(N)]
            for i, c in enumerate(batch.components):
                c.bind_apply(pos[i], in1[i], rts[i])
            ep = [0]

            def one_pass():
                ep[0] += 1
                lib.rbd_batch_apply(batch._h, 0, ep[0])
            t_batch = med_us(one_pass, reps)
            assert np.allclose(pos[N - 1], out[N - 1], atol=1e-12)
            # CPU: N single-robot passes on one thread
            orc = oracle.Oracle(d)
            oq = qj.copy(); oo = np.zeros((N, d.n_out, 7))
            t_cpu = med_us(lambda: OL.orc_apply_loop(orc._w, N, p(oq), p(root), p(oo)), max(3, reps // 10))
            assert np.allclose(oo[:, :, :3], out[:, :, :3], atol=1e-9)
            rows[N] = {"host_contiguous_pageable_us": t_page, "host_contiguous_pinned_us": t_pin, "batch_scattered_us": t_batch,
                       "cpu_per_robot_loop_us": t_cpu, "cpu_us_per_robot": t_cpu / N,
                       "gpu_us_per_robot_pageable": t_page / N}
            del ws, batch
        # crossover: smallest measured N where the pageable host call beats the CPU loop, and a linear-fit estimate
        ns = sorted(rows)
        cross = next((n for n in ns if rows[n]["host_contiguous_pageable_us"] < rows[n]["cpu_per_robot_loop_us"]), None)
        fixed = rows[1]["host_contiguous_pageable_us"]; per_cpu = rows[1024]["cpu_us_per_robot"]
        per_gpu = (rows[32768]["host_contiguous_pageable_us"] - fixed) / 32767
        est = fixed / max(1e-9, per_cpu - per_gpu)
        res[name] = {"n_out": d.n_out, "rows": rows, "crossover_measured_N": cross, "crossover_estimate_N": est,
                     "fixed_cost_us": fixed, "cpu_us_per_robot": per_cpu, "gpu_marginal_us_per_robot": per_gpu}
    line = json.dumps({"what": "apply (FK + frame placements + Rigid3 poses) per propagation pass, microseconds, median",
                       "host_threads_for_cpu": 1, "results": res})
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
