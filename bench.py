#!/usr/bin/env python
"""bench.py — dynamics evals/sec (FK + joint Jacobians + CRBA + RNEA) on B200, next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload talos|solo|panda]

One "step" = one pass of the hot path (the fused eval kernel behind ``rbd_eval_soa``) over one batch of synthetic
robot states.  Per-GPU batch is fixed (weak scaling): Talos-reduced at N = 1,048,576 instances per GPU
(BASELINE.json configs[3]; 8 GPUs = the 8M-instance configs[4]).  Instances are independent, so ranks shard them by
contiguous ranges and no collective runs on the data path (torch.distributed only for the barrier / max-over-ranks).

Prints ONE JSON line (rank 0).  Keys follow the driver contract plus ``roofline`` and ``cpu_baseline``:
  value     whole-job evals/s with inputs already resident in HBM (device tiled-SoA), CUDA-event timed
  e2e       the same metric through the reference-facing host entry point ``rbd_eval_host`` with pinned HOST buffers in
            the reference's instance-major layout: H2D of q,v,a and D2H of oMi,J,M,tau inside the timed region
  roofline  HBM roofline of the fused eval kernel: algorithmic bytes per launch / measured launch time vs the measured
            copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  oracle/ (C restatement of the pinocchio algorithms, "port") on all host threads, bounded sample

``--impl reference`` times that CPU restatement as the reference arm (the reference itself — SOFA + pinocchio — is not
installable in this image).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dynamics evals/sec (FK+Jac+CRBA+RNEA)"
UNIT = "evals/s"
WORKLOADS = {"talos": ("talos_reduced", 1 << 20), "solo": ("solo12", 1 << 18), "panda": ("panda7", 1 << 16)}
L2_BYTES = 126 << 20


def _traffic(robot: str, n: int, tile: int):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None
    when no capture exists for this exact workload."""
    p = os.path.join(ROOT, "profiles", "eval_traffic.json")
    try:
        with open(p) as f:
            rec = json.load(f).get(f"{robot}:{n}:tile{tile}")
        return (float(rec["dram_bytes_per_launch"]), rec["source"]) if rec else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    _BITS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
             0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
             0x100: "display_clock_setting"}

    def _sample(self):
        if self._nvml is not None:
            try:
                n = self._nvml
                self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                try:
                    r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self._BITS.items():
                    if r & bit:
                        self.reasons.add(name)
                return
            except Exception:
                pass
        try:
            import subprocess
            o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in o.strip().split(",")]
            self.samples.append(int(f[0])); self.max_mhz = int(f[1])
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.01)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------
def cpu_arm(desc, seconds: float, nthreads: int = 0):
    """Times the oracle (CPU restatement, instance-major, all outputs) on `nthreads` host threads for about
    `seconds` of wall time: repeated passes over a bounded batch (outputs <= ~3 GB so that it fits any host).
    Returns (evals/s, threads, sample description, instances timed, seconds timed)."""
    import ctypes as C
    import numpy as np
    import oracle
    from sofa.rigidbodydynamics_b200 import model as mdl
    from sofa.rigidbodydynamics_b200._capi import make_desc
    threads = nthreads if nthreads > 0 else oracle.lib().orc_max_threads()
    n0 = 512 * threads
    q, v, a = mdl.random_state(desc, n0, seed=0x5EED)
    oracle.eval_batch(desc, q, v, a, threads)                       # warm-up (page faults, thread start)
    t = time.perf_counter(); oracle.eval_batch(desc, q, v, a, threads); dt = max(time.perf_counter() - t, 1e-6)
    cap = max(n0, int(3e9 // desc.eval_bytes()))
    n = int(min(max(n0, n0 * seconds / dt), cap))
    n -= n % threads
    passes = max(1, int(round(seconds / (dt * n / n0))))
    q, v, a = mdl.random_state(desc, n, seed=0x5EED)
    cdesc, _keep = make_desc(desc)
    outs = [np.zeros((n, k)) for k in (12 * (desc.njoints - 1), 6 * desc.nv, desc.nv * (desc.nv + 1) // 2, desc.nv)]
    for o in outs:
        o.fill(0.0)                                                 # touch pages outside the timed region
    L = oracle.lib()
    used = threads
    t = time.perf_counter()
    for _ in range(passes):
        used = L.orc_eval_batch(C.addressof(cdesc), n, q.ctypes.data, v.ctypes.data, a.ctypes.data,
                                outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, outs[3].ctypes.data,
                                threads)
    dt = time.perf_counter() - t
    return (n * passes / dt, int(used),
            f"{passes} pass(es) over {n} {desc.name} instances, all outputs, {dt:.2f} s wall", n * passes, dt)


def run_reference(args, desc, n_cfg):
    """Reference arm: the CPU implementation of the path on the host cores (oracle port; see module docstring)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    threads = oracle.lib().orc_max_threads()
    per_step = []
    sample = ""
    n_s = 0
    for s in range(args.warmup + args.steps):
        val, used, sample, n_s, dt = cpu_arm(desc, 4.0 if s >= args.warmup else 1.0, threads)
        if s >= args.warmup:
            per_step.append((n_s, dt))
    tot_n = sum(x for x, _ in per_step); tot_t = sum(t for _, t in per_step)
    value = tot_n / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(per_step)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{desc.name} N={n_cfg} per GPU (bounded CPU sample per step: {sample})",
                       "robot": desc.name, "note": desc.note},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{len(per_step)} step(s), each {sample}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="talos", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="instances per GPU (default: the BASELINE size of the workload)")
    ap.add_argument("--tile", type=int, default=32, help="tiled-SoA tile (32 = AoSoA fast path, 0 = plain SoA, tile = n)")
    ap.add_argument("--block", type=int, default=0, help="eval kernel block size override (0 = automatic)")
    ap.add_argument("--no-tmem", action="store_true", help="keep the whole sweep state in shared memory (A/B knob)")
    ap.add_argument("--f32", action="store_true",
                    help="secondary: also time the optional FP32 mode (1e-4 tolerance) and report it under 'fp32_mode'")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()

    from sofa.rigidbodydynamics_b200 import model as mdl
    robot, n_default = WORKLOADS[args.workload]
    desc = mdl.ROBOTS[robot]()
    n = args.n or n_default
    if args.impl == "reference":
        return run_reference(args, desc, n)

    import numpy as np
    import torch
    import torch.distributed as dist
    if args.warmup < 3:
        args.warmup = 3                                              # timing rule: W >= 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner to stdout when
    # the communicator is created), so file descriptor 1 is pointed at stderr for the rest of the run and the JSON line
    # goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from sofa.rigidbodydynamics_b200.dynamics import BatchedDynamics
    dyn = BatchedDynamics(desc, n, device=local, tile=(args.tile or None))
    if args.block:
        dyn.ws.set_eval_block(args.block)
    if args.no_tmem:
        dyn.ws.set_eval_tmem(False)
    I = dyn.model.info
    nq, nv = I.nq, I.nv
    bytes_per_eval = int(I.eval_bytes)

    # synthetic states, generated on the device for the resident leg (seeded per rank: SURVEY.md §8(d))
    g = torch.Generator(device=dev); g.manual_seed(0x5EED + rank)
    q_im = (torch.rand((n, nq), generator=g, dtype=torch.float64, device=dev) * 2 - 1) * np.pi
    for i in range(1, desc.njoints):
        t, o = int(desc.jtype[i]), int(desc.idx_q[i])
        if t == mdl.J_FF:
            q_im[:, o:o + 3] = torch.rand((n, 3), generator=g, dtype=torch.float64, device=dev) * 2 - 1
            w = torch.randn((n, 4), generator=g, dtype=torch.float64, device=dev)
            q_im[:, o + 3:o + 7] = w / w.norm(dim=1, keepdim=True)
    v_im = torch.rand((n, nv), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    a_im = torch.rand((n, nv), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    q, v, a = dyn.to_soa(q_im), dyn.to_soa(v_im), dyn.to_soa(a_im)
    out = dyn.alloc_outputs()
    step_bytes = n * bytes_per_eval
    # timing rule: inputs larger than L2, or flush L2 between timed iterations (a write larger than L2)
    flush = None
    if step_bytes <= 2 * L2_BYTES:
        flush = torch.empty(3 * L2_BYTES // 8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        dyn.eval(q, v, a, out)

    # ---- device-resident leg ----------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = dyn.ws.launch_count
    with ClockSampler(local) as clk:
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for e0, e1 in evs:
            if flush is not None:
                flush.fill_(1.0)
            e0.record(); step(); e1.record()
        t1.record()
        barrier()
    launches = dyn.ws.launch_count - l0
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / max(1, len(evs))
    # with an L2 flush between steps the job time is the sum of the per-step times (the flush is not part of the path)
    total_ms = t0.elapsed_time(t1) if flush is None else kern_ms * len(evs)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = world * n * args.steps / (total_ms * 1e-3)

    # quick self-check of the timed outputs against the oracle (not timed): 8 strided instances
    check_err = None
    if rank == 0:
        import oracle
        idx = torch.arange(0, n, max(1, n // 8), device=dev)[:8]
        ref, _ = oracle.eval_batch(desc, q_im[idx].cpu().numpy(), v_im[idx].cpu().numpy(), a_im[idx].cpu().numpy(), 1)
        check_err = 0.0
        for k in ("oMi", "J", "M", "tau"):
            got = dyn.to_instance_major(out[k])[idx].cpu().numpy()
            check_err = max(check_err, float(np.max(np.abs(got - ref[k])) / max(1.0, float(np.max(np.abs(ref[k]))))))
        assert check_err <= 1e-10, f"bench outputs differ from the oracle: {check_err:.3e}"

    # ---- optional FP32 mode (secondary; the headline stays FP64) ------------------------------------------------
    fp32 = None
    if args.f32:
        q32, v32, a32 = (x.float() for x in (q, v, a))
        K32 = {"oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": nv}
        o32 = {k: torch.empty((dyn.ntiles, K32[k], dyn.tile), dtype=torch.float32, device=dev) for k in K32}

        def step32():
            dyn.ws.eval_soa_f32(n, dyn.tile, q32, v32, a32, o32["oMi"], o32["J"], o32["M"], o32["tau"])

        for _ in range(args.warmup):
            step32()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.steps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); step32(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms32 = sum(ts) / len(ts)
        err32 = float((o32["tau"].double() - out["tau"]).abs().max() / max(1.0, float(out["tau"].abs().max())))
        fp32 = {"ms_per_step": ms32, "evals_per_s": n / (ms32 * 1e-3), "bytes_per_eval": bytes_per_eval // 2,
                "hbm_frac": (n * (bytes_per_eval // 2) / (ms32 * 1e-3) / 1e9) / _peaks()[0],
                "max_rel_err_tau_vs_fp64": err32, "tolerance": 1e-4,
                "eval_storage": dyn.ws.eval_plan(True, True, f32=True)}
        del q32, v32, a32, o32

    # ---- end-to-end leg: host buffers through rbd_eval_host ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hq, hv, ha = (x.cpu().pin_memory() for x in (q_im, v_im, a_im))
        del q, v, a, out
        torch.cuda.empty_cache()
        K = {"oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": nv}
        ho = {k: torch.empty((n, K[k]), dtype=torch.float64, pin_memory=True) for k in K}
        h2d = 8 * n * (nq + 2 * nv); d2h = 8 * n * sum(K.values())

        def e2e_step():
            dyn.ws.eval_host(n, hq, hv, ha, ho["oMi"], ho["J"], ho["M"], ho["tau"])

        e2e_step()
        barrier()
        t = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * args.e2e_steps / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "api": "rbd_eval_host (pinned host buffers, instance-major; all of oMi,J,M,tau copied back)"}
        if rank == 0:
            import oracle
            ref, _ = oracle.eval_batch(desc, hq[:4].numpy(), hv[:4].numpy(), ha[:4].numpy(), 1)
            for k in K:
                err = float(np.max(np.abs(ho[k][:4].numpy() - ref[k])) / max(1.0, float(np.max(np.abs(ref[k])))))
                assert err <= 1e-10, f"e2e outputs differ from the oracle: {k} {err:.3e}"

    # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        val, used, sample, _, _ = cpu_arm(desc, args.cpu_seconds)
        cpu = {"value": val, "unit": UNIT, "cores": used, "kind": "port", "sample": sample}

    if rank == 0:
        # models whose joints are all RX / RY / RZ / free-flyer run the rbd::simple instantiation of the same sweep
        simple = all(int(t) in (mdl.J_RX, mdl.J_RY, mdl.J_RZ, mdl.J_FF) for t in desc.jtype[1:])
        kname = "eval_kernel_simple" if simple else "eval_kernel"
        peak, peak_src = _peaks()
        achieved = step_bytes / (kern_ms * 1e-3) / 1e9
        traffic, traffic_src = _traffic(desc.name, n, dyn.tile)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{desc.name} N={n} per GPU, fused FK+Jac+CRBA+RNEA, all outputs written",
                           "robot": desc.name, "instances_per_gpu": n, "layout": f"tiled SoA, tile={dyn.tile}",
                           "eval_storage": dyn.ws.eval_plan(True, True),
                           "l2": (f"inputs+outputs {step_bytes / 2**20:.0f} MiB per step > 126 MiB L2, no flush needed" if flush is None
                                  else f"inputs+outputs {step_bytes / 2**20:.0f} MiB per step: L2 flushed between timed steps "
                                       f"by a {3 * L2_BYTES / 2**20:.0f} MiB fill"),
                           "note": desc.note, "sharding": f"instances, contiguous ranges, {world} rank(s), no collective"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": step_bytes, "peak_source": peak_src,
                             "kernel": f"rbd::{kname}<true,true,{dyn.tile if dyn.tile == 32 else 0}>", "kernel_ms": kern_ms,
                             "bytes_per_eval": bytes_per_eval},
                "cpu_baseline": cpu, "oracle_check_rel_err": check_err}
        if fp32 is not None:
            line["fp32_mode"] = fp32
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
