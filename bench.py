#!/usr/bin/env python
"""bench.py — dynamics evals/sec (FK + joint Jacobians + CRBA + RNEA) on B200, next to the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload talos|solo|panda]

One "step" = one pass of the hot path (the fused eval kernel behind ``rbd_eval_soa``) over one batch of synthetic
robot states.  Per-GPU batch is fixed (weak scaling): Talos-reduced at N = 1,048,576 instances per GPU
(BASELINE.json configs[3]; 8 GPUs = the 8M-instance configs[4]).  Instances are independent, so ranks shard them by
contiguous ranges and no collective runs on the data path (torch.distributed only for the barrier / max-over-ranks).

Prints ONE JSON line (rank 0).  Keys follow the driver contract plus ``roofline``, ``cpu_baseline`` and ``secondary``:
  value     whole-job evals/s with inputs already resident in HBM (device tiled-SoA), CUDA-event timed
  e2e       the same metric through the reference-facing host entry points with pinned HOST buffers in the reference's
            instance-major layout, H2D of the inputs and D2H of the outputs inside the timed region.  The headline
            ``e2e.value`` is ``rbd_eval_host_sparse`` with ALL outputs (oMi, J, M, tau; M in the support-only format:
            the structural zeros of the packed triangle carry no information and are not shipped);
            ``e2e.variants`` also times ``rbd_eval_host`` (packed M, round 1's figure) and the Mass / ForceField
            selection (M + tau only)
  roofline  HBM roofline of the fused eval kernel (algorithmic bytes per launch / measured launch time vs the measured
            copy bandwidth in MEASURED_PEAKS.json) and, under ``fp64``, the FP64 roofline (flops per eval counted by
            the ORC_COUNT_FLOPS build of the oracle vs the DFMA peak measured live by ``rbd_probe_fp64_peak``); the
            kernel name is read back from the launch (``rbd_last_kernel_name``)
  cpu_baseline  oracle/ (C restatement of the pinocchio algorithms, "port") on all host threads over the SAME N in
            reused 32768-instance output buffers, plus a single-thread figure on a bounded sample
  secondary the other BASELINE configs and the Mapping path, each with its own time, algorithmic bytes, HBM fraction,
            kernel name and clock sample: Panda N=65,536 and Solo N=262,144 fused eval; apply / applyJ / applyJT (and
            the constraint-row, applyDJT, Mass / ForceField, crba, rnea entry points) for Panda, Solo and Talos

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
CPU_CHUNK = 32768          # instances per reused output buffer of the CPU arm


def _profile_record(robot: str, n: int, tile: int):
    """Per-launch figures of the dominant kernel from the committed `ncu --set full` capture (profiles/), or {}."""
    p = os.path.join(ROOT, "profiles", "eval_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(f"{robot}:{n}:tile{tile}") or {}
    except (OSError, ValueError):
        return {}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int, period: float = 0.01):
        self.index = index
        self.period = period
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
            self._stop.wait(self.period)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._sample()                                             # at least one sample inside the region
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port over the SAME N as the GPU arm, outputs in reused CPU_CHUNK-instance buffers
class CpuArm:
    def __init__(self, desc, n: int, seed: int = 0x5EED):
        import ctypes as C
        import numpy as np
        import oracle
        from sofa.rigidbodydynamics_b200 import model as mdl
        from sofa.rigidbodydynamics_b200._capi import make_desc
        self.C, self.desc, self.n = C, desc, int(n)
        self.L = oracle.lib()
        self.max_threads = self.L.orc_max_threads()
        self.q, self.v, self.a = mdl.random_state(desc, self.n, seed=seed)
        self.cdesc, self._keep = make_desc(desc)
        self.chunk = min(CPU_CHUNK, self.n)
        K = (12 * (desc.njoints - 1), 6 * desc.nv, desc.nv * (desc.nv + 1) // 2, desc.nv)
        self.outs = [np.zeros((self.chunk, k)) for k in K]
        for o in self.outs:
            o.fill(0.0)                                             # touch the pages outside any timed region

    def run(self, n: int, threads: int) -> int:
        """One pass over the first n instances (all outputs) on `threads` threads; returns the threads used."""
        C, d = self.C, self.desc
        used = threads
        for i0 in range(0, n, self.chunk):
            cn = min(self.chunk, n - i0)
            used = self.L.orc_eval_batch(C.addressof(self.cdesc), cn, self.q[i0:].ctypes.data, self.v[i0:].ctypes.data,
                                         self.a[i0:].ctypes.data, self.outs[0].ctypes.data, self.outs[1].ctypes.data,
                                         self.outs[2].ctypes.data, self.outs[3].ctypes.data, threads)
        return int(used)

    def time_passes(self, seconds: float, threads: int, n: int = 0):
        """Whole passes over n instances (default: all N) for about `seconds`; returns (evals/s, threads, passes, s)."""
        n = n or self.n
        self.run(min(n, 4 * self.chunk), threads)                   # warm-up: thread start, caches
        t = time.perf_counter(); used = self.run(n, threads); first = time.perf_counter() - t
        passes, total = 1, first
        while total + first <= seconds:
            t = time.perf_counter(); self.run(n, threads); total += time.perf_counter() - t
            passes += 1
        return n * passes / total, used, passes, total


def cpu_baseline(desc, n: int, seconds: float):
    arm = CpuArm(desc, n)
    val, used, passes, dt = arm.time_passes(seconds, arm.max_threads)
    n1 = min(n, 2 * CPU_CHUNK)
    v1, _, p1, dt1 = arm.time_passes(min(seconds, 4.0), 1, n1)
    return {"value": val, "unit": UNIT, "cores": used, "kind": "port", "same_config": True,
            "sample": f"{passes} full pass(es) over the workload's {n} {desc.name} instances, all outputs written to "
                      f"reused {arm.chunk}-instance buffers, {dt:.2f} s wall",
            "single_thread": {"value": v1, "unit": UNIT, "cores": 1,
                              "sample": f"{p1} pass(es) over the first {n1} instances, {dt1:.2f} s wall"},
            "note": "pinocchio-3.3.1-equivalent CPU restatement (oracle/rbd_oracle.c: generic scalar C, FK recomputed by "
                    "computeJointJacobians, crba and rnea as in pinocchio), not pinocchio itself"}


def run_reference(args, desc, n_cfg):
    """Reference arm: the CPU implementation of the path on the host cores (oracle port; see module docstring).
    One step = one full pass over the workload's N instances on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(desc, n_cfg)
    threads = arm.max_threads
    per_step = []
    used = threads
    for s in range(args.warmup + args.steps):
        t = time.perf_counter()
        used = arm.run(n_cfg if s >= args.warmup else min(n_cfg, 4 * arm.chunk), threads)
        dt = time.perf_counter() - t
        if s >= args.warmup:
            per_step.append(dt)
    tot_t = sum(per_step)
    value = n_cfg * len(per_step) / tot_t
    sample = (f"{len(per_step)} step(s), each one full pass over {n_cfg} {desc.name} instances, all outputs written to "
              f"reused {arm.chunk}-instance buffers")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(per_step)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{desc.name} N={n_cfg}, FK+Jac+CRBA+RNEA, all outputs written (CPU restatement, "
                                   f"{used} host threads)", "robot": desc.name, "instances": n_cfg, "note": desc.note},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "same_config": True,
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU measurement helpers
def _time_launches(torch, fn, steps, warmup, flush, dev_index):
    """Median CUDA-event time (ms) of `fn` (one entry-point call) over `steps` launches after `warmup`, L2 flushed
    between launches when `flush` is given; clocks sampled during the timed loop; kernel name read back."""
    import numpy as np
    from sofa.rigidbodydynamics_b200 import api
    for _ in range(max(3, warmup)):
        fn()
    kernel = api.last_kernel_name()
    torch.cuda.synchronize()
    ts = []
    with ClockSampler(dev_index, period=0.002) as clk:
        for _ in range(steps):
            if flush is not None:
                flush.fill_(1.0)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), clk.summary(), kernel


def _entry(n, ms, bytes_per_instance, peak, clocks, kernel):
    gbps = n * bytes_per_instance / (ms * 1e-3) / 1e9
    return {"ms": ms, "instances_per_s": n / (ms * 1e-3), "bytes_per_instance": bytes_per_instance,
            "algorithmic_bytes_per_launch": n * bytes_per_instance, "achieved_GBps": gbps, "frac": gbps / peak,
            "kernel": kernel, "clocks": clocks}


def measure_config(torch, robot: str, n: int, steps: int, dev_index: int, with_eval: bool = True, tile: int = 32):
    """Secondary measurements of one (robot, N): the fused eval (unless it is the headline) and the Mapping / Mass /
    ForceField entry points on device tiled-SoA buffers, L2 flushed between launches."""
    import numpy as np
    from sofa.rigidbodydynamics_b200 import api, model as mdl
    d = mdl.ROBOTS[robot]()
    dev = torch.device("cuda", dev_index)
    peak, _ = _peaks()
    nt = (n + tile - 1) // tile
    m = api.Model(d); ws = api.Workspace(m, n, dev_index)
    ws.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    I = m.info
    g = torch.Generator(device=dev); g.manual_seed(1)
    R = lambda K: torch.rand((nt, K, tile), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    E = lambda K: torch.empty((nt, K, tile), dtype=torch.float64, device=dev)
    ff = bool(I.has_freeflyer_root)
    flush = torch.empty(3 * L2_BYTES // 8, dtype=torch.float64, device=dev)
    q = R(I.nq) * np.pi
    if ff:
        q[:, 0:3] /= np.pi
        q[:, 3:7] /= q[:, 3:7].norm(dim=1, keepdim=True)
    v = R(I.nv); dx = R(I.nv)
    out = {}
    if with_eval:
        oMi, J, M, tau = E(I.omi_fields), E(I.jac_fields), E(I.mass_fields), E(I.nv)
        ms, clk, kern = _time_launches(torch, lambda: ws.eval_soa(n, tile, q, v, dx, oMi, J, M, tau), steps, 3, flush, dev_index)
        out["eval"] = _entry(n, ms, int(I.eval_bytes), peak, clk, kern)
        out["eval"]["evals_per_s"] = out["eval"]["instances_per_s"]
        Mv = E(I.mass_support_fields)
        ms, clk, kern = _time_launches(torch, lambda: ws.eval_soa_sparse(n, tile, q, v, dx, oMi, J, Mv, tau), steps, 3, flush, dev_index)
        out["eval_sparseM"] = _entry(n, ms, int(I.eval_bytes) - 8 * (I.mass_fields - I.mass_support_fields), peak, clk, kern)
        del oMi, J, M, tau, Mv
    qj = q[:, 7:].contiguous() if ff else q
    root = q[:, :7].contiguous() if ff else None
    dq = R(I.nv_joints); rootv = R(6) if ff else None
    w = R(6 * I.n_out)
    pos, vel = E(7 * I.n_out), E(6 * I.n_out)
    tauj = torch.zeros((nt, I.nv_joints, tile), dtype=torch.float64, device=dev)
    rt = torch.zeros((nt, 6, tile), dtype=torch.float64, device=dev) if ff else None
    # constraint rows (KCM.inl:515-639): two rows per instance, two non-zeros per row on random outputs
    nrows, nnz = 2 * n, 4 * n
    row_ptr = torch.arange(0, nnz + 1, 2, dtype=torch.int32, device=dev)
    row_inst = (torch.arange(0, nrows, dtype=torch.int32, device=dev) // 2).contiguous()
    col = torch.randint(0, int(I.n_out), (nnz,), generator=g, dtype=torch.int32, device=dev)
    val = torch.rand((nnz, 6), generator=g, dtype=torch.float64, device=dev) * 2 - 1
    rows = torch.empty((nrows, I.nv), dtype=torch.float64, device=dev)
    keep = torch.empty((nrows,), dtype=torch.int32, device=dev)
    res = torch.zeros((nt, I.nv, tile), dtype=torch.float64, device=dev)
    Mp = E(I.mass_fields)
    calls = {"apply": lambda: ws.apply_soa(n, tile, qj, root, pos),
             "applyJ": lambda: ws.applyJ_soa(n, tile, dq, rootv, vel),
             "applyJT": lambda: ws.applyJT_soa(n, tile, w, tauj, rt),
             "applyJT_rows": lambda: ws.applyJT_rows_dev(nrows, row_ptr, row_inst, col, val, rows, keep),
             "applyDJT": lambda: ws.applyDJT_soa(n, tile, dq, rootv, w, 1.0, tauj, rt),
             "addMDx": lambda: ws.addMDx_soa(n, tile, q, dx, 1.0, res),
             "addForce": lambda: ws.addForce_soa(n, tile, q, v, res),
             "rnea": lambda: ws.rnea_soa(n, tile, q, v, dx, res),
             "crba": lambda: ws.crba_soa(n, tile, q, Mp),
             "aba": lambda: ws.aba_soa(n, tile, q, v, dx, res),
             "minv": lambda: ws.minv_soa(n, tile, q, Mp)}
    if 16 * I.nv * I.nv * n <= 40 << 30:   # both derivative matrices resident (Talos 2^20: 24 GB)
        DQm, DVm = E(I.nv * I.nv), E(I.nv * I.nv)
        calls["rnea_derivatives"] = lambda: ws.rnea_derivatives_soa(n, tile, q, v, dx, DQm, DVm, None)
        calls["rnea_derivatives_sparse"] = lambda: ws.rnea_derivatives_soa_sparse(n, tile, q, v, dx, DQm, DVm)
    # algorithmic bytes per instance: inputs read + outputs written (+ cached q for J / JT, += read-modify-write)
    B = {"apply": 8 * (I.nq + 7 * I.n_out + I.nq),            # q in, poses out, q cache out
         "applyJ": 8 * (I.nq + I.nv + 6 * I.n_out),
         "applyJT": 8 * (I.nq + 6 * I.n_out + 2 * I.nv),
         "applyJT_rows": 2 * (8 * I.nq + 2 * (4 + 48) + 8 * I.nv + 4 + 8),   # per instance: 2 rows x (q, 2 nnz, row, keep, ptr)
         "applyDJT": 8 * (I.nq + I.nv + 6 * I.n_out + 2 * I.nv),
         "addMDx": 8 * (I.nq + I.nv + 2 * I.nv),
         "addForce": 8 * (I.nq + I.nv + 2 * I.nv),
         "rnea": 8 * (I.nq + 2 * I.nv + I.nv),
         "crba": 8 * (I.nq + I.mass_fields),
         "aba": 8 * (I.nq + 3 * I.nv),           # q, v, tau in, ddq out (the per-joint scratch records are not algorithmic)
         "minv": 8 * (I.nq + I.mass_fields),     # q in, packed upper triangle of M^-1 out
         "rnea_derivatives": 8 * (I.nq + 2 * I.nv + 2 * I.nv * I.nv),   # q, v, a in, dtau_dq and dtau_dv (dense, zeros included) out
         "rnea_derivatives_sparse": 8 * (I.nq + 2 * I.nv + 2 * len(m.drnea_pattern()[1]))}   # ... structurally non-zero entries only
    for name, fn in calls.items():
        try:
            ms, clk, kern = _time_launches(torch, fn, steps, 3, flush, dev_index)
            out[name] = _entry(n, ms, B[name], peak, clk, kern)
        except api.RbdError as ex:   # an entry point that refuses this model / batch: recorded, the others still run
            out[name] = {"error": str(ex)}
    # layout adapters of the host path (Conversions.h at batch scale): instance-major <-> tile-32, widest array (M)
    Kw = int(I.mass_fields)
    aos = torch.empty((n, Kw), dtype=torch.float64, device=dev)
    for name, fn in (("layout_soa_to_aos_M", lambda: ws.soa_to_aos(n, Kw, tile, Mp, aos)),
                     ("layout_aos_to_soa_M", lambda: ws.aos_to_soa(n, Kw, tile, aos, Mp))):
        ms, clk, _ = _time_launches(torch, fn, steps, 3, flush, dev_index)
        out[name] = _entry(n, ms, 16 * Kw, peak, clk, "rbd::transpose32_kernel" if tile == 32 else "rbd::soa_to_aos / aos_to_soa kernel")
    # FP64 roofline next to the HBM one for the entries the flop-counting oracle build covers (the separate CRBA / RNEA
    # entry points are compute-bound: a fraction of the HBM roofline says little about them)
    try:
        import oracle
        fc = oracle.flop_counts(d)
        peak_tf = api.probe_fp64_peak(dev_index)
        for name, part in (("eval", "eval"), ("eval_sparseM", "eval"), ("crba", "crba"), ("rnea", "rnea")):
            if name in out and peak_tf > 0:
                ach = fc[part]["flops"] * out[name]["instances_per_s"] / 1e12
                out[name]["fp64"] = {"flops_per_instance": fc[part]["flops"], "achieved_tflops": ach, "peak_tflops": peak_tf,
                                     "frac": ach / peak_tf,
                                     "source": "oracle -DORC_COUNT_FLOPS count of the restated reference algorithm (its own FK "
                                               "pass included) x measured rate / rbd_probe_fp64_peak"}
                out[name]["binding"] = "fp64" if ach / peak_tf > out[name]["frac"] else "hbm"
    except Exception as e:  # the annotation must never cost the measurement
        out["fp64_annotation_error"] = str(e)
    return {"robot": d.name, "n": n, "tile": tile, "n_out": int(I.n_out), "l2": "flushed between timed launches",
            "note": d.note, "kernels": out}


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
    ap.add_argument("--no-arm", action="store_true", help="serial chains: the generic chain sweep instead of the unrolled arm kernel (A/B knob)")
    ap.add_argument("--f32", action="store_true",
                    help="secondary: also time the optional FP32 mode (1e-4 tolerance) and report it under 'fp32_mode'")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--secondary-steps", type=int, default=30)
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

    from sofa.rigidbodydynamics_b200 import api
    from sofa.rigidbodydynamics_b200.dynamics import BatchedDynamics
    dyn = BatchedDynamics(desc, n, device=local, tile=(args.tile or None))
    if args.block:
        dyn.ws.set_eval_block(args.block)
    if args.no_tmem:
        dyn.ws.set_eval_tmem(False)
    if args.no_arm:
        dyn.ws.set_eval_arm(False)
    I = dyn.model.info
    nq, nv = I.nq, I.nv
    bytes_per_eval = int(I.eval_bytes)
    eval_storage = dyn.ws.eval_plan(True, True)

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
    kernel_name = api.last_kernel_name()                             # the instantiation the launch went to
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
                "max_rel_err_tau_vs_fp64": err32, "tolerance": 1e-4, "kernel": api.last_kernel_name(),
                "eval_storage": dyn.ws.eval_plan(True, True, f32=True)}
        del q32, v32, a32, o32

    # ---- end-to-end leg: host buffers through the reference-facing host entry points ----------------------------
    e2e = None
    if not args.no_e2e:
        hq, hv, ha = (x.cpu().pin_memory() for x in (q_im, v_im, a_im))
        del q, v, a, out
        torch.cuda.empty_cache()
        K = {"oMi": I.omi_fields, "J": I.jac_fields, "M": I.mass_fields, "tau": nv}
        ho = {k: torch.empty((n, K[k]), dtype=torch.float64, pin_memory=True) for k in K}
        hMv = torch.empty((n, I.mass_support_fields), dtype=torch.float64, pin_memory=True)
        in_all = 8 * n * (nq + 2 * nv)
        variants = {
            # name: (call, h2d bytes, d2h bytes, api description)
            "sparse_all": (lambda: dyn.ws.eval_host_sparse(n, hq, hv, ha, ho["oMi"], ho["J"], hMv, ho["tau"]),
                           in_all, 8 * n * (K["oMi"] + K["J"] + I.mass_support_fields + nv),
                           "rbd_eval_host_sparse: all of oMi, J, M (support-only format), tau copied back"),
            "packed_all": (lambda: dyn.ws.eval_host(n, hq, hv, ha, ho["oMi"], ho["J"], ho["M"], ho["tau"]),
                           in_all, 8 * n * sum(K.values()),
                           "rbd_eval_host: all of oMi, J, M (packed upper triangle incl. structural zeros), tau"),
            "sparse_M_tau": (lambda: dyn.ws.eval_host_sparse(n, hq, hv, ha, None, None, hMv, ho["tau"]),
                             in_all, 8 * n * (I.mass_support_fields + nv),
                             "rbd_eval_host_sparse: M + tau only (what a Mass / ForceField component consumes)"),
        }
        res = {}
        for name, (fn, h2d, d2h, what) in variants.items():
            fn()
            barrier()
            t = time.perf_counter()
            for _ in range(args.e2e_steps):
                fn()
            barrier()
            dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            sec = float(dt.item())
            res[name] = {"value": world * n * args.e2e_steps / sec, "unit": UNIT, "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "api": what,
                         "d2h_GBps_per_gpu": d2h * args.e2e_steps / sec / 1e9}
        head = res["sparse_all"]
        e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
               "d2h_bytes_per_step": head["d2h_bytes_per_step"], "steps": args.e2e_steps,
               "api": head["api"] + " (pinned host buffers, instance-major)", "variants": res}
        if rank == 0:
            import oracle
            ref, _ = oracle.eval_batch(desc, hq[:4].numpy(), hv[:4].numpy(), ha[:4].numpy(), 1)
            packed = dyn.model.mass_packed_index()
            for k in K:
                err = float(np.max(np.abs(ho[k][:4].numpy() - ref[k])) / max(1.0, float(np.max(np.abs(ref[k])))))
                assert err <= 1e-10, f"e2e outputs differ from the oracle: {k} {err:.3e}"
            err = float(np.max(np.abs(hMv[:4].numpy() - ref["M"][:, packed])) / max(1.0, float(np.max(np.abs(ref["M"])))))
            assert err <= 1e-10, f"e2e support-only M differs from the oracle: {err:.3e}"
        del ho, hMv, hq, hv, ha

    # ---- secondary configs + Mapping path (rank 0, N=1 only) ----------------------------------------------------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        del dyn
        torch.cuda.empty_cache()
        secondary = {}
        for rb, nn in (("panda7", 1 << 16), ("solo12", 1 << 18), ("talos_reduced", 1 << 20)):
            headline = (rb == desc.name and nn == n)
            try:  # the secondary object is a rider: it must never cost the headline line
                secondary[f"{rb}:N={nn}"] = measure_config(torch, rb, nn, args.secondary_steps, local, with_eval=not headline)
            except Exception as ex:
                secondary[f"{rb}:N={nn}"] = {"error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()

    # ---- measured FP64 peak + CPU baseline (rank 0) --------------------------------------------------------------
    fp64 = None
    if rank == 0:
        import oracle
        fc = oracle.flop_counts(desc)
        peak_tf = api.probe_fp64_peak(local)
        prof = _profile_record(desc.name, n, args.tile or n)
        ach = fc["eval"]["flops"] * (n / (kern_ms * 1e-3)) / 1e12
        fp64 = {"flops_per_eval": fc["eval"]["flops"], "ops": fc["eval"], "by_algorithm": {k: fc[k]["flops"] for k in fc},
                "flops_source": "oracle -DORC_COUNT_FLOPS (adds + muls + divs + sqrts of the restated reference "
                                "algorithm: FK recomputed by computeJointJacobians, crba, rnea; sincos counted apart)",
                "peak_tflops": peak_tf, "peak_source": "measured live: rbd_probe_fp64_peak (DFMA chains on every SM)",
                "achieved_tflops": ach, "frac": ach / peak_tf if peak_tf > 0 else None,
                "kernel_fp64_flops_per_eval": prof.get("fp64_flops_per_eval"),
                "kernel_fp64_source": prof.get("fp64_source")}
        if prof.get("fp64_flops_per_eval"):
            # what the FP64 pipe really executes (the world-frame sweep does one FK pass, not three)
            k_ach = float(prof["fp64_flops_per_eval"]) * (n / (kern_ms * 1e-3)) / 1e12
            fp64["kernel_achieved_tflops"] = k_ach
            fp64["kernel_frac"] = k_ach / peak_tf if peak_tf > 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(desc, n, args.cpu_seconds)

    if rank == 0:
        peak, peak_src = _peaks()
        achieved = step_bytes / (kern_ms * 1e-3) / 1e9
        prof = _profile_record(desc.name, n, args.tile or n)
        hbm_frac = achieved / peak
        binding = "hbm"
        if fp64 and fp64["frac"] is not None and fp64["frac"] > hbm_frac:
            binding = "fp64"
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{desc.name} N={n} per GPU, fused FK+Jac+CRBA+RNEA, all outputs written",
                           "robot": desc.name, "instances_per_gpu": n, "layout": f"tiled SoA, tile={args.tile or n}",
                           "eval_storage": eval_storage,
                           "l2": (f"inputs+outputs {step_bytes / 2**20:.0f} MiB per step > 126 MiB L2, no flush needed" if flush is None
                                  else f"inputs+outputs {step_bytes / 2**20:.0f} MiB per step: L2 flushed between timed steps "
                                       f"by a {3 * L2_BYTES / 2**20:.0f} MiB fill"),
                           "note": desc.note, "sharding": f"instances, contiguous ranges, {world} rank(s), no collective"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
                "roofline": {"bound": binding, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": hbm_frac, "traffic": prof.get("dram_bytes_per_launch"),
                             "traffic_source": prof.get("source"),
                             "algorithmic_bytes_per_launch": step_bytes, "peak_source": peak_src,
                             "kernel": kernel_name, "kernel_source": "read back from the launch (rbd_last_kernel_name)",
                             "kernel_ms": kern_ms, "bytes_per_eval": bytes_per_eval, "fp64": fp64,
                             "binding": f"{binding} (the roofline with the larger achieved fraction binds; "
                                        f"hbm {hbm_frac:.3f}" + (f", fp64 {fp64['frac']:.3f})" if fp64 and fp64["frac"] is not None else ")")},
                "cpu_baseline": cpu, "oracle_check_rel_err": check_err, "secondary": secondary}
        if fp32 is not None:
            line["fp32_mode"] = fp32
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
