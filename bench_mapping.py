#!/usr/bin/env python
"""bench_mapping.py — the secondary measurements of bench.py for ONE (robot, N), on their own (NOT the driver contract;
bench.py is, and its ``secondary`` object carries the same figures for the three BASELINE robots): the fused eval
(packed and support-only M), apply -> applyJ -> applyJT on device tiled-SoA buffers with one wrench per mapped output,
the constraint-row applyJT (two rows per instance, two non-zeros each), applyDJT, the Mass / ForceField products and
the separate CRBA / RNEA entry points.  One JSON line: time, algorithmic HBM bandwidth, kernel name (read back from the
launch) and a clock sample per entry point."""
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
    import torch
    import bench
    print(json.dumps(bench.measure_config(torch, args.robot, args.n, args.steps, 0, with_eval=True, tile=args.tile)))


if __name__ == "__main__":
    main()
