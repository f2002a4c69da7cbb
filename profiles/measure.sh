#!/bin/bash
# usage (on the GPU box, from the repo root): profiles/measure.sh <tag>
# One round of measurements: GPU tests, the bench line (+ FP32 mode), the reference arm, the mapping benches, the ncu
# launch list of the bench command and one `ncu --set full` capture of each hot kernel.  Everything lands in
# gpurun_out/<tag>_*; the summaries kept under profiles/ are made from these by profiles/ncu_summary.py.
T="$1"; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/${T}_box.txt; nproc >> $O/${T}_box.txt
(time python -m pytest tests -m gpu -x -q) > $O/${T}_tests.log 2>&1; tail -2 $O/${T}_tests.log
python bench.py --steps 20 --warmup 3 --f32 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_reference_arm.json
for r in talos_reduced:1048576 solo12:262144 panda7:65536; do
  python bench_mapping.py --robot ${r%%:*} --n ${r##*:} > $O/${T}_bench_mapping_${r%%:*}.json 2>> $O/${T}_bench.err
done
for w in solo panda; do python bench.py --workload $w --steps 20 --warmup 3 --no-e2e --no-cpu > $O/${T}_bench_${w}.json 2>> $O/${T}_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/${T}_launches.log 2>&1
profiles/capture.sh "eval_kernel" ${T}_eval -- python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu
profiles/capture.sh "applyJT_ring" ${T}_applyJT -- python bench_mapping.py --robot talos_reduced --n 1048576 --steps 3
profiles/capture.sh "apply_kernel" ${T}_apply -- python bench_mapping.py --robot talos_reduced --n 1048576 --steps 3
profiles/capture.sh "applyJ_kernel" ${T}_applyJ -- python bench_mapping.py --robot talos_reduced --n 1048576 --steps 3
rm -f $O/${T}_*_cuda.csv   # the largest page; the summaries use raw + sass (+ cuda when present)
