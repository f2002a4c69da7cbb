#!/bin/bash
# usage (on the GPU box, from the repo root): profiles/measure_r2.sh <tag>
# Round-2 measurement set: GPU tests, the bench line (with its `secondary` object: every BASELINE config + the Mapping
# path), the reference arm, the ncu launch list of the bench command, one `ncu --set full` capture per hot kernel and
# the FP64 instruction counts of the three eval instantiations.  Everything lands in gpurun_out/<tag>_*; the summaries
# kept under profiles/ are made from these by profiles/ncu_summary.py.
T="$1"; O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/${T}_box.txt; nproc >> $O/${T}_box.txt
(time python -m pytest tests -m gpu -x -q) > $O/${T}_tests.log 2>&1; tail -2 $O/${T}_tests.log
python bench.py --steps 20 --warmup 3 --f32 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench.err; cut -c1-300 $O/${T}_bench_1gpu.json
python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err; cut -c1-200 $O/${T}_bench_reference_arm.json
Q="--steps 2 --warmup 1 --no-e2e --no-cpu --no-secondary"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --secondary-steps 2 > $O/${T}_launches.log 2>&1
profiles/capture.sh "eval_kernel_ffroot8" ${T}_eval_talos -- python bench.py $Q
profiles/capture.sh "eval_kernel_ffroot<" ${T}_eval_solo -- python bench.py --workload solo $Q
profiles/capture.sh "eval_arm_kernel" ${T}_eval_panda -- python bench.py --workload panda $Q
profiles/capture.sh "aba2_kernel" ${T}_aba2_talos -- python experiments/time_aba.py talos_reduced
profiles/capture.sh "minv_kernel" ${T}_minv_talos -- python experiments/time_minv.py talos_reduced
profiles/capture.sh "drnea_kernel" ${T}_drnea_talos -- python experiments/time_drnea.py talos_reduced:262144
profiles/capture.sh "applyJT_arm_kernel" ${T}_applyJT_panda -- python bench_mapping.py --robot panda7 --n 65536 --steps 3
python experiments/time_drnea.py > $O/${T}_time_drnea.json 2>&1
python experiments/time_aba.py > $O/${T}_time_aba.json 2>&1
python experiments/time_minv.py > $O/${T}_time_minv.json 2>&1
profiles/capture.sh "applyJT_rows_cached" ${T}_rows_talos -- python bench_mapping.py --robot talos_reduced --n 1048576 --steps 3
profiles/capture.sh "transpose32v_kernel<\(bool\)1>|transpose32v_kernel<true>" ${T}_soa_to_aos -- python bench_mapping.py --robot talos_reduced --n 1048576 --steps 3
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum
for w in talos solo panda; do
  ncu --metrics $M --clock-control none --kernel-name-base demangled -k "regex:eval_.*kernel" -s 2 -c 1 --csv --log-file $O/${T}_fp64_${w}.csv \
      python bench.py --workload $w $Q > /dev/null 2>&1
done
rm -f $O/${T}_*_cuda.csv   # the largest page; the summaries use raw + sass
