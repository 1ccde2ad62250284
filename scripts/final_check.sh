#!/bin/bash
# One-call GPU check used at the end of round 1: A/B of the L2-prefetch variant of the per-aggregate
# kernels, then the full GPU test-suite and the bench with the faster setting.
mkdir -p gpurun_out
us() { grep -o "= [0-9.]* us/it" "$1" | tail -1 | grep -o "[0-9.]*" | head -1; }
FEMB_TL_PREFETCH=0 timeout 60 python scripts/twolevel_bench.py 56 56 54 2>&1 | grep -v symbolic | tail -1 > gpurun_out/r1_pf_off.log
FEMB_TL_PREFETCH=1 timeout 60 python scripts/twolevel_bench.py 56 56 54 2>&1 | grep -v symbolic | tail -1 > gpurun_out/r1_pf_on.log
cat gpurun_out/r1_pf_off.log gpurun_out/r1_pf_on.log
OFF=$(us gpurun_out/r1_pf_off.log); ON=$(us gpurun_out/r1_pf_on.log)
PF=$(python -c "print(1 if float('${ON:-999}') < 0.99 * float('${OFF:-1}') else 0)")
echo "prefetch off ${OFF} us/it, on ${ON} us/it -> FEMB_TL_PREFETCH=${PF}" | tee gpurun_out/r1_pf_choice.log
export FEMB_TL_PREFETCH=$PF
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r1_tests_final3.log
timeout 200 python bench.py > gpurun_out/r1_bench_final3.json 2> gpurun_out/r1_bench_final3.err
python -c "
import json; d=json.load(open('gpurun_out/r1_bench_final3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['pcg']['iterations'], d['pcg']['ms_per_iteration'], d['pcg']['coarse_setup_ms'], d['modal']['ms'], d['cpu_baseline']['value'], d['gpu_launches'])"
