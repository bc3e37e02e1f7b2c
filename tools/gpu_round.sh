#!/bin/bash
# One GPU box: the whole gpu test-suite, the default bench line, then the ncu passes of the headline step
# (launch list + one --set full capture of the step's kernels) for profiles/.  Usage: bash tools/gpu_round.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --configs="
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(copy_chunks|jac_cells_cg|hess_cells_cg|zero_fill)' -c 30 -f \
    -o gpurun_out/${tag}_step $B > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_step.ncu-rep --page raw --csv > gpurun_out/${tag}_step_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_step.ncu-rep
