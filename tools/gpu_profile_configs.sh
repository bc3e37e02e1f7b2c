#!/bin/bash
# One GPU box: ncu --set full of every callback kernel of BASELINE configs (default 3 4 5), exported as raw CSV for
# tools/ncu_configs_summary.py.  The .ncu-rep files are deleted on the box unless KEEP_REP=1 (gpurun copies back at
# most 64 MiB).  Usage (under gpurun): bash tools/gpu_profile_configs.sh <tag> [configs...]
tag=${1:-r02}; shift
cfgs=${@:-3 4 5}
mkdir -p gpurun_out
for k in $cfgs; do
  python tools/profile_configs.py --config $k > gpurun_out/${tag}_cfg${k}_plain.log 2>&1 || { echo "plain run failed for $k"; tail -5 gpurun_out/${tag}_cfg${k}_plain.log; continue; }
  ncu --set full --clock-control none -k regex:'k_(vector|jac|hess|zero|add|sum|copy)' -c ${NCU_COUNT:-40} -f \
      -o gpurun_out/${tag}_cfg${k} python tools/profile_configs.py --config $k > gpurun_out/${tag}_cfg${k}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_cfg${k}.ncu-rep --page raw --csv > gpurun_out/${tag}_cfg${k}_raw.csv 2>/dev/null
  [ -n "$KEEP_REP" ] || rm -f gpurun_out/${tag}_cfg${k}.ncu-rep
  echo "config $k: $(grep -c k_ gpurun_out/${tag}_cfg${k}_raw.csv) kernels captured"
done
