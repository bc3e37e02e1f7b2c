#!/bin/bash
# A/B helper: bash tools/ab_run.sh "<label>:<ENV=..;ENV2=..>" ... -- runs tools/bench_configs.py --only "$ONLY" per variant
ONLY=${ONLY:-5:}
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  ( IFS=';'; for e in $envs; do [ -n "$e" ] && export "$e"; done
    echo "== $label ($envs)"
    python tools/bench_configs.py --only "$ONLY" --steps ${STEPS:-20} 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v['ms'],4) for k,v in d['callbacks'].items()}, max(d['max_rel_err_vs_oracle_slab'].values()))
    elif 'rror' in l: print(l.strip())
" )
done
