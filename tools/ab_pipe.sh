# vector kernels: two-deep software pipeline (state gathers one tile ahead) x launch bounds, configs 2, 3, 4, 5
for lib in pipe0 default pipe_m3 pipe_m2; do
  if [ "$lib" = default ]; then unset PDENLP_CUDA_LIB; else export PDENLP_CUDA_LIB=$PWD/tools/ab/libpdenlp_$lib.so; fi
  for c in 2 3 4 5; do
    echo "== lib=$lib cfg=$c"
    python tools/bench_configs.py --only "$c:" --steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v['ms'],4) for k,v in d['callbacks'].items() if k in ('obj','grad!','cons!')}, max(d['max_rel_err_vs_oracle_slab'].values()))
    elif 'rror' in l: print(l.strip())
"
  done
done
