# reference-tensor vector kernels on config 5: L2 prefetch of the next tile's gathers (variant built with -DPDENLP_KVEC_PREFETCH=1)
for rep in 1 2; do
for lib in default pf; do
  if [ "$lib" = default ]; then unset PDENLP_CUDA_LIB; else export PDENLP_CUDA_LIB=$PWD/tools/ab/libpdenlp_$lib.so; fi
  echo "== lib=$lib"
  python tools/bench_configs.py --only "5:" --steps 20 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v['ms'],4) for k,v in d['callbacks'].items()}, max(d['max_rel_err_vs_oracle_slab'].values()))
    elif 'rror' in l: print(l.strip())
"
done
done
