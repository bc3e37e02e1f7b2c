#!/bin/bash
# A/B of the reference-tensor vector kernels on config 5: launch-bounds variants x threads per CTA
for lib in default tools/ab/libpdenlp_mb1.so tools/ab/libpdenlp_mb3.so; do
  for thr in 256 128; do
    if [ "$lib" = default ]; then unset PDENLP_CUDA_LIB; else export PDENLP_CUDA_LIB=$PWD/$lib; fi
    export PDENLP_KVEC_THREADS=$thr
    echo "== lib=$lib threads=$thr"
    python tools/bench_configs.py --only "5:" --steps 10 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v['ms'],4) for k,v in d['callbacks'].items()}, d['max_rel_err_vs_oracle_slab'])
    elif 'rror' in l: print(l.strip())
"
  done
done
