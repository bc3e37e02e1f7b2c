# hess_coord!: zero fill of the constraint block as extra CTAs of the chunk-copy launch (default; PDENLP_NO_MERGE_FILL=1 = separate) vs its own PDL launch
python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "membrane or template" 2>&1 | tail -2
for rep in 1 2; do
for v in sep merged; do
  if [ $v = sep ]; then export PDENLP_NO_MERGE_FILL=1; else unset PDENLP_NO_MERGE_FILL; fi
  for n in 724 2048; do
    python bench.py --n $n --steps 200 --warmup 10 --no-cpu --no-e2e --configs= 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v n=$n', round(d['ms_per_step'],4), round(d['jac_ms'],4), round(d['hess_ms'],4))"
  done
done
done
