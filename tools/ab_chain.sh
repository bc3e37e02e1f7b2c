# chunk copy chained behind the generic-cell kernel (PDL) vs auxiliary-stream fork/join; n = 724 is the 1/8-slab size
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2; do
for v in chain nochain; do
  if [ $v = nochain ]; then export PDENLP_NO_CHAIN=1; else unset PDENLP_NO_CHAIN; fi
  for n in 724 2048; do
    python bench.py --n $n --steps 200 --warmup 10 --no-cpu --no-e2e --configs= 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v n=$n', round(d['ms_per_step'],4), round(d['jac_ms'],4), round(d['hess_ms'],4))"
  done
done
done
