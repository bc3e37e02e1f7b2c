python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graph" 2>&1 | tail -4
for v in "graphs:" "nographs:PDENLP_NO_GRAPHS=1"; do
  for c in 4 5; do ONLY="$c:" STEPS=40 bash tools/ab_run.sh "cfg$c-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"; done
done
