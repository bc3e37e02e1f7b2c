python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in "tile:PDENLP_TPL_MODE=tile" "cell:PDENLP_TPL_MODE=cell"; do
  for c in 2 5; do ONLY="$c:" STEPS=30 bash tools/ab_run.sh "cfg$c-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"; done
done
