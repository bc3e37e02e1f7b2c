# template modes on configs 2 and 5 (same box, one process per variant): tile templates (round-2 baseline), cell-granular
# templates with / without the auxiliary stream
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
for v in "tile:PDENLP_TPL_MODE=tile" "cell:PDENLP_TPL_MODE=cell" "cell-noaux:PDENLP_NO_AUX_STREAM=1"; do
  ONLY="2:" STEPS=20 bash tools/ab_run.sh "cfg2-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"
  ONLY="5:" STEPS=20 bash tools/ab_run.sh "cfg5-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"
done
done
