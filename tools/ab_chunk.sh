python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
for v in "c32:PDENLP_CHUNK_CELLS=32" "c64:PDENLP_CHUNK_CELLS=64" "c48:PDENLP_CHUNK_CELLS=48" "c16:PDENLP_CHUNK_CELLS=16"; do
  for c in 2 5; do ONLY="$c:" STEPS=30 bash tools/ab_run.sh "cfg$c-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"; done
done
ONLY="4:" STEPS=30 bash tools/ab_run.sh "cfg4-fusedkappa:"
