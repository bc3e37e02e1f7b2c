# configs 3, 4, 5: all callbacks, with and without the auxiliary stream
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in "aux:" "noaux:PDENLP_NO_AUX_STREAM=1"; do
  for c in 3 4 5; do ONLY="$c:" STEPS=30 bash tools/ab_run.sh "cfg$c-$v"; done
done
