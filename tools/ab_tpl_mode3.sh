# cell-granular templates: chunk order and diagnostics (skip generic / skip chunks); PDENLP_NO_SELFCHECK for the skips
for v in "order0:PDENLP_CG_ORDER=0" "order1:PDENLP_CG_ORDER=1" "order1occ1:PDENLP_CG_ORDER=1;PDENLP_CG_OCC=1" "order1occ4:PDENLP_CG_ORDER=1;PDENLP_CG_OCC=4" \
   "skipgen0:PDENLP_CG_ORDER=2;PDENLP_NO_SELFCHECK=1" "skipgen1:PDENLP_CG_ORDER=3;PDENLP_NO_SELFCHECK=1" "skipchunks:PDENLP_CG_ORDER=4;PDENLP_NO_SELFCHECK=1"; do
  ONLY="2:" STEPS=20 bash tools/ab_run.sh "cfg2-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*//"
  ONLY="5:" STEPS=20 bash tools/ab_run.sh "cfg5-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*//"
done
