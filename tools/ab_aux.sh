# config 5: how the kappa-block kernels share the SMs with the chunk copy (auxiliary stream)
for v in "base:" "bg1:PDENLP_KAPPA_BG_CTAS=1" "bg2:PDENLP_KAPPA_BG_CTAS=2" "bg3:PDENLP_KAPPA_BG_CTAS=3" "hi:PDENLP_AUX_PRIO=hi" "hi-bg1:PDENLP_AUX_PRIO=hi;PDENLP_KAPPA_BG_CTAS=1" "hi-bg2:PDENLP_AUX_PRIO=hi;PDENLP_KAPPA_BG_CTAS=2" "bg1-t128:PDENLP_KAPPA_BG_CTAS=2;PDENLP_KVEC_THREADS=128"  "bg4-t128:PDENLP_KAPPA_BG_CTAS=4;PDENLP_KVEC_THREADS=128"; do
  ONLY="5:" STEPS=20 bash tools/ab_run.sh "cfg5-$v" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"
done
