# threads per CTA of the chunk copy (more CTAs in flight per SM hide the per-CTA record load)
for rep in 1 2; do
for t in 256 128 64; do
  for c in 2 5; do ONLY="$c:" STEPS=30 bash tools/ab_run.sh "cfg$c-t$t:PDENLP_COPY_THREADS=$t" 2>&1 | sed -e "s/'obj'.*'jac_coord/'jac_coord/" -e "s/'jprod.*}/}/"; done
done
done
