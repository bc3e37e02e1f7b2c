# cell-granular templates: occupancy / PDL variants + ncu launch list (durations, DRAM bytes) of configs 2 and 5
for v in "occ1:PDENLP_CG_OCC=1" "occ3:PDENLP_CG_OCC=3" "nopdl:PDENLP_NO_PDL=1" "occ1nopdl:PDENLP_CG_OCC=1;PDENLP_NO_PDL=1"; do
  ONLY="2:" STEPS=20 bash tools/ab_run.sh "cfg2-$v"
  ONLY="5:" STEPS=20 bash tools/ab_run.sh "cfg5-$v"
done
for k in 2 5; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_(jac|hess|zero|vector_[KM])' -c 30 --csv \
  --log-file gpurun_out/r02n_cfg${k}_cell_launches.csv python tools/profile_configs.py --config $k > /dev/null 2>&1
done
