# e2e step time against the number of panel upload chunks (bench.py, C3).  Diagnostic.
for n in ${CHUNKS:-8 10 12 16 20}; do
  IBDGEM_PANEL_CHUNKS=$n timeout 120 python bench.py --no-cpu-baseline > gpurun_out/sw_$n.json 2> gpurun_out/sw_err.log
done
