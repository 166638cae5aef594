# ncu --set full of ONE full-size pair launch of the headline bench, with the per-instruction source page.  One GPU, ~1.5 min.
tag=${1:-ncu1}
mkdir -p gpurun_out/$tag
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 1 -f -o gpurun_out/$tag/prof_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/$tag/ncu_stdout.log 2>&1
ncu -i gpurun_out/$tag/prof_tc.ncu-rep --page raw --csv > gpurun_out/$tag/prof_tc_raw.csv 2>/dev/null
ncu -i gpurun_out/$tag/prof_tc.ncu-rep --page source --csv --print-source sass > gpurun_out/$tag/prof_tc_source.csv 2>/dev/null
rm -f gpurun_out/$tag/prof_tc.ncu-rep
ls -la gpurun_out/$tag
