# Round evidence on one B200: sanitizer logs, ncu launch list + full captures of the top kernels, bench lines.  ~12 GPU-minutes.
tag=${1:-r2_evidence}
out=gpurun_out/$tag
mkdir -p $out
# ---- compute-sanitizer (small sizes: smoke() and the two exact-given-stash tests) ----
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py --smoke > $out/${tool}_smoke.log 2>&1; echo "rc=$?" >> $out/${tool}_smoke.log
done
for tool in racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact_given_stash" > $out/${tool}_stash.log 2>&1; echo "rc=$?" >> $out/${tool}_stash.log
done
tail -n 3 $out/*check*.log
# ---- launch list of the default bench command (per-launch durations; shares must agree with the event timings) ----
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $out/launches_bench.log 2>&1
# ---- full captures: pair + heads (full-size launches), query + composite kernels ----
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 2 -f -o $out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_tc.log 2>&1
ncu -i $out/prof_tc.ncu-rep --page raw --csv > $out/prof_tc_raw.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -k regex:"k_march_count_s|k_knn_fill_s|k_composite_fwd" -s 3 -c 3 -f -o $out/prof_query python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_query.log 2>&1
ncu -i $out/prof_query.ncu-rep --page raw --csv > $out/prof_query_raw.csv 2>/dev/null
# ---- bench lines ----
python bench.py --steps 20 --warmup 5 --verify > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --precision f16x3 > $out/bench_n1_f16x3.json 2> /dev/null
ls -la $out
