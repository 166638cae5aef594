# Final evidence of a build on one B200 (shorter than gpu_evidence.sh: the training kernels' sanitizer runs are not repeated).
tag=${1:-r2_final}
out=gpurun_out/$tag
mkdir -p $out
for tool in racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py --smoke > $out/${tool}_smoke.log 2>&1; echo "rc=$?" >> $out/${tool}_smoke.log
done
tail -n 3 $out/*check*.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $out/launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 2 -f -o $out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_tc.log 2>&1
ncu -i $out/prof_tc.ncu-rep --page raw --csv > $out/prof_tc_raw.csv 2>/dev/null
timeout 1200 ncu --set full --clock-control none -k regex:"k_march_count_s|k_knn_fill_s|k_composite_fwd" -s 3 -c 3 -f -o $out/prof_query python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_query.log 2>&1
ncu -i $out/prof_query.ncu-rep --page raw --csv > $out/prof_query_raw.csv 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
python bench.py --steps 20 --warmup 5 --verify > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --workload train --steps 40 --warmup 5 > $out/bench_train_n1.json 2> /dev/null
python tools/timeline_pair.py > $out/timeline_pair.txt 2>&1
python tools/timeline_heads.py > $out/timeline_heads.txt 2>&1
ls -la $out
python -c "
import json
d=json.loads(open('$out/bench_n1.json').read().strip().splitlines()[-1])
print('N=1', round(d['ms_per_step'],2), 'ms', round(d['value']/1e6,2), 'Mrays/s e2e', round(d['e2e']['value']/1e6,2), 'frac', round(d['roofline']['frac'],4), 'hbm', round(d['roofline']['hbm_path']['frac'],4), d['clocks'], d.get('verify',{}).get('knn_bit_exact'), d.get('verify',{}).get('image_max_abs_err'))
print('cpu', d['cpu_baseline'])
r=json.loads(open('$out/bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline'])
"
