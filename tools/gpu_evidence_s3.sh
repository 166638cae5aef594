# Final evidence of the session-3 build of round 2 on one B200: whole GPU suite, default bench line, reference arm, ncu launch list,
# ncu --set full of the full-size pair + heads launches, phase timelines.  Usage: bash tools/gpu_evidence_s3.sh <tag>
tag=${1:-r2_s3}
out=gpurun_out/$tag
mkdir -p $out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 4 $out/pytest_gpu.log
timeout -s KILL 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $out/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 2 -c 2 -f -o $out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_tc.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_tc.ncu-rep --page raw --csv > $out/prof_tc_raw.csv 2>/dev/null
ncu -i $out/prof_tc.ncu-rep --page source --csv --print-source sass > $out/prof_tc_source.csv 2>/dev/null
rm -f $out/prof_tc.ncu-rep
timeout -s KILL 300 python tools/timeline_pair.py > $out/timeline_pair.txt 2>&1
timeout -s KILL 300 python tools/timeline_heads.py > $out/timeline_heads.txt 2>&1
python - <<PY
import json
d=json.loads(open("$out/bench_n1.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("N=1", round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,2), "Mrays/s e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(r["frac"],4), "hbm", round(r["hbm_path"]["frac"],4), d["clocks"])
print("train", d["secondary"]["train"]["ms_per_step"], "decode", d["secondary"]["decode"]["value"], "cpu", d["cpu_baseline"])
try:
    q=json.loads(open("$out/bench_ref.json").read().strip().splitlines()[-1]); print("ref", q["value"], q["cpu_baseline"])
except Exception as e: print("ref failed", e)
PY
