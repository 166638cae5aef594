# Quick GPU check of a pair-kernel change: smoke (vs the oracle), precision tests, a short bench line, the pair timeline.
# Usage: bash tools/gpu_r3.sh <tag>     (every step under its own kill-timeout: a barrier bug must not hang the box)
tag=${1:-r3}
out=gpurun_out/$tag
mkdir -p $out
timeout -s KILL 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 $out/smoke.log
timeout -s KILL 600 python -m pytest tests/test_gpu_precision.py -x -q -m gpu > $out/pytest_precision.log 2>&1; echo "precision rc=$?"; tail -n 4 $out/pytest_precision.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$out/bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("ms/step", round(d["ms_per_step"],2), "Mrays/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "pair share", round(r["share_of_step"],3), "heads share", round(r["heads_share_of_step"],3), "frac", round(r["frac"],3), d["clocks"], r["hbm_path"].get("stage_ms_per_step"))
except Exception as e: print("bench failed", e, open("$out/bench.err").read()[-1500:])
PY
timeout -s KILL 300 python tools/timeline_pair.py > $out/timeline_pair.txt 2>&1; cat $out/timeline_pair.txt | head -40
