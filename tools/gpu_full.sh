# Whole GPU suite + the driver's default bench line + the reference arm.  Usage: bash tools/gpu_full.sh <tag>
tag=${1:-full}
out=gpurun_out/$tag
mkdir -p $out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 6 $out/pytest_gpu.log
timeout -s KILL 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; tail -c 600 $out/bench.err
python - <<PY
import json
try:
    d=json.loads(open("$out/bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("ms/step", round(d["ms_per_step"],2), "Mrays/s", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "frac", round(r["frac"],3), "hbm", round(r["hbm_path"]["frac"],4), d["clocks"])
    print("train", d["secondary"]["train"]["ms_per_step"], "decode", d["secondary"]["decode"]["value"], "cpu", d["cpu_baseline"])
except Exception as e: print("bench failed", e)
PY
