# GPU-box check of the current tree: parity tests, the whole GPU suite, bench lines.  Usage: bash tools/gpu_check.sh <tag> [quick]
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests/test_gpu_precision.py tests/test_gpu_voxel_mode.py -q -m gpu -s > $out/pytest_precision.log 2>&1; echo "rc=$?" >> $out/pytest_precision.log
grep -E "probe|max err|max \||passed|failed|Error|rc=" $out/pytest_precision.log | tail -n 30
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --precision f16x3 > $out/bench_f16x3.json 2> $out/bench_f16x3.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --precision f16+e4m3x2 > $out/bench_f8.json 2> $out/bench_f8.err
python - <<PY
import json
for n in ("f16x3","f8"):
    try:
        d=json.loads(open("$out/bench_%s.json"%n).read().strip().splitlines()[-1])
        r=d["roofline"]; print(n, "ms/step", round(d["ms_per_step"],2), "Mrays/s", round(d["value"]/1e6,2), "pair share", round(r["share_of_step"],3), "heads share", round(r["heads_share_of_step"],3), "frac", round(r["frac"],3), d["clocks"], r["hbm_path"].get("stage_ms_per_step"))
    except Exception as e: print(n, "failed", e, open("$out/bench_%s.err"%n).read()[-1500:])
PY
if [ "$2" = "quick" ]; then exit 0; fi
python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s > $out/pytest_configs.log 2>&1; echo "rc=$?" >> $out/pytest_configs.log
tail -n 15 $out/pytest_configs.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_configs.py --deselect tests/test_gpu_precision.py > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log
tail -n 8 $out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 600 $out/bench_ref.err
cat $out/bench_ref.json | head -c 1200
