# A/B of the tensor-memory operand form of the pair kernel (NPCD_TC_TS=1 default / 0) on ONE box.  Usage: bash tools/gpu_ts.sh <tag>
tag=${1:-ts}
out=gpurun_out/$tag
mkdir -p $out
timeout -s KILL 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $out/smoke.log
timeout -s KILL 600 python -m pytest tests/test_gpu_precision.py -x -q -m gpu > $out/pytest_precision.log 2>&1; echo "precision rc=$?"; tail -n 12 $out/pytest_precision.log
for round in 1 2; do
  for ts in 1 0; do
    NPCD_TC_TS=$ts timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/bench_ts${ts}_$round.json 2> $out/bench_ts${ts}_$round.err
    python -c "
import json; d=json.loads(open('$out/bench_ts${ts}_$round.json').read().strip().splitlines()[-1]); r=d['roofline']
print('TS=$ts', $round, 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), 'pair', round(r['share_of_step']*d['ms_per_step'],1), 'heads', round(r['heads_share_of_step']*d['ms_per_step'],1), d['clocks']['sm_mhz'])" || tail -3 $out/bench_ts${ts}_$round.err
  done
done
timeout -s KILL 300 python tools/timeline_pair.py > $out/timeline_pair.txt 2>&1; head -40 $out/timeline_pair.txt
