# A/B of prebuilt library variants on ONE box (box-to-box clock differences are as large as the effects looked for):
# usage: bash tools/gpu_ab.sh <tag> <variant> [<variant> ...]   ("default" = the in-tree library; others: build/variants/libnpcd_<v>.so)
tag=$1; shift
mkdir -p gpurun_out/$tag
V=neural-point-cloud-diffusion_b200/build/variants
for round in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then unset NPCD_LIB_PATH; else export NPCD_LIB_PATH=$V/libnpcd_$v.so; fi
    timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/$tag/bench_${v}_$round.json 2> gpurun_out/$tag/bench_${v}_$round.err
    python -c "
import json; d=json.loads(open('gpurun_out/$tag/bench_${v}_$round.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$v', $round, 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), 'pair', round(r['share_of_step']*d['ms_per_step'],1), 'heads', round(r['heads_share_of_step']*d['ms_per_step'],1), d['clocks']['sm_mhz'])" || tail -3 gpurun_out/$tag/bench_${v}_$round.err
  done
done
unset NPCD_LIB_PATH
timeout -s KILL 300 python tools/timeline_pair.py > gpurun_out/$tag/timeline_pair.txt 2>&1; head -40 gpurun_out/$tag/timeline_pair.txt
