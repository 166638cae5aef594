# Development aid: timeline + bench of the field kernels for build variants (extra -D switches), one GPU.
# usage: bash tools/gpu_variants.sh <tag> "<defines of variant 1>" "<defines of variant 2>" ...   ("" = default build)
tag=$1; shift
mkdir -p gpurun_out/$tag
i=0
for defs in "$@"; do
  export NPCD_NVCC_DEFINES="$defs"
  echo "=== variant $i: '$defs'"
  python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build_$i.log 2>&1 || tail -5 gpurun_out/$tag/build_$i.log
  timeout 600 python -m pytest tests/test_gpu_precision.py -x -q -m gpu 2>&1 | tail -4
  timeout 300 python tools/timeline_pair.py 2>&1 | tail -30
  for p in f16+e4m3x2 f16x3; do
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --precision $p 2>/dev/null > gpurun_out/$tag/bench_${i}_$p.json
    python -c "import json,sys; d=json.loads(open('gpurun_out/$tag/bench_${i}_$p.json').read()); r=d['roofline']; print(d['kernels']['precision'], 'ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), 'pair', round(r['share_of_step']*d['ms_per_step'],1), 'heads', round(r['heads_share_of_step']*d['ms_per_step'],1), d['clocks']['sm_mhz'])"
  done
  i=$((i+1))
done
