# Development aid: the ray-coherent query kernels (impl 3) against the others: parity tests, then per-stage times of the 251-view step.
mkdir -p gpurun_out/q
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants or knn or full_size or rays" 2>&1 | tail -5
for impl in 0 3; do
  NPCD_QUERY_IMPL=$impl timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/q/bench_$impl.err > gpurun_out/q/bench_$impl.json
  python -c "import json; d=json.loads(open('gpurun_out/q/bench_$impl.json').read()); h=d['roofline']['hbm_path']; print('impl $impl ms', round(d['ms_per_step'],2), {k: round(v,3) for k,v in h['stage_ms_per_step'].items()}, 'frac', round(h['frac'],4))" || tail -3 gpurun_out/q/bench_$impl.err
done
