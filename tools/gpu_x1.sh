mkdir -p gpurun_out/x1
timeout 900 python -m pytest tests/test_gpu_precision.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -30
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/x1/bench_default.json 2> gpurun_out/x1/bench_default.err
python -c "
import json
d=json.loads(open('gpurun_out/x1/bench_default.json').read().strip().splitlines()[-1])
print('default ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), d['clocks']['sm_mhz'], d.get('precision_variants'))
" || tail -5 gpurun_out/x1/bench_default.err
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --precision f16+e4m3 --verify > gpurun_out/x1/bench_x1.json 2> gpurun_out/x1/bench_x1.err
python -c "
import json
d=json.loads(open('gpurun_out/x1/bench_x1.json').read().strip().splitlines()[-1])
print('x1 ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['clocks']['sm_mhz'], d['verify']['image_max_abs_err'], d['verify']['knn_bit_exact'], d['roofline']['frac'], d['dtype'][:60])
" || tail -5 gpurun_out/x1/bench_x1.err
