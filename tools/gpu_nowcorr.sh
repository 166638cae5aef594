# Precision experiment: inference field kernels without the weight-rounding correction product (1.5 tensor passes per product).
mkdir -p gpurun_out/nw
export NPCD_LIB_PATH=$PWD/neural-point-cloud-diffusion_b200/build/variants/libnpcd_nowcorr.so
timeout 600 python -m pytest tests/test_gpu_precision.py -q -m gpu 2>&1 | tail -25
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary --verify > gpurun_out/nw/bench.json 2> gpurun_out/nw/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/nw/bench.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],2), 'Mrays/s', round(d['value']/1e6,2), d['clocks'], d['verify']['image_max_abs_err'], d['verify']['knn_bit_exact'])
" || tail -5 gpurun_out/nw/bench.err
