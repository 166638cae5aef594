# Development aid: timing-only ablations of the inference field kernels (prebuilt by tools/build_variants.py), one GPU.
# usage: bash tools/gpu_ablate.sh <tag> <variant> ...      ("default" = the in-tree library)
tag=$1; shift
mkdir -p gpurun_out/$tag
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
for v in "$@"; do
  if [ "$v" = default ]; then unset NPCD_LIB_PATH; else export NPCD_LIB_PATH=$PWD/neural-point-cloud-diffusion_b200/build/variants/libnpcd_$v.so; fi
  echo "=== $v"
  timeout 300 python tools/timeline_pair.py > gpurun_out/$tag/timeline_$v.txt 2>&1; head -2 gpurun_out/$tag/timeline_$v.txt
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/$tag/bench_$v.err > gpurun_out/$tag/bench_$v.json
  python -c "import json,sys; d=json.loads(open('gpurun_out/$tag/bench_$v.json').read()); r=d['roofline']; print('   ms', round(d['ms_per_step'],2), 'pair', round(r['share_of_step']*d['ms_per_step'],1), 'heads', round(r['heads_share_of_step']*d['ms_per_step'],1), 'MHz', d['clocks']['sm_mhz'])" || tail -3 gpurun_out/$tag/bench_$v.err
done
unset NPCD_LIB_PATH
