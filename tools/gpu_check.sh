# GPU-box check of the current tree: config-size parity tests, the whole GPU suite, bench lines.  Usage: bash tools/gpu_check.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s > $out/pytest_configs.log 2>&1; echo "rc=$?" >> $out/pytest_configs.log
tail -n 15 $out/pytest_configs.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_configs.py > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log
tail -n 8 $out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 600 $out/bench_ref.err
cat $out/bench_ref.json | head -c 1200
