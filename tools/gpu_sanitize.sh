set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2/gpuinfo.txt
nproc >> gpurun_out/r2/gpuinfo.txt; free -g >> gpurun_out/r2/gpuinfo.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python __graft_entry__.py --smoke > gpurun_out/r2/racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2/racecheck_smoke.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 30 python __graft_entry__.py --smoke > gpurun_out/r2/synccheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2/synccheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python __graft_entry__.py --smoke > gpurun_out/r2/memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2/memcheck_smoke.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact_given_stash" > gpurun_out/r2/racecheck_stash.log 2>&1; echo "rc=$?" >> gpurun_out/r2/racecheck_stash.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "exact_given_stash" > gpurun_out/r2/synccheck_stash.log 2>&1; echo "rc=$?" >> gpurun_out/r2/synccheck_stash.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
tail -3 gpurun_out/r2/*.log
