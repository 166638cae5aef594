# Multi-GPU check (run under `gpurun --gpus N`): the 2-GPU sharded-step test and the default bench line at N ranks.
n=${1:-2}
tag=${2:-r2_n$n}
mkdir -p gpurun_out/$tag
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k two_gpu > gpurun_out/$tag/pytest_two_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/$tag/pytest_two_gpu.log
tail -n 6 gpurun_out/$tag/pytest_two_gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/$tag/bench_n$n.json 2> gpurun_out/$tag/bench_n$n.err
tail -c 1200 gpurun_out/$tag/bench_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$tag/bench_n$n.json").read().strip().splitlines()[-1])
print("render", d["n_gpus"], round(d["ms_per_step"],2), "ms", round(d["value"]/1e6,2), "Mrays/s", [ (r["S"], round(r["ms_per_step"],2)) for r in d["workload_stats"]["per_rank"]])
s=d["secondary"]
print("train", round(s["train"]["ms_per_step"],3), "ms/step", s["train"]["ms_per_step_rank0"], "host", round(s["train"]["host_ms_per_step_rank0"],2))
print("decode", round(s["decode"]["ms_per_step"],2), "ms", round(s["decode"]["value"]/1e6,2), "Mrays/s")
PY
