# cta_group::2 build of the tensor-memory-form kernels (prebuilt variant k2ts) against the default build on ONE box.
tag=${1:-k2ts}
out=gpurun_out/$tag
mkdir -p $out
V=neural-point-cloud-diffusion_b200/build/variants
NPCD_LIB_PATH=$V/libnpcd_k2ts.so timeout -s KILL 300 python __graft_entry__.py --smoke > $out/smoke_k2.log 2>&1; echo "k2ts smoke rc=$?"; tail -n 2 $out/smoke_k2.log
NPCD_LIB_PATH=$V/libnpcd_k2ts.so timeout -s KILL 600 python -m pytest tests/test_gpu_precision.py -x -q -m gpu > $out/pytest_precision_k2.log 2>&1; echo "k2ts precision rc=$?"; tail -n 5 $out/pytest_precision_k2.log
bash tools/gpu_ab.sh $tag default k2ts
