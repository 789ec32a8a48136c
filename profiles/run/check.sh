# usage: bash profiles/run/check.sh <tag> [bench-steps]   (on the GPU box, via gpurun)
TAG=${1:-x}; STEPS=${2:-20}
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
