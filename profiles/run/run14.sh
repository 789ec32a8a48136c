# two GPUs: the IPC peer-exchange test, then the bench line at N = 2
timeout 250 python -m pytest tests/test_gpu_peer.py -x -q -k two_gpus > gpurun_out/pytest_peer2.log 2>&1; tail -3 gpurun_out/pytest_peer2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_final.log 2> gpurun_out/bench_n2_final.err; tail -c 300 gpurun_out/bench_n2_final.err
