# peer-mailbox exchange: tests on two GPUs, then the bench line at N = 2
timeout 500 python -m pytest tests/test_gpu_peer.py -x -q -s > gpurun_out/pytest_peer.log 2>&1; tail -15 gpurun_out/pytest_peer.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_peer.log 2> gpurun_out/bench_n2_peer.err; tail -c 600 gpurun_out/bench_n2_peer.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_n2_peer.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); e=d['elbo']
    print('N=2 value', d['value']/1e6, 'elbo', e['value'], 'nccl', e.get('nccl_all_reduce_steps_per_s'), 'sync', e.get('host_synchronised_every_step_steps_per_s'), e['collective'][:80], e['last_loss'], e.get('nccl_last_loss'))
PY
