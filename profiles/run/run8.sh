# diagnose the two-GPU peer worker: run it outside pytest with a traceback dump after 60 s
python - <<'PY'
import sys
sys.path.insert(0, "tests")
import test_gpu_peer as t
src = "import faulthandler, sys\nfaulthandler.dump_traceback_later(60, repeat=False, file=sys.stderr)\n" + t._WORKER.format(root="/root/repo")
open("/tmp/peer_worker.py", "w").write(src)
PY
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29631 VBFEM_PEER_TIMEOUT_MS=5000 WORLD_SIZE=2
RANK=0 timeout 150 python /tmp/peer_worker.py > gpurun_out/peer_w0.log 2>&1 &
RANK=1 timeout 150 python /tmp/peer_worker.py > gpurun_out/peer_w1.log 2>&1 &
wait
tail -5 gpurun_out/peer_w0.log; tail -5 gpurun_out/peer_w1.log
