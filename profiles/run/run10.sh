timeout 150 python profiles/warp2_check.py 4096 20 2>&1 | tail -12
export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
timeout 100 python profiles/timeline_warp.py 4096 adj > gpurun_out/tl_warp2_adj.log 2>&1
timeout 100 python profiles/timeline_warp.py 148 adj > gpurun_out/tl_warp2_adj_1persm.log 2>&1
grep "reverse pass\|mean" gpurun_out/tl_warp2_adj.log gpurun_out/tl_warp2_adj_1persm.log
