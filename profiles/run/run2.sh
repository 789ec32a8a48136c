export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
timeout 100 python profiles/timeline_warp.py 4096 fwd
timeout 100 python profiles/timeline_warp.py 4096 adj
unset VBFEM_LIB
timeout 100 python profiles/prof_target_warp.py && timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp -s 1 -c 1 -o gpurun_out/prof_warp_adj python profiles/prof_target_warp.py > gpurun_out/ncu_warp.log 2>&1; tail -2 gpurun_out/ncu_warp.log
