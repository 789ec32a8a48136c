# warp kernel 2: ncu full capture of one fused launch, then the clock64 phase budget (profiling build)
timeout 100 python profiles/prof_target_warp.py adj > gpurun_out/plain_warp2.log 2>&1 && timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp2 -s 1 -c 1 -o gpurun_out/prof_warp2_adj python profiles/prof_target_warp.py adj > gpurun_out/ncu_warp2_adj.log 2>&1; tail -1 gpurun_out/ncu_warp2_adj.log
export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
timeout 100 python profiles/timeline_warp.py 4096 adj > gpurun_out/tl_warp2_adj.log 2>&1
timeout 100 python profiles/timeline_warp.py 148 adj > gpurun_out/tl_warp2_adj_1persm.log 2>&1
timeout 100 python profiles/timeline_warp.py 4096 fwd > gpurun_out/tl_warp2_fwd.log 2>&1
cat gpurun_out/tl_warp2_adj.log
