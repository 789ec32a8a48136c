export VBFEM_WARP_NW=12 VBFEM_WARP_BATCH=30
timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp -s 1 -c 1 -o gpurun_out/prof_warp_fwd python profiles/prof_target_warp.py fwd > gpurun_out/ncu_warp_fwd.log 2>&1; tail -1 gpurun_out/ncu_warp_fwd.log
timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp -s 1 -c 1 -o gpurun_out/prof_warp_adj2 python profiles/prof_target_warp.py adj > gpurun_out/ncu_warp_adj2.log 2>&1; tail -1 gpurun_out/ncu_warp_adj2.log
