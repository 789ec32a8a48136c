timeout 100 python profiles/prof_target_warp.py adj > gpurun_out/plain_warp.log 2>&1 && timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp -s 1 -c 1 -o gpurun_out/prof_warp_adj_final2 python profiles/prof_target_warp.py adj > gpurun_out/ncu_warp_adj_final2.log 2>&1; tail -1 gpurun_out/ncu_warp_adj_final2.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pre_ncu.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
timeout 100 python profiles/timeline_warp.py 4096 adj > gpurun_out/tl_warp_adj.log 2>&1
timeout 100 python profiles/timeline_warp.py 148 adj > gpurun_out/tl_warp_adj_1persm.log 2>&1
timeout 100 python profiles/timeline_warp.py 4096 fwd > gpurun_out/tl_warp_fwd.log 2>&1
