# final captures of the round: GPU tests, bench line, reference arm, launch list, ncu full of the warp kernel (fused), timelines
bash profiles/run/check.sh final 20
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pre_ncu.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 100 python profiles/prof_target_warp.py adj > gpurun_out/plain_warp2.log 2>&1 && timeout 250 ncu --set full --clock-control none --import-source on -k regex:fem_warp2 -s 1 -c 1 -o gpurun_out/prof_warp2_adj python profiles/prof_target_warp.py adj > gpurun_out/ncu_warp2_adj.log 2>&1; tail -1 gpurun_out/ncu_warp2_adj.log
export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
timeout 100 python profiles/timeline_warp.py 4096 adj > gpurun_out/tl_warp2_adj.log 2>&1
timeout 100 python profiles/timeline_warp.py 148 adj > gpurun_out/tl_warp2_adj_1persm.log 2>&1
timeout 100 python profiles/timeline_warp.py 4096 fwd > gpurun_out/tl_warp2_fwd.log 2>&1
unset VBFEM_LIB
timeout 200 python profiles/misc_timings.py > gpurun_out/misc_timings.log 2>&1; tail -3 gpurun_out/misc_timings.log
