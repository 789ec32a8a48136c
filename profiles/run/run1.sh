set -x
export VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_tl.so
python profiles/timeline_panel.py 296 fwd > gpurun_out/tl80_fwd.log 2>&1
python profiles/timeline_panel.py 296 adj > gpurun_out/tl80_adj.log 2>&1
unset VBFEM_LIB
VBFEM_FORCE_PANEL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_panel20.log 2>&1
python profiles/prof_target_80.py 296 adj 2 > gpurun_out/plain80.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fem_panel -s 1 -c 1 -o gpurun_out/prof80_adj python profiles/prof_target_80.py 296 adj 2 > gpurun_out/ncu80.log 2>&1
tail -2 gpurun_out/ncu80.log
