# panel kernel 2 as the config-4 default: GPU tests, bench line, ncu full capture of one fused launch (296 samples)
bash profiles/run/check.sh p2 20
timeout 100 python profiles/prof_target_80.py 296 adj 2 > gpurun_out/plain80_p2.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:fem_panel2 -s 1 -c 1 -o gpurun_out/prof80_p2_adj python profiles/prof_target_80.py 296 adj 2 > gpurun_out/ncu80_p2.log 2>&1; tail -1 gpurun_out/ncu80_p2.log
timeout 100 python profiles/panel_check.py 1024 5 > gpurun_out/panel_check_p2.log 2>&1; tail -6 gpurun_out/panel_check_p2.log
