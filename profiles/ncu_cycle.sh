#!/bin/bash
# usage: profiles/ncu_cycle.sh <tag>   -- ncu --set full of the twisted kernel on 1184 samples, then per-line attribution
set -e
TAG=$1
cd /root/repo
/usr/local/graft/bin/gpurun --timeout 900 -- "python profiles/prof_target.py 1184 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fem_front -s 1 -c 1 -o gpurun_out/prof_$TAG python profiles/prof_target.py 1184 2 > gpurun_out/ncu_$TAG.log 2>&1; tail -n 1 gpurun_out/plain.log; tail -n 1 gpurun_out/ncu_$TAG.log" 2>&1 | tail -4
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv 2>/dev/null > /tmp/src_$TAG.csv
rm -rf /tmp/cub_$TAG && mkdir -p /tmp/cub_$TAG && cd /tmp/cub_$TAG
cuobjdump -xelf all /root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem.so >/dev/null 2>&1
nvdisasm -g -c *.cubin > dis.txt 2>/dev/null
cd /root/repo
python profiles/attribute_samples.py /tmp/src_$TAG.csv /tmp/cub_$TAG/dis.txt fem_front_kernelILi25ELi128ELi1E ${2:-30}
