echo "== pairs (default build)"; timeout 150 python profiles/warp2_check.py 4096 20 2>&1 | tail -13
echo "== sequential diagonal block (A/B build)"; VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_seq.so timeout 150 python profiles/warp2_check.py 4096 20 2>&1 | tail -6
