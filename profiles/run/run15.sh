echo "== 2x2 pivot blocks (default build)"; timeout 100 python profiles/panel_check.py 1024 5 2>&1 | tail -9
echo "== column by column (A/B build)"; VBFEM_LIB=/root/repo/variational-bayesian-inference-for-computational-mechanics_b200/csrc/libvbfem_seq.so timeout 100 python profiles/panel_check.py 1024 5 2>&1 | tail -3
