"""Turn one `ncu --set full` report into the three tracked summaries: <out>_details.txt (details page),
<out>_raw_selected.txt (the counters DESIGN.md quotes + stall cycles per issued instruction) and
<out>_top_sass.txt (top SASS lines by stall samples with their source lines, via top_sass.py).
  python profiles/summarise_ncu.py <report.ncu-rep> <out prefix> <kernel substring for nvdisasm> "<header line>" """
import csv, io, os, subprocess, sys, tempfile

rep, out, kname, header = sys.argv[1:5]
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "variational-bayesian-inference-for-computational-mechanics_b200", "csrc", "libvbfem.so")
SELECTED = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def run(*a):
    return subprocess.run(a, check=True, capture_output=True, text=True).stdout


open(out + "_details.txt", "w").write(run("ncu", "-i", rep, "--page", "details"))
rows = list(csv.reader(io.StringIO(run("ncu", "-i", rep, "--page", "raw", "--csv"))))
names, units, vals = rows[0], rows[1], rows[2]
tab = {n: (u, v) for n, u, v in zip(names, units, vals)}
with open(out + "_raw_selected.txt", "w") as f:
    f.write("# " + header + "\n")
    for k in SELECTED:
        if k in tab:
            f.write(f"{k:100s} {tab[k][0]:10s} {tab[k][1]}\n")
    f.write("# stall cycles per issued instruction (> 0.15)\n")
    for k in sorted(tab):
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            try:
                v = float(tab[k][1].replace(",", ""))
            except ValueError:
                continue
            if v > 0.15:
                f.write(f"{k:100s} {v:.6f}\n")
with tempfile.TemporaryDirectory() as td:
    src = os.path.join(td, "src.csv")
    open(src, "w").write(run("ncu", "-i", rep, "--page", "source", "--csv"))
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(LIB)], cwd=td, check=True, capture_output=True)
    cubins = [os.path.join(td, x) for x in os.listdir(td) if x.endswith(".cubin")]
    dis = os.path.join(td, "dis.txt")
    open(dis, "w").write(run("nvdisasm", "-g", "-c", *cubins))
    open(out + "_top_sass.txt", "w").write(run(sys.executable, os.path.join(HERE, "top_sass.py"), src, dis, kname, "40"))
print("wrote", out + "_{details,raw_selected,top_sass}.txt")
