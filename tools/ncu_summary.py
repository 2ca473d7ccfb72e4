"""Text summary of an .ncu-rep (run where ncu is installed; no GPU needed):
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx_summary.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
print(f"# {rep}\nkernel: {d.get('Kernel Name', ('', '?'))[1]}   grid {d.get('launch__grid_size', ('', '?'))[1]} x block "
      f"{d.get('launch__block_size', ('', '?'))[1]}, regs/thread {d.get('launch__registers_per_thread', ('', '?'))[1]}, "
      f"dyn smem/block {d.get('launch__shared_mem_per_block_dynamic', ('', '?'))[1]} {d.get('launch__shared_mem_per_block_dynamic', ('', ''))[0]}")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_atom.sum"]
for w in want:
    if w in d:
        print(f"{w:70s} {d[w][1]:>18s} {d[w][0]}")
print("\n# warp stall reasons (cycles per issued instruction)")
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = float(d[h][1])
        if v >= 0.02:
            print(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:6.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None
hd = None
out = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hd = r
        continue
    if hd is None or len(r) < len(hd) or r[0] == "":
        continue
    try:
        samp = float(r[hd.index("# Samples")] or 0)
        inst = float(r[hd.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    a = out.setdefault((cur, int(r[0]), r[1].strip()[:100]), [0, 0])
    a[0] += inst
    a[1] += samp
ti = sum(v[0] for v in out.values()) or 1
ts = sum(v[1] for v in out.values()) or 1
print(f"\n# hottest source lines ({ti / 1e6:.1f}M warp instructions, {int(ts)} stall samples)")
for k, v in sorted(out.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{k[0][:18]:18s}:{k[1]:4d} inst {v[0] / ti * 100:5.1f}%  samples {v[1] / ts * 100:5.1f}%  | {k[2]}")
