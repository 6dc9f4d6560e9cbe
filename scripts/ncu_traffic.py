"""kernels_full.ncu-rep (scripts/ncu_gemm.py under `ncu --set full`) -> per-kernel summary text + the DRAM traffic
per AWQ GEMM launch that bench.py reports as roofline.traffic.   python scripts/ncu_traffic.py REP OUT_TXT OUT_JSON"""
import csv, io, json, subprocess, sys
rep, out_txt, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def b(r, name):
    return float(r[idx[name]]) * SCALE[units[idx[name]]]
cols = ["gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
lines, awq = [], []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    rd, wr = b(r, "dram__bytes_read.sum"), b(r, "dram__bytes_write.sum")
    lines.append(f"{r[idx['ID']]:>3} {name[:58]:58s} dur {float(r[idx[cols[0]]]):8.2f} {units[idx[cols[0]]]}  dram rd {rd/1e6:9.3f} MB wr {wr/1e6:7.3f} MB  "
                 f"dram% {float(r[idx[cols[1]]]):5.1f} tensor% {float(r[idx[cols[2]]]):5.1f} issue% {float(r[idx[cols[3]]]):5.1f} "
                 f"fma% {float(r[idx[cols[4]]]):5.1f} alu% {float(r[idx[cols[5]]]):5.1f} regs {r[idx[cols[6]]]} grid {r[idx[cols[7]]]} block {r[idx[cols[8]]]}")
    if "gemm_kernel<16, 1>" in name:
        awq.append((rd, wr))
open(out_txt, "w").write("ncu --set full --clock-control none (cold caches, serialised): python scripts/ncu_gemm.py, M = 13\n"
                         "AWQ launches in order: q|k|v x2, o x2, gate|up x2, down x2; then lm_head x2, tree attention x2, (accept, gather) x2\n"
                         + "\n".join(lines) + "\n")
second = awq[1::2]                                   # the second launch of each shape
traffic = sum(rd + wr for rd, wr in second) / len(second)
json.dump({"awq_gemm_traffic_bytes_per_launch": traffic, "launches_averaged": len(second),
           "per_shape_read_bytes": [rd for rd, _ in second], "per_shape_write_bytes": [wr for _, wr in second],
           "source": out_txt + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(out_json, "w"), indent=1)
print("\n".join(lines)); print("traffic per AWQ launch", traffic)
