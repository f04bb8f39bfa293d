"""Digest of an ncu report for profiles/: header line, the raw-page metrics of tools/ncu_summary.py (plus the L2 / L1 byte
counters north_star names) and the per-function instruction shares of tools/ncu_funcs.py.
usage: python tools/ncu_digest.py REPORT.ncu-rep "header text" > profiles/NAME.txt"""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
rep, header = sys.argv[1], sys.argv[2]
print("# " + header)
print("# source: " + rep)
sys.stdout.flush()
out = subprocess.run([sys.executable, os.path.join(here, "ncu_summary.py"), rep], capture_output=True, text=True).stdout
print(out, end="")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
import csv
rows = list(csv.reader(raw.splitlines()))
if len(rows) > 2:
    hdr, units, r = rows[0], rows[1], rows[2]
    # achieved L2 traffic: `--set full` carries the L2 -> L1 return bytes and the L2 -> crossbar bytes (lts__t_bytes itself is not in the set)
    for k in ("l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "l1tex__m_l1tex2xbar_write_bytes.sum.per_second",
              "derived__lts__lts2xbar_bytes.sum.per_second", "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed",
              "LTS.TriageCompute.lts__average_t_sector_hit_rate_realtime.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
              "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_op_read.sum", "l1tex__t_bytes.sum", "l1tex__t_bytes.sum.per_second",
              "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "sm__inst_executed_pipe_fp32.avg.pct_of_peak_sustained_active",
              "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.pct"):
        if k in hdr:
            i = hdr.index(k)
            print("%-85s %-12s %s" % (k, units[i], r[i]))
print()
print("# executed warp instructions per function (tools/ncu_funcs.py; '__launch_bounds__' = the body of the kernel)")
print(subprocess.run([sys.executable, os.path.join(here, "ncu_funcs.py"), rep], capture_output=True, text=True).stdout, end="")
