"""Key metrics of every kernel in an `ncu --set full` report, as text (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "free-text header" > profiles/roundN/prof.summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__block_size", "launch__grid_size"]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
name_i = hdr.index("Kernel Name")
for n, r in enumerate(rows[2:]):
    print("launch %d: %s" % (n, r[name_i][:110]))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("    %-68s %-16s %s" % (w, units[i], r[i]))
