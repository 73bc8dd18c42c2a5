"""Print selected metrics per launch from an `ncu --page raw --csv` export (one row per kernel launch).
usage: python tools/ncu_rows.py file.csv [metric substrings...]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__cycles_active.avg", "lts__t_bytes.sum"]  # fmt: skip
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = sys.argv[2:] or WANT
cols = [i for i, h in enumerate(hdr) if any(w in h for w in want)]
name_i = hdr.index("Kernel Name")
for r in rows[2:]:
    print(r[name_i].split("(")[0][-40:])
    for i in cols:
        print(f"    {hdr[i]:75s} {r[i]:>16s} {units[i]}")
