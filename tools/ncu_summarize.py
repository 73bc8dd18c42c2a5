"""Turn the ncu artefacts a gpurun call brought back into the small text/JSON summaries committed under profiles/.
  python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/r01_launches_v3.txt
  python tools/ncu_summarize.py full gpurun_out/decode_v3_full.ncu-rep profiles/r01_decode_v3_ncu_full.txt [profiles/decode_traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("ns") else v * 1e3 if unit.startswith("ms") else v
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {n} launches, {tot / 1e3:.1f} ms total",
           f"# source: {src}", f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share %':>8s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k[:70]:70s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:10.2f} {v[1] / tot * 100:8.2f}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:12]))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]  # fmt: skip


def full(src, dst, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"# ncu --set full --clock-control none --import-source on; {len(data)} captured launch(es) of {data[0][hdr.index('Kernel Name')].split('(')[0]}",
           f"# source: {src} (ncu -i ... --page raw --csv)"]
    vals = {}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals[w] = (units[i], [r[i] for r in data])
            out.append(f"{w:90s} {units[i]:14s} " + "  ".join(r[i] for r in data))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    rd = [float(v) * scale[vals["dram__bytes_read.sum"][0]] for v in vals["dram__bytes_read.sum"][1]]
    wr = [float(v) * scale[vals["dram__bytes_write.sum"][0]] for v in vals["dram__bytes_write.sum"][1]]
    tr = sum(a + b for a, b in zip(rd, wr)) / len(rd)
    out.append(f"dram traffic per launch (read + write, mean): {tr:.0f} bytes")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    if traffic_json:
        json.dump({"dram_bytes_per_launch": tr, "source": dst}, open(traffic_json, "w"))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
