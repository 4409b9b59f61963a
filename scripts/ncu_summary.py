"""Summarise an `ncu --set full` report: per kernel (averaged over the captured launches) duration, DRAM bytes, pipe
utilisation, instruction counts.  usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.csv [TRAFFIC.json]
The traffic JSON ({kernel: {dram_bytes_per_launch, launches_captured}}) is what bench.py reports as `roofline.traffic`."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    ("us", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_pct_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
    ("lts_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("inst_M", "smsp__inst_executed.sum"), ("regs", "launch__registers_per_thread"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
]
SCALE = {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").strip()
        a = agg.setdefault(name, {"n": 0})
        a["n"] += 1
        for key, m in METRICS:
            if m not in col or r[col[m]] in ("", "n/a"):
                continue
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if key.endswith("_MB") or key == "us":
                v *= SCALE.get(u, 1.0)
            if key == "inst_M":
                v *= 1e-6
            a[key] = a.get(key, 0.0) + v
    keys = [k for k, _ in METRICS]
    with open(out, "w") as f:
        f.write("kernel,launches," + ",".join(keys) + "\n")
        for name, a in agg.items():
            f.write(name + "," + str(a["n"]) + "," + ",".join("%.3f" % (a.get(k, float("nan")) / a["n"]) for k in keys) + "\n")
    print(open(out).read())
    if len(sys.argv) > 3:
        try:
            tj = json.load(open(sys.argv[3]))
        except Exception:  # noqa: BLE001
            tj = {}
        for name, a in agg.items():
            base = re.sub(r"<.*", "", name)
            tj[base] = {"dram_bytes_per_launch": 1e6 * (a.get("dram_read_MB", 0) + a.get("dram_write_MB", 0)) / a["n"],
                        "launches_captured": a["n"]}
        json.dump(tj, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
