"""Summarise `ncu --set full` reports (read here, no GPU needed) into a small JSON for profiles/.
    python scripts/ncu_summary.py out.json name1=report1.ncu-rep[:kernel-regex] name2=...
Every entry keeps the metrics the roofline lines cite (duration, DRAM bytes, pipe / DRAM utilisation, occupancy)."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
           "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct"]


def rows(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    lines = [l for l in out.splitlines() if not l.startswith("==")]
    r = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def main():
    dst = sys.argv[1]
    result = {"units": {}}
    for arg in sys.argv[2:]:
        name, spec = arg.split("=", 1)
        report, _, pattern = spec.partition(":")
        hdr, units, data = rows(report)
        picked = []
        for row in data:
            d = dict(zip(hdr, row))
            if pattern and not re.search(pattern, d.get("Kernel Name", "")):
                continue
            entry = {"Kernel Name": re.sub(r"\(.*", "", d.get("Kernel Name", "")).replace("void ", "").replace(
                "stemseg::<unnamed>::", "").replace("unnamed>::", ""), "Grid Size": d.get("Grid Size", "")}
            for m in METRICS:
                if m in d:
                    val, unit = d[m], units[hdr.index(m)]
                    if unit.lower().endswith("byte") and unit != "Mbyte":          # one unit per file: Mbyte
                        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Gbyte": 1e3, "Tbyte": 1e6}[unit]
                        val, unit = "%.6f" % (float(val.replace(",", "")) * scale), "Mbyte"
                    entry[m] = val
                    result["units"][m] = unit
            picked.append(entry)
        result[name] = picked[0] if len(picked) == 1 else picked
    with open(dst, "w") as f:
        json.dump(result, f, indent=1)
    print(json.dumps(result, indent=1)[:3000])


if __name__ == "__main__":
    main()
