"""Export the metrics the bench / DESIGN quote from an `ncu --set full` report into a small JSON.
    python tools/ncu_kernel_json.py report.ncu-rep "command line that produced it" > profiles/xxx.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * mult.get(unit, 1)


def main():
    rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                d[h] = f"{v} {u}".strip()
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                d[h + ".bytes"] = to_bytes(v, u)
        launches.append(d)
    traffic = [x["dram__bytes_read.sum.bytes"] + x["dram__bytes_write.sum.bytes"] for x in launches]
    print(json.dumps({"command": cmd, "report": rep, "dram_traffic_bytes_per_launch_mean": sum(traffic) / max(len(traffic), 1),
                      "launches": launches}, indent=1))


if __name__ == "__main__":
    main()
