"""Turns an .ncu-rep (brought back in gpurun_out/) into the small JSON summaries committed under profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.json> [--what "..."] [--bytes ALGORITHMIC_BYTES_PER_LAUNCH] [--kernel REGEX]
One entry per profiled launch: duration, DRAM bytes, throughput, occupancy, issue / pipe utilisation, stall reasons and,
with --bytes, the achieved algorithmic bandwidth and its fraction of MEASURED_PEAKS.json's HBM peak."""
import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
    "lts__t_sectors_srcunit_tex_aperture_peer.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--what", default="")
    ap.add_argument("--bytes", type=float, default=None, help="algorithmic bytes per launch")
    ap.add_argument("--kernel", default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    peak = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    launches = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        name = rec.get("Kernel Name", "")
        if a.kernel and not re.search(a.kernel, name):
            continue
        e = {"kernel": name}
        for k in KEEP:
            if k in rec and rec[k] != "":
                e[k] = {"value": rec[k], "unit": units[hdr.index(k)]}
        stalls = {}
        for k in hdr:
            m = STALL.match(k)
            if m and rec.get(k, "") not in ("", "0"):
                stalls[m.group(1)] = round(float(rec[k]), 3)
        e["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        try:
            ms = float(rec["gpu__time_duration.sum"]) * TO_MS[units[hdr.index("gpu__time_duration.sum")]]
            rd = float(rec["dram__bytes_read.sum"]) * TO_BYTES[units[hdr.index("dram__bytes_read.sum")]]
            wr = float(rec["dram__bytes_write.sum"]) * TO_BYTES[units[hdr.index("dram__bytes_write.sum")]]
            e["duration_ms"] = ms
            e["dram_bytes_per_launch"] = rd + wr
            e["dram_gbs"] = (rd + wr) / (ms * 1e-3) / 1e9
            if a.bytes:
                e["algorithmic_bytes_per_launch"] = a.bytes
                e["achieved_gbs"] = a.bytes / (ms * 1e-3) / 1e9
                e["traffic_over_algorithmic"] = (rd + wr) / a.bytes
                if peak:
                    e["frac_of_measured_hbm_peak"] = e["achieved_gbs"] / peak
        except (KeyError, ValueError):
            pass
        launches.append(e)
    out = {"_what": a.what, "_report": os.path.basename(a.report), "measured_hbm_peak_gbs": peak, "launches": launches}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    for e in launches:
        print(e["kernel"][:60], e.get("duration_ms"), e.get("dram_gbs"), e.get("frac_of_measured_hbm_peak"))


if __name__ == "__main__":
    main()
