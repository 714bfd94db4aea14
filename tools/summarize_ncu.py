"""Turn gpurun_out ncu artefacts into the small, tracked summaries under profiles/."""
import collections
import csv
import re
import subprocess
import sys

def short(name):
    name = name.replace("void heon::", "").replace("heon::", "")
    return re.sub(r"\(.*", "", re.sub(r"<.*", "", name)).strip()

def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum" or "at::" in r[ix["Kernel Name"]]:
            continue
        a = agg.setdefault(short(r[ix["Kernel Name"]]), [0, 0.0])
        a[0] += 1; a[1] += float(r[ix["Metric Value"]]) / 1000.0
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,avg_us,share_of_step\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k},{n},{t:.1f},{t/n:.1f},{t/tot:.3f}\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write("kernel," + ",".join(w for w in WANT if w in ix) + ",top_stalls\n")
        for r in rows[2:]:
            st = sorted([(float(r[ix[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls], reverse=True)[:4]
            f.write(short(r[ix["Kernel Name"]]) + "," + ",".join(f"{r[ix[w]]} {units[ix[w]]}".strip() for w in WANT if w in ix)
                    + "," + " ".join(f"{n}={v:.2f}" for v, n in st) + "\n")

if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (launches if kind == "launches" else full)(src, dst)
