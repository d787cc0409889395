"""Turn the raw ncu outputs of profiles/run_ncu.sh (gpurun_out/<round>_launches.csv, <round>_prof.ncu-rep) into the
small tracked summaries profiles/<round>_launch_list_summary.md and profiles/<round>_ncu_summary.md.
Usage (CPU box, no GPU needed):  python profiles/summarise.py r01c"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "gpurun_out")


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()


def launch_list():
    rows = list(csv.reader(open(os.path.join(OUT, f"{R}_launches.csv"))))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mn, mu, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0.0, 0])
    for r in data:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] in ("ns", "nsecond") else v * 1000 if r[mu] in ("ms", "msecond") else v
        a = agg[short(r[kn])]; a[0] += v; a[1] += 1
    tot = sum(v[0] for v in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{R}_launch_list_summary.md"), "w") as f:
        f.write(f"# {R}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, see profiles/run_ncu.sh)\n\n"
                "Command: `python bench.py --steps 6 --warmup 3 --pretrain 96 --no-graph --skip-cpu` (early training, ~580k samples/step,\n"
                "dense gradients; per-launch times are cold-cache and serialised -- compare SHARES with bench.py's `kernels_us`).\n\n"
                "| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|\n")
        for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {k} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |\n")


METRICS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
           ("l1tex__t_sector_hit_rate.pct", "L1 hit %")]


def full():
    rep = os.path.join(OUT, f"{R}_prof.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    with open(os.path.join(ROOT, "profiles", f"{R}_ncu_summary.md"), "w") as f:
        f.write(f"# {R}: `ncu --set full --clock-control none --import-source on` of one training step's kernels\n\n"
                "Same command as the launch list.  One row per captured launch (first capture of each kernel).\n\n| kernel | "
                + " | ".join(m[1] for m in METRICS) + " |\n|---|" + "---|" * len(METRICS) + "\n")
        seen = set()
        for r in rows[2:]:
            name = short(r[kn])
            if "march_test" in name or "composite_test" in name:       # render rounds differ a lot: keep every capture
                gi = hdr.index("launch__grid_size") if "launch__grid_size" in hdr else None
                name = f"{name} #{sum(1 for x in seen if x.startswith(name))}"
            if name in seen:
                continue
            seen.add(name)
            vals = []
            for m, _ in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    v = r[i]
                    try:
                        v = f"{float(v.replace(',', '')):.1f}"
                    except ValueError:
                        pass
                    vals.append(f"{v} {units[i]}".replace("register/thread", "").replace("Mbyte", "MB").strip())
                else:
                    vals.append("-")
            f.write(f"| {name} | " + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    if os.path.exists(os.path.join(OUT, f"{R}_launches.csv")):      # the render capture has no launch list
        launch_list()
    full()
    print("written", R)
