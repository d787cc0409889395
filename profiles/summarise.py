"""Turn the raw ncu outputs of profiles/run_ncu.sh (gpurun_out/<round>_launches.csv, <round>_prof.ncu-rep,
<round>_ncu_units.json) into the small tracked summaries profiles/<round>_launch_list_summary.md,
profiles/<round>_ncu_summary.md and -- with `--metrics-json` -- profiles/ncu_metrics.json: the per-unit DRAM / L2 / L1
counters of every hot kernel that bench.py scales to its own unit counts (instead of constants typed in by hand).
Usage (CPU box, no GPU needed):  python profiles/summarise.py r02 [--metrics-json]"""
import collections
import csv
import json
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
        units_p = os.path.join(OUT, f"{R}_ncu_units.json")
        units = json.load(open(units_p)) if os.path.exists(units_p) else {}
        f.write(f"# {R}: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, see profiles/run_ncu.sh)\n\n"
                f"Command: the one in profiles/run_ncu.sh; captured region = {units.get('steps', '?')} eager training steps at steady state "
                f"({units.get('samples', '?')} samples, {units.get('alive', '?')} alive) + the two memory probes.\n"
                "Per-launch times are cold-cache and serialised -- compare SHARES with bench.py's `kernels_us`.\n\n"
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
           ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sectors.sum", "L2 sectors"), ("lts__t_requests.sum", "L2 requests")]


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


# kernel name (as ncu prints it) -> (C-ABI entry point bench.py times, unit the counters are divided by)
KMAP = [("hashgrid_fw_kernel", "b2n_hashgrid_fw", "sample"), ("hashgrid_bw_kernel", "b2n_hashgrid_bw", "alive_sample"),
        ("field_mlp_fw_kernel", "b2n_field_mlp_fw", "sample"), ("field_mlp_bw_kernel", "b2n_field_mlp_bw", "alive_sample"),
        ("composite_loss_fwbw_kernel", "b2n_composite_loss_fwbw", "sample"), ("adam_kernel", "b2n_adam_step", "param"),
        ("membench_gather_kernel", "b2n_membench_gather", "load"), ("membench_read_kernel", "b2n_membench_read", "byte")]
FIELDS = [("dram_bytes", ("dram__bytes_read.sum", "dram__bytes_write.sum")),
          ("lts_sectors", ("lts__t_sectors.sum",)), ("lts_requests", ("lts__t_requests.sum",)),
          ("lts_sectors_tex_read", ("lts__t_sectors_srcunit_tex_op_read.sum",)),
          ("lts_requests_tex_read", ("lts__t_requests_srcunit_tex_op_read.sum",)),
          ("l1_ld_sectors", ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",)),
          ("l1_ld_sectors_hit", ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",)),
          ("l1_ld_requests", ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",))]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "sector": 1.0, "request": 1.0, "": 1.0,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def metrics_json():
    """profiles/ncu_metrics.json: counters of the LAST captured launch of every hot kernel (steady state, warm L2 as far as
    ncu's replay allows) divided by the number of units that launch processed."""
    rep = os.path.join(OUT, f"{R}_prof.ncu-rep")
    units = json.load(open(os.path.join(OUT, f"{R}_ncu_units.json")))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, un = rows[0], rows[1]
    kn = hdr.index("Kernel Name")

    def val(r, metric):
        if metric not in hdr:
            return None
        i = hdr.index(metric)
        try:
            return float(r[i].replace(",", "")) * SCALE.get(un[i], 1.0)
        except ValueError:
            return None
    n_units = {"sample": units["samples"], "alive_sample": units["alive"], "param": units["params"],
               "load": units["gather_probe_loads"], "byte": units["read_probe_bytes"]}
    out = dict(round=R, config=units.get("config"), source=f"gpurun_out/{R}_prof.ncu-rep via profiles/summarise.py",
               units=units, kernels={})
    for kname, api, unit in KMAP:
        cand = [r for r in rows[2:] if short(r[kn]).startswith(kname)]
        if not cand:
            continue
        r = cand[-1]
        e = dict(unit=unit, units_in_launch=n_units[unit], time_us=val(r, "gpu__time_duration.sum"))
        for key, ms in FIELDS:
            vs = [val(r, m) for m in ms]
            e[key + "_per_unit"] = (sum(vs) / n_units[unit]) if all(v is not None for v in vs) else None
        if e.get("l1_ld_sectors_per_unit"):
            e["l1_hit_rate"] = e["l1_ld_sectors_hit_per_unit"] / e["l1_ld_sectors_per_unit"]
        e["l2_requests_per_s"] = e["lts_requests_per_unit"] * n_units[unit] / (e["time_us"] * 1e-6) if e["time_us"] and \
            e.get("lts_requests_per_unit") else None
        out["kernels"][api] = e
    with open(os.path.join(ROOT, "profiles", "ncu_metrics.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in out["kernels"].items()}, indent=1))


if __name__ == "__main__":
    if os.path.exists(os.path.join(OUT, f"{R}_launches.csv")):      # the render capture has no launch list
        launch_list()
    full()
    if "--metrics-json" in sys.argv:
        metrics_json()
    print("written", R)
