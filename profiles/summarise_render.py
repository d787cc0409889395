"""profiles/run_ncu_render.sh's capture -> profiles/<round>_render_ncu_summary.md (CPU box, no GPU needed).
Usage: python profiles/summarise_render.py r02"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
rep = os.path.join(ROOT, "gpurun_out", f"{R}_render.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 global-load requests"),
        ("lts__t_requests.sum", "L2 requests"), ("lts__t_sectors.sum", "L2 sectors"),
        ("sm__cycles_elapsed.max", "SM cycles")]
kn = hdr.index("Kernel Name")
with open(os.path.join(ROOT, "profiles", f"{R}_render_ncu_summary.md"), "w") as f:
    f.write(f"# {R}: whole-ray renderer, `ncu --set full --clock-control none` (profiles/run_ncu_render.sh)\n\n"
            "One 800x800 frame (640,000 rays) of the c2-like scene after 600 training steps: the pre-pass and the persistent\n"
            "kernel.  Times under ncu are serialised and cold-cache; the event-timed numbers are in profiles/README.md.\n\n")
    for r in data:
        name = r[kn].split("(")[0].replace("void ", "")
        f.write(f"## {name}\n\n| metric | value |\n|---|---|\n")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                f.write(f"| {label} | {r[i]} {units[i]} |\n")
        f.write("\n")
print(open(os.path.join(ROOT, "profiles", f"{R}_render_ncu_summary.md")).read())
