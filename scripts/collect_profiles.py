"""Turns the files `scripts/final_capture_1gpu.sh <prefix>` left in gpurun_out/ into the tracked evidence under profiles/:
bench lines, phase timers, DRAM csvs, launch list + its summary, and ncu summaries / hot lines read from the .ncu-rep files
(needs `ncu` on PATH; runs on the CPU-only build box).   python scripts/collect_profiles.py r3"""
import collections, csv, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = sys.argv[1] if len(sys.argv) > 1 else "r3"
G, O = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def cp(name):
    src = os.path.join(G, f"{P}_{name}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(O, f"{P}_{name}"))


for n in ["bench_c3_1gpu.json", "bench_c4_1gpu.json", "bench_c2_1gpu.json", "bench_c1_1gpu.json", "bench_reference_arm_c3.json", "bench_c3_2gpu.json",
          "bench_c3_4gpu.json", "bench_c3_8gpu.json", "bench_c5_8gpu.json", "dram_c3_fullM.csv", "dram_c5_M1184.csv", "launches_c3.csv",
          "phase_timers_C2.txt", "phase_timers_C3.txt", "phase_timers_C4.txt", "phase_timers_C5.txt"]:
    cp(n)

# launch list summary
rows = list(csv.DictReader(l for l in open(os.path.join(O, f"{P}_launches_c3.csv")) if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r["Kernel Name"].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"]) / 1e6
tot = sum(a[1] for a in agg.values())
with open(os.path.join(O, f"{P}_launch_list_summary_c3.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu-baseline (C3, 1 GPU, FINAL build of round 2)\n")
    f.write("cold-cache, serialised launches: compare SHARES, not absolutes\n")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n:60s} launches={a[0]:4d} total_ms={a[1]:12.3f} share={100 * a[1] / tot:6.2f}%\n")

# ncu summaries + hot lines
for c, cmd in (("c3", "python scripts/probe.py C3 1184 --reps 1  (C3 problem shape, M = 1184 = 8 trajectories per CTA; rbo_rollout_kernel)"),
               ("c5", "python scripts/probe.py C5 296 --reps 1  (C5 problem shape n=1000 d=20 h=2, M = 296; large-n variant rbo_rollout_kernel_largen)")):
    rep = os.path.join(G, f"{P}_{c}_rollout.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(f"/tmp/{P}_{c}_raw.csv", "w").write(raw)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    open(f"/tmp/{P}_{c}_src.csv", "w").write(src)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), f"/tmp/{P}_{c}_raw.csv"], capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_hot_lines.py"), f"/tmp/{P}_{c}_src.csv", "24"], capture_output=True, text=True).stdout
    plain = open(os.path.join(G, f"{P}_plain_{c}.log")).read().strip()
    notes = os.path.join(O, f"{P}_{c}_ncu_notes.txt")
    with open(os.path.join(O, f"{P}_{c}_ncu_summary.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on -k regex:rbo_rollout_kernel -c 1: {cmd}; B200, FINAL build of round 2\n")
        f.write("plain run of the same command before the capture: " + plain[plain.index("kernel_ms"):] + "\n")
        f.write(summ)
        if os.path.exists(notes):
            f.write(open(notes).read())
    open(os.path.join(O, f"{P}_{c}_rollout_hot_lines.txt"), "w").write(hot)
print("profiles updated for prefix", P)
