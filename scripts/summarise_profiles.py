"""Turns gpurun_out/launches_*.csv and gpurun_out/prof_*.ncu-rep into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", f"prof_coarse_{tag}b.ncu-rep")
out = {}
if os.path.exists(launches):
    rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("=="))]
    hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict(); n_per = collections.Counter(); order = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum": continue
        name = r[ci["Kernel Name"]].split("(")[0]
        v = float(r[ci["Metric Value"]].replace(",", "")); unit = r[ci["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        per[name] = per.get(name, 0.0) + us; n_per[name] += 1
    tot = sum(per.values())
    with open(os.path.join(ROOT, "profiles", f"launches_{tag}_summary.md"), "w") as f:
        f.write(f"# Launch list summary ({tag})\n\nCommand: `" + os.environ.get("LAUNCH_CMD", "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline") + "`\n"
                "(B200, per-launch times are cold-cache and serialised: compare SHARES). Includes the database upload kernels of the setup.\n\n"
                "| kernel | launches | total us | share | us / launch |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(per.items(), key=lambda x: -x[1]):
            f.write(f"| `{k}` | {n_per[k]} | {v:.1f} | {100*v/tot:.1f}% | {v/n_per[k]:.1f} |\n")
        f.write(f"\ntotal {tot:.1f} us over {sum(n_per.values())} launches\n")
    print(open(os.path.join(ROOT, "profiles", f"launches_{tag}_summary.md")).read())
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
    with open(os.path.join(ROOT, "profiles", f"ncu_k_match_coarse_{tag}.md"), "w") as f:
        f.write(f"# ncu --set full, k_match_coarse ({tag})\n\nCommand: `" + os.environ.get("FULL_CMD", "ncu --set full --clock-control none --import-source on -k regex:k_match_coarse -s 2 -c 1 python scripts/gpu_match_bench.py") + "` "
                "(" + os.environ.get("FULL_WHAT", "1 M descriptors x 2000 queries") + ", k=4, B200). Times under the profiler are not bench values.\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in keys:
            if k in d: f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        h = srows[1]; ci = {x: i for i, x in enumerate(h)}
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot = collections.Counter(); top = []
        for r in srows[2:]:
            s = float(r[ci["# Samples"]])
            for k in stalls: tot[k] += float(r[ci[k]])
            top.append((s, r[ci["Source"]].strip(), r[ci["Instructions Executed"]]))
        f.write("\n## warp stall samples (whole kernel)\n\n| reason | samples |\n|---|---:|\n")
        for k, v in tot.most_common(8): f.write(f"| {k} | {int(v)} |\n")
        f.write("\n## hottest SASS instructions\n\n| samples | executed | SASS |\n|---:|---:|---|\n")
        for s, srcl, ex in sorted(top, key=lambda x: -x[0])[:14]: f.write(f"| {int(s)} | {ex} | `{srcl[:90]}` |\n")
    print(open(os.path.join(ROOT, "profiles", f"ncu_k_match_coarse_{tag}.md")).read())
