"""Rebuild profiles/r1_ncu_summary.md from the ncu outputs brought back in gpurun_out/."""
import collections, csv, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
out = []
P = out.append
P("# Round 1 ncu summaries (B200, sm_100a, driver 580, CUDA 12.9)\n")
P("All captures with `--clock-control none`; numbers taken under ncu are never bench values. Kernel names: "
  "`Turbo` = the fast-mode arithmetic of the hot kernels, `Exact` = bit-exact mode (DESIGN.md section 4).\n")

def launch_table(path, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        agg[name][0] += 1
        agg[name][1] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
    tot = sum(v[1] for v in agg.values())
    P(f"`{cmd}`\n")
    P("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        P(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % |")
    integ = sum(v[1] for k, v in agg.items() if "integrate" in k or "free_flight" in k)
    P(f"\nIntegrate + free-flight kernels: {100 * integ / tot:.1f} % of the serialised GPU time of the first "
      f"{sum(v[0] for v in agg.values())} launches.\n")

def metric_table(path, title, cmd, keep=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "warp instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write")]
    P(f"## {title}\n\n`{cmd}`\n")
    P("| kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 1))
    seen = set()
    stall_lines = []
    st = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        key = (name, r[idx["launch__grid_size"]])
        if key in seen or (keep and not keep(name)):
            continue
        seen.add(key)
        cells = []
        for m, _ in want:
            v = r[idx[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {units[idx[m]]}".strip())
        P(f"| `{name}` | " + " | ".join(cells) + " |")
        vals = sorted(((float(r[idx[h]]), h) for h in st if r[idx[h]]), reverse=True)[:5]
        stall_lines.append(f"* `{name}` grid {key[1]}: " + ", ".join(
            f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in vals))
    P("\nTop warp stall reasons (cycles per issued instruction):\n")
    out.extend(stall_lines)
    P("")

P("## 1. Launch list of the bench command\n")
launch_table(os.path.join(G, "launches_r1b.csv"),
             "ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 1 "
             "--batch 131072 --max-time 1e6 --no-cpu --exact-step 0 --e2e-steps 1 --warmup-e2e 0")
P("Raw list: `r1_launches_bench_small.csv`. Under ncu the kernels run one at a time and cold; in the real run the "
  "kernels of one pass execute concurrently on separate streams. `bench.py` reports `integrate_share_of_step` = 0.99 for "
  "the same command from CUDA events.\n")
metric_table(os.path.join(G, "prof_r1_turbo_full.ncu-rep"), "2. Dominant kernel at full occupancy (`--set full`)",
             "NS=4,8 MODES=fast ncu --set full --clock-control none --import-source on -k regex:k_integrate_thread -c 6 "
             "python scripts/bench_bins.py   (151552 trajectories with exactly 4 (8) moving particles, 4096-step cap)",
             keep=lambda n: "<4," in n or "<8," in n)
metric_table(os.path.join(G, "prof_r1_final.ncu-rep"), "3. The kernels inside the bench command (131072-slot pool, under-filled GPU)",
             "ncu --set full --clock-control none --import-source on -k regex:'k_integrate_thread|k_free_flight|k_integrate_hybrid' "
             "-s 30 -c 14 python bench.py --steps 2 --warmup 1 --batch 131072 --max-time 1e6 ...")
metric_table(os.path.join(G, "prof_r1a.ncu-rep"), "4. Before the Turbo arithmetic and the hybrid kernels (same full-occupancy benchmark, earlier in the round)",
             "NS=4,16 MODES=fast ncu --set full ... python scripts/bench_bins.py", keep=lambda n: "<4," in n or "warp" in n)
P("""## Reading

* `k_integrate_thread<4, Turbo>` at full occupancy is FP32-issue bound: issue slots ~80 % busy, top stall `not_selected`
  (more eligible warps than issue slots), L2 hit > 97 %, DRAM traffic tens of MB per launch for 2.5e9 particle-steps (HBM idle:
  state in registers, table in L1/L2). Section 4 shows the same kernel before the Turbo arithmetic: 444 SASS instructions per
  step; now 336 (142 FFMA, 64 FMUL, 43 FADD+..., 10 MUFU, 6 LDG, 6 F2I, 6 VIMNMX).
* achieved (full occupancy, A = 4) = (100*4 + 10*6 + 14*1.0) flops x 19.35 M warp-steps x 32 / duration = 0.51 of the
  72.4 TFLOP/s FFMA peak measured in the same process; what is left is instruction mix (FMUL/FADD count one flop per issue
  slot, ~20 % of the slots are the table lookup's integer/convert work and register moves).
* `k_free_flight<Turbo>`: 26.5 instructions per particle-step (4-step unrolled), a single dependent chain per thread.
* `k_integrate_warp<1, Fast>` of section 4 (n = 16: 13.3e9 warp instructions) is what the hybrid kernels replaced for 9..12
  moving particles (6-10x) and what partner lists now thin out above 16.
""")
open(os.path.join(ROOT, "profiles", "r1_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
