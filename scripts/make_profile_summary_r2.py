"""Rebuild profiles/r2_ncu_summary.md from the ncu outputs of round 2 brought back in gpurun_out/
(and copy the raw CSVs next to it)."""
import collections, csv, os, re, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
out = []
P = out.append


def src(f):
    """A capture as brought back in gpurun_out/, else the copy already committed under profiles/."""
    return os.path.join(G, f) if os.path.exists(os.path.join(G, f)) else os.path.join(PROF, f)


def read_rows(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]
    raise SystemExit(f"{path}: no header")


def short(name):
    name = re.sub(r"\((heroam::)?DevView.*", "", name).replace("void ", "").replace("heroam::", "")
    return re.sub(r"\((int|bool)\)", "", name)


def launch_table(path, cmd):
    hdr, rows = read_rows(path)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        a = agg[short(r[idx["Kernel Name"]])]
        a[0] += 1
        a[1] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[idx["Metric Unit"]], 1e-6)
    tot = sum(v[1] for v in agg.values())
    P(f"`{cmd}`\n")
    P("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        P(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % |")
    integ = sum(v[1] for k, v in agg.items() if "integrate" in k or "free_flight" in k)
    P(f"\nIntegrate + free-flight kernels: {100 * integ / tot:.1f} % of the serialised GPU time of the first "
      f"{sum(v[0] for v in agg.values())} launches ({tot / 1e3:.2f} s serialised).\n")


def metric_table(path, title, cmd):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "warp instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write")]
    P(f"## {title}\n\n`{cmd}`\n")
    P("| kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 1))
    st = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    stall_lines = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = short(r[idx["Kernel Name"]])
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
        stall_lines.append(f"* `{name}` grid {r[idx['launch__grid_size']]}: " + ", ".join(
            f"{h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in vals))
    P("\nTop warp stall reasons (cycles per issued instruction):\n")
    out.extend(stall_lines)
    P("")


def gather_table(paths, cmd):
    P(f"`{cmd}`\n")
    P("| table | kernel | grid | time | issue active | gather requests | sectors / request | L1 sector throughput | L1 hit | L2 throughput | L2 hit | "
      "DRAM read | DRAM read rate | of HBM peak (6 551 GB/s) |\n|" + "---|" * 14)
    for label, path in paths:
        hdr, rows = read_rows(path)
        idx = {h: i for i, h in enumerate(hdr)}
        per = collections.OrderedDict()
        for r in rows:
            key = (r[idx["ID"]], short(r[idx["Kernel Name"]]), r[idx["Grid Size"]].strip("()").split(",")[0])
            per.setdefault(key, {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
        for (_, name, grid), m in per.items():
            t = m["gpu__time_duration.sum"] * 1e-9
            req, sec = m["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"], m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
            P(f"| {label} | `{name}` | {grid} | {t * 1e3:.2f} ms | {m['smsp__issue_active.avg.pct_of_peak_sustained_active']:.0f} % | {req:.3g} | "
              f"{sec / req:.1f} | {m['l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum'] / t / 1e12:.2f} TB/s | {m['l1tex__t_sector_hit_rate.pct']:.1f} % | "
              f"{m['lts__t_bytes.sum'] / t / 1e12:.2f} TB/s | {m['lts__t_sector_hit_rate.pct']:.1f} % | {m['dram__bytes_read.sum'] / 1e9:.2f} GB | "
              f"{m['dram__bytes_read.sum'] / t / 1e9:.0f} GB/s | {100 * m['dram__bytes_read.sum'] / t / 6551e9:.1f} % |")
    P("")


P("# Round 2 ncu summaries (B200, sm_100a, driver 580, CUDA 12.9)\n")
P("All captures with `--clock-control none`, each after the same command had exited 0 without ncu; numbers taken under ncu are never "
  "bench values. `Turbo` = the fast-mode arithmetic of the hot kernels; the trailing `0/1` of `k_integrate_thread` = the L1-prefetch "
  "variant of the latency-bound regime, of `k_free_flight` = clock set at landing (DESIGN.md sections 3-4).\n")
P("## 1. Launch list of the bench command\n")
SMALL = "--steps 2 --warmup 3 --batch 131072 --max-time 1e6 --no-cpu --exact-step 0 --relaxed-step 0 --closed-step 0 --extras 0 --e2e-steps 1 --warmup-e2e 0"
launch_table(src("r2_launches_bench_small.csv"),
             "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py " + SMALL)
P("Raw list: `r2_launches_bench_small.csv`.  As in round 1 this reduced configuration (131 072 slots, max_time 1e6) over-weights the rare "
  "wide bins when serialised: each of their launches lasts as long as its dependency chain however few trajectories it holds, while in "
  "the real run the ~20 kernels of a pass execute concurrently on separate streams.\n")
P("## 2. Launch list of a full-size pool run (the bulk of a run: its first 9000 launches = ~460 passes)\n")
launch_table(src("r2_launches_pool.csv"),
             "N=2000000 MODE=fast ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv python scripts/pool_once.py")
P("Raw list: `r2_launches_pool.csv`.  The same run without ncu: 12.6 s for 1047 passes (5.03e11 particle-steps/s with its drain): "
  "serialised, its first 460 passes alone take 28.5 s -- the kernels of a pass overlap more than two-fold.\n")
FULL = ("N=2000000 CAP=40000 ncu --set full --clock-control none --import-source on -k regex:'k_integrate_thread|k_free_flight|"
        "k_pool_resolve|k_integrate_hybrid|k_integrate_smem' -s 17 -c 22 python scripts/pool_once.py   (2e6 trajectories of config C2 "
        "in the pool, passes 2-3)")
metric_table(src("r2_prof_pool_final_raw.csv"), "3. The hot kernels at full occupancy, inside a pool run (`--set full`), final register allocations", FULL)
P("Raw: `r2_prof_pool_final_raw.csv`.  Captured on commit babdd72, where the step body was written inside `k_integrate_thread`; the "
  "shipped build moves it into `integrate_thread_segment` (same arithmetic, same register allocations within 4, 2-5 % fewer "
  "instructions per step -- the loop sizes in `r2_ptxas.md` are the shipped ones: A = 2: 128 instead of 135, A = 4: 322 instead of 333, "
  "A = 8: 950 instead of 974); a 1e6-trajectory run takes 8.26-8.28 s with it against 8.29-8.46 s (`r2_forms.log`: inline form, segment form, twice).\n")
metric_table(src("r2_prof_pool_raw.csv"),
             "3b. The same capture of an intermediate build: the register finding (before)", FULL)
P("Raw: `r2_prof_pool_raw.csv`.  This build differed from the one above in the text around the step loop only (time-slice bookkeeping "
  "of a dropped experiment in the kernel's view structure): ptxas chose 96 / 128 / 168 registers for A = 3 / 4 / 6 where the final build has 112 / 146 / 228 "
  "(`r2_ptxas.md`).  Compare the rows of the two tables at equal grid: the smaller allocations keep fewer gathers in flight "
  "(long_scoreboard per instruction) and are slower although more warps are resident; the two-particle kernel is the opposite case "
  "(72 registers by `__launch_bounds__`, 80 when ptxas is left alone) and its cap was kept.\n")
P("## 4. The gather path: T-ref (6.4 MB, L1/L2-resident) against T-large (26.6 M entries = 106 MB)\n")
gather_table([("T-ref", src("r2_gather_pool_tlarge0.csv")), ("T-large", src("r2_gather_pool_tlarge1.csv"))],
             "TLARGE={0,1} N=2000000 CAP=40000 ncu --kernel-name-base demangled --metrics gpu__time_duration.sum,smsp__inst_executed.sum,"
             "l1tex__t_{requests,sectors,bytes}_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,"
             "lts__t_sector_hit_rate.pct,dram__bytes_{read,write}.sum,... -k regex:'k_integrate_thread<.int.[234],' -s 6 -c 6 "
             "python scripts/pool_once.py")
P(open(os.path.join(PROF, "r2_reading.md")).read() if os.path.exists(os.path.join(PROF, "r2_reading.md")) else "")
open(os.path.join(PROF, "r2_ncu_summary.md"), "w").write("\n".join(out) + "\n")
for f in ("r2_launches_bench_small.csv", "r2_launches_pool.csv", "r2_prof_pool_raw.csv", "r2_prof_pool_final_raw.csv",
          "r2_gather_pool_tlarge0.csv", "r2_gather_pool_tlarge1.csv"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(PROF, f))
print("\n".join(out))
