"""Summarise a per-pass trace of a pooled run (stderr of any run with HEROAM_TRACE=1, e.g.
`HEROAM_TRACE=1 N=4000000 python scripts/trace_pool.py 2> trace.txt`): cumulative time, pass time and
bin populations at a few passes, and where the run switches to the drain regime.

    python scripts/analyze_trace.py trace.txt [pass numbers ...]
"""
import re
import sys

PAT = re.compile(r"pass (\d+) live (\d+) (?:behind \d+ )?bins ((?:\d+ )+) ?([\d.]+) ms  psteps (\d+) free (\d+)")
# bins as printed (from bin 1): 1..12 thread/hybrid kernels by moving count, 13 smem16, 14-16 warp kernels,
# 17-19 whole-trajectory free flight (short, medium, long), 20 split-off loners, 21.. cores of split trajectories


def load(path):
    runs, cur = [], []
    for line in open(path):
        m = PAT.match(line)
        if not m:
            continue
        row = (int(m.group(1)), int(m.group(2)), [int(x) for x in m.group(3).split()], float(m.group(4)), int(m.group(5)), int(m.group(6)))
        if row[0] == 1 and cur:
            runs.append(cur)
            cur = []
        cur.append(row)
    if cur:
        runs.append(cur)
    return runs


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    runs = load(sys.argv[1])
    run = max(runs, key=len)  # the warm-up run comes first and is short
    wanted = [int(x) for x in sys.argv[2:]] or sorted({1, len(run) // 8, len(run) // 4, len(run) // 2, 3 * len(run) // 4, len(run)} - {0})
    total = sum(r[3] for r in run)
    print(f"{len(run)} passes, {total / 1e3:.2f} s in the integrate kernels, {run[-1][4]:.4g} particle-steps "
          f"({run[-1][5] / max(1, run[-1][4]):.0%} in free flight)")
    cum, drain_at = 0.0, None
    for r in run:
        b = r[2]
        in_kernels = sum(b[:16]) + sum(b[20:])
        if drain_at is None and r[0] > 1 and in_kernels < 148 * 768:
            drain_at = (r[0], cum)
        cum += r[3]
        if r[0] in wanted:
            print(f"  pass {r[0]:5d}  t = {cum / 1e3:6.2f} s  {r[3]:6.1f} ms  narrow 2/3/4 = {b[1]}/{b[2]}/{b[3]}  wider = {sum(b[4:16])}  "
                  f"flights s/m/l = {b[16]}/{b[17]}/{b[18]}  loners = {b[19]}  cores = {sum(b[20:])}")
    if drain_at:
        print(f"  fewer than one wave of trajectories from pass {drain_at[0]} on (t = {drain_at[1] / 1e3:.2f} s): "
              f"the drain takes {(total - drain_at[1]) / 1e3:.2f} s = {(total - drain_at[1]) / total:.0%} of the run")


if __name__ == "__main__":
    main()
