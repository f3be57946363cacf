"""SASS evidence for the hot kernels (no GPU needed): cuobjdump of the built library ->
profiles/r2_sass_<kernel>.txt (full listing, the per-step loop marked, opcode histogram of that loop)
and profiles/r2_ptxas.md (registers / stack / shared memory of every kernel).

    python scripts/make_sass_summary.py
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sphere_roaming_b200", "libheroam_gpu.so")
OUT = os.path.join(ROOT, "profiles")

KERNELS = {
    "k_free_flight_Turbo": "_ZN6heroam13k_free_flightINS_5TurboELb1EEEvNS_7DevViewEPK5uint2j",
    "k_free_flight_TurboClosed": "_ZN6heroam13k_free_flightINS_11TurboClosedELb1EEEvNS_7DevViewEPK5uint2j",
    "k_integrate_thread_2_Turbo": "_ZN6heroam18k_integrate_threadILi2ENS_5TurboELb0EEEvNS_7DevViewEPKjj",
    "k_integrate_thread_4_Turbo": "_ZN6heroam18k_integrate_threadILi4ENS_5TurboELb0EEEvNS_7DevViewEPKjj",
    "k_integrate_thread_8_Turbo": "_ZN6heroam18k_integrate_threadILi8ENS_5TurboELb0EEEvNS_7DevViewEPKjj",
    "k_integrate_thread_4_Exact": "_ZN6heroam18k_integrate_threadILi4ENS_5ExactELb0EEEvNS_7DevViewEPKjj",
}


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), re.sub(r"\s+", " ", m.group(2))))
    os.makedirs(OUT, exist_ok=True)
    rows = []
    for short, mangled in KERNELS.items():
        ins = funcs.get(mangled)
        if not ins:
            print("missing", short, file=sys.stderr)
            continue
        loops = []
        for addr, text in ins:
            m = re.search(r"BRA.*?(0x[0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
        # the step loop: the innermost of the longest backward branches (an outer loop contains it)
        loops.sort(key=lambda ab: ab[1] - ab[0], reverse=True)
        step = loops[0] if loops else (0, 0)
        for ab in loops[1:]:
            if ab[0] >= step[0] and ab[1] <= step[1] and (ab[1] - ab[0]) > 0.4 * (step[1] - step[0]):
                step = ab
        body = [t for a, t in ins if step[0] <= a <= step[1]]
        ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for t in body)
        with open(os.path.join(OUT, f"r2_sass_{short}.txt"), "w") as f:
            f.write(f"# {demangle(mangled)}\n# cuobjdump -sass sphere_roaming_b200/libheroam_gpu.so (sm_100a), {len(ins)} instructions\n")
            f.write(f"# step loop: {step[0]:#06x} .. {step[1]:#06x}, {len(body)} instructions per trip\n")
            f.write("# opcode histogram of the step loop: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n\n")
            for addr, text in ins:
                mark = ">>" if step[0] <= addr <= step[1] else "  "
                f.write(f"{mark} /*{addr:04x}*/ {text};\n")
        rows.append((short, len(ins), len(body), ops))
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    usage = []
    for i, line in enumerate(res):
        m = re.search(r"Function (\S+):", line)
        if m and i + 1 < len(res):
            u = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", res[i + 1]))
            usage.append((demangle(m.group(1)), u))
    with open(os.path.join(OUT, "r2_ptxas.md"), "w") as f:
        f.write("# Round 2: resource usage of every kernel (`cuobjdump -res-usage`, sm_100a) and the step loops of the hot ones\n\n")
        f.write("## Step loops (`scripts/make_sass_summary.py`; full listings in `r2_sass_*.txt`)\n\n")
        f.write("| kernel | instructions | per trip of the step loop | FFMA | FMUL | FADD | MUFU | LDG | other |\n|---|---|---|---|---|---|---|---|---|\n")
        for short, n, nb, ops in rows:
            known = sum(ops.get(k, 0) for k in ("FFMA", "FMUL", "FADD", "MUFU", "LDG"))
            f.write(f"| `{short}` | {n} | {nb} | {ops.get('FFMA', 0)} | {ops.get('FMUL', 0)} | {ops.get('FADD', 0)} | {ops.get('MUFU', 0)} | "
                    f"{ops.get('LDG', 0)} | {nb - known} |\n")
        f.write("\nThe free-flight loop is unrolled four times (divide by 4 for one step).  No `UTMA*`/`UTC*MMA`/`HMMA` anywhere: the path has "
                "no tile movement and no contraction (DESIGN.md section 3); the table gathers are `LDG.E.CONSTANT` (read-only path).\n\n")
        f.write("## Registers / stack / shared memory\n\n| kernel | registers | stack [B] | shared [B] |\n|---|---|---|---|\n")
        for name, u in sorted(usage):
            if "k_fma_peak" in name:
                continue
            f.write(f"| `{name[:110]}` | {u.get('REG', '?')} | {u.get('STACK', '?')} | {u.get('SHARED', '?')} |\n")
    print("wrote", len(rows), "listings and r2_ptxas.md")


if __name__ == "__main__":
    main()
