"""Static evidence for the sm_100a kernels, no GPU needed: per kernel the ptxas resource line (registers,
spills, barriers, static shared memory) and the counts of the SASS mnemonics that show which hardware path a
kernel takes (tcgen05 MMA / TMEM loads, TMA bulk-tensor loads, vector reductions, fp64 arithmetic ...).

    python tools/static_evidence.py > profiles/r1e_static.md
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from avlmaps_b200 import _build  # noqa: E402

# mnemonic prefix -> what it proves (B200_PROFILING.md names UTCHMMA / UTMALDG / LDTM as the tcgen05 / TMA evidence)
WATCH = OrderedDict([
    ("UTCHMMA", "tcgen05.mma (kind::f16)"), ("UTCBAR", "tcgen05.commit -> mbarrier"), ("UTCATOMSWS", "tcgen05.alloc / dealloc (TMEM columns)"),
    ("LDTM", "tcgen05.ld (TMEM -> registers)"), ("STTM", "tcgen05.st (registers -> TMEM)"),
    ("UTMALDG", "TMA cp.async.bulk.tensor load"), ("UTMAPF", "TMA L2 prefetch"), ("SYNCS", "mbarrier arrive / try_wait"),
    ("UCGABAR", "cluster barrier"), ("REDG.E.ADD.F32x4", "red.global.add.v4.f32 (16-byte vector reduction)"),
    ("REDG", "red.global, all forms (fire-and-forget atomic)"), ("ATOMG", "returning global atomic"),
    ("ATOMS", "shared-memory atomic"), ("REDUX", "redux.sync warp reduction"), ("DFMA", "fp64 fma"), ("DMUL", "fp64 mul"),
    ("DADD", "fp64 add"), ("HMMA", "legacy mma.sync (should be 0)"), ("LDG.E.128", "128-bit global load"),
    ("STG.E.128", "128-bit global store"),
])


def demangle(names):
    out = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\((int|bool)\)", "", n)           # screen_kernel<(int)2> -> screen_kernel<2>
        n = re.sub(r"\((?![^()]*\)>).*$", "", n)        # drop the parameter list
        n = n.replace("<unnamed>::", "")
        short.append(n.replace("avl::", ""))
    return dict(zip(names, short))


def ptxas_table():
    rows = []
    for src in _build.SOURCES:
        cmd = [_build._nvcc(), *_build.NVCC_FLAGS, "-Xptxas=-v", "-c", str(_build.CSRC / src), "-o", "/dev/null"]
        err = subprocess.run(cmd, capture_output=True, text=True).stderr
        cur = None
        for line in err.splitlines():
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
            if m:
                cur = {"file": src, "name": m.group(1), "spill": "0/0", "stack": 0}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["spill"] = int(m.group(1)), f"{m.group(2)}/{m.group(3)}"
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(.*)", line)
            if m:
                cur["regs"], cur["bars"] = int(m.group(1)), int(m.group(2) or 0)
                sm = re.search(r"(\d+) bytes smem", m.group(3))
                cur["smem"] = int(sm.group(1)) if sm else 0
    return rows


def sass_counts():
    txt = subprocess.run(["cuobjdump", "-sass", str(_build.LIB)], capture_output=True, text=True).stdout
    per = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op.startswith(w) and (len(op) == len(w) or op[len(w)] in "._"):
                    cur[w] += 1
    return per


def main():
    _build.build()
    rows = ptxas_table()
    names = demangle([r["name"] for r in rows])
    sass = sass_counts()
    print("# Static evidence (ptxas -v, cuobjdump -sass) -- `python tools/static_evidence.py`, nvcc 12.9, sm_100a\n")
    print("No GPU involved: what the compiler emitted for every kernel of `libavlmaps_b200.so`. Timings are in the")
    print("`*_summary.md` files; this file shows which hardware path each kernel is on and that nothing spills.\n")
    print("## 1. Resources per kernel\n")
    print("| file | kernel | registers | spill st/ld (B) | stack (B) | barriers | static smem (B) | SASS instructions |")
    print("|---|---|---:|---:|---:|---:|---:|---:|")
    for r in rows:
        c = sass.get(r["name"], Counter())
        print(f"| `{r['file']}` | `{names[r['name']]}` | {r.get('regs', '?')} | {r['spill']} | {r['stack']} | {r.get('bars', 0)} | "
              f"{r.get('smem', 0)} | {c['_total']} |")
    print("\n## 2. Hardware-path mnemonics (count of SASS instructions per kernel; kernels with none of them omitted)\n")
    used = [w for w in WATCH if any(c[w] for c in sass.values())]
    print("| kernel | " + " | ".join(f"`{w}`" for w in used) + " |")
    print("|---|" + "---:|" * len(used))
    for r in rows:
        c = sass.get(r["name"], Counter())
        if not any(c[w] for w in used):
            continue
        print(f"| `{names[r['name']]}` | " + " | ".join(str(c[w]) if c[w] else "" for w in used) + " |")
    print("\nLegend: " + "; ".join(f"`{w}` = {WATCH[w]}" for w in used) + ".")
    absent = [w for w in WATCH if w not in used]
    if absent:
        print("\nAbsent from every kernel: " + "; ".join(f"`{w}` ({WATCH[w]})" for w in absent) + ".")


if __name__ == "__main__":
    main()
