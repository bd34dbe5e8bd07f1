"""Per-kernel SASS comparison of one translation unit between a git revision and the working tree (no GPU needed).
Used to show that adding an opt-in kernel or a host-side change leaves the already-measured kernels byte-identical:

    python tools/sass_diff.py build_path.cu            # HEAD vs working tree
    python tools/sass_diff.py sim_screen.cu 47e7594    # a given revision vs working tree

Exit code 1 if any kernel present on both sides differs."""
from __future__ import annotations

import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from avlmaps_b200 import _build  # noqa: E402


def sass_of(cu: Path, include_root: Path, out: Path):
    flags = [f for f in _build.NVCC_FLAGS if f not in ("--cudart", "static")]
    subprocess.run([_build._nvcc(), *flags, "-I", str(include_root), "-c", str(cu), "-o", str(out)], check=True,
                   capture_output=True, text=True)
    txt = subprocess.run(["cuobjdump", "-sass", str(out)], capture_output=True, text=True, check=True).stdout
    mangled = re.findall(r"Function : (\S+)", txt)
    pretty = subprocess.run(["cu++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
    # the anonymous-namespace hash in the mangled name depends on the file path: compare by the demangled signature
    names = {m: re.sub(r"\((int|bool)\)", "", re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", p))
             for m, p in zip(mangled, pretty)}
    funcs, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[m.group(1)]
            funcs[cur] = []
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur is not None and m:
            funcs[cur].append(m.group(1).strip())
    return funcs


def main():
    name = sys.argv[1]
    rev = sys.argv[2] if len(sys.argv) > 2 else "HEAD"
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        old_root = td / "old"
        (old_root / "avlmaps_b200" / "csrc").mkdir(parents=True)
        (old_root / "include").mkdir()
        listing = subprocess.run(["git", "-C", str(ROOT), "ls-tree", "-r", "--name-only", rev, "avlmaps_b200/csrc", "include"],
                                 capture_output=True, text=True, check=True).stdout.split()
        for rel in listing:
            (old_root / rel).write_bytes(subprocess.run(["git", "-C", str(ROOT), "show", f"{rev}:{rel}"], capture_output=True,
                                                        check=True).stdout)
        old = sass_of(old_root / "avlmaps_b200" / "csrc" / name, old_root, td / "old.o")
        new = sass_of(_build.CSRC / name, ROOT, td / "new.o")
    differ = 0
    short_of = lambda k: re.sub(r"^void ", "", k).split("(")[0][-70:]  # noqa: E731
    for k in sorted(set(old) | set(new)):
        short = short_of(k)
        if k not in old:
            twins = [short_of(o) for o in old if o not in new and old[o] == new[k]]
            if twins:
                print(f"same   {short}  ({len(new[k])}; was {twins[0]}: renamed, instructions identical)")
            else:
                print(f"NEW    {short}  ({len(new[k])} instructions)")
        elif k not in new:
            if not any(new[n] == old[k] for n in new if n not in old):
                print(f"GONE   {short}")
        elif old[k] == new[k]:
            print(f"same   {short}  ({len(new[k])})")
        else:
            differ += 1
            print(f"DIFFER {short}  ({len(old[k])} -> {len(new[k])})")
    sys.exit(1 if differ else 0)


if __name__ == "__main__":
    main()
