#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of libb200kge's objects.

    python scripts/sass_diff.py snapshot <dir>     # dump normalised per-function SASS of kge_b200/build/*.o
    python scripts/sass_diff.py compare <dir>      # compare the current build against a snapshot

Used to show that a change leaves the device code of already-validated kernels bit-identical (instruction
text with addresses stripped); kernels that only exist on one side are listed separately."""
import glob
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fns, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                fns[name] = body
            # anonymous-namespace symbols carry a hash of the source PATH: drop it so that builds in different
            # directories compare equal
            name, body = re.sub(r"_GLOBAL__N__[0-9a-f]+_(\d+_\w+?_cu)_[0-9a-f]+", r"_GLOBAL__N__\1", m.group(1)), []
            continue
        if name is None or re.match(r"^\s*/\*[0-9a-f]+\*/\s*$", line):     # encoding-only lines
            continue
        body.append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line).rstrip())
    if name:
        fns[name] = body
    return {k: hashlib.sha256("\n".join(v).encode()).hexdigest() for k, v in fns.items()}


def snapshot():
    snap = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "kge_b200", "build", "*.o"))):
        snap[os.path.basename(obj)] = functions(obj)
    return snap


def main():
    mode, d = sys.argv[1], sys.argv[2]
    path = os.path.join(d, "sass_functions.json")
    if mode == "snapshot":
        os.makedirs(d, exist_ok=True)
        json.dump(snapshot(), open(path, "w"), indent=1)
        print("wrote", path)
        return 0
    old, new = json.load(open(path)), snapshot()
    bad = 0
    for obj, fns in old.items():
        for fn, h in fns.items():
            cur = new.get(obj, {}).get(fn)
            if cur is None:
                print("MISSING ", obj, fn[:100]); bad += 1
            elif cur != h:
                print("CHANGED ", obj, fn[:100]); bad += 1
    added = sum(1 for obj, fns in new.items() for fn in fns if fn not in old.get(obj, {}))
    print(f"{sum(len(f) for f in old.values())} baseline kernels: {bad} changed/missing; {added} new kernels")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
