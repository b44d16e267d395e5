#!/usr/bin/env python3
"""tools/sass_hashes.py <lib.so> [--write profiles/validated_sass.json [note]] -- per-kernel md5 of the SASS
(cuobjdump -sass, addresses and comments stripped).  Used to pin which device code actually ran on
hardware: a kernel whose hash differs from profiles/validated_sass.json has changed since the last
GPU validation (tests/test_validated_sass.py)."""
import hashlib
import json
import re
import subprocess
import sys


def kernel_hashes(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            d[cur] = hashlib.md5()
            continue
        if cur is None or re.match(r"\s*//", ln):
            continue
        d[cur].update(re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().encode() + b"\n")
    return {k: v.hexdigest() for k, v in sorted(d.items())}


if __name__ == "__main__":
    h = kernel_hashes(sys.argv[1])
    if len(sys.argv) > 3 and sys.argv[2] == "--write":
        note = sys.argv[4] if len(sys.argv) > 4 else "md5 of each kernel's SASS (tools/sass_hashes.py)"
        with open(sys.argv[3], "w") as f:
            json.dump({"_note": note, "kernels": h}, f, indent=1)
    for k, v in h.items():
        print(v, k)
