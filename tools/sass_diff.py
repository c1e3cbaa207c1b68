#!/usr/bin/env python3
"""Which kernels of two builds of libneucor_b200.so differ?  Compares the SASS of every function (addresses and encodings
stripped) — used to confine a change to the kernels it means to touch when no GPU is at hand to re-validate the others.
usage: tools/sass_diff.py old.so new.so"""
import hashlib
import re
import subprocess
import sys


def functions(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and "/*" in line:
            t = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            if t:
                d[cur].append(t)
    return {k: (hashlib.md5("\n".join(v).encode()).hexdigest()[:10], len(v)) for k, v in d.items()}


if __name__ == "__main__":
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    for k in sorted(set(a) | set(b)):
        if a.get(k) != b.get(k):
            print("DIFF", k, a.get(k), b.get(k))
    print(len(a), len(b), "functions;", sum(1 for k in a if a.get(k) == b.get(k)), "identical")
