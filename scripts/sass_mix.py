#!/usr/bin/env python
"""Static memory / synchronisation instruction mix per kernel of libcm_b200.so:
   cuobjdump -sass convergent-matrix_b200/libcm_b200.so | python scripts/sass_mix.py > profiles/r2/sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        if re.match(r"(LDG|STG|LDS|STS|RED|ATOM|UBLK|LDGSTS|MATCH|BAR|UTMA|SYNCS|CCTL)", op):
            cnt[cur][op] += 1
print("# memory / sync instruction mix per kernel of libcm_b200.so (cuobjdump -sass, sm_100a): static counts")
for k in sorted(cnt):
    d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    print(re.sub(r"\(.*", "", d))
    print("   " + "  ".join(f"{o} x{n}" for o, n in sorted(cnt[k].items())))
