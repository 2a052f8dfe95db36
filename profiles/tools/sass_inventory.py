"""SASS mnemonic counts per kernel of libmsfec_b200.so:  python profiles/tools/sass_inventory.py [lib] > table.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mpi-msfec_b200/libmsfec_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
cols = ["DMMA", "DFMA", "LDGSTS", "UBLKCP", "UBLKPF", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL"]
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = demangle(m.group(1))
        name = re.sub(r"\((int|bool)\)", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*", "", name).split("::")[-1]
        cur = name; cnt[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        cnt[cur]["total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                cnt[cur][c] += 1
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k in sorted(cnt):
    c = cnt[k]
    print(f"| `{k}` | {c['total']} | " + " | ".join(str(c[x]) for x in cols) + " |")
