#!/usr/bin/env python
"""Tensor-core / TMEM / bulk-copy SASS mnemonics of every kernel of libmepp_b200.so, plus the issue loop of the scan:
python tools/sass_evidence.py > profiles/r02_sass_tcgen05.txt   (CPU only: cuobjdump on the built library)"""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mepp_b200", "libmepp_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
WANT = ("UTCIMMA", "UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "USETMAXREG", "UTCALLOC", "ELECT", "SYNCS")
counts, fn, k, body = collections.OrderedDict(), None, -1, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        fn = re.sub(r"\(.*", "", names[k])
        counts[fn] = collections.Counter(); body[fn] = []
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn:
        ins = m.group(2).strip()
        body[fn].append((m.group(1), ins))
        op = [w for w in ins.split() if not w.startswith("@")][0]
        base = op.split(".")[0]
        if base in WANT:
            counts[fn][op if base != "SYNCS" else "SYNCS.*"] += 1
print("# cuobjdump -sass mepp_b200/libmepp_b200.so (sm_100a): tcgen05 / TMEM / bulk-copy (TMA engine) / mbarrier mnemonics per kernel")
for fn, c in counts.items():
    if any(b in "".join(c) for b in ("UTC", "LDTM", "UBLKCP")):
        print(f"\n{fn}")
        for op, n in sorted(c.items()):
            print(f"    {n:4d}  {op}")
for want, title in (("tc_scan4i_kernel<true, 0, 1>", "issue loop of one N-tile (uniform datapath: descriptors advance by UIADD3, products back to back)"),):
    for fn in body:
        if fn.endswith(want):
            idx = [i for i, (_, ins) in enumerate(body[fn]) if "UTCIMMA" in ins]
            print(f"\n## {fn}: {title}")
            for a, ins in body[fn][idx[0] - 12: idx[-1] + 8]:
                print(f"    /*{a}*/  {ins}")
            idx = [i for i, (_, ins) in enumerate(body[fn]) if "LDTM" in ins][:4]
            print(f"\n## {fn}: TMEM reads of a full N-tile (epilogue)")
            for a, ins in body[fn][idx[0] - 6: idx[-1] + 3]:
                print(f"    /*{a}*/  {ins}")
