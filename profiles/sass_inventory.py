#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel size and counts of the Blackwell tensor / TMA / TMEM mnemonics in libgp_b200.so,
plus the lines of the CTA-pair forward edge kernel that carry them.

    python profiles/sass_inventory.py > profiles/r02_sass_inventory.md

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG / UTMAPF = cp.async.bulk.tensor load /
store / prefetch, UTCBAR = tcgen05.commit, LDGSTS = cp.async, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graph-physics_b200", "graphphysics_b200", "lib", "libgp_b200.so")
OPS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "LDGSTS", "SYNCS", "HMMA", "REDG", "ATOMG", "RED")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.splitlines()
    kern, counts, lines = None, collections.OrderedDict(), collections.defaultdict(list)
    for ln in sass:
        m = re.search(r"Function : (\S+)", ln)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            continue
        if kern and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            counts[kern]["n"] += 1
            for o in OPS:
                if re.search(r"\b" + o + r"(\b|\.)", ln):
                    counts[kern][o] += 1
                    lines[kern].append(ln.rstrip())
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    short = lambda n: re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", ""))
    tot = collections.Counter()
    rows = []
    for (k, c), n in zip(counts.items(), names):
        for o in OPS:
            tot[o] += c[o]
        rows.append((short(n), c, k))
    print("# SASS inventory of `libgp_b200.so` (sm_100a), round 2\n")
    print("`python profiles/sass_inventory.py` (cuobjdump -sass).  Library totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS) + ".\n")
    print("The only global reductions (`REDG` / `ATOMG`) are integer: the bucket counters of the graph-layout / mesh kernels "
          "(`gp_csr_from_coo`, `gp_coalesce_*`, `gp_world_pairs_*`). The phase-profiling counters of the MLP kernels exist only in the "
          "`-DGP_MLP_PROF` tuning build. No float atomics, no `HMMA`.\n")
    print("| kernel | instructions | " + " | ".join(OPS[:9]) + " |\n|---|---|" + "---|" * 9)
    for n, c, _ in sorted(rows, key=lambda r: -r[1]["n"]):
        if any(c[o] for o in OPS[:8]):
            print(f"| `{n[:72]}` | {c['n']} | " + " | ".join(str(c[o]) for o in OPS[:9]) + " |")
    for want in ("edge_fwd2_kernel", "mlp_bwd_kernel<128, 1>"):
        for n, c, k in rows:
            if n.endswith(want):
                print(f"\n## `{want}`: the lines carrying tensor-core / TMA / TMEM instructions\n\n```")
                for ln in lines[k]:
                    if not re.search(r"\bSYNCS|\bLDGSTS", ln):
                        print(re.sub(r"\s+", " ", ln.strip())[:150])
                print("```")
                break


if __name__ == "__main__":
    main()
