#!/usr/bin/env python
"""Per-instruction view of one kernel of an .ncu-rep: executed-count levels, dynamic
opcode mix and the instructions with the most stall samples."""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, kernel, topn=16):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["Instructions Executed"]].isdigit()]
    g = lambda r, k: int(r[ix[k]] or 0)  # noqa: E731
    tot = sum(g(r, "Instructions Executed") for r in data)
    S = sum(g(r, "# Samples") for r in data)
    print(f"kernel {kernel}: static {len(data)} instr, executed {tot} warp-instr, {S} samples")
    lv = collections.Counter()
    for r in data:
        lv[g(r, "Instructions Executed")] += 1
    print("executed-count levels (count x static instrs, share):")
    for n, c in sorted(lv.items(), reverse=True)[:12]:
        print(f"   {n:>12d} x {c:4d}  {n * c / tot:6.3f}")
    c = collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
        c[m.group(2) if m else "?"] += g(r, "Instructions Executed")
    print("opcode mix:", ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in c.most_common(18)))
    print("top stall samples:")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:topn]:
        print(f"   {g(r, '# Samples'):6d} exec {g(r, 'Instructions Executed'):>10d}  {r[ix['Source']][:58]:58s} "
              f"long_sb {r[ix['stall_long_sb']]} wait {r[ix['stall_wait']]} short_sb {r[ix['stall_short_sb']]} "
              f"math {r[ix['stall_math']]} mio {r[ix['stall_mio']]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 16)
