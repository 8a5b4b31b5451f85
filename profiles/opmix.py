"""Dynamic opcode mix (warp instructions executed) of the first kernel in an `ncu --page source --csv` dump."""
import collections
import csv
import re
import sys


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    I = {h: i for i, h in enumerate(hdr)}
    body, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            if body:
                break
            continue
        if r[0] in seen:
            break
        seen.add(r[0])
        body.append(r)
    tot, ops = 0, collections.Counter()
    for r in body:
        e = int(r[I["Instructions Executed"]] or 0)
        s = re.sub(r"^@!?U?P\d+\s+", "", r[I["Source"]].strip())
        ops[s.split()[0].split(".")[0]] += e
        tot += e
    print(rows[0][1][:70], "| warp instructions", tot, "| static", len(body))
    print("  " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(top)))
    lsu = sum(ops[o] for o in ("SHFL", "LDS", "STS", "LDG", "REDG", "RED", "STG", "LDL", "STL", "ATOMG", "ATOMS", "LDSM"))
    print(f"  load/store-pipe instructions {100 * lsu / tot:.1f}%")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
