"""Stall-reason totals and the hottest SASS lines of the first kernel in an `ncu --page source --csv` dump."""
import collections
import csv
import sys


def main(path, top=14):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    I = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    body, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) or r[0] in seen or not r[0].startswith("0x"):
            if body:
                break
            continue
        seen.add(r[0])
        body.append(r)
    tot, samples = collections.Counter(), 0
    for r in body:
        samples += int(r[I["# Samples"]] or 0)
        for c in stall_cols:
            v = r[I[c]]
            if v and v != "0":
                tot[c] += int(v)
    print(rows[0][1][:100], "| SASS instrs", len(body), "| samples", samples)
    print("  stalls:", ", ".join(f"{c[6:]} {100 * v / samples:.1f}%" for c, v in tot.most_common(8)))
    tags = sum(int(r[I["L1 Tag Requests Global"]] or 0) for r in body)
    l2s = sum(int(r[I["L2 Theoretical Sectors Global"]] or 0) for r in body)
    print(f"  L1 tag requests global {tags}, L2 theoretical sectors global {l2s}")
    for r in sorted(body, key=lambda r: -int(r[I["# Samples"]] or 0))[:top]:
        print(f"   {r[I['# Samples']]:>6} {r[I['Source']].strip()[:80]:80s} exec {r[I['Instructions Executed']]} "
              f"tags {r[I['L1 Tag Requests Global']]} l2sec {r[I['L2 Theoretical Sectors Global']]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
