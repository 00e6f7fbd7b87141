"""Print metrics of an `ncu --page raw --csv` dump whose names match a regex (one kernel per row)."""
import csv, re, sys
path, pat = sys.argv[1], re.compile(sys.argv[2])
minv = float(sys.argv[3]) if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            try:
                fv = float(v.replace(",", ""))
            except ValueError:
                fv = None
            if minv is not None and (fv is None or fv < minv):
                continue
            print(f"  {h:95s} {u:14s} {v}")
