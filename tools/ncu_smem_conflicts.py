"""Shared-memory wavefronts (ideal vs. excessive) per SASS instruction from an ncu report: python tools/ncu_smem_conflicts.py rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
tot = exc = 0
for r in rows[2:]:
    w = int(r[ix["L1 Wavefronts Shared"]] or 0); e = int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    tot += w; exc += e
    if e > 0:
        print(f"{r[ix['Source']].strip()[:60]:60s} exec {r[ix['Instructions Executed']]:>10s} wavefronts {w:>10d} excessive {e:>10d}")
print("total wavefronts", tot, "excessive", exc)
