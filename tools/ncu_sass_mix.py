"""Opcode mix and stall hot spots from an ncu report's source page: python tools/ncu_sass_mix.py rep.ncu-rep [nbuckets]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; nb = int(sys.argv[2]) if len(sys.argv) > 2 else 24
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); samples = collections.Counter()
N = len(body)
ex_total = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
print("static SASS instructions:", N, " executed warp-inst:", ex_total)
for r in body:
    op = r[ix["Source"]].split()[0] if not r[ix["Source"]].startswith("@") else r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    tot[op] += int(r[ix["Instructions Executed"]] or 0)
print("opcode mix (% of executed):")
print("  " + "  ".join(f"{k}:{100*v/ex_total:.1f}" for k, v in tot.most_common(28)))
# buckets along the address space
size = (N + nb - 1) // nb
print("bucket  static  exec%  samples  no_inst  math  short_sb long_sb wait dispatch not_sel mio")
for b in range(nb):
    seg = body[b*size:(b+1)*size]
    if not seg: break
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    g = lambda k: sum(int(r[ix[k]] or 0) for r in seg)
    first = seg[0][ix["Source"]][:28]
    print(f"{b:3d} {len(seg):6d} {100*ex/ex_total:6.1f} {g('# Samples'):7d} {g('stall_no_inst'):7d} {g('stall_math'):6d} {g('stall_short_sb'):6d} {g('stall_long_sb'):6d} {g('stall_wait'):6d} {g('stall_dispatch'):6d} {g('stall_not_selected'):6d} {g('stall_mio'):5d}  {first}")
