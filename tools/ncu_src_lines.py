"""Per-source-line executed warp instructions and stall samples from an ncu report (compiled with -lineinfo, captured with
--import-source on):  python tools/ncu_src_lines.py rep.ncu-rep [symbols] [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nsym = float(sys.argv[2]) if len(sys.argv) > 2 else 0; minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; hl = r; continue
    if r[0] in ("Function Name", "Kernel Name"): continue
    if hdr and r[0].isdigit():
        # source text may hold unescaped quotes / commas: index the metric columns from the END of the row
        def g(k):
            v = r[hl.index(k) - len(hl)]
            return int(v) if v.lstrip("-").isdigit() else 0
        lines.append((cur, int(r[0]), r[1].strip()[:100], g("Instructions Executed"), g("# Samples"),
                      {k: g(k) for k in ("stall_dispatch", "stall_math", "stall_wait", "stall_short_sb", "stall_mio", "stall_barrier", "stall_no_inst", "stall_long_sb", "stall_not_selected", "stall_selected")}))
tot = sum(l[3] for l in lines); stot = sum(l[4] for l in lines)
print(f"executed {tot}" + (f" ({tot / nsym:.0f} per symbol)" if nsym else "") + f", samples {stot}")
print(f"{'file:line':24s} {'inst%':>6s} {'/sym':>7s} {'smp%':>6s} {'s/i':>5s}  disp math wait ssb mio bar noi | source")
for f, ln, src, ex, sm, st in lines:
    if 100 * ex / tot < minpct and 100 * sm / stot < minpct: continue
    ratio = (sm / stot) / (ex / tot) if ex else 0
    print(f"{f + ':' + str(ln):24s} {100 * ex / tot:6.2f} {ex / nsym if nsym else 0:7.1f} {100 * sm / stot:6.2f} {ratio:5.2f}  "
          f"{st['stall_dispatch']:4d} {st['stall_math']:4d} {st['stall_wait']:4d} {st['stall_short_sb']:4d} {st['stall_mio']:4d} {st['stall_barrier']:4d} {st['stall_no_inst']:4d} | {src}")
