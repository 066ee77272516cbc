"""Per-opcode executed counts and stall samples from an ncu report's source page:
    python tools/ncu_src_ops.py rep.ncu-rep [symbols]   (symbols: divide executed counts to get warp-instr per OFDM symbol)"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; nsym = float(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
REASONS = ["stall_dispatch", "stall_math", "stall_wait", "stall_short_sb", "stall_mio", "stall_barrier", "stall_no_inst", "stall_long_sb", "stall_not_selected", "stall_selected"]
ex = collections.Counter(); st = collections.defaultdict(collections.Counter)
for r in body:
    src = r[ix["Source"]].split()
    op = src[1] if src[0].startswith("@") else src[0]
    parts = op.split(".")
    key = parts[0] + ("." + parts[1] if len(parts) > 1 and parts[0] in ("IMAD", "MUFU", "LDS", "STS", "BAR", "I2F", "I2FP", "F2I", "SHFL", "LDG", "FFMA") and parts[1] in ("WIDE", "HI", "LG2", "SQRT", "RSQ", "RCP", "SIN", "COS", "EX2", "64", "128", "SYNC", "U32", "F32", "BFLY", "SAT", "IDX", "DOWN", "MOV", "SHL", "IADD") else "")
    n = int(r[ix["Instructions Executed"]] or 0)
    ex[key] += n
    for k in REASONS:
        st[key][k] += int(r[ix[k]] or 0)
tot = sum(ex.values())
print(f"static {len(body)}  executed warp-inst {tot}" + (f"  per symbol {tot / nsym:.0f}" if nsym else ""))
print(f"{'op':12s} {'exec%':>6s} {'/sym':>7s} | " + " ".join(f"{k[6:]:>8s}" for k in REASONS))
tots = collections.Counter()
for k, v in ex.most_common(40):
    print(f"{k:12s} {100 * v / tot:6.2f} {v / nsym if nsym else 0:7.1f} | " + " ".join(f"{st[k][q]:8d}" for q in REASONS))
for k in ex:
    for q in REASONS:
        tots[q] += st[k][q]
print(f"{'TOTAL':12s} {100.0:6.2f} {tot / nsym if nsym else 0:7.1f} | " + " ".join(f"{tots[q]:8d}" for q in REASONS))
