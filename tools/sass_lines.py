"""Static SASS instruction count per source line (innermost inlined location) of one kernel in an executable / .so / cubin:
    python tools/sass_lines.py <binary> <substring of the mangled kernel name> [min_count]
Needs -lineinfo.  Prints lines ordered by first appearance in the instruction stream, with running address (KB)."""
import re, subprocess, sys, tempfile, os, glob
binary, pat = sys.argv[1], sys.argv[2]
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=d, capture_output=True)
txt = ""
for c in glob.glob(os.path.join(d, "*.cubin")):
    txt += subprocess.run(["nvdisasm", "--print-line-info", c], capture_output=True, text=True).stdout
secs = re.split(r"\n(?=\.text\.)", txt)
for s in secs:
    head = s.split("\n", 1)[0]
    if not head.startswith(".text.") or pat not in head:
        continue
    cur = "?"; order = []; cnt = {}; first = {}; n = 0
    for line in s.split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            n += 1
            if cur not in cnt: cnt[cur] = 0; first[cur] = n; order.append(cur)
            cnt[cur] += 1
    print(head[:150]); print(f"static instructions: {n} ({n * 16 / 1024:.1f} KB)")
    for k in order:
        if cnt[k] >= minc: print(f"  first@{first[k] * 16 / 1024:6.1f} KB  {cnt[k]:5d}  {k}")
    break
