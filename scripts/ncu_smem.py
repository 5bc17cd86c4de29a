"""shared-memory wavefronts per SASS instruction (ideal vs excessive): python scripts/ncu_smem.py rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
iw, ie, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared Ideal")
data = []; tw = te = 0
for n, r in enumerate(rows[1:]):
    try: w = int(r[iw] or 0); e = int(r[ie] or 0); i = int(r[ii] or 0)
    except ValueError: continue
    if w: data.append((e, w, i, n, r[isrc].strip()[:70], r[iex])); tw += w; te += e
print("total wavefronts", tw, "excessive", te)
for e, w, i, n, src, ex in sorted(data, reverse=True)[:top]:
    print("#%5d excess %8d total %8d ideal %8d ex=%-8s %s" % (n, e, w, i, ex, src))
