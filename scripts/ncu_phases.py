"""stall samples between barriers (phases) of the first kernel in an ncu report"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ib = hdr.index("stall_barrier")
acc = 0; accb = 0; first = 0; tot = 0; execd = 0
for n, r in enumerate(rows[1:]):
    try: s = int(r[isamp]); b = int(r[ib] or 0); e = int(r[iex] or 0)
    except ValueError: continue
    acc += s - b; accb += b; tot += s; execd += e
    if "BAR." in r[isrc] or "EXIT" in r[isrc]:
        print("#%5d-%5d  non-barrier samples %6d  barrier samples %6d  warp-instr %9d   ends with %s" % (first, n, acc, accb, execd, r[isrc].strip()[:40]))
        acc = 0; accb = 0; first = n + 1; execd = 0
print("total", tot)
