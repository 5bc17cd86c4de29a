"""hot SASS instructions of an ncu report by stall samples: python scripts/ncu_hot.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
lines = out.splitlines()
# first kernel only
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for n, r in enumerate(rows[1:]):
    try: s = int(r[isamp])
    except ValueError: continue
    tot += s
    data.append((s, n, r))
print("total samples", tot, "instructions", len(data))
for s, n, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print("%6d %5.1f%% #%5d ex=%-8s %-60s %s" % (s, 100.0 * s / tot, n, r[iex], r[isrc][:60], st))
