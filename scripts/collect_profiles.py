"""copy what the judge reads from gpurun_out/<tag>/ (scratch, written by scripts/gpu_round.sh) and
gpurun_out/<tag>_scale/ (scripts/gpu_scale.sh) into profiles/ (tracked): bench lines, launch list with per-kernel
shares, ncu summaries of the two kernels, per-opcode counts, DRAM traffic of the dominant kernel, scaling lines.
python scripts/collect_profiles.py r02"""
import csv, io, json, os, shutil, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
GS = os.path.join(ROOT, "gpurun_out", tag + "_scale")


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise ValueError(path)


for src, dst in (("bench_1ant.log", "bench_1ant"), ("bench_8ant.log", "bench_8ant"), ("bench_ref.log", "bench_ref"),
                 ("bench_1ant_nobatch.log", "bench_1ant_launch_per_segment"), ("bench_1ant_clean.log", "bench_1ant_clean_input"),
                 ("bench_1ant_nbit8.log", "bench_1ant_nbit8"), ("exe_60s.log", "exe_60s")):
    p = os.path.join(G, src)
    if os.path.exists(p):
        json.dump(last_json(p), open(os.path.join(P, "%s_%s.json" % (tag, dst)), "w"), indent=1)
for src, dst in (("launches.csv", "launches_final.csv"), ("h2d_rate.log", "h2d_rate_ubench.txt"), ("k1_opcodes.csv", "k1_opcodes.csv"),
                 ("kernel_times_serialised.log", "kernel_times_serialised.txt"), ("k2_trace.log", "k2_chunk_trace.txt"),
                 ("exe_60s.err", "exe_60s_profile_table.txt"), ("box.txt", "box.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, "%s_%s" % (tag, dst)))

# per-kernel shares of the launch list
shares = {}
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "launches.csv")) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
tot = 0.0
for r in rows[1:]:
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    d = shares.setdefault(r[ik][:60], {"launches": 0, "us": 0.0})
    d["launches"] += 1; d["us"] += v; tot += v
for d in shares.values():
    d["mean_us"] = d["us"] / d["launches"]; d["share"] = d["us"] / tot; del d["us"]


def summary(rep, out, title):
    s = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], stdout=subprocess.PIPE, text=True).stdout
    ph = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_phases.py"), rep], stdout=subprocess.PIPE, text=True).stdout
    open(out, "w").write("# %s\n# ncu --set full --clock-control none --import-source on, one launch of bench.py --steps 1 --warmup 3 --seconds-per-step 4 (scripts/gpu_round.sh)\n" % title + s
                         + "-- stall samples between barriers (scripts/ncu_phases.py)\n" + ph)
    return s


k1 = summary(os.path.join(G, "prof_k1.ncu-rep"), os.path.join(P, "%s_k1_final_ncu_summary.txt" % tag),
             "%s final: vf_k1_pipelined on the bench workload (1 antenna, 10 segments of 1024 FFTs per launch, rfi_mode 2, 37%% of time steps masked)" % tag)
summary(os.path.join(G, "prof_k2.ncu-rep"), os.path.join(P, "%s_k2_final_ncu_summary.txt" % tag),
        "%s final: vf_k2_normalise<2,1> on the bench workload" % tag)
rd = wr = None
for line in k1.splitlines():
    f = line.split()
    if line.startswith("dram__bytes_read.sum"): rd = float(f[1]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[f[2]]
    if line.startswith("dram__bytes_write.sum"): wr = float(f[1]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[f[2]]
json.dump({"source": "profiles/%s_k1_final_ncu_summary.txt (ncu --set full, one launch, 1 antenna, rfi_mode 2)" % tag,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch_1ant": rd + wr,
           "segments_per_launch": last_json(os.path.join(G, "bench_1ant.log"))["roofline"]["segments_per_launch"],
           "note": "reads = the input samples, once; writes = the detected-power tiles between the two kernels (42 MB per segment on this workload), which a "
                   "launch over several segments no longer keeps in the 126 MB L2; they are not algorithmic bytes (25 862 144 per segment: samples in, packed filterbank out)",
           "launch_shares": shares}, open(os.path.join(P, "k1_traffic.json"), "w"), indent=1)
print(json.dumps(shares, indent=1))

# ---- multi-GPU session -------------------------------------------------------
if os.path.isdir(GS):
    for f in sorted(os.listdir(GS)):
        src = os.path.join(GS, f)
        if f.startswith(("scale_", "coadd16_", "ref_")) and f.endswith(".log"):
            try:
                json.dump(last_json(src), open(os.path.join(P, "%s_%s.json" % (tag, f[:-4])), "w"), indent=1)
            except ValueError:
                print("no JSON line in", src)
    lines = []
    for n in (1, 2, 4, 8):
        src = os.path.join(GS, "h2d_concurrent_%d.log" % n)
        if os.path.exists(src):
            lines += [l for l in open(src).read().splitlines() if l.startswith("ranks")]
    if lines:
        open(os.path.join(P, "%s_h2d_concurrent.txt" % tag), "w").write(
            "# scripts/ubench/h2d_concurrent.py: every rank copies a pinned 256 MB buffer to its GPU 40 times, all ranks together\n" + "\n".join(lines) + "\n")
    for f in ("box_8gpu.txt", "pytest_multirank_8gpu.log", "pytest_multirank_2gpu.log"):
        if os.path.exists(os.path.join(GS, f)):
            shutil.copy(os.path.join(GS, f), os.path.join(P, "%s_%s" % (tag, f.replace(".log", ".txt"))))
