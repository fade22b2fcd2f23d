"""Aggregate ncu warp-stall samples per CUDA source line.
usage: python profiles/src_hotspots.py report.ncu-rep [top_n]   (needs -lineinfo builds and --import-source on)"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None
agg = collections.Counter(); src = {}; stall = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].strip().isdigit(): continue
    try:
        n = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    key = (cur_file, int(r[0]))
    agg[key] += n; src[key] = r[1].strip()
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: stall[key][h[6:]] += int(r[i])
            except Exception: pass
tot = sum(agg.values())
print("total samples", tot)
for (f, ln), n in agg.most_common(top):
    st = ", ".join(f"{k}:{v}" for k, v in stall[(f, ln)].most_common(3))
    print(f"{n:7d} {100 * n / tot:5.1f}%  {f}:{ln:<5d} {src[(f, ln)][:80]:80s} [{st}]")
