"""Summarise an ncu report's source page: top CUDA source lines by stall samples / instructions.
usage: python tools/ncu_src.py report.ncu-rep [function-regex] [n]"""
import csv, subprocess, sys, io, re, collections
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "."; n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict()
cur_file = cur_fn = None; hdr = None; seen_fn = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not re.search(pat, cur_fn or ""): continue
    if not r[0].strip().isdigit(): continue
    ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    key = (cur_file.split("/")[-1], int(r[0]))
    a = agg.setdefault(key, [0.0, 0.0, r[1][:100], collections.Counter()])
    try:
        a[0] += float(r[ss] or 0); a[1] += float(r[ie] or 0)
        for i in stall_cols:
            v = float(r[i] or 0)
            if v: a[3][hdr[i]] += v
    except ValueError: pass
tot_s = sum(a[0] for a in agg.values()) or 1; tot_i = sum(a[1] for a in agg.values()) or 1
print(f"samples={tot_s:.0f} inst={tot_i/1e6:.1f}M  (all captured launches of functions matching /{pat}/)")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = ",".join(f"{k[6:]}:{v/a[0]*100:.0f}" for k, v in a[3].most_common(2)) if a[0] else ""
    print(f"{a[0]/tot_s*100:5.1f}% smp {a[1]/tot_i*100:5.1f}% ins {key[0][:12]:>12s}:{key[1]:<4d} {top:28s} {a[2]}")
