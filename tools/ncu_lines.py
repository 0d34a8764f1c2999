"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) >= 2 and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0] != "":
        d = dict(zip(hdr, r))
        try: s = int(d["# Samples"])
        except Exception: continue
        if s > 0:
            st = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
            agg.append((s, cur.split('/')[-1], r[0], r[1].strip()[:86], st))
tot = sum(a[0] for a in agg)
print("total samples", tot)
for s, f, ln, src, st in sorted(agg, reverse=True)[:top]:
    t3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*s/tot:5.1f}% {f}:{ln:>4} {src:86s} {t3}")
