"""Contiguous SASS ranges of (roughly) constant execution count inside an address window of a kernel:
where the executed warp instructions of a phase go.
    python tools/ncu_blocks.py rep nparticles lo hi [min_slots]"""
import csv, subprocess, sys
rep, npart, lo, hi = sys.argv[1], float(sys.argv[2]), int(sys.argv[3], 16), int(sys.argv[4], 16)
mins = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; base = None; blocks = []; cur = None
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or not r or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r)); a = int(r[0], 16)
    if base is None: base = a
    off = a - base
    if not (lo < off <= hi): continue
    n = int(d["Instructions Executed"] or 0); t = int(d["Thread Instructions Executed"] or 0); s = int(d["# Samples"] or 0)
    if cur is None or abs(n - cur["n0"]) > 0.15 * max(n, cur["n0"], 1):
        cur = dict(start=off, n0=n, ninst=0, tot=0, thr=0, samp=0, first=d["Source"].strip()[:40]); blocks.append(cur)
    cur["ninst"] += 1; cur["tot"] += n; cur["thr"] += t; cur["samp"] += s; cur["end"] = off
for b in blocks:
    sl = b["tot"] / npart * 32
    if sl >= mins:
        print(f"{b['start']:#07x}-{b['end']:#07x} {b['ninst']:5d} instrs x {b['n0']/npart*32:6.3f}/round = {sl:7.1f} slots/particle lanes {b['thr']/max(1,b['tot']):4.1f} samp {b['samp']:6d}  {b['first']}")
