"""Split a kernel's SASS at its BAR.SYNC instructions and report executed instructions / samples per region.
    python tools/ncu_regions.py gpurun_out/prof.ncu-rep <nparticles>"""
import collections, csv, subprocess, sys
rep, npart = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; regs = []; cur = collections.Counter(); base = None
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or not r or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r)); a = int(r[0], 16)
    if base is None: base = a
    s = d["Source"].strip()
    cur["inst"] += int(d["Instructions Executed"] or 0); cur["thr"] += int(d["Thread Instructions Executed"] or 0)
    cur["samp"] += int(d["# Samples"] or 0)
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_no_inst"):
        cur[k] += int(d.get(k) or 0)
    if "BAR.SYNC" in s or "EXIT" in s:
        regs.append((a - base, s, cur)); cur = collections.Counter()
tot = sum(c["inst"] for _, _, c in regs); ts = sum(c["samp"] for _, _, c in regs)
for off, s, c in regs:
    if c["inst"] == 0: continue
    st = ", ".join(f"{k[6:]} {100*c[k]/max(1,c['samp']):.0f}%" for k in c if k.startswith("stall_") and c[k] > 0.08 * c["samp"])
    print(f"..{off:#07x} {s[:28]:28s} inst {100*c['inst']/tot:5.1f}% ({c['inst']/npart*32:7.1f}/particle) lanes {c['thr']/max(1,c['inst']):4.1f} samp {100*c['samp']/ts:5.1f}%  {st}")
