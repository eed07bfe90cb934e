"""Opcode histogram of a kernel in an ncu report (warp instructions executed per opcode, per particle).
    python tools/ncu_ops.py gpurun_out/prof.ncu-rep <nparticles> [top] [lo_offset hi_offset]
lo/hi: byte offsets from the kernel's first instruction (the region ends tools/ncu_regions.py prints), hex."""
import collections, csv, subprocess, sys
rep, npart = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lo, hi = (int(sys.argv[4], 16), int(sys.argv[5], 16)) if len(sys.argv) > 5 else (0, 1 << 60)
base = None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
ops = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or not r or not r[0].startswith("0x"): continue
    d = dict(zip(hdr, r))
    if base is None: base = int(r[0], 16)
    if not (lo < int(r[0], 16) - base <= hi): continue
    s = d["Source"].strip()
    parts = s.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LD", "ST", "ATOM", "RED", "MUFU", "SHFL", "BAR")) else "")
    n = int(d["Instructions Executed"] or 0)
    ops[op] += n; thr[op] += int(d["Thread Instructions Executed"] or 0); smp[op] += int(d["# Samples"] or 0)
tot = sum(ops.values()); ts = sum(smp.values())
print(f"warp instructions {tot}  = {tot/npart*32:.0f} thread slots / particle; samples {ts}")
for op, n in ops.most_common(top):
    print(f"{op:22s} {100*n/tot:6.2f}%  {n/npart*32:8.1f} slots/particle  lanes {thr[op]/max(n,1):5.1f}  samp {100*smp[op]/ts:5.1f}%")
