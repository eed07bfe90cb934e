"""Per-source-line view of an ncu report, read here (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <mangled-kernel-substring> [top]

Joins the SASS page of the report (per-instruction samples and stall reasons) with the line table
of the in-tree library (nvdisasm -g on the cubin extracted from libosiris_b200.so) by instruction
offset, and prints the hottest source lines with their dominant stall reasons.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "osiris_b200", "libosiris_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
line_of = {}
for cub in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    if pat not in dis:
        continue
    cur_fn, cur_line, infn = None, None, False
    for ln in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = pat in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur_line

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels may be in the report: take the blocks whose name matches
agg = collections.defaultdict(lambda: collections.Counter())
hdr, base, use = None, None, False
demangled = None
for r in rows:
    if r and r[0] == "Kernel Name":
        use = True  # reports here hold one kernel family; keep all
        demangled = r[1]
        hdr, base = None, None
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if not use or hdr is None or not r or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    d = dict(zip(hdr, r))
    a = agg[key]
    a["samples"] += int(d.get("# Samples", 0) or 0)
    a["inst"] += int(d.get("Instructions Executed", 0) or 0)
    a["thread_inst"] += int(d.get("Thread Instructions Executed", 0) or 0)
    a["smem_wavefronts"] += int(d.get("L1 Wavefronts Shared", 0) or 0)
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k:
            a[k] += int(v or 0)
tot = sum(a["samples"] for a in agg.values()) or 1
tot_inst = sum(a["inst"] for a in agg.values()) or 1
print(f"kernel: {demangled}\ntotal samples {tot}, warp instructions {tot_inst}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'lanes':>5s} {'smemWF%':>7s}  top stalls")
tot_wf = sum(a["smem_wavefronts"] for a in agg.values()) or 1
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v), key=lambda kv: -kv[1])[:3]
    lanes = a["thread_inst"] / a["inst"] if a["inst"] else 0
    name = f"{key[0]}:{key[1]}" if key else "?"
    print(f"{name:28s} {100*a['samples']/tot:6.2f} {100*a['inst']/tot_inst:6.2f} {lanes:5.1f} {100*a['smem_wavefronts']/tot_wf:7.2f}  "
          + ", ".join(f"{k} {100*v/max(1,a['samples']):.0f}%" for k, v in st))
