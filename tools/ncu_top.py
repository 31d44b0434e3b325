"""Summarises an `ncu --page source --csv` dump: instruction mix by opcode and the hottest SASS lines."""
import csv
import collections
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
mix, smp = collections.Counter(), collections.Counter()
tot_i = tot_s = 0
for r in body:
    op = r[ci["Source"]].split()[0 if not r[ci["Source"]].strip().startswith("@") else 1]
    op = op.split(".")[0]
    n = int(float(r[ci["Instructions Executed"]] or 0))
    s = int(float(r[ci["Warp Stall Sampling (All Samples)"]] or 0))
    mix[op] += n
    smp[op] += s
    tot_i += n
    tot_s += s
print(f"kernel: {rows[0][1][:100] if rows and len(rows[0]) > 1 else ''}")
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("opcode           instr      %instr   %samples")
for op, n in mix.most_common(22):
    print(f"{op:14s} {n:10d}   {100.0 * n / max(tot_i, 1):6.2f}   {100.0 * smp[op] / max(tot_s, 1):6.2f}")
print("hottest SASS lines (samples, instr, source)")
body.sort(key=lambda r: -float(r[ci["Warp Stall Sampling (All Samples)"]] or 0))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in body[:int(sys.argv[2]) if len(sys.argv) > 2 else 18]:
    top = sorted(((float(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{r[ci['Warp Stall Sampling (All Samples)']]:>6s} {r[ci['Instructions Executed']]:>9s}  {r[ci['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
