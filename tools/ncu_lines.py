"""Summarise an ncu report per CUDA source line: stall samples, instructions, L1/shared wavefronts.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    lines.append(r)


def col(name):
    return hdr.index(name)


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


c_s, c_i = col("# Samples"), col("Instructions Executed")
c_ws, c_wi = col("L1 Wavefronts Shared"), col("L1 Wavefronts Shared Ideal")
c_tg = col("L1 Tag Requests Global")
c_l2 = col("L2 Theoretical Sectors Global")
tot_s = sum(num(r[c_s]) for r in lines) or 1
tot_i = sum(num(r[c_i]) for r in lines) or 1
tot_w = sum(num(r[c_ws]) for r in lines) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}, shared wavefronts {tot_w:.0f}, "
      f"ideal {sum(num(r[c_wi]) for r in lines):.0f}, global tag requests {sum(num(r[c_tg]) for r in lines):.0f}, "
      f"L2 sectors {sum(num(r[c_l2]) for r in lines):.0f}")
print(f"{'line':>5} {'samp%':>6} {'inst%':>6} {'shWave':>10} {'shIdeal':>10} {'glTag':>10} {'L2sect':>10}  source")
for r in sorted(lines, key=lambda r: -num(r[c_s]))[:top]:
    print(f"{r[0]:>5} {100 * num(r[c_s]) / tot_s:6.2f} {100 * num(r[c_i]) / tot_i:6.2f} {num(r[c_ws]):10.0f} "
          f"{num(r[c_wi]):10.0f} {num(r[c_tg]):10.0f} {num(r[c_l2]):10.0f}  {r[1].strip()[:100]}")
