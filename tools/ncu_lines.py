"""Per-source-line stall samples from an ncu report captured with --import-source on (kernels built with -lineinfo):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top N]
Uses `ncu --page source --csv --print-source sass,cuda`: every SASS row carries its sample count, the first SASS row of
a source line carries the line number and text; samples are summed per line."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
hdr, cur, agg, tot = None, None, {}, 0
for r in csv.reader(io.StringIO(out)):
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        cur = (r[0], r[1][:100])
    try:
        s = int(r[6])
    except ValueError:
        continue
    tot += s
    if cur:
        agg[cur] = agg.get(cur, 0) + s
print("samples:", tot)
for (line, text), v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
    print(f"{100 * v / max(tot, 1):5.1f}%  L{line:>5}  {text}")
