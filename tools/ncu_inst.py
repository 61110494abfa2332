"""Per-source-line executed-instruction counts (and stall samples) from an ncu report captured with --import-source on:
    python tools/ncu_inst.py report.ncu-rep kernel-regex [top N]
Sums the SASS rows of every source line of the FIRST matching kernel."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern] if kern != "-" else ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
agg, tot, tots = {}, 0, 0
fname, fn, first, take = "", None, None, False
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        fn = r[1]
        if first is None:
            first = fn
        take = fn == first
        continue
    if not take or len(r) < 8 or r[0] in ("", "Line No"):
        continue                      # SASS rows are already summed into their source-line row
    try:
        s = int(r[6]); ie = int(r[7])
    except ValueError:
        continue
    tot += ie; tots += s
    a = agg.setdefault((fname, r[0], r[1][:90]), [0, 0])
    a[0] += ie; a[1] += s
print(first[:100] if first else None)
print("warp instructions:", tot, "samples:", tots)
for (f, line, text), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100 * v[0] / max(tot, 1):5.1f}% inst {100 * v[1] / max(tots, 1):5.1f}% smp  {f}:{line:>4}  {text}")
