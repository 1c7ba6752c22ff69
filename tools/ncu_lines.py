"""Per-source-line stall samples of one kernel from an ncu report captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; lines = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or not r[0]: continue
    try: lines.append((int(r[6]), int(r[7]), cur_file, int(r[0]), r[1].strip()[:110]))
    except ValueError: pass
tot = sum(l[0] for l in lines)
print("total samples", tot)
for smp, inst, f, ln, text in sorted(lines, reverse=True)[:top]:
    print(f"{100.0 * smp / tot:5.1f}%  inst {inst:>10}  {f}:{ln:<4} {text}")
