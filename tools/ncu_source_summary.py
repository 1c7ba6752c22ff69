"""Summarise an `ncu --page source --csv` dump: stall samples per opcode, per stall reason and per elimination step."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = his[0]
end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
ia = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
tot = sum(int(r[isamp] or 0) for r in data)
print(len(data), 'SASS instructions; total stall samples', tot)
byop = collections.Counter(); exop = collections.Counter()
for r in data:
    toks = r[ia].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    byop[op] += int(r[isamp] or 0); exop[op] += int(r[iex] or 0)
for op, c in byop.most_common(16):
    print(f"  {op:24s} samples {c:7d} ({100 * c / tot:5.1f}%)  warp-instr executed {exop[op]}")
print('stall reasons (all samples):')
for name in ['stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_no_inst', 'stall_not_selected',
             'stall_selected', 'stall_dispatch', 'stall_branch_resolving', 'stall_barrier', 'stall_lg']:
    if name in hdr:
        i = hdr.index(name)
        v = sum(int(r[i] or 0) for r in data)
        print(f"  {name:24s} {v:8d} ({100 * v / tot:5.1f}%)")
cum = 0; marks = []
for idx, r in enumerate(data):
    cum += int(r[isamp] or 0)
    if 'SHFL.BFLY' in r[ia] and (not marks or idx - marks[-1][0] > 60):
        marks.append((idx, cum))
prev = 0; pidx = 0
print('samples per region (regions delimited by the first SHFL.BFLY of each elimination step):')
for idx, c in marks:
    print(f"  instr {pidx:5d}..{idx:5d}: {c - prev:7d} ({100 * (c - prev) / tot:5.1f}%)"); prev = c; pidx = idx
print(f"  instr {pidx:5d}..end  : {tot - prev:7d} ({100 * (tot - prev) / tot:5.1f}%)")
