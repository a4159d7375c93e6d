"""Per-kernel SASS regions of an `ncu --page source --csv --print-source sass` export: consecutive instructions with the same
execution count are one block; prints each block's share of the executed warp-instructions and of the stall samples.
usage: ncu_sass_blocks2.py file.csv kernel-substring [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = None; cols = None; K = {}
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name': kern = r[1]; K.setdefault(kern, []); cols = None; continue
    if r[0] == 'Address': cols = r; continue
    if kern and cols and r[0].startswith('0x'): K[kern].append(dict(zip(cols, r)))
ins = [v for k, v in K.items() if sys.argv[2] in k][0]
seen = set(); uniq = []
for i in ins:   # some exports list the kernel's instructions twice
    if i['Address'] in seen: continue
    seen.add(i['Address']); uniq.append(i)
ins = uniq
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
tot = sum(int(i['Instructions Executed']) for i in ins); ts = sum(int(i['# Samples']) for i in ins)
print(f"total {tot} warp-instr, {len(ins)} SASS, {ts} samples")
blocks = []
for idx, i in enumerate(ins):
    e = int(i['Instructions Executed'])
    if blocks and abs(blocks[-1][0] - e) <= 0.02 * max(e, 1): blocks[-1][1].append(i)
    else: blocks.append([e, [i], idx])
for e, b, idx in blocks:
    w = sum(int(x['Instructions Executed']) for x in b); s = sum(int(x['# Samples']) for x in b)
    if w / tot < minshare and s / ts < minshare: continue
    ops = {}
    for x in b:
        t = x['Source'].split(); o = t[1] if t[0].startswith('@') else t[0]; o = o.split('.')[0]; ops[o] = ops.get(o, 0) + 1
    top = ' '.join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:7])
    print(f"@{idx:5d} exec {e:9d} x{len(b):4d}  instr {100*w/tot:5.1f}%  samples {100*s/ts:5.1f}%  [{top}]")
