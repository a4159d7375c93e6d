"""Group consecutive SASS instructions of one kernel by execution count: shows which code regions
(loops) account for the executed warp-instructions. usage: ncu_sass_blocks.py file.csv kernel [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = None; cols = None; ins = []
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name': kern = r[1]; continue
    if r[0] == 'Address': cols = r; continue
    if kern and sys.argv[2] in kern and cols and r[0].startswith('0x'): ins.append(dict(zip(cols, r)))
tot = sum(int(i['Instructions Executed']) for i in ins)
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
blocks = []
for i in ins:
    e = int(i['Instructions Executed'])
    if blocks and abs(blocks[-1][0] - e) <= 0.02 * max(e, 1): blocks[-1][1].append(i)
    else: blocks.append([e, [i]])
print(f"total {tot} warp-instr, {len(ins)} SASS")
for e, b in blocks:
    w = sum(int(x['Instructions Executed']) for x in b)
    if w / tot < minshare: continue
    samp = sum(int(x['# Samples']) for x in b)
    ops = {}
    for x in b:
        o = x['Source'].split()[0 if not x['Source'].strip().startswith('@') else 1].split('.')[0]
        ops[o] = ops.get(o, 0) + 1
    top = ' '.join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
    print(f"  exec/instr {e:9d} x {len(b):4d} instr = {100*w/tot:5.1f}%  samples {samp:6d}  [{top}]")
