"""Summarise `ncu --page source --csv` output: per kernel, opcode histogram and the hottest SASS."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
kern = None; cols = None; data = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name': kern = r[1]; data[kern] = []; continue
    if r[0] == 'Address': cols = r; continue
    if kern and cols and r[0].startswith('0x'): data[kern].append(dict(zip(cols, r)))
flt = sys.argv[2] if len(sys.argv) > 2 else ''
for k, ins in data.items():
    if flt and flt not in k: continue
    tot = sum(int(i['Instructions Executed']) for i in ins)
    samp = sum(int(i['# Samples']) for i in ins) or 1
    print(f"=== {k}: {len(ins)} SASS, {tot} warp-instr executed, {samp} samples")
    op = collections.Counter(); ops = collections.Counter()
    for i in ins:
        o = i['Source'].split()[0] if not i['Source'].strip().startswith('@') else i['Source'].split()[1]
        o = o.split('.')[0]
        op[o] += int(i['Instructions Executed']); ops[o] += int(i['# Samples'])
    for o, c in op.most_common(28):
        print(f"   {o:12s} {c:12d} {100*c/tot:5.1f}%   samples {100*ops[o]/samp:5.1f}%")
    print("   -- hottest by stall samples")
    for i in sorted(ins, key=lambda i: -int(i['# Samples']))[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
        print(f"   {int(i['# Samples']):6d} {100*int(i['# Samples'])/samp:5.1f}%  exec {int(i['Instructions Executed']):9d}  {i['Source'].strip()[:90]}")
