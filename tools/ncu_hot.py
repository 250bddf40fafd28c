"""Hot SASS of an ncu source page dump.  usage: python tools/ncu_hot.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples", tot_s, "total warp-instr", tot_i, "SASS lines", len(data))
by_op = collections.Counter(); by_op_s = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    by_op[op] += int(r[ix["Instructions Executed"]] or 0); by_op_s[op] += int(r[ix["# Samples"]] or 0)
print("-- by opcode: warp-instr%  samples%")
for op, n in by_op.most_common(25):
    print(f"   {op:12s} {100*n/tot_i:6.2f}%  {100*by_op_s[op]/tot_s:6.2f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("-- hottest lines by samples")
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/tot_s:5.1f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
