"""Hot SASS of an ncu source page dump (first kernel in the file, or the k-th).
usage: python tools/ncu_hot.py src.csv [top] [kernel_index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[kidx]
hi = starts[kidx + 1] if kidx + 1 < len(starts) else len(rows)
print(rows[lo][1][:120])
hdr = rows[lo + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
num = lambda r, k: int(float(r[ix[k]] or 0))
tot_s = sum(num(r, "# Samples") for r in data)
tot_i = sum(num(r, "Instructions Executed") for r in data)
print("total samples", tot_s, "total warp-instr", tot_i, "SASS lines", len(data))
by_op = collections.Counter(); by_op_s = collections.Counter()
for r in data:
    toks = r[ix["Source"]].split()
    op = toks[0] if toks else "?"
    if op.startswith("@") and len(toks) > 1: op = toks[1]
    op = op.split(".")[0]
    by_op[op] += num(r, "Instructions Executed"); by_op_s[op] += num(r, "# Samples")
print("-- by opcode: warp-instr%  samples%")
for op, n in by_op.most_common(25):
    print(f"   {op:12s} {100*n/max(tot_i,1):6.2f}%  {100*by_op_s[op]/max(tot_s,1):6.2f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("-- hottest lines by samples")
order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((num(r, c), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {num(r,'# Samples'):7d} {100*num(r,'# Samples')/max(tot_s,1):5.1f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
