import csv,sys,re
rows=list(csv.reader(open(sys.argv[1])))
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
lo=starts[0]; hi=starts[1] if len(starts)>1 else len(rows)
hdr=rows[lo+1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[lo+2:hi] if len(r)==len(hdr)]
pat=re.compile(sys.argv[2]) if len(sys.argv)>2 else re.compile(r"BAR\.SYNC|FENCE|LDTM|UTCHMMA|UTMALDG|SYNCS\.ARRIVE|STG|EXIT|MUFU.EX2|MUFU.RCP|TRYWAIT")
tot=sum(int(float(r[ix["# Samples"]] or 0)) for r in data)
acc=0; last=None; n_mufu=0
for i,r in enumerate(data):
    src=r[ix["Source"]].strip()
    s=int(float(r[ix["# Samples"]] or 0))
    m=pat.search(src)
    key=m.group(0) if m else None
    if key and not (key==last and key in ("MUFU.EX2","STG","LDTM","UTCHMMA","UTMALDG")):
        print(f"{i:5d} +{acc:6d} ({100*acc/tot:5.1f}%)  | {src[:80]}  exec={r[ix['Instructions Executed']]}")
        acc=0
    last=key if key else last
    acc+=s
print("tail",acc)
