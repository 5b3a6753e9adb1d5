import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None
d=collections.defaultdict(list)
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        rr=dict(zip(hdr,r))
        try: d[rr["Kernel Name"][:28]].append(float(rr["Metric Value"])/1000)
        except: pass
for k,v in d.items(): print("%-30s n=%3d avg %7.1f min %7.1f max %7.1f us"%(k,len(v),sum(v)/len(v),min(v),max(v)))
