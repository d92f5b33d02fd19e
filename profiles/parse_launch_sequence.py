import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size") if "Grid Size" in hdr else None
seq=[(r[ki].split('(')[0].replace('mwx::<unnamed>::','').replace('void ',''), float(r[vi].replace(',',''))/1e3, r[gi] if gi is not None else '') for r in rows[1:]]
# find last occurrence of k_xr_update and print the iteration before it
idx=[i for i,(k,_,_) in enumerate(seq) if k.startswith('k_xr_update')]
a,b=idx[-2],idx[-1]
tot=0
for k,t,g in seq[a+1:b+1]:
    print(f"{k[:40]:40s} {t:9.1f} us grid {g}"); tot+=t
print('iteration total us', tot)
