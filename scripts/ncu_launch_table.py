import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5 and not r[0].startswith("==")]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{"k":r[ki][:44]})[r[mi]]=float(r[vi].replace(",",""))
for k,v in d.items():
    t=v.get("gpu__time_duration.sum",0)/1e3; rd=v.get("dram__bytes_read.sum",0); wr=v.get("dram__bytes_write.sum",0)
    print(f"{k:>4} {v['k']:46s} {t:10.1f} us  read {rd/1e6:9.1f} MB  write {wr/1e6:9.1f} MB  {(rd+wr)/t/1e3 if t else 0:8.1f} GB/s")
