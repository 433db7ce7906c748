import csv,collections,sys,subprocess,io
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 25
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr,units,r=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','smsp__cycles_active.avg','sm__cycles_elapsed.max','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_uniform.sum','sm__inst_executed_pipe_cbu.sum','sm__inst_executed_pipe_adu.sum','sm__inst_executed_pipe_xu.sum']
for k in keys:
    if k in hdr:
        i=hdr.index(k); print(f'{k:80s} {units[i]:14s} {r[i]}')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=collections.Counter(); data=[]
for rr in rows[2:]:
    if len(rr)<len(hdr) or rr[0]=='Address': break
    data.append(rr)
    for s in stalls:
        try: tot[s]+=int(rr[idx[s]])
        except: pass
T=sum(tot.values())
print('instructions',len(data),'samples',T)
for s,v in tot.most_common(12): print(f'  {s:28s} {v:7d} {100*v/T:5.1f}%')
top=sorted(data,key=lambda x:-int(x[idx['# Samples']] or 0))[:topn]
for x in top: print(x[idx['# Samples']], x[0][-5:], x[1][:70], {s[6:]:x[idx[s]] for s in stalls if x[idx[s]] not in ('0','')})
