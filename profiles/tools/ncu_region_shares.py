import csv, sys, collections
def f(x):
    try: return float(x.replace(',',''))
    except: return 0.0
regions=[(284,370,'mlp_eval'),(388,410,'ckpt'),(417,455,'gates_only'),(456,510,'stage/bulk'),(530,560,'pl.init'),(561,615,'pl.update'),(616,663,'pl.build'),(664,720,'pl.advance'),(721,775,'pl.eval'),(776,823,'pl_evals'),(824,928,'warp reduce/tables'),(929,987,'lat_hidden/x0'),(988,1023,'out_put'),(1024,1160,'fwd kernel body'),(1164,1217,'sweep.events'),(1218,1235,'sweep.add'),(1236,1340,'sweep.finish'),(1341,1447,'lat_epilogue'),(1448,1800,'bwd kernel body'),(0,283,'vec helpers')]
rows=list(csv.reader(open(sys.argv[1])))
cur=None; curfile=''
agg=collections.defaultdict(lambda: collections.defaultdict(lambda:[0,0,collections.Counter()]))
hdr=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': curfile=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': cur=r[1][12:52]; hdr=None; continue
    if r[0]=='Line No': hdr=r; ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples'); stall={h:i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}; continue
    if hdr is None or len(r)!=len(hdr) or not r[0].strip().isdigit(): continue
    ln=int(r[0])
    if curfile=='slode_mlp_kernels.cuh':
        reg=next((n for a,b,n in regions if a<=ln<=b),'other')
    else: reg=curfile
    a=agg[cur][reg]; a[0]+=f(r[ci]); a[1]+=f(r[si])
    for h,i in stall.items(): a[2][h]+=f(r[i])
for k,d in agg.items():
    ti=sum(v[0] for v in d.values()); ts=sum(v[1] for v in d.values())
    print('=====',k,'instr %.4g samples %d'%(ti,ts))
    for reg,v in sorted(d.items(), key=lambda kv:-kv[1][1]):
        top=', '.join('%s %.0f%%'%(h[6:],100*c/max(v[1],1)) for h,c in v[2].most_common(4))
        print('  %-18s %5.1f%% ins %5.1f%% smp | %s'%(reg,100*v[0]/ti,100*v[1]/ts,top))
