"""Per-function instruction counts and warp-stall samples of the first kernel in an .ncu-rep captured with --import-source on (device functions
are found by their definitions in qp_twisted.cuh; other files are reported whole).  python profiles/ncu_functions.py <report.ncu-rep>"""
import csv,re,collections,subprocess,sys
rep=sys.argv[1]
def I(x):
    try: return int(x)
    except: return 0
a=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda'],capture_output=True,text=True).stdout
b=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass,cuda'],capture_output=True,text=True).stdout
import io
src=[r[1] for r in csv.reader(io.StringIO(a)) if r and r[0].isdigit()]
rows=list(csv.reader(io.StringIO(b)))
fn=None; fmap={}
for i,l in enumerate(src,1):
    m=re.match(r'^(?:template.*>\s*)?__device__.*?\b(\w+)\s*\(',l) or re.match(r'^__global__.*?\b(\w+)\s*\(',l)
    if m: fn=m.group(1)
    fmap[i]=fn
keys=['Instructions Executed','Warp Stall Sampling (All Samples)','stall_long_sb','stall_no_inst','stall_short_sb','stall_wait','stall_math','stall_barrier','stall_branch_resolving','stall_mio','stall_selected']
agg={k:collections.Counter() for k in keys}; sass=collections.Counter()
cur=None;hdr=None;curfile=None
for r in rows:
    if r and r[0] in('File Path','File Name'): curfile=r[1].split('/')[-1]; continue
    if r and r[0]=='Line No': hdr=r; continue
    if not hdr or not r: continue
    if r[0].isdigit(): cur=(curfile,int(r[0])); continue
    if r[0]=='' and len(r)>7:
        d=dict(zip(hdr[2:],r[2:]))
        f=fmap.get(cur[1]) if cur[0]=='qp_twisted.cuh' else cur[0]
        for k in keys: agg[k][f]+=I(d.get(k))
        sass[f]+=1
ti=sum(agg[keys[0]].values()); ts=sum(agg[keys[1]].values())
print('inst',ti,'samples',ts,'sass',sum(sass.values()))
print('%-22s %6s %6s %6s | %s'%('fn','inst%','samp%','cyc/i',' '.join(k.replace('stall_','')[:6] for k in keys[2:])))
for f,v in agg[keys[0]].most_common(28):
    s=agg[keys[1]][f]
    print('%-22s %6.1f %6.1f %6.2f | %s'%(f,100*v/ti,100*s/ts,(s/ts)/(v/ti), ' '.join('%6.1f'%(100*agg[k][f]/max(s,1)) for k in keys[2:])))
