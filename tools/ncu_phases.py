"""Group an `ncu --page source --print-source cuda,sass --csv` dump of k_render by kernel phase (source-line ranges given as
start:name pairs) and list the heaviest lines by executed instructions."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
cur_file=None; hdr=None; agg=collections.defaultdict(lambda:[0,0,0]); text={}; cur=None
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No':
        hdr=r; ii,si,ti=hdr.index('Instructions Executed'),hdr.index('# Samples'),hdr.index('Thread Instructions Executed'); continue
    if hdr and len(r)==len(hdr):
        if r[0].isdigit(): cur=(cur_file,int(r[0])); text[cur]=r[1]; continue
        if r[2]:
            try: agg[cur][0]+=int(r[ii] or 0); agg[cur][1]+=int(r[si] or 0); agg[cur][2]+=int(r[ti] or 0)
            except ValueError: pass
T=sum(v[0] for v in agg.values()); S=sum(v[1] for v in agg.values())
print('total warp-inst', T, 'samples', S)
marks=[]
src=open(sys.argv[2]).read().splitlines() if len(sys.argv)>2 else []
for i,l in enumerate(src,1):
    if '// ----' in l or '//@' in l: marks.append((i,l.strip()[:70]))
marks=[(0,'top')]+marks+[(10**9,'end')]
g=collections.defaultdict(lambda:[0,0]); other=collections.defaultdict(lambda:[0,0])
for (f,l),v in agg.items():
    if f=='mh_render.cu':
        for (a,n),(b,_) in zip(marks,marks[1:]):
            if a<=l<b: g[(a,n)][0]+=v[0]; g[(a,n)][1]+=v[1]; break
    else: other[(f,l)][0]+=v[0]; other[(f,l)][1]+=v[1]
for k in sorted(g): print(f'{k[0]:5d} {k[1]:72s} inst {100*g[k][0]/T:5.1f}% samp {100*g[k][1]/S:5.1f}%')
for k,v in sorted(other.items(), key=lambda kv:-kv[1][0])[:10]: print(k, f'inst {100*v[0]/T:5.1f}% samp {100*v[1]/S:5.1f}%', text.get(k,'')[:60])
print('per-line top by inst')
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 25]: print(f'{k[0]}:{k[1]} inst {100*v[0]/T:5.1f}% samp {100*v[1]/S:5.1f}% act {v[2]/max(v[0],1):4.1f} | {text.get(k,"")[:90].strip()}')
