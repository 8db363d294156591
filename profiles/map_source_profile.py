import csv, re, sys, difflib, collections, bisect
sass_file, prof_csv, srcroot = sys.argv[1], sys.argv[2], sys.argv[3]
# parse disassembly
ins=[]; cur=[]
for l in open(sass_file):
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        cur.append((m.group(1),int(m.group(2))))
        continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m:
        ins.append((int(m.group(1),16), m.group(2).strip(), list(cur)))
        cur=[]
# carry forward line info when an instruction has none
last=[]
for k,(a,t,c) in enumerate(ins):
    if c: last=c
    else: ins[k]=(a,t,last)
rows=[]
with open(prof_csv) as f:
    r=csv.reader(f); next(r); hdr=next(r); ix={h:i for i,h in enumerate(hdr)}
    for row in r:
        if len(row)<40: continue
        g=lambda k:int(row[ix[k]] or 0)
        rows.append(dict(src=row[ix['Source']].strip(), ie=g('Instructions Executed'), te=g('Thread Instructions Executed'), s=g('# Samples'), noi=g('stall_no_inst'), lsb=g('stall_long_sb'), wait=g('stall_wait'), ssb=g('stall_short_sb'), br=g('stall_branch_resolving'), conf=g('L1 Conflicts Shared N-Way') if 'L1 Conflicts Shared N-Way' in ix else 0))
op=lambda t: re.sub(r'^@!?U?P\d+\s+','',t).split()[0] if t else ''
A=[op(t) for a,t,c in ins]; B=[op(r['src']) for r in rows]
sm=difflib.SequenceMatcher(None,A,B,autojunk=False)
mapping={}
for i,j,n in sm.get_matching_blocks():
    for k in range(n): mapping[j+k]=i+k
print('matched',len(mapping),'of',len(rows),file=sys.stderr)
# function ranges per file
funcs={}
def load(fn):
    if fn in funcs: return funcs[fn]
    starts=[];names=[]
    try:
        for n,l in enumerate(open(fn),1):
            m=re.match(r'^(?:template\s*<[^>]*>\s*)?(?:static\s+|__device__\s+|__host__\s+|__forceinline__\s+|__noinline__\s+|inline\s+|__global__\s+|constexpr\s+)*[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$',l)
            if m and not l.startswith(' ') and not l.startswith('//') and m.group(1) not in ('if','for','while','switch','return','sizeof','static_assert','defined'):
                starts.append(n);names.append(m.group(1))
    except Exception as e: pass
    funcs[fn]=(starts,names); return funcs[fn]
def fname(fn,line):
    s,nm=load(fn)
    k=bisect.bisect_right(s,line)-1
    return nm[k] if k>=0 else '?'
agg=collections.defaultdict(lambda: collections.Counter())
aggline=collections.defaultdict(lambda: collections.Counter())
tot=collections.Counter()
for j,r in enumerate(rows):
    i=mapping.get(j)
    chain=ins[i][2] if i is not None else []
    # innermost entry = first; pick innermost in our repo
    key='?'; lk='?'
    for fn,line in chain:
        if 'csrc' in fn or 'include/r' in fn:
            key=fname(fn,line); lk=f"{fn.split('/')[-1]}:{line}"; break
    for k in ('ie','te','s','noi','lsb','wait','ssb','br','conf'):
        agg[key][k]+=r[k]; aggline[lk][k]+=r[k]; tot[k]+=r[k]
    agg[key]['n']+=1
print('TOTAL',dict(tot))
print(f"{'function':34s} {'sass':>6s} {'inst%':>6s} {'lanes':>5s} {'smpl%':>6s} {'noi%':>5s} {'lsb%':>5s} {'wait%':>5s} {'ssb%':>5s} {'br%':>5s}")
for k,v in sorted(agg.items(),key=lambda kv:-kv[1]['s'])[:60]:
    s=max(v['s'],1)
    print(f"{k:34s} {v['n']:6d} {100*v['ie']/tot['ie']:6.2f} {v['te']/max(v['ie'],1):5.1f} {100*v['s']/tot['s']:6.2f} {100*v['noi']/s:5.1f} {100*v['lsb']/s:5.1f} {100*v['wait']/s:5.1f} {100*v['ssb']/s:5.1f} {100*v['br']/s:5.1f}")
print()
for k,v in sorted(aggline.items(),key=lambda kv:-kv[1]['s'])[:70]:
    s=max(v['s'],1)
    print(f"{k:28s} inst% {100*v['ie']/tot['ie']:5.2f} lanes {v['te']/max(v['ie'],1):5.1f} smpl% {100*v['s']/tot['s']:5.2f} noi {100*v['noi']/s:4.0f} lsb {100*v['lsb']/s:4.0f} wait {100*v['wait']/s:4.0f}")
