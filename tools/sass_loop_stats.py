"""Per-kernel statistics of the innermost multiply/square loops in the built objects (cuobjdump -sass):
counts of IMAD.WIDE (wide), IMAD.HI (hi), 32-bit IMAD (lo) and IMAD-encoded moves/adds, with an estimate of
FMA-heavy-pipe cycles per loop iteration (wide 4, hi 6, lo 2). Usage: python tools/sass_loop_stats.py build/field_*.o"""
import re,sys,collections,subprocess
def analyze(obj):
    txt=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout
    fn=None; ins=collections.defaultdict(list)
    for line in txt.split('\n'):
        m=re.search(r'Function : (\S+)',line)
        if m: fn=m.group(1); continue
        m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);',line)
        if m and fn: ins[fn].append((int(m.group(1),16),m.group(3),m.group(4)))
    for fn,L in ins.items():
        addr={a:i for i,(a,_,_) in enumerate(L)}
        loops=[]
        for i,(a,op,args) in enumerate(L):
            if op.startswith('BRA'):
                m=re.search(r'0x([0-9a-f]+)',args)
                if m:
                    t=int(m.group(1),16)
                    if t<a and t in addr: loops.append((addr[t],i))
        # innermost loops with many IMAD
        res=[]
        for (s,e) in loops:
            if any(s<=s2 and e2<=e and (s2,e2)!=(s,e) for (s2,e2) in loops): continue
            c=collections.Counter(op for _,op,_ in L[s:e+1])
            wide=sum(v for k,v in c.items() if k.startswith('IMAD.WIDE'))
            if wide<10: continue
            hi=sum(v for k,v in c.items() if k.startswith('IMAD.HI'))
            lo=sum(v for k,v in c.items() if k in('IMAD','IMAD.X'))
            mov=sum(v for k,v in c.items() if k.startswith('IMAD.MOV') or k.startswith('IMAD.IADD') or k.startswith('IMAD.SHL'))
            res.append((e-s+1,wide,hi,lo,mov))
        short=re.sub(r'.*anemoi_kernelINS_\d+F_','',fn)[:22]
        print(short,[dict(n=r[0],wide=r[1],hi=r[2],lo=r[3],imadmisc=r[4],fmaheavy_cyc=r[1]*4+r[2]*6+r[3]*2+r[4]*2) for r in res])
for o in sys.argv[1:]: analyze(o)
