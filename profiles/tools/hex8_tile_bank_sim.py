import itertools
def sts64_wf(addrs_by_thread):
    # addrs: list of 32 entries (double index or None); half-warps of 16
    tot=0
    for h in range(2):
        cnt={}
        for t in range(16*h,16*h+16):
            a=addrs_by_thread[t]
            if a is None: continue
            cnt.setdefault(a%16,set()).add(a)
        tot+=max([len(v) for v in cnt.values()],default=0)
    return tot
def lds128_wf(addrs):  # addrs in doubles, 16B aligned
    tot=0
    for q in range(4):
        cnt={}
        for t in range(8*q,8*q+8):
            a=addrs[t]
            if a is None: continue
            cnt.setdefault((a//2)%8,set()).add(a)
        tot+=max([len(v) for v in cnt.values()],default=0)
    return tot
def evaluate(phys, estride):
    w=0; n=0
    for d in range(5):
        for i in range(3):
            for k in range(3):
                direct=[]; mirror=[]
                for t in range(32):
                    el,a=divmod(t,8)
                    if d==4 and a>=4: direct.append(None); mirror.append(None); continue
                    bb=(a+d)&7
                    direct.append(el*estride+phys(a,i*24+3*bb+k))
                    mirror.append(el*estride+phys(bb,k*24+3*a+i))
                w+=sts64_wf(direct); n+=1
                if d>0: w+=sts64_wf(mirror); n+=1
    r=0
    for c in range(36):
        addrs=[]
        for t in range(32):
            el,a=divmod(t,8)
            addrs.append(el*estride+phys(a,2*c))
        r+=lds128_wf(addrs)
    return w,n,r
def mk(S,sw):
    def phys(A,o):
        u=o>>1
        return A*S+((u^sw(A,u))<<1)+(o&1)
    return phys
cands={
 'cur S72 xor(A>>1)': mk(72,lambda A,u:(A>>1)),
 'S72 none': mk(72,lambda A,u:0),
 'S74 none': mk(74,lambda A,u:0),
 'S72 xor A': mk(72,lambda A,u:A&3),
 'S72 xor ((A>>1)|((A&1)<<2))': mk(72,lambda A,u:((A>>1)|((A&1)<<2))),
}
for es in (584,578,600):
  for name,ph in cands.items():
    # check bijection
    seen=set(ph(A,o) for A in range(8) for o in range(72))
    ok=len(seen)==576
    print(es,name,'bij',ok,'max',max(seen), evaluate(ph,es))
print('search')
best=[]
for pe in itertools.permutations(range(4)):
    for po in itertools.permutations(range(4)):
        s=[0]*8
        for j in range(4): s[2*j]=pe[j]; s[2*j+1]=po[j]
        ph=mk(72,lambda A,u,s=s:s[A])
        w,n,r=evaluate(ph,584)
        best.append((w+r,w,r,s))
best.sort()
print(best[:5])
# also unit-level swizzles using more bits: u ^ (s_A) with s_A in 0..7 restricted to units<32? skip
