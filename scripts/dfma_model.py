import random
def model(N, a, b):
    M=2*N
    al=[(a>>(32*i))&0xffffffff for i in range(N)]
    bl=[(b>>(32*i))&0xffff for i in range(N)]
    bh=[(b>>(32*i+16))&0xffff for i in range(N)]
    H=0x433<<52
    def biasE(j): return 0 if j==0 else (0xBCD00000 if j==1 else 0xBCCFBCCF)
    def biasO(j): return (0xBCCFBCCF<<16) if j==M-2 else 0
    E=[];O=[]
    for j in range(M-1):
        e=biasE(j); o=biasO(j)
        for i in range(max(0,j-N+1),min(N-1,j)+1):
            e+=al[i]*bl[j-i]; o+=al[i]*bh[j-i]
            assert e<2**52 and o<2**52
        E.append(H+e); O.append(H+o)
    # S1
    S1=[0]*M; c=0
    for k in range(M):
        v=(E[k]&0xffffffff if k<M-1 else 0)+((E[k-1]>>32) if k>=1 else 0)+c
        S1[k]=v&0xffffffff; c=v>>32
    S2=[0]*M; c=0
    for k in range(M):
        v=(O[k]&0xffffffff if k<M-1 else 0)+((O[k-1]>>32) if k>=1 else 0)+c
        S2[k]=v&0xffffffff; c=v>>32
    Sh=[0]*M
    for k in range(M):
        lo=S2[k-1] if k>=1 else 0
        Sh[k]=((S2[k]<<16)|(lo>>16))&0xffffffff
    T=0;c=0
    for k in range(M):
        v=S1[k]+Sh[k]+c
        T|=(v&0xffffffff)<<(32*k); c=v>>32
    return T
for N,p in ((12,0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab),(8,0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001)):
    for t in range(2000):
        a=random.randrange(p); b=random.randrange(p)
        if t==0: a=b=(1<<(32*N))-1
        if t==1: a=b=p-1
        if t==2: a=0
        assert model(N,a,b)==a*b, (N,t)
print("ok")
