import random
M32=0xffffffff
R_LIMBS=[0x00000001,0xffffffff,0xfffe5bfe,0x53bda402,0x09a1d805,0x3339d808,0x299d7d48,0x73eda753]
R=sum(l<<(32*i) for i,l in enumerate(R_LIMBS))
N=8
class CC:
    def __init__(s): s.c=0
    def add_cc(s,a,b): v=a+b; s.c=v>>32; return v&M32
    def addc_cc(s,a,b): v=a+b+s.c; s.c=v>>32; return v&M32
    def addc(s,a,b): v=a+b+s.c; return v&M32   # carry out dropped (must be 0 where it matters)
    def mad_lo_cc(s,a,b,c): v=((a*b)&M32)+c; s.c=v>>32; return v&M32
    def madc_lo_cc(s,a,b,c): v=((a*b)&M32)+c+s.c; s.c=v>>32; return v&M32
    def madc_hi_cc(s,a,b,c): v=((a*b)>>32)+c+s.c; s.c=v>>32; return v&M32
    def madc_hi(s,a,b,c):
        v=((a*b)>>32)+c+s.c; assert v>>32==0, "overflow in madc_hi"; return v&M32
def reduce_step_special(cc,E,O):
    E0=E[0]
    m=(E0*M32)&M32
    nz=1 if E0!=0 else 0
    c=1 if E0==M32 else 0
    add1=(E0+nz)&M32
    add2=(m-nz+c)&M32
    O[1]=cc.add_cc(O[1],add2)
    for j in range(2,N-2,2):
        O[j]=cc.madc_lo_cc(R_LIMBS[j+1],m,O[j]); O[j+1]=cc.madc_hi_cc(R_LIMBS[j+1],m,O[j+1])
    O[N-2]=cc.madc_lo_cc(R_LIMBS[N-1],m,O[N-2]); O[N-1]=cc.madc_hi(R_LIMBS[N-1],m,O[N-1])
    E[0]=0
    E[1]=cc.add_cc(E[1],add1)
    for j in range(2,N,2):
        E[j]=cc.madc_lo_cc(R_LIMBS[j],m,E[j]); E[j+1]=cc.madc_hi_cc(R_LIMBS[j],m,E[j+1])
    v=O[N-1]+cc.c; assert v>>32==0; O[N-1]=v
def mul_inline(a,b):
    cc=CC()
    al=[(a>>(32*i))&M32 for i in range(N)]; bl=[(b>>(32*i))&M32 for i in range(N)]
    E=[0]*N;O=[0]*N
    bi=bl[0]
    for j in range(0,N,2):
        O[j]=(al[j+1]*bi)&M32; O[j+1]=(al[j+1]*bi)>>32
        E[j]=(al[j]*bi)&M32; E[j+1]=(al[j]*bi)>>32
    reduce_step_special(cc,E,O)
    for i in range(1,N):
        bi=bl[i]
        nE=O[:]; nO=[0]*N
        nE[0]=cc.add_cc(nE[0],E[1])
        for j in range(0,N-2,2):
            nO[j]=cc.madc_lo_cc(al[j+1],bi,E[j+2]); nO[j+1]=cc.madc_hi_cc(al[j+1],bi,E[j+3])
        nO[N-2]=cc.madc_lo_cc(al[N-1],bi,0); nO[N-1]=cc.madc_hi(al[N-1],bi,0)
        nE[0]=cc.mad_lo_cc(al[0],bi,nE[0]); nE[1]=cc.madc_hi_cc(al[0],bi,nE[1])
        for j in range(2,N,2):
            nE[j]=cc.madc_lo_cc(al[j],bi,nE[j]); nE[j+1]=cc.madc_hi_cc(al[j],bi,nE[j+1])
        v=nO[N-1]+cc.c; assert v>>32==0; nO[N-1]=v
        reduce_step_special(cc,nE,nO)
        E=nE;O=nO
    r=[0]*N
    r[0]=cc.add_cc(E[1],O[0])
    for j in range(1,N-1): r[j]=cc.addc_cc(E[j+1],O[j])
    r[N-1]=cc.addc(0,O[N-1])
    v=sum(l<<(32*i) for i,l in enumerate(r))
    if v>=R: v-=R
    return v
Rinv=pow(1<<256,-1,R)
random.seed(1)
cases=[(0,0),(R-1,R-1),(1,1),(R-1,1),(M32,M32),((1<<255)-1 if (1<<255)-1<R else R-2,R-1)]
for t in range(20000):
    if t<len(cases): a,b=cases[t]
    else:
        a=random.randrange(R); b=random.randrange(R)
        if t%7==0: a&=~M32            # low limb zero
        if t%11==0: a|=M32            # low limb all ones
        if t%13==0: b=(b&~M32)|1
    assert mul_inline(a,b)==a*b*Rinv%R,(t,hex(a),hex(b))
print("ok")
