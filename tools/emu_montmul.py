"""Emulates the PTX carry-chain sequence used by fe_mul in csrc/field.cuh (design check).
E[k] sits at column k, D[k] at column k+1; T = E + D*2^32; roles swap every iteration."""
import random
M32 = 0xFFFFFFFF
class CC:
    def __init__(s): s.cf = 0
    def add_cc(s,a,b): t=a+b; s.cf=t>>32; return t&M32
    def addc_cc(s,a,b): t=a+b+s.cf; s.cf=t>>32; return t&M32
    def addc(s,a,b): t=a+b+s.cf; return t&M32
    def mad_lo_cc(s,a,b,c): t=((a*b)&M32)+c; s.cf=t>>32; return t&M32
    def madc_lo_cc(s,a,b,c): t=((a*b)&M32)+c+s.cf; s.cf=t>>32; return t&M32
    def madc_hi_cc(s,a,b,c): t=((a*b)>>32)+c+s.cf; s.cf=t>>32; return t&M32
    def madc_hi(s,a,b,c): t=((a*b)>>32)+c+s.cf; assert t>>32==0; return t&M32

def limbs(x,n=8): return [(x>>(32*i))&M32 for i in range(n)]
def val(l): return sum(v<<(32*i) for i,v in enumerate(l))

def montmul(a,b,p,inv):
    cc=CC()
    A=limbs(a); B=limbs(b); Pm=limbs(p)
    E=[0]*8; D=[0]*8
    for i in range(8):
        bi=B[i]
        # --- roles: E even-aligned (col k), D odd-aligned (col k+1) ---
        # D += a_odd*bi
        D[0]=cc.mad_lo_cc(A[1],bi,D[0]); D[1]=cc.madc_hi_cc(A[1],bi,D[1])
        D[2]=cc.madc_lo_cc(A[3],bi,D[2]); D[3]=cc.madc_hi_cc(A[3],bi,D[3])
        D[4]=cc.madc_lo_cc(A[5],bi,D[4]); D[5]=cc.madc_hi_cc(A[5],bi,D[5])
        D[6]=cc.madc_lo_cc(A[7],bi,D[6]); D[7]=cc.madc_hi(A[7],bi,D[7])
        # E += a_even*bi ; carry -> D[7]
        E[0]=cc.mad_lo_cc(A[0],bi,E[0]); E[1]=cc.madc_hi_cc(A[0],bi,E[1])
        E[2]=cc.madc_lo_cc(A[2],bi,E[2]); E[3]=cc.madc_hi_cc(A[2],bi,E[3])
        E[4]=cc.madc_lo_cc(A[4],bi,E[4]); E[5]=cc.madc_hi_cc(A[4],bi,E[5])
        E[6]=cc.madc_lo_cc(A[6],bi,E[6]); E[7]=cc.madc_hi_cc(A[6],bi,E[7])
        D[7]=cc.addc(D[7],0)
        m=(E[0]*inv)&M32
        D[0]=cc.mad_lo_cc(Pm[1],m,D[0]); D[1]=cc.madc_hi_cc(Pm[1],m,D[1])
        D[2]=cc.madc_lo_cc(Pm[3],m,D[2]); D[3]=cc.madc_hi_cc(Pm[3],m,D[3])
        D[4]=cc.madc_lo_cc(Pm[5],m,D[4]); D[5]=cc.madc_hi_cc(Pm[5],m,D[5])
        D[6]=cc.madc_lo_cc(Pm[7],m,D[6]); D[7]=cc.madc_hi(Pm[7],m,D[7])
        E[0]=cc.mad_lo_cc(Pm[0],m,E[0]); E[1]=cc.madc_hi_cc(Pm[0],m,E[1])
        E[2]=cc.madc_lo_cc(Pm[2],m,E[2]); E[3]=cc.madc_hi_cc(Pm[2],m,E[3])
        E[4]=cc.madc_lo_cc(Pm[4],m,E[4]); E[5]=cc.madc_hi_cc(Pm[4],m,E[5])
        E[6]=cc.madc_lo_cc(Pm[6],m,E[6]); E[7]=cc.madc_hi_cc(Pm[6],m,E[7])
        D[7]=cc.addc(D[7],0)
        assert E[0]==0
        # shift right one column: newE = D + stray E[1] (with carry), newD[k]=E[k+2]
        nE=[0]*8
        nE[0]=cc.add_cc(D[0],E[1])
        for k in range(1,7): nE[k]=cc.addc_cc(D[k],0)
        nE[7]=cc.addc(D[7],0)
        nD=[E[2],E[3],E[4],E[5],E[6],E[7],0,0]
        E,D=nE,nD
    # merge
    r=[0]*8
    r[0]=E[0]
    r[1]=cc.add_cc(E[1],D[0])
    for k in range(2,8): r[k]=cc.addc_cc(E[k],D[k-1])
    assert cc.cf==0 and D[7]==0
    v=val(r)
    assert v<2*p
    if v>=p: v-=p
    return v

P=0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R=0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
for mod in (P,R):
    inv=(-pow(mod,-1,1<<32))%(1<<32)
    print(hex(inv))
    Rinv=pow(1<<256,-1,mod)
    for t in range(3000):
        a=random.randrange(mod) if t>10 else mod-1
        b=random.randrange(mod) if t>5 else mod-1
        assert montmul(a,b,mod,inv)==a*b*Rinv%mod
print("ok")
