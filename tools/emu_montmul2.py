"""Fused-shift variant actually used in field.cuh: arrays X,Y swap roles each iteration."""
import random
from emu_montmul import CC, limbs, val, M32, P, R
def montmul(a,b,p,inv):
    cc=CC(); A=limbs(a); B=limbs(b); Pm=limbs(p)
    X=[0]*8; Y=[0]*8
    for i in range(8):
        bi=B[i]
        Y[0]=cc.add_cc(Y[0],X[1])
        X[0]=cc.madc_lo_cc(A[1],bi,X[2]); X[1]=cc.madc_hi_cc(A[1],bi,X[3])
        X[2]=cc.madc_lo_cc(A[3],bi,X[4]); X[3]=cc.madc_hi_cc(A[3],bi,X[5])
        X[4]=cc.madc_lo_cc(A[5],bi,X[6]); X[5]=cc.madc_hi_cc(A[5],bi,X[7])
        X[6]=cc.madc_lo_cc(A[7],bi,0);    X[7]=cc.madc_hi(A[7],bi,0)
        Y[0]=cc.mad_lo_cc(A[0],bi,Y[0]); Y[1]=cc.madc_hi_cc(A[0],bi,Y[1])
        Y[2]=cc.madc_lo_cc(A[2],bi,Y[2]); Y[3]=cc.madc_hi_cc(A[2],bi,Y[3])
        Y[4]=cc.madc_lo_cc(A[4],bi,Y[4]); Y[5]=cc.madc_hi_cc(A[4],bi,Y[5])
        Y[6]=cc.madc_lo_cc(A[6],bi,Y[6]); Y[7]=cc.madc_hi_cc(A[6],bi,Y[7])
        X[7]=cc.addc(X[7],0)
        m=(Y[0]*inv)&M32
        X[0]=cc.mad_lo_cc(Pm[1],m,X[0]); X[1]=cc.madc_hi_cc(Pm[1],m,X[1])
        X[2]=cc.madc_lo_cc(Pm[3],m,X[2]); X[3]=cc.madc_hi_cc(Pm[3],m,X[3])
        X[4]=cc.madc_lo_cc(Pm[5],m,X[4]); X[5]=cc.madc_hi_cc(Pm[5],m,X[5])
        X[6]=cc.madc_lo_cc(Pm[7],m,X[6]); X[7]=cc.madc_hi(Pm[7],m,X[7])
        Y[0]=cc.mad_lo_cc(Pm[0],m,Y[0]); Y[1]=cc.madc_hi_cc(Pm[0],m,Y[1])
        Y[2]=cc.madc_lo_cc(Pm[2],m,Y[2]); Y[3]=cc.madc_hi_cc(Pm[2],m,Y[3])
        Y[4]=cc.madc_lo_cc(Pm[4],m,Y[4]); Y[5]=cc.madc_hi_cc(Pm[4],m,Y[5])
        Y[6]=cc.madc_lo_cc(Pm[6],m,Y[6]); Y[7]=cc.madc_hi_cc(Pm[6],m,Y[7])
        X[7]=cc.addc(X[7],0)
        assert Y[0]==0
        X,Y=Y,X
    # after swap: X = E-role (X[0]==0), Y = D-role
    r=[0]*8
    r[0]=cc.add_cc(Y[0],X[1])
    for k in range(1,7): r[k]=cc.addc_cc(Y[k],X[k+1])
    r[7]=cc.addc(Y[7],0)
    v=val(r); assert v<2*p
    return v-p if v>=p else v
if __name__=="__main__":
    for mod in (P,R):
        inv=(-pow(mod,-1,1<<32))%(1<<32); Rinv=pow(1<<256,-1,mod)
        for t in range(5000):
            a=random.randrange(mod) if t>10 else mod-1
            b=random.randrange(mod) if t>5 else mod-1
            assert montmul(a,b,mod,inv)==a*b*Rinv%mod
    print("ok")
