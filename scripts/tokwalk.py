import sys, numpy as np
sys.path.insert(0, '.')
from tests.fixtures import load_golden_stream
LAY=[0x0000,0x0008,0x0019,0x0091,0x012A,0x01A2,0x02A1,0x0A21,0x123B,0x12B3,0x13B2,0x1B32,0x23B1,0x2B31,0x3B21,0xB321]
g = load_golden_stream(sys.argv[1] if len(sys.argv)>1 else "music_cbr128_bs2048")
d = g["data"]; N=g["block_size"]
nyb = lambda i: (int(d[i>>1]) >> (4*(i&1))) & 15
pos = 0
for blk in range(3):
    h0 = nyb(pos); pos+=1
    wc = h0 | (nyb(pos)<<4) if h0&8 else h0|0x10
    if h0&8: pos+=1
    for c in range(2):
        pat = LAY[wc>>4]
        while True:
            S = N>>(pat&7)
            q = nyb(pos); pos+=1
            if q==14: q+=nyb(pos); pos+=1
            start=pos; left=S; kind=0; draws=0
            if q!=29:
              while left>0:
                v=nyb(pos)
                if v==0: n=nyb(pos+1)+1; pos+=2; left-=n
                elif v==1: n=(nyb(pos+1)<<4|nyb(pos+2))+33; pos+=3; left-=n
                elif v==8: n=((nyb(pos+1)<<4|nyb(pos+2))<<1|(nyb(pos+3)&1))+16; pos+=4; left-=n; draws+=n
                elif v==15:
                    q2=nyb(pos+1)
                    if q2<14: pos+=2
                    elif q2==14:
                        if nyb(pos+2)==15: pos+=3; kind=1; break
                        pos+=3
                    else: pos+=5; kind=2; draws+=left; break
                else: pos+=1; left-=1
            print(f"blk {blk} ch {c} S {S} q {q} start nyb {start} (page {start//2048} tp {start%2048}) end {pos} (page {pos//2048} rel {pos%2048}) kind {kind} coefs_before_stop {S-left if kind else S} draws {draws}")
            pat>>=4
            if not pat: break
    pos=(pos+1)&~1
