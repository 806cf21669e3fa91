"""numpy model of the tcgen05 K-sdft partial sums (sdft_tc_kernel.cu): 3xTF32 products, f32 accumulation in the
tensor core with round-to-nearest ("rn") or truncation ("rz"), accumulation groups of G samples, outer rotation in
f32, combine in f64.  Prints (max, median) of max_k |X - X_ref| / max_k |X_ref| over 200 frames of window group 1.
On the B200 G = 16 and 32 meet the parity bounds and G = 64 does not: the hardware follows the "rz" rows
(DESIGN.md section 3).

    python scripts/tc_accumulation_model.py
"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pitchvis_b200 import synth
rng=np.random.default_rng(0)
audio = synth.polyphonic_chords(4.0, 22050.0, seed=7).astype(np.float32)
N=4096; H=368; q=N//H; rem=N%H; k_lo=3; nk=229
ks=np.arange(k_lo,k_lo+nk)
def tf32_rna(x):
    u=x.view(np.uint32).astype(np.uint64)
    u=(u+0x1000)&0xffffe000
    return u.astype(np.uint32).view(np.float32)
def rz32(x64):
    # round float64 toward zero to float32
    f=x64.astype(np.float32)
    over=np.abs(f.astype(np.float64))>np.abs(x64)
    f2=np.nextafter(f,np.float32(0))
    return np.where(over,f2,f)
def run(G, mode):
    # rows: chunks
    T=200
    rows=T+q+1
    x=audio[1000:1000+rows*H+16].astype(np.float32)
    X=np.zeros((rows,nk),np.complex128); R=np.zeros((rows,nk),np.complex128)
    xs=x[:rows*H].reshape(rows,H)
    hi=tf32_rna(xs.copy()); lo=tf32_rna((xs-hi).astype(np.float32))
    # group boundaries: [0,rem) cut into G-size (multiple of 16), then [rem,H)
    bounds=[]
    def cut(a,b):
        s=a
        while s<b:
            e=min(s+G,b); bounds.append((s,e)); s=e
    cut(0,rem); cut(rem,H)
    acc_re=np.zeros((rows,nk),np.float32); acc_im=np.zeros((rows,nk),np.float32)
    def f32(v): return v.astype(np.float32)
    for (s,e) in bounds:
        L=e-s
        b=np.arange(L)
        ang=-2*np.pi*((ks[None,:]*b[:,None])%N)/N
        Bre=f32(np.cos(ang)); Bim=f32(np.sin(ang))
        outs=[]
        for Bm in (Bre,Bim):
            bh=tf32_rna(Bm.copy()); bl=tf32_rna((Bm-bh).astype(np.float32))
            d=np.zeros((rows,nk),np.float32)
            for k8 in range(0,L,8):
                sl=slice(s+k8,s+min(k8+8,L)); bs=slice(k8,min(k8+8,L))
                for (A,Bq) in ((lo,bh),(hi,bl),(hi,bh)):
                    p=A[:,sl].astype(np.float64)@Bq[bs].astype(np.float64)
                    t=d.astype(np.float64)+p
                    d=rz32(t) if mode=='rz' else t.astype(np.float32)
            outs.append(d)
        Sre,Sim=outs
        anga=-2*np.pi*((ks*s)%N)/N
        Are=f32(np.cos(anga)); Aim=f32(np.sin(anga))
        if s==rem and rem!=0:
            R[:,:]=acc_re.astype(np.float64)+1j*acc_im.astype(np.float64)
        # complex fma in f32 (emulate with float32 ops, fused-ish)
        acc_re=f32(acc_re.astype(np.float64)+Are.astype(np.float64)*Sre) ; acc_re=f32(acc_re.astype(np.float64)-Aim.astype(np.float64)*Sim)
        acc_im=f32(acc_im.astype(np.float64)+Are.astype(np.float64)*Sim) ; acc_im=f32(acc_im.astype(np.float64)+Aim.astype(np.float64)*Sre)
    C=acc_re.astype(np.float64)+1j*acc_im.astype(np.float64)
    # combine in f64 (isolating partial-sum error)
    i=np.arange(q+1)
    ph=np.exp(-2j*np.pi*((ks[None,:]*(i[:,None]*H))%N)/N)
    Xt=np.zeros((T,nk),np.complex128)
    for t in range(T):
        Xt[t]=(ph[:q]*C[t:t+q]).sum(0)+ph[q]*R[t+q]
    # reference
    ref=np.zeros((T,nk),np.complex128)
    for t in range(T):
        w=x[t*H:t*H+N].astype(np.float64)
        ref[t]=np.fft.rfft(w)[ks]
    err=np.abs(Xt-ref).max(1)/np.abs(ref).max(1)
    return err.max(), np.median(err)
for G in (16, 32, 64, 128, 368):
    for mode in ('rn','rz'):
        print(G,mode,run(G,mode))
