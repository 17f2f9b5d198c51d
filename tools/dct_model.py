"""numpy model of the kernel DCT algorithm (index maps + math), checked vs scipy.
Mirrors csrc/dct_core.cuh phase by phase; vectorised over butterfly id u (== thread work item)."""
import numpy as np, scipy.fftpack as fp

def radices(M):
    r=[]; m=M
    while m%8==0 and m>=8: r.append(8); m//=8
    while m%4==0 and m>=4: r.append(4); m//=4
    while m%2==0 and m>=2: r.append(2); m//=2
    assert m==1
    return r

def dif_forward(z, rad):
    """in-place DIF; z complex[M] natural -> digit-reversed"""
    M=len(z); Lb=M
    for r in rad:
        st=Lb//r
        u=np.arange(M//r); B=(u//st)*Lb; j=u%st
        idx=B[:,None]+j[:,None]+np.arange(r)[None,:]*st           # [u][q]
        x=z[idx]
        q=np.arange(r); p=np.arange(r)
        D=np.exp(-2j*np.pi*np.outer(q,p)/r)                        # [q][p]
        y=x@D                                                      # [u][p]
        tw=np.exp(-2j*np.pi*(j[:,None]*p[None,:])/Lb)
        z[idx]=y*tw
        Lb=st
    return z

def dit_inverse(z, rad):
    """in-place DIT inverse (unnormalised, e^{+}); digit-reversed in -> natural out. exact reverse of dif_forward"""
    M=len(z)
    Lbs=[]; Lb=M
    for r in rad: Lbs.append(Lb); Lb//=r
    for r,Lb in reversed(list(zip(rad,Lbs))):
        st=Lb//r
        u=np.arange(M//r); B=(u//st)*Lb; j=u%st
        idx=B[:,None]+j[:,None]+np.arange(r)[None,:]*st
        p=np.arange(r); q=np.arange(r)
        tw=np.exp(+2j*np.pi*(j[:,None]*p[None,:])/Lb)
        x=z[idx]*tw                                                # [u][p]
        D=np.exp(+2j*np.pi*np.outer(p,q)/r)                        # [p][q]
        z[idx]=x@D
    return z

def pos_of_freq(k, M, rad):
    """position of frequency k after dif_forward"""
    pos=0; Lb=M; kk=k
    for r in rad:
        p=kk%r; kk//=r
        pos+=p*(Lb//r); Lb//=r
    return pos

def makhoul(x):
    N=len(x); v=np.empty(N); v[:N//2]=x[0::2]; v[N-1-np.arange(N//2)]=x[1::2]; return v
def imakhoul(v):
    N=len(v); x=np.empty(N); x[0::2]=v[:N//2]; x[1::2]=v[N-1-np.arange(N//2)]; return x

def dct2(x):
    N=len(x); M=N//2; rad=radices(M)
    v=makhoul(x); z=(v[0::2]+1j*v[1::2]).copy()
    z=dif_forward(z,rad)
    P=np.array([pos_of_freq(k,M,rad) for k in range(M)])
    Z=lambda k: z[P[k]]
    C=np.empty(N); s=np.sqrt(2/N)
    om=lambda m: np.exp(-1j*np.pi*m/(2*N))
    Z0=Z(0); C[0]=np.sqrt(1/N)*(Z0.real+Z0.imag); C[M]=s*np.cos(np.pi/4)*(Z0.real-Z0.imag)
    if M>=2:
        k=M//2; A=om(k)*np.conj(Z(k)); C[k]=s*A.real; C[N-k]=-s*A.imag
    for k in range(1,M//2):
        a=Z(k); b=np.conj(Z(M-k))
        E=(a+b)/2; O=(a-b)/(2j); tO=om(4*k)*O
        A=om(k)*(E+tO); A2=om(M-k)*np.conj(E-tO)
        C[k]=s*A.real; C[N-k]=-s*A.imag; C[M-k]=s*A2.real; C[M+k]=-s*A2.imag
    return C

def dct3(C):
    N=len(C); M=N//2; rad=radices(M)
    P=np.array([pos_of_freq(k,M,rad) for k in range(M)])
    z=np.empty(M,complex); s=np.sqrt(2/N)
    om=lambda m: np.exp(-1j*np.pi*m/(2*N))
    V0=C[0]*np.sqrt(N); VM=C[M]/(s*np.cos(np.pi/4))
    z[P[0]]=(V0+VM)/2+1j*(V0-VM)/2
    if M>=2:
        k=M//2; A=(C[k]-1j*C[N-k])/s; V=np.conj(om(k))*A; z[P[k]]=np.conj(V)
    for k in range(1,M//2):
        A=(C[k]-1j*C[N-k])/s; A2=(C[M-k]-1j*C[M+k])/s
        V=np.conj(om(k))*A; V2=np.conj(om(M-k))*A2
        E=(V+np.conj(V2))/2; O=(V-np.conj(V2))*np.conj(om(4*k))/2
        z[P[k]]=E+1j*O; z[P[M-k]]=np.conj(E)+1j*np.conj(O)
    z=dit_inverse(z,rad)/M
    v=np.empty(N); v[0::2]=z.real; v[1::2]=z.imag
    return imakhoul(v)

if __name__=="__main__":
    rng=np.random.default_rng(1)
    for N in (4,8,16,32,64,128,256,512,1024,2048):
        x=rng.random(N)
        C=dct2(x); R=fp.dct(x,norm='ortho')
        xi=dct3(R); 
        print(N, radices(N//2), np.abs(C-R).max(), np.abs(xi-x).max())
