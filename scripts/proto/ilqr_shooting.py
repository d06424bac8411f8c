"""Prototype: AL + box-iLQR on the backward-Euler grid (single aircraft, input cost)."""
import numpy as np, time, sys
g=9.81
def rollout(s0,U,h,w):
    N=len(U)+1; S=np.zeros((N,3)); S[0]=s0
    for i in range(1,N):
        phi,v=U[i-1]; psi=S[i-1,2]+h*g*np.tan(phi)/v
        S[i]=(S[i-1,0]+h*(v*np.cos(psi)-w[0]), S[i-1,1]+h*(v*np.sin(psi)-w[1]), psi)
    return S
def total(S,U,p1,lam,rho,kv,kb,vsp,nrm):
    c=S[-1]-p1
    cost=nrm*np.sum(kv*(U[:,1]-vsp)**2+kb*U[:,0]**2)
    return cost+lam@c+0.5*rho*c@c, cost, c
def boxqp2(H,q,lo,hi):
    """min 1/2 d'Hd+q'd, lo<=d<=hi, 2-D.  Returns d, free mask."""
    d=-np.linalg.solve(H,q)
    if np.all(d>=lo) and np.all(d<=hi): return d,np.array([True,True])
    best=None
    for k in (0,1):          # clamp coordinate k at a bound, solve other
        for bk in (lo[k],hi[k]):
            j=1-k; dj=-(q[j]+H[j,k]*bk)/H[j,j]; free=np.array([False,False]); 
            if dj<lo[j]: dj=lo[j]
            elif dj>hi[j]: dj=hi[j]
            else: free[j]=True
            d=np.zeros(2); d[k]=bk; d[j]=dj
            val=0.5*d@H@d+q@d
            if best is None or val<best[0]: best=(val,d,free)
    return best[1],best[2]
def solve(s0,p1,N,h,w,kv,kb,vsp,lb,ub,U,verbose=True):
    nrm=1.0/N; lam=np.zeros(3); rho=10.; mu=1e-6
    S=rollout(s0,U,h,w); J,cost,c=total(S,U,p1,lam,rho,kv,kb,vsp,nrm); nit=0
    for outer in range(30):
        for it in range(200):
            # backward
            Wx=lam+rho*(S[-1]-p1); Wxx=rho*np.eye(3)
            ks=np.zeros((N-1,2)); Ks=np.zeros((N-1,2,3)); dV=0.
            for i in range(N-1,0,-1):
                phi,v=U[i-1]; psi=S[i,2]; sp,cp=np.sin(psi),np.cos(psi)
                a=h*g/(v*np.cos(phi)**2); b=-h*g*np.tan(phi)/v**2
                A=np.array([[1,0,-h*v*sp],[0,1,h*v*cp],[0,0,1]])
                B=np.array([[-h*v*sp*a, h*cp-h*v*sp*b],[h*v*cp*a, h*sp+h*v*cp*b],[a,b]])
                lu=nrm*np.array([2*kb*phi,2*kv*(v-vsp)]); luu=nrm*np.diag([2*kb,2*kv])
                Qs=A.T@Wx; Qu=B.T@Wx+lu; Qss=A.T@Wxx@A; Qus=B.T@Wxx@A; Quu=B.T@Wxx@B+luu+mu*np.eye(2)
                d,free=boxqp2(Quu,Qu,lb-U[i-1],ub-U[i-1])
                K=np.zeros((2,3))
                if free.all(): K=-np.linalg.solve(Quu,Qus)
                elif free.any():
                    j=int(np.argmax(free)); K[j]=-Qus[j]/Quu[j,j]
                ks[i-1]=d; Ks[i-1]=K
                dV+=d@Qu+0.5*d@Quu@d
                Wx=Qs+K.T@Quu@d+K.T@Qu+Qus.T@d; Wxx=Qss+K.T@Quu@K+K.T@Qus+Qus.T@K; Wxx=0.5*(Wxx+Wxx.T)
            # forward with line search
            alpha=1.0; ok=False
            for ls in range(12):
                Sn=np.zeros_like(S); Sn[0]=s0; Un=np.zeros_like(U)
                for i in range(1,N):
                    u=U[i-1]+alpha*ks[i-1]+Ks[i-1]@(Sn[i-1]-S[i-1]); u=np.clip(u,lb,ub); Un[i-1]=u
                    psi=Sn[i-1,2]+h*g*np.tan(u[0])/u[1]
                    Sn[i]=(Sn[i-1,0]+h*(u[1]*np.cos(psi)-w[0]), Sn[i-1,1]+h*(u[1]*np.sin(psi)-w[1]), psi)
                Jn,costn,cn=total(Sn,Un,p1,lam,rho,kv,kb,vsp,nrm)
                if Jn<J-1e-4*alpha*abs(dV)*0 - 0 and Jn<J: ok=True; break
                alpha*=0.5
            nit+=1
            if not ok: mu=min(mu*10,1e6); 
            else:
                dJ=J-Jn; S,U,J,cost,c=Sn,Un,Jn,costn,cn; mu=max(mu/3,1e-9)
                if dJ<1e-12*max(1,abs(J)): break
            if not ok and mu>=1e6: break
        if verbose: print(outer,'inner',it+1,'cost %.8e |c| %.2e rho %.0f mu %.1e'%(cost,np.abs(c).max(),rho,mu))
        if np.abs(c).max()<1e-8: break
        lam=lam+rho*c; rho=min(rho*5,1e8) if np.abs(c).max()>1e-7 else rho
        J,cost,c=total(S,U,p1,lam,rho,kv,kb,vsp,nrm); mu=1e-6
    return S,U,nit,cost,c
for (p1,t1,hz,kb,phm,vb) in (((80.,30.,0.),10.,20.,1.,40.,(9.,15.)), ((0.,30.,np.pi),10.,10.,0.,30.,(9.,14.)), ((0.,30.,np.pi),20.,50.,0.,30.,(9.,14.))):
    N=int(t1*hz)+1; h=t1/(N-1)
    lb=np.array([-np.deg2rad(phm),vb[0]]); ub=np.array([np.deg2rad(phm),vb[1]])
    U=np.tile([0.1 if kb==0 else 0.0,12.],(N-1,1))
    t0=time.time(); S,U,nit,cost,c=solve(np.zeros(3),np.array(p1),N,h,(0.,0.),1.,kb,12.,lb,ub,U)
    print('N',N,'iters',nit,'cost',cost + (1.0/N)*0,'|c|',np.abs(c).max(),'time',time.time()-t0)
