"""Monte Carlo behind DESIGN.md section 8: distinct 128-byte lines touched by one warp-wide window-row load of the gather (4 x 4 columns x two
depths hugging a randomly oriented plane), for the row-major slice (8 pair entries per line) and for 2-D tiled lines (4 entries x 2 rows,
2 entries x 4 rows).  All lanes active (the kernel predicates some off): 7.4 vs 5.7 lines.  python tools/sim_slice_lines.py"""
import numpy as np
rng=np.random.default_rng(0)
rho=1.9; K=4
def rand_rot():
    q=rng.normal(size=4); q/=np.linalg.norm(q)
    a,b,c,d=q
    return np.array([[a*a+b*b-c*c-d*d,2*(b*c-a*d),2*(b*d+a*c)],[2*(b*c+a*d),a*a-b*b+c*c-d*d,2*(c*d-a*b)],[2*(b*d-a*c),2*(c*d+a*b),a*a-b*b-c*c+d*d]])
def lines_rowmajor(i,j,pitch=268): # entry (i,j): 8 entries per line along j
    return set(zip(i.tolist(), (j//8).tolist()))
def lines_tiled(i,j): # line = 4 entries x 2 rows
    return set(zip((i//2).tolist(), (j//4).tolist()))
res={'row':[], 'tile':[], 'tile24':[], 'lanes':[]}
for trial in range(4000):
    M=rand_rot(); e1,e2,n=M[:,0],M[:,1],M[:,2]
    # dominant axis of the normal = depth axis d; permute so d is index 2
    d=np.argmax(np.abs(n)); perm=[k for k in range(3) if k!=d]+[d]
    e1p,e2p,npm=e1[perm],e2[perm],n[perm]
    # stick origin random; 4x4 columns; each column: depths where |h|<=r; two lanes per column (parity), step s
    org=rng.uniform(-100,100,size=3); 
    # choose origin depth so plane crosses: solve h=0 at centre column
    ca,cb=org[0]+1.5,org[1]+1.5
    dc=-(ca*npm[0]+cb*npm[1])/npm[2]
    base_d=np.floor(dc)-3
    for s in range(4):
        I=[];J=[];act=[]
        for lb in range(4):
            for la in range(4):
                a=org[0]+la; b=org[1]+lb
                c0=-(a*npm[0]+b*npm[1])/npm[2]; hw=1.9/abs(npm[2])
                lo=int(np.ceil(c0-hw)); hi=int(np.floor(c0+hw))
                for par in range(2):
                    tau=lo+par+2*s
                    if tau>hi: continue
                    u=np.array([a,b,tau]); al=u@e1p; be=u@e2p; h=u@npm
                    jw=int(np.ceil(al-rho)); iw=int(np.ceil(be-rho))
                    I.append(iw);J.append(jw); act.append((al-jw,be-iw,h))
        if len(I)<8: continue
        I=np.array(I);J=np.array(J)
        # inner row t=1, pair q=0 (entry j) -> count lines over lanes (ignore predication)
        for t in (1,):
            for q in (0,):
                i=I+t+1000; j=J+2*q+1000
                res['row'].append(len(lines_rowmajor(i,j)))
                res['tile'].append(len(lines_tiled(i,j)))
                res['tile24'].append(len(set(zip((i//4).tolist(),(j//2).tolist()))))
                res['lanes'].append(len(I))
for k,v in res.items(): print(k, np.mean(v))
