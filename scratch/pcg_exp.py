import time, sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from visual_underwater_slam_b200 import synthetic
from oracle import lm

def build(n_poses, n_lm, n_loops, seed, iters_before):
    d = synthetic.make_trajectory_graph(n_poses, seed=seed, n_landmarks=n_lm, n_loops=n_loops, pixel_noise=1.0)
    prob = d['graph'].to_problem(d['initial'])
    lay = lm.Layout(prob); vals = lm.values_of(prob)
    lam = 1e-5
    for it in range(iters_before):
        J,b = lm.linearize(prob, vals, lay)
        delta = lm.solve_damped(J,b,lam,lay)
        vals = lm.retract(vals, delta, lay); lam/=10
    J,b = lm.linearize(prob, vals, lay)
    return prob, lay, J, b, lam

def reduced_system(lay, J, b, lam):
    H = (J.T@J).tocsc() + lam*sp.identity(lay.n, format='csc'); g = J.T@b
    n = lay.n; N = lay.nx
    # node ordering: for node i: X_i(6), V_i(3); then bias
    perm = np.empty(9*N+6, dtype=np.int64)
    for c in range(6): perm[c:9*N:9] = lay.ox + 6*np.arange(N) + c
    for c in range(3): perm[6+c:9*N:9] = lay.ov + 3*np.arange(N) + c
    perm[9*N:] = np.arange(6)
    lmi = np.arange(lay.ol, lay.ov)
    Hc = H[perm][:,perm]; gc = g[perm]
    if lay.nl:
        E = H[perm][:,lmi]; C = H[lmi][:,lmi].tocsr()
        nl=lay.nl; Cd=np.zeros((nl,3,3)); Cc=C.tocoo(); Cd[Cc.row//3,Cc.row%3,Cc.col%3]=Cc.data
        Ci=np.linalg.inv(Cd)
        ii=(3*np.arange(nl)[:,None,None]+np.arange(3)[None,:,None])+np.zeros((1,1,3),dtype=np.int64)
        jj=(3*np.arange(nl)[:,None,None]+np.arange(3)[None,None,:])+np.zeros((1,3,1),dtype=np.int64)
        Cinv=sp.csc_matrix((Ci.ravel(),(ii.ravel(),jj.ravel())),shape=(3*nl,3*nl))
        ECi=(E@Cinv).tocsc()
        S=(Hc-ECi@E.T).tocsc(); gs=gc-ECi@g[lmi]
    else:
        S=Hc.tocsc(); gs=gc
    return S, gs

def band_part(S, N, k):
    """keep blocks (i,j) of the node part with super-node index |I-J|<=1 where I=i//k (block tridiagonal over supernodes); drop bias coupling"""
    Sc = S.tocoo()
    r, c = Sc.row, Sc.col
    node = (r < 9*N) & (c < 9*N)
    I = (r//9)//k; Jn = (c//9)//k
    keep = node & (np.abs(I-Jn) <= 1)
    bias = (r >= 9*N) & (c >= 9*N)
    keep |= bias
    return sp.csc_matrix((Sc.data[keep], (r[keep], c[keep])), shape=S.shape)

def pcg(S, g, Minv, tol=1e-10, maxit=5000):
    x = np.zeros_like(g); r = g.copy(); z = Minv(r); p = z.copy(); rz = r@z
    g0 = np.linalg.norm(g); hist=[]
    for it in range(maxit):
        Ap = S@p; a = rz/(p@Ap); x += a*p; r -= a*Ap
        rn = np.linalg.norm(r)/g0; hist.append(rn)
        if rn < tol: break
        z = Minv(r); rz2 = r@z; p = z + (rz2/rz)*p; rz = rz2
    return x, it+1, hist

def run(name, n_poses, n_lm, n_loops, iters_before):
    prob, lay, J, b, lam = build(n_poses, n_lm, n_loops, 11, iters_before)
    S, gs = reduced_system(lay, J, b, lam)
    N = lay.nx
    xs = lm._splu_sym(S).solve(gs)
    print(f"== {name}: N={N} lam={lam:g} dim={S.shape[0]} nnz={S.nnz}")
    # (a) block-Jacobi
    D = band_part(S, N, 1)
    # pure block diag: supernode |I-J|==0
    Sc=S.tocoo(); keep=((Sc.row//9)==(Sc.col//9)); Dj=sp.csc_matrix((Sc.data[keep],(Sc.row[keep],Sc.col[keep])),shape=S.shape)
    luj=lm._splu_sym(Dj)
    x,it,h=pcg(S,gs,luj.solve, maxit=3000); print('block-jacobi iters',it,'final',h[-1], 'err vs direct', np.linalg.norm(x-xs)/np.linalg.norm(xs))
    for k in [1,5,10]:
        M = band_part(S,N,k)
        try:
            lu = lm._splu_sym(M)
            # check SPD-ness roughly via solve of random
            x,it,h=pcg(S,gs,lu.solve, maxit=500); print(f'tridiag-supernode k={k} (no bias coupling) iters',it,'final',h[-1],'err', np.linalg.norm(x-xs)/np.linalg.norm(xs))
        except Exception as e:
            print('k',k,'failed',e)
        # with bias bordering: M2 = band + bias border (exact)
        Sc=S.tocoo(); r,c=Sc.row,Sc.col
        border=((r>=9*N)^(c>=9*N))
        M2 = M + sp.csc_matrix((Sc.data[border],(r[border],c[border])),shape=S.shape)
        lu2=lm._splu_sym(M2)
        x,it,h=pcg(S,gs,lu2.solve, maxit=500); print(f'   + bias border exact: iters',it,'final',h[-1],'err', np.linalg.norm(x-xs)/np.linalg.norm(xs))

if __name__=='__main__':
    which = sys.argv[1]
    if which=='c1': 
        run('C1 first iter', 2000, 0, 50, 0); run('C1 late iter', 2000,0,50,4)
    if which=='c2':
        run('C2 first iter', 1500, 6000, 0, 0); run('C2 late', 1500, 6000, 0, 4)
