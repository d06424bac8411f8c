"""Prototype 2: primal log-barrier Newton (interior point) on the full collocation NLP, physical inputs with bounds,
exact Lagrangian Hessian + barrier Hessian, sparse KKT, fraction-to-boundary, l1 merit.  CPU / SciPy -- R&D for round 2."""
import sys, time, numpy as np, sympy as sp, scipy.sparse as ss, scipy.sparse.linalg as sl
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'drone-sim-python_b200'))
from d2d_b200 import opty_utils as orc      # triangle() initial guess only
g = 9.81
psi, phi, v, lx, ly, lp, h, kvel, kbank, vsp, nrm = sp.symbols('psi phi v lx ly lp h kvel kbank vsp nrm')
fx, fy, fp = h * v * sp.cos(psi), h * v * sp.sin(psi), h * g * sp.tan(phi) / v
cost = nrm * (kvel * (v - vsp) ** 2 + kbank * phi ** 2)
Lnode = cost - lx * fx - ly * fy - lp * fp
var = [psi, phi, v]
args = [psi, phi, v, lx, ly, lp, h, kvel, kbank, vsp, nrm]
f_inc = sp.lambdify(args, [fx, fy, fp], 'numpy')
f_jac = sp.lambdify(args, list(sp.Matrix([fx, fy, fp]).jacobian(var)), 'numpy')
f_cost = sp.lambdify(args, cost, 'numpy')
f_cgrad = sp.lambdify(args, [sp.diff(cost, q) for q in var], 'numpy')
f_hess = sp.lambdify(args, list(sp.hessian(Lnode, var)), 'numpy')
B = lambda q, N: np.broadcast_to(np.asarray(q, float), (N,))

def solve(N, hh, p0, p1, wind, pb, vb, kv, kb, vspv, z0, tol=1e-8, maxit=300, verbose=False):
    n, m, rb = 5 * N, 3 * (N - 1) + 6, 3 * (N - 1)
    idx = lambda k: 5 * np.arange(N) + k
    par = lambda X, L: (X[:, 2], X[:, 3], X[:, 4], L[:, 0], L[:, 1], L[:, 2], hh, kv, kb, vspv, 1.0 / N)
    rows, cols, vals = [], [], []
    for k in range(3):
        rows += list(3 * np.arange(N - 1) + k) * 2; cols += list(idx(k)[1:]) + list(idx(k)[:-1]); vals += [1.0] * (N - 1) + [-1.0] * (N - 1)
        rows += [rb + k, rb + 3 + k]; cols += [k, 5 * (N - 1) + k]; vals += [1.0, 1.0]
    Jlin = ss.csr_matrix((vals, (rows, cols)), shape=(m, n))
    lo = np.full(n, -np.inf); hi = np.full(n, np.inf)
    lo[idx(3)], hi[idx(3)], lo[idx(4)], hi[idx(4)] = pb[0], pb[1], vb[0], vb[1]
    bnd = np.isfinite(lo)
    z = z0.copy(); z[bnd] = np.clip(z[bnd], lo[bnd] + 0.05 * (hi[bnd] - lo[bnd]), hi[bnd] - 0.05 * (hi[bnd] - lo[bnd]))
    lam = np.zeros(m); Z0 = np.zeros((N, 3))
    def con_cost(z):
        X = z.reshape(N, 5)
        inc = np.array([B(q, N) for q in f_inc(*par(X, Z0))])
        c = np.zeros(m)
        c[:rb] = ((X[1:, :3] - X[:-1, :3]).T - inc[:, 1:] + np.array([hh * wind[0], hh * wind[1], 0.0])[:, None]).T.ravel()
        c[rb:rb + 3] = X[0, :3] - p0; c[rb + 3:] = X[-1, :3] - p1
        return c, float(np.sum(B(f_cost(*par(X, Z0)), N)))
    def barrier(z, mu): return -mu * (np.log(z[bnd] - lo[bnd]).sum() + np.log(hi[bnd] - z[bnd]).sum())
    mu, nu, reg, t0, total = 1e-1, 10.0, 1e-8, time.time(), 0
    for outer in range(12):
        for it in range(maxit):
            total += 1
            X = z.reshape(N, 5)
            c, cost = con_cost(z)
            L = np.zeros((N, 3)); L[1:] = lam[:rb].reshape(N - 1, 3)
            Jl = f_jac(*par(X, Z0)); Jn = np.array([[B(Jl[3 * a + b], N) for b in range(3)] for a in range(3)])
            rows, cols, vals = [], [], []
            for a in range(3):
                for b in range(3):
                    rows += list(3 * np.arange(N - 1) + a); cols += list(idx(2 + b)[1:]); vals += list(-Jn[a, b, 1:])
            J = Jlin + ss.csr_matrix((vals, (rows, cols)), shape=(m, n))
            gradf = np.zeros(n)
            cg = [B(q, N) for q in f_cgrad(*par(X, Z0))]
            gradf[idx(2)], gradf[idx(3)], gradf[idx(4)] = cg
            gb = np.zeros(n); hb = np.zeros(n)
            gb[bnd] = -mu / (z[bnd] - lo[bnd]) + mu / (hi[bnd] - z[bnd]); hb[bnd] = mu / (z[bnd] - lo[bnd]) ** 2 + mu / (hi[bnd] - z[bnd]) ** 2
            Hl = f_hess(*par(X, L)); Hn = np.array([[B(Hl[3 * a + b], N) for b in range(3)] for a in range(3)])
            rows, cols, vals = [], [], []
            for a in range(3):
                for b in range(3):
                    rows += list(idx(2 + a)); cols += list(idx(2 + b)); vals += list(Hn[a, b])
            H0 = ss.csr_matrix((vals, (rows, cols)), shape=(n, n)) + ss.diags(hb)
            rd = gradf + gb + J.T @ lam
            err = max(np.abs(rd).max(), np.abs(c).max())
            if verbose and it % 5 == 0: print(f"  mu {mu:.0e} it {it:3d} cost {cost:.8e} |c| {np.abs(c).max():.2e} |rd| {np.abs(rd).max():.2e} reg {reg:.1e}")
            if err < max(tol, 10 * mu if mu > 1e-9 else tol): break
            while True:
                H = H0 + reg * ss.identity(n)
                K = ss.bmat([[H, J.T], [J, -1e-10 * ss.identity(m)]], format='csc')
                sol = sl.spsolve(K, -np.concatenate([rd, c]))
                dz, dl = sol[:n], sol[n:]
                if np.isfinite(sol).all() and dz @ (H @ dz) > 1e-12 * (dz @ dz): break
                reg = max(reg * 10, 1e-6)
                if reg > 1e8: break
            # fraction to the boundary
            amax = 1.0
            neg = bnd & (dz < 0); pos = bnd & (dz > 0)
            if neg.any(): amax = min(amax, 0.995 * np.min((z[neg] - lo[neg]) / -dz[neg]))
            if pos.any(): amax = min(amax, 0.995 * np.min((hi[pos] - z[pos]) / dz[pos]))
            nu = max(nu, 2 * np.abs(lam + dl).max())
            phi0 = cost + barrier(z, mu) + nu * np.abs(c).sum()
            dphi = (gradf + gb) @ dz - nu * np.abs(c).sum()
            alpha = amax
            while alpha > 1e-8:
                zt = z + alpha * dz; ct, costt = con_cost(zt)
                if costt + barrier(zt, mu) + nu * np.abs(ct).sum() <= phi0 + 1e-4 * alpha * dphi: break
                alpha *= 0.5
            z = z + alpha * dz; lam = lam + alpha * dl
            reg = max(reg / 3, 1e-8) if alpha > 0.5 * amax else min(reg * 3, 1e4)
        if mu <= 1e-9: break
        mu = max(mu * 0.2, 1e-9) if mu > 1e-9 else mu
    c, cost = con_cost(z)
    return z.reshape(N, 5), total, cost, np.abs(c).max(), time.time() - t0

if __name__ == "__main__":
    cases = {"exp_0": (101, 0.1, (0, 0, 0), (0, 30, np.pi), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 0., 12.),
             "c3": (1001, 0.02, (0, 0, 0), (0, 30, np.pi), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 0., 12.),
             "exp_1": (101, 0.1, (0, 0, 0), (100, 0, 0), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 50., 12.),
             "bank": (201, 0.05, (0, 0, 0), (80, 30, 0), (-np.deg2rad(40), np.deg2rad(40)), (9., 15.), 1., 1., 12.)}
    for nm, (N, hh, p0, p1, pb, vb, kv, kb, vspv) in cases.items():
        xg, yg, psig, phig, vg = orc.triangle(p0[:2], p1[:2], 12., (N - 1) * hh, N, go_left=-1.)
        z0 = np.zeros((N, 5)); z0[:, 0], z0[:, 1], z0[:, 2], z0[:, 3], z0[:, 4] = xg, yg, np.unwrap(psig), 0.0, 12.0
        X, it, cost, cm, dt = solve(N, hh, np.array(p0, float), np.array(p1, float), (0., 0.), pb, vb, kv, kb, vspv, z0.ravel(), verbose="-v" in sys.argv)
        print(f"{nm}: {it} Newton iterations, {dt:.2f}s, cost {cost:.8e}, |c| {cm:.2e}, phi in [{X[:,3].min():.3f},{X[:,3].max():.3f}] v in [{X[:,4].min():.2f},{X[:,4].max():.2f}]")
