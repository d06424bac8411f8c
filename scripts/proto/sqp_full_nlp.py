"""Prototype: Newton-type SQP on the full collocation NLP (states + inputs), exact Lagrangian Hessian (node-block diagonal),
sparse KKT solve, l1-merit line search, input bounds by the sin substitution.  CPU / SciPy only -- R&D for round 2."""
import sys, time, numpy as np, sympy as sp, scipy.sparse as ss, scipy.sparse.linalg as sl
g = 9.81
# node-level nonlinear part of the Lagrangian: defects are  x_i - x_{i-1} - h (v cos psi - wx) etc.
psi, tf, tv, lx, ly, lp, h, mf, hf, mv, hv, kvel, kbank, vsp, nrm = sp.symbols('psi tf tv lx ly lp h mf hf mv hv kvel kbank vsp nrm')
phi = mf + hf * sp.sin(tf); v = mv + hv * sp.sin(tv)
fx, fy, fp = h * v * sp.cos(psi), h * v * sp.sin(psi), h * g * sp.tan(phi) / v          # "increments" (without wind)
cost = nrm * (kvel * (v - vsp) ** 2 + kbank * phi ** 2)
Lnode = cost - lx * fx - ly * fy - lp * fp
var = [psi, tf, tv]
args = [psi, tf, tv, lx, ly, lp, h, mf, hf, mv, hv, kvel, kbank, vsp, nrm]
f_inc = sp.lambdify(args, [fx, fy, fp], 'numpy')
f_jac = sp.lambdify(args, list(sp.Matrix([fx, fy, fp]).jacobian(var)), 'numpy')
f_cost = sp.lambdify(args, cost, 'numpy')
f_cgrad = sp.lambdify(args, [sp.diff(cost, q) for q in var], 'numpy')
f_hess = sp.lambdify(args, list(sp.hessian(Lnode, var)), 'numpy')

def solve(N, hh, p0, p1, wind, pb, vb, kv, kb, vspv, z0=None, tol=1e-8, maxit=200, verbose=True):
    mfv, hfv, mvv, hvv = (pb[0] + pb[1]) / 2, (pb[1] - pb[0]) / 2, (vb[0] + vb[1]) / 2, (vb[1] - vb[0]) / 2
    n = 5 * N                                        # per node: x, y, psi, theta_phi, theta_v
    ix = lambda i: 5 * i; iy = lambda i: 5 * i + 1; ip = lambda i: 5 * i + 2; itf = lambda i: 5 * i + 3; itv = lambda i: 5 * i + 4
    m = 3 * (N - 1) + 6
    par = lambda ps, a, b, lxx, lyy, lpp: (ps, a, b, lxx, lyy, lpp, hh, mfv, hfv, mvv, hvv, kv, kb, vspv, 1.0 / N)
    z = np.zeros(n) if z0 is None else z0.copy()
    lam = np.zeros(m)
    # constant (linear) part of the constraint Jacobian
    rows, cols, vals = [], [], []
    for i in range(1, N):
        r = 3 * (i - 1)
        for k, idx in enumerate((ix, iy, ip)):
            rows += [r + k, r + k]; cols += [idx(i), idx(i - 1)]; vals += [1.0, -1.0]
    rb = 3 * (N - 1)
    for k, idx in enumerate((ix, iy, ip)):
        rows += [rb + k, rb + 3 + k]; cols += [idx(0), idx(N - 1)]; vals += [1.0, 1.0]
    Jlin = ss.csr_matrix((vals, (rows, cols)), shape=(m, n))
    node = np.arange(1, N)
    def evaluate(z, lam):
        X = z.reshape(N, 5)
        zero = np.zeros(N)
        inc = np.array(f_inc(*par(X[:, 2], X[:, 3], X[:, 4], zero, zero, zero)))          # (3, N)
        c = np.zeros(m)
        c[:rb] = ((X[1:, :3] - X[:-1, :3]).T - inc[:, 1:] + np.array([hh * wind[0], hh * wind[1], 0.0])[:, None]).T.ravel()
        c[rb:rb + 3] = X[0, :3] - p0; c[rb + 3:] = X[-1, :3] - p1
        cost = np.sum(f_cost(*par(X[:, 2], X[:, 3], X[:, 4], zero, zero, zero)))
        return c, cost, inc
    def merit(z, nu):
        c, cost, _ = evaluate(z, lam)
        return cost + nu * np.abs(c).sum(), c, cost
    nu, reg = 10.0, 1e-8
    t0 = time.time()
    for it in range(maxit):
        X = z.reshape(N, 5)
        zero = np.zeros(N)
        c, cost, inc = evaluate(z, lam)
        # multipliers per node for the three defect rows that involve node i's nonlinear terms (rows of node i, i >= 1)
        L = np.zeros((N, 3)); L[1:] = lam[:rb].reshape(N - 1, 3)
        Jl = f_jac(*par(X[:, 2], X[:, 3], X[:, 4], zero, zero, zero))                         # 9 entries of d inc / d (psi, tf, tv)
        Jn = np.array([[np.broadcast_to(Jl[3 * a + b], (N,)) for b in range(3)] for a in range(3)], dtype=float)
        rows, cols, vals = [], [], []
        for a in range(3):
            for b, idx in enumerate((ip, itf, itv)):
                rows += list(3 * (node - 1) + a); cols += list(idx(node)); vals += list(-Jn[a, b, 1:])
        J = Jlin + ss.csr_matrix((vals, (rows, cols)), shape=(m, n))
        gradf = np.zeros(n)
        cg = np.array([np.broadcast_to(q, (N,)) for q in f_cgrad(*par(X[:, 2], X[:, 3], X[:, 4], zero, zero, zero))], dtype=float)
        gradf[2::5], gradf[3::5], gradf[4::5] = cg[0], cg[1], cg[2]
        # node 0's inputs only enter the cost; its increments are not constraints -> multipliers 0 there
        Hl = f_hess(*par(X[:, 2], X[:, 3], X[:, 4], L[:, 0], L[:, 1], L[:, 2]))
        Hn = np.array([[np.broadcast_to(Hl[3 * a + b], (N,)) for b in range(3)] for a in range(3)], dtype=float)
        kkt_res = max(np.abs(gradf + J.T @ lam).max(), np.abs(c).max())
        if verbose and (it % 5 == 0 or kkt_res < tol):
            print(f"it {it:3d} cost {cost:.8e} |c| {np.abs(c).max():.2e} |grad L| {np.abs(gradf + J.T @ lam).max():.2e} reg {reg:.1e} nu {nu:.1e}")
        if kkt_res < tol:
            break
        while True:                                   # inertia-style regularisation: raise reg until the step is a descent direction for the merit
            rows, cols, vals = [], [], []
            for a, ia in enumerate((ip, itf, itv)):
                for b, ib in enumerate((ip, itf, itv)):
                    rows += list(ia(np.arange(N))); cols += list(ib(np.arange(N))); vals += list(Hn[a, b])
            H = ss.csr_matrix((vals, (rows, cols)), shape=(n, n)) + reg * ss.identity(n)
            K = ss.bmat([[H, J.T], [J, -1e-10 * ss.identity(m)]], format='csc')
            sol = sl.spsolve(K, -np.concatenate([gradf + J.T @ lam, c]))
            dz, dl = sol[:n], sol[n:]
            curv = dz @ (H @ dz)
            if np.isfinite(sol).all() and curv > 1e-12 * (dz @ dz):
                break
            reg = max(reg * 10, 1e-6)
            if reg > 1e6: break
        lam_new = lam + dl
        nu = max(nu, 2 * np.abs(lam_new).max())
        m0, _, _ = merit(z, nu)
        dmer = gradf @ dz - nu * np.abs(c).sum()
        alpha = 1.0
        while alpha > 1e-6:
            m1, _, _ = merit(z + alpha * dz, nu)
            if m1 <= m0 + 1e-4 * alpha * dmer: break
            alpha *= 0.5
        z = z + alpha * dz; lam = lam + alpha * dl
        reg = max(reg / 3, 1e-8) if alpha == 1.0 else min(reg * 2, 1e3)
    X = z.reshape(N, 5)
    return X, it, cost, np.abs(c).max(), time.time() - t0

if __name__ == "__main__":
    cases = {"exp_0": (101, 0.1, (0, 0, 0), (0, 30, np.pi), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 0., 12.),
             "c3": (1001, 0.02, (0, 0, 0), (0, 30, np.pi), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 0., 12.),
             "exp_1": (101, 0.1, (0, 0, 0), (100, 0, 0), (-np.deg2rad(30), np.deg2rad(30)), (9., 14.), 1., 50., 12.),
             "bank": (201, 0.05, (0, 0, 0), (80, 30, 0), (-np.deg2rad(40), np.deg2rad(40)), (9., 15.), 1., 1., 12.)}
    for nm, (N, hh, p0, p1, pb, vb, kv, kb, vspv) in cases.items():
        # start: straight line between the end points, heading along it, mid inputs
        import os
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'drone-sim-python_b200'))
        from d2d_b200 import opty_utils as orc      # triangle() initial guess only
        xg, yg, psig, phig, vg = orc.triangle(p0[:2], p1[:2], 12., (N - 1) * hh, N, go_left=-1.)
        z0 = np.zeros((N, 5)); z0[:, 0], z0[:, 1], z0[:, 2] = xg, yg, np.unwrap(psig)
        z0[:, 3] = 0.0; z0[:, 4] = np.arcsin(np.clip((12. - (vb[0] + vb[1]) / 2) / ((vb[1] - vb[0]) / 2), -0.99, 0.99))
        X, it, cost, cm, dt = solve(N, hh, np.array(p0, float), np.array(p1, float), (0., 0.), pb, vb, kv, kb, vspv, z0.ravel(), verbose="-v" in sys.argv)
        print(f"{nm}: {it} iterations, {dt:.2f}s, cost {cost:.8e}, |c| {cm:.2e}")
