"""Minimal python-control stand-in: lqr() exactly as python-control computes it
without slycot (scipy CARE, K = R^-1 B^T X)."""
import numpy as np, scipy.linalg

def lqr(A, B, Q, R):
    A, B, Q, R = (np.asarray(m, dtype=float) for m in (A, B, Q, R))
    X = scipy.linalg.solve_continuous_are(A, B, Q, R)
    K = np.linalg.solve(R, B.T @ X)
    E = np.linalg.eigvals(A - B @ K)
    return K, X, E
