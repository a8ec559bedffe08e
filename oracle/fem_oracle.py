"""CPU restatement (test infrastructure only) of the reference's Dirichlet pre-processing,
src/polysolve/linear/FEMSolver.cpp:97-156 (dirichlet_solve), :303-331 (prefactorize), :345-372 (prefactorized solve).
Follows the reference statement by statement: N = indicator vector, g = f - (1 - N) .* (A (N .* f)), the matrix
rebuilt from the triplets whose row and column are both free plus the triplets (k, k, N(k)) for every k
(setFromTriplets sums duplicates, so a free diagonal keeps a_kk + 0 and a Dirichlet diagonal becomes 1)."""
import numpy as np
import scipy.sparse as sp


def dirichlet_system(A, f, dirichlet_nodes):
    A = sp.csc_matrix(A)
    n = A.shape[0]
    N = np.zeros(n)
    N[np.asarray(dirichlet_nodes, dtype=np.int64)] = 1.0                       # :108-113
    g = f - (1.0 - N) * (A @ (N * f))                                          # :115
    coo = A.tocoo()
    keep = (N[coo.row] != 1) & (N[coo.col] != 1)                               # :139
    rows = np.concatenate([coo.row[keep], np.arange(n)])
    cols = np.concatenate([coo.col[keep], np.arange(n)])
    vals = np.concatenate([coo.data[keep], N])                                 # :146-149
    At = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()             # :151-152
    At.sum_duplicates()
    return At, g
