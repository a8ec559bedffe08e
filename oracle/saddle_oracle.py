"""TEST INFRASTRUCTURE ONLY. Python restatement of the reference's SaddlePointSolver
(src/polysolve/linear/SaddlePointSolver.cpp:112-290), the caller that reaches the hot path through
Solver::create(asymmetric_solver_name) / Solver::create(symmetric_solver_name) (:168-169). Statement by statement:
factorize (:112-148) builds the diagonally scaled blocks As, Bs, Cs and Ss = Cs - Bs^T Bs; solve (:152-286) runs the
outer loop of inner solves + the small least-squares combination. `make_solver(role)` returns an object with the
polysolve::linear::Solver interface (analyze_pattern / factorize / solve); the tests pass the "CUDA" backend."""
import numpy as np
import scipy.sparse as sp


class SaddlePointSolver:
    def __init__(self, make_solver, max_iter=50, conv_tol=1e-8):
        self.make_solver, self.max_iter, self.conv_tol = make_solver, max_iter, conv_tol
        self.num_iterations, self.final_res_norm = 0, 0.0

    def analyze_pattern(self, A, precond_num):          # SaddlePointSolver.hpp:36
        self.precond_num = precond_num

    def factorize(self, Ain):                           # :112-148
        p = self.precond_num
        assert p > 0
        self.Ain = sp.csc_matrix(Ain)
        A, B, C = self.Ain[:p, :p], self.Ain[:p, p:], self.Ain[p:, p:]
        self.Wm = sp.diags(1.0 / np.sqrt(A.diagonal()), format="csc")      # :130-134
        self.Wc = sp.identity(C.shape[0], format="csc")                    # :136-137
        self.As = sp.csc_matrix(self.Wm @ A @ self.Wm)
        self.Bs = sp.csc_matrix(self.Wm @ B @ self.Wc)
        self.BsT = sp.csc_matrix(self.Bs.T)
        self.Cs = sp.csc_matrix(self.Wc @ C @ self.Wc)
        self.Ss = sp.csc_matrix(self.Cs - self.BsT @ self.Bs)              # :144

    def solve(self, rhs, result):                       # :152-286
        p = self.precond_num
        Rm, Rc = rhs[:p], rhs[p:]
        Rms, Rcs = self.Wm @ Rm, self.Wc @ Rc
        cur_m, cur_c = Rms.copy(), Rcs.copy()
        yu, yp, Rmu, Rmp, Rcu, Rcp = [], [], [], [], [], []
        asym, sym = self.make_solver("asymmetric"), self.make_solver("symmetric")
        sym.analyze_pattern(self.Ss, self.Ss.shape[0])
        sym.factorize(self.Ss)
        i = 0
        while i < self.max_iter:
            yu.append(np.zeros(Rm.size))
            yp.append(np.zeros(Rc.size))
            asym.analyze_pattern(self.As, self.As.shape[0])                 # :187-190
            asym.factorize(self.As)
            asym.solve(cur_m, yu[i])
            Rcst = cur_c - self.BsT @ yu[i]                                 # :194
            sym.solve(Rcst, yp[i])                                          # :199
            Rmst = cur_m - self.Bs @ yp[i]                                  # :203
            yu[i][:] = 0
            asym.solve(Rmst, yu[i])                                         # :208-209
            Rmu.append(self.As @ yu[i])
            Rmp.append(self.Bs @ yp[i])
            Rcu.append(self.BsT @ yu[i])
            Rcp.append(self.Cs @ yp[i])
            k = i + 1
            M = np.zeros((2 * k, 2 * k))
            b = np.zeros(2 * k)
            for a in range(k):                                              # :225-238
                for c in range(k):
                    M[a, c] = Rmu[a] @ Rmu[c] + Rcu[a] @ Rcu[c]
                    M[a, k + c] = Rmu[a] @ Rmp[c] + Rcu[a] @ Rcp[c]
                    M[k + a, c] = Rmp[a] @ Rmu[c] + Rcp[a] @ Rcu[c]
                    M[k + a, k + c] = Rmp[a] @ Rmp[c] + Rcp[a] @ Rcp[c]
                b[a] = Rms @ Rmu[a] + Rcs @ Rcu[a]
                b[k + a] = Rms @ Rmp[a] + Rcs @ Rcp[a]
            alpha = np.linalg.lstsq(M, b, rcond=None)[0]                    # A.ldlt().solve(b), :253
            au, ap = alpha[:k], alpha[k:]
            yuf = sum(au[j] * yu[j] for j in range(k))                      # compute_solution, :24-49
            ypf = sum(ap[j] * yp[j] for j in range(k))
            result[:p] = self.Wm @ yuf
            result[p:] = self.Wc @ ypf
            self.final_res_norm = float(np.linalg.norm(self.Ain @ result - rhs))
            if self.final_res_norm < self.conv_tol:                         # :267-270
                break
            cur_m, cur_c = Rms.copy(), Rcs.copy()
            for j in range(k):                                              # :277-283
                cur_m -= au[j] * Rmu[j] + ap[j] * Rmp[j]
                cur_c -= au[j] * Rcu[j] + ap[j] * Rcp[j]
            i += 1
        self.num_iterations = i
        return result
