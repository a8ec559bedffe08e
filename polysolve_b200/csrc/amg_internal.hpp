// Building blocks of the AMG setup shared by the single-GPU hierarchy (amg.cu) and the row-partitioned one (amg_dist.cu).
#pragma once
#include "amg.hpp"

#include <functional>
#include <vector>

namespace psb {

struct Temp
{
    DevBuf<unsigned char> buf;
    void *get(size_t bytes)
    {
        buf.alloc(bytes, false, 256);
        return buf.p;
    }
};

template <typename T>
T d2h(Ctx &c, const T *p)
{
    T v;
    PSB_CUDA(cudaMemcpyAsync(&v, p, sizeof(T), cudaMemcpyDeviceToHost, c.stream));
    PSB_CUDA(cudaStreamSynchronize(c.stream));
    return v;
}

struct AmgLevel
{
    CsrDev Aown;
    const CsrDev *A = nullptr;
    CsrDev P, R;
    int n = 0;
    long long n_pad = 0;
    DevBuf<double> dinv, w, f, u, ualt, t, cp;
    DevBuf<int> agg;
    int n_agg = 0;
    // block mode: Ahat = Dblk^-1 A is the smoother's operator (Asm), dinvb the inverted diagonal blocks, bh = Dblk^-1 rhs
    CsrDev Ahat;
    const CsrDev *Asm = nullptr;
    DevBuf<double> dinvb, bh;
    DevBuf<int> agg_node;
    double rho = 0, cheb_d = 0, cheb_c = 0, omega = 0;
    std::vector<double> alpha, beta;
    int mis_rounds = 0;
    // direct coarse solve (amg direct_coarse = true): Z = A^-1 dense, (z_np x z_np)
    DevBuf<double> Zinv;
    int z_np = 0;
    bool direct = false;
    double t_relax = 0, t_agg = 0, t_prolong = 0, t_transpose = 0, t_ap = 0, t_rap = 0; // setup phase wall-clock, ms
};

// What a row partition adds to the per-level setup: the global index of the first local row (start vector of the power
// iteration), the halo push that precedes every multiplication, and a max over the ranks for the Gershgorin bound. The
// dot products of the power iteration all-reduce inside their kernels (grid_reduce) unless Ctx::comm_local is set.
struct SetupHooks
{
    long long row0 = 0;
    std::function<void(const double *)> push;
    std::function<double(double)> allmax;
};

void exclusive_scan_int(Ctx &c, Temp &tmp, const int *in, int *out, long long n);
void exclusive_scan_ll(Ctx &c, Temp &tmp, const long long *in, long long *out, long long n);
// T = M^T (stable: the rows of T keep the row order of M)
void transpose(Ctx &c, Temp &tmp, const CsrDev &M, CsrDev &T);
// C = A * B, columns of A index the rows of B; rows of C sorted by column
void spgemm(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C);
// the two implementations behind spgemm(): shared-memory hashing (spgemm.cu; false = a row exceeds its largest bin) and
// expand - radix sort - compress (amg.cu; any matrix). PSB200_SPGEMM=sort in the environment forces the second (A/B runs).
bool spgemm_hash(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C);
void spgemm_sort(Ctx &c, Temp &tmp, const CsrDev &A, const CsrDev &B, int ncolsB, CsrDev &C);
bool &spgemm_use_hash();
// D^-1 (or the inverted diagonal blocks + Ahat), spectral radius, Chebyshev coefficients, work vectors of one level.
// seed_index: level number in the whole hierarchy (start vector of the power iteration)
void setup_relaxation(Ctx &c, const AmgParams &prm, AmgLevel &L, int seed_index, const SetupHooks *hooks = nullptr);
// Gershgorin bound of D^-1 A over the local rows (block mode: block norms, needs L.dinvb of setup_relaxation)
double gershgorin_rho(Ctx &c, const AmgParams &prm, const AmgLevel &L);
// MIS-2 (or imposed) aggregates of the square matrix Asq (the level matrix; on a row partition its diagonal block):
// fills L.agg / L.agg_node / L.n_agg / L.mis_rounds. imposed: per node, may be null.
void build_aggregates(Ctx &c, Temp &tmp, const AmgParams &prm, const CsrDev &Asq, double eps_strong, const std::vector<int> *imposed, AmgLevel &L);
// P = (I - omega D_f^-1 A_f) P_tent from the square matrix Asq and L.agg. Block mode smooths with Dblk^-1 Asq: pass the
// level's Ahat when Asq is the level matrix itself, nullptr to have it computed from Asq.
void build_prolongation(Ctx &c, Temp &tmp, const AmgParams &prm, const CsrDev &Asq, const CsrDev *Ahat, double eps_strong, double omega, AmgLevel &L);
// dense.cu: Z = A^-1 of an SPD level matrix by blocked Cholesky on fp64 tensor cores; x = Z f
void dense_inverse_build(Ctx &c, const CsrDev &A, DevBuf<double> &Z, int &np_out);
void dense_inverse_apply(Ctx &c, int n, int np, const double *Z, const double *f, double *x, const int *done);
double wall_ms(cudaStream_t st);
// one smoother application on a level (see amg.cu); after_step pushes the halo of every new iterate on a row partition
struct FusedPush; // push_epi.cuh: with it the Chebyshev steps push their boundary rows from the SpMV epilogue
void relax_level(Ctx &ctx, const AmgParams &prm, AmgLevel &L, bool fine, const double *rhs, double *&x, double *&x_alt, bool x_is_zero,
                 const int *done, const std::function<void(const double *)> *after_step, const FusedPush *fused = nullptr);
// D = M[local rows, local columns] of a row-partitioned matrix (columns >= M.nl dropped); src[k] = position in M (dist.cu)
void extract_diag_block(Ctx &ctx, const CsrDev &M, CsrDev &D, DevBuf<int> &src);

} // namespace psb
