// Device-resident Neo-Hookean Problem for BASELINE config 5 (include/psb200_problems.h): the user side of
// polysolve::nonlinear::Problem (reference src/polysolve/nonlinear/Problem.hpp:49-67) evaluated on the GPU, with the
// Hessian assembled by a CUDA kernel straight into the compressed-column pattern the linear solver analysed.
//
//   energy   : one thread per tetrahedron, fixed-order two-stage sum
//   gradient : one thread per node, gathers over the node's incident (tet, local vertex) list
//   Hessian  : one thread per 3 x 3 block (node pair), gathers over the tets that contain the pair:
//              K_ab = vol [ mu (gN_a . gN_b) I + (mu - lambda ln J) p_b p_a^T + lambda p_a p_b^T ],  p_a = F^-T gN_a
// No floating-point atomics anywhere: results are bit-reproducible, which the multi-GPU Newton driver relies on (every
// rank evaluates the same x and must take the same line-search branch).
#include "../../include/psb200_problems.h"
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <limits>
#include <string>
#include <vector>

struct psb200_nh
{
    int device = 0;
    long long nn = 0, nt = 0, nb = 0; // nodes, tets, blocks
    double mu = 1, lam = 1;
    std::string err;
    cudaStream_t st = nullptr;
    std::vector<int32_t> outer, inner; // scalar CSC pattern (host)
    // device
    psb::DevBuf<int> tets, n2t_ptr, n2t, blk_ptr, blk_row, inc_ptr, inc, d_outer;
    psb::DevBuf<double> X, G, vol, x, g, vals, partial;
    psb::DevBuf<unsigned char> fixed;
};

namespace {

thread_local std::string g_nh_create_error;

inline int nblk(long long n, int t = 256) { return (int)std::max<long long>(1, (n + t - 1) / t); }

struct TetState
{
    double F[9], Finv[9], J;
};
// F = I + sum_a u_a gN_a^T
__device__ __forceinline__ void tet_F(const int *__restrict__ tets, const double *__restrict__ G, const double *__restrict__ x, long long t, TetState &s)
{
    double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
        const int v = tets[4 * t + a];
        const double u0 = x[3 * (long long)v], u1 = x[3 * (long long)v + 1], u2 = x[3 * (long long)v + 2];
        const double g0 = G[12 * t + 3 * a], g1 = G[12 * t + 3 * a + 1], g2 = G[12 * t + 3 * a + 2];
        F[0] += u0 * g0; F[1] += u0 * g1; F[2] += u0 * g2;
        F[3] += u1 * g0; F[4] += u1 * g1; F[5] += u1 * g2;
        F[6] += u2 * g0; F[7] += u2 * g1; F[8] += u2 * g2;
    }
    const double c00 = F[4] * F[8] - F[5] * F[7], c01 = F[5] * F[6] - F[3] * F[8], c02 = F[3] * F[7] - F[4] * F[6];
    const double J = F[0] * c00 + F[1] * c01 + F[2] * c02;
    const double id = 1.0 / J;
#pragma unroll
    for (int e = 0; e < 9; ++e)
        s.F[e] = F[e];
    s.J = J;
    s.Finv[0] = c00 * id;
    s.Finv[1] = (F[2] * F[7] - F[1] * F[8]) * id;
    s.Finv[2] = (F[1] * F[5] - F[2] * F[4]) * id;
    s.Finv[3] = c01 * id;
    s.Finv[4] = (F[0] * F[8] - F[2] * F[6]) * id;
    s.Finv[5] = (F[2] * F[3] - F[0] * F[5]) * id;
    s.Finv[6] = c02 * id;
    s.Finv[7] = (F[1] * F[6] - F[0] * F[7]) * id;
    s.Finv[8] = (F[0] * F[4] - F[1] * F[3]) * id;
}

__global__ void nh_energy_kernel(long long nt, const int *__restrict__ tets, const double *__restrict__ G, const double *__restrict__ vol,
                                 const double *__restrict__ x, double mu, double lam, double *__restrict__ w)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt)
        return;
    TetState s;
    tet_F(tets, G, x, t, s);
    double e;
    if (!(s.J > 0.0))
        e = INFINITY;
    else
    {
        double f2 = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k)
            f2 += s.F[k] * s.F[k];
        const double lj = log(s.J);
        e = vol[t] * (0.5 * mu * (f2 - 3.0) - mu * lj + 0.5 * lam * lj * lj);
    }
    w[t] = e;
}
// fixed-order sum: CTA b adds the chunk [b * chunk, (b + 1) * chunk) thread-strided, tree in shared memory
__global__ void nh_sum1_kernel(long long n, const double *__restrict__ w, long long chunk, double *__restrict__ partial)
{
    __shared__ double sm[256];
    const long long b0 = (long long)blockIdx.x * chunk, b1 = min(n, b0 + chunk);
    double s = 0;
    for (long long i = b0 + threadIdx.x; i < b1; i += 256)
        s += w[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1)
    {
        if ((int)threadIdx.x < o)
            sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0)
        partial[blockIdx.x] = sm[0];
}
__global__ void nh_sum2_kernel(int nparts, const double *__restrict__ partial, double *out)
{
    if (threadIdx.x || blockIdx.x)
        return;
    double s = 0;
    for (int i = 0; i < nparts; ++i)
        s += partial[i];
    *out = s;
}

__global__ void nh_gradient_kernel(long long nn, const int *__restrict__ tets, const double *__restrict__ G, const double *__restrict__ vol,
                                   const int *__restrict__ n2t_ptr, const int *__restrict__ n2t, const double *__restrict__ x,
                                   const unsigned char *__restrict__ fixed, double mu, double lam, double *__restrict__ g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nn)
        return;
    double g0 = 0, g1 = 0, g2 = 0;
    for (int e = n2t_ptr[v]; e < n2t_ptr[v + 1]; ++e)
    {
        const long long t = n2t[e] >> 2;
        const int a = n2t[e] & 3;
        TetState s;
        tet_F(tets, G, x, t, s);
        const double lj = log(s.J), cf = lam * lj - mu; // P = mu F + (lambda ln J - mu) F^-T
        const double n0 = G[12 * t + 3 * a], n1 = G[12 * t + 3 * a + 1], n2 = G[12 * t + 3 * a + 2];
        // (F^-T gN)_r = sum_l Finv[l][r] gN_l
        const double p0 = s.Finv[0] * n0 + s.Finv[3] * n1 + s.Finv[6] * n2;
        const double p1 = s.Finv[1] * n0 + s.Finv[4] * n1 + s.Finv[7] * n2;
        const double p2 = s.Finv[2] * n0 + s.Finv[5] * n1 + s.Finv[8] * n2;
        const double w = vol[t];
        g0 += w * (mu * (s.F[0] * n0 + s.F[1] * n1 + s.F[2] * n2) + cf * p0);
        g1 += w * (mu * (s.F[3] * n0 + s.F[4] * n1 + s.F[5] * n2) + cf * p1);
        g2 += w * (mu * (s.F[6] * n0 + s.F[7] * n1 + s.F[8] * n2) + cf * p2);
    }
    g[3 * v] = fixed[3 * v] ? 0.0 : g0;
    g[3 * v + 1] = fixed[3 * v + 1] ? 0.0 : g1;
    g[3 * v + 2] = fixed[3 * v + 2] ? 0.0 : g2;
}

// one thread per block (row node i, column node j): 9 values into the CSC slots of columns 3 j .. 3 j + 2
__global__ void nh_hessian_kernel(long long nn, const int *__restrict__ tets, const double *__restrict__ G, const double *__restrict__ vol,
                                  const int *__restrict__ blk_ptr, const int *__restrict__ blk_row, const int *__restrict__ inc_ptr,
                                  const int *__restrict__ inc, const int *__restrict__ outer, const double *__restrict__ x,
                                  const unsigned char *__restrict__ fixed, double mu, double lam, long long nb, double *__restrict__ vals)
{
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb)
        return;
    // column node j of block b: the last j with blk_ptr[j] <= b
    long long lo = 0, hi = nn;
    while (hi - lo > 1)
    {
        const long long mid = (lo + hi) >> 1;
        if (blk_ptr[mid] <= b)
            lo = mid;
        else
            hi = mid;
    }
    const long long j = lo;
    const long long i = blk_row[b];
    const int q = (int)(b - blk_ptr[j]);
    double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = inc_ptr[b]; e < inc_ptr[b + 1]; ++e)
    {
        const long long t = inc[e] >> 4;
        const int a = (inc[e] >> 2) & 3, bb = inc[e] & 3; // a: local index of the row node, bb: of the column node
        TetState s;
        tet_F(tets, G, x, t, s);
        const double lj = log(s.J);
        const double *na = G + 12 * t + 3 * a, *nbv = G + 12 * t + 3 * bb;
        double pa[3], pb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            pa[r] = s.Finv[r] * na[0] + s.Finv[3 + r] * na[1] + s.Finv[6 + r] * na[2];
            pb[r] = s.Finv[r] * nbv[0] + s.Finv[3 + r] * nbv[1] + s.Finv[6 + r] * nbv[2];
        }
        const double gg = na[0] * nbv[0] + na[1] * nbv[1] + na[2] * nbv[2];
        const double w = vol[t], c1 = mu - lam * lj;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                K[3 * r + c] += w * ((r == c ? mu * gg : 0.0) + c1 * pb[r] * pa[c] + lam * pa[r] * pb[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
            double v = K[3 * r + c];
            const bool fr = fixed[3 * i + r], fc = fixed[3 * j + c];
            if (fr || fc)
                v = (i == j && r == c) ? 1.0 : 0.0;
            vals[outer[3 * j + c] + 3 * q + r] = v;
        }
}

void evaluate_energy(psb200_nh &h, double *out)
{
    const int parts = 1024;
    const long long chunk = (h.nt + parts - 1) / parts;
    h.partial.alloc(h.nt + parts + 8);
    double *w = h.partial.p, *partial = h.partial.p + h.nt, *res = h.partial.p + h.nt + parts;
    nh_energy_kernel<<<nblk(h.nt), 256, 0, h.st>>>(h.nt, h.tets.p, h.G.p, h.vol.p, h.x.p, h.mu, h.lam, w);
    nh_sum1_kernel<<<parts, 256, 0, h.st>>>(h.nt, w, chunk, partial);
    nh_sum2_kernel<<<1, 32, 0, h.st>>>(parts, partial, res);
    PSB_CUDA(cudaGetLastError());
    PSB_CUDA(cudaMemcpyAsync(out, res, sizeof(double), cudaMemcpyDeviceToHost, h.st));
    PSB_CUDA(cudaStreamSynchronize(h.st));
}

template <class F>
int nh_guarded(psb200_nh_handle h, F &&f)
{
    if (!h)
        return 1;
    try
    {
        h->err.clear();
        psb::DeviceScope ds(h->device, true);
        psb::AllocScope as(h->st);
        f(*h);
        return 0;
    }
    catch (const std::exception &e)
    {
        h->err = e.what();
        return 2;
    }
}

void upload_x(psb200_nh &h, const double *x)
{
    PSB_CUDA(cudaMemcpyAsync(h.x.p, x, sizeof(double) * 3 * (size_t)h.nn, cudaMemcpyHostToDevice, h.st));
}

} // namespace

extern "C" {

const char *psb200_nh_last_error(psb200_nh_handle h) { return h ? h->err.c_str() : g_nh_create_error.c_str(); }

int psb200_nh_create(psb200_nh_handle *out, int64_t n_nodes, const double *X, int64_t n_tets, const int32_t *tets, double mu, double lambda,
                     const uint8_t *fixed, int device)
{
    if (!out || !X || !tets || n_nodes <= 0 || n_tets <= 0 || 3 * n_nodes > 0x7fffffffLL / 32)
        return 1;
    *out = nullptr;
    auto *h = new psb200_nh();
    try
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw std::runtime_error("psb200_nh: no CUDA device available; the device-resident Problem has no CPU fallback");
        if (device >= 0)
            PSB_CUDA(cudaSetDevice(device));
        PSB_CUDA(cudaGetDevice(&h->device));
        PSB_CUDA(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
        psb::AllocScope as(h->st);
        h->nn = n_nodes;
        h->nt = n_tets;
        h->mu = mu;
        h->lam = lambda;
        const long long nn = n_nodes, nt = n_tets;
        // per-tet shape gradients and volume
        std::vector<double> G((size_t)12 * nt), vol((size_t)nt);
        for (long long t = 0; t < nt; ++t)
        {
            const int32_t *v = tets + 4 * t;
            for (int a = 0; a < 4; ++a)
                if (v[a] < 0 || v[a] >= nn)
                    throw std::invalid_argument("psb200_nh: tetrahedron vertex out of range");
            double D[9]; // columns = edge vectors
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r)
                    D[3 * r + c] = X[3 * (size_t)v[c + 1] + r] - X[3 * (size_t)v[0] + r];
            const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
            const double det = D[0] * c00 + D[1] * c01 + D[2] * c02;
            if (!(det > 0))
                throw std::invalid_argument("psb200_nh: tetrahedron " + std::to_string(t) + " is degenerate or negatively oriented");
            const double id = 1.0 / det;
            const double inv[9] = {c00 * id, (D[2] * D[7] - D[1] * D[8]) * id, (D[1] * D[5] - D[2] * D[4]) * id,
                                   c01 * id, (D[0] * D[8] - D[2] * D[6]) * id, (D[2] * D[3] - D[0] * D[5]) * id,
                                   c02 * id, (D[1] * D[6] - D[0] * D[7]) * id, (D[0] * D[4] - D[1] * D[3]) * id};
            double *g = G.data() + 12 * t;
            for (int k = 0; k < 3; ++k)
            {
                g[3 + k] = inv[k];      // gN_1 = row 0 of D^-1
                g[6 + k] = inv[3 + k];
                g[9 + k] = inv[6 + k];
                g[k] = -(inv[k] + inv[3 + k] + inv[6 + k]);
            }
            vol[t] = det / 6.0;
        }
        // node -> incident (tet, local vertex)
        std::vector<int> n2t_ptr((size_t)nn + 1, 0), n2t((size_t)4 * nt);
        for (long long t = 0; t < nt; ++t)
            for (int a = 0; a < 4; ++a)
                n2t_ptr[tets[4 * t + a] + 1]++;
        for (long long v = 0; v < nn; ++v)
            n2t_ptr[v + 1] += n2t_ptr[v];
        {
            std::vector<int> cur(n2t_ptr.begin(), n2t_ptr.end() - 1);
            for (long long t = 0; t < nt; ++t)
                for (int a = 0; a < 4; ++a)
                    n2t[cur[tets[4 * t + a]]++] = (int)(t * 4 + a);
        }
        // blocks (column node j, row node i) with their incident (tet, a, b): sort the 16 nt pairs by (j, i)
        if (nt > (0x7fffffffLL >> 4))
            throw std::invalid_argument("psb200_nh: too many tetrahedra for the int32 incidence code");
        std::vector<std::pair<unsigned long long, int>> pairs;
        pairs.reserve((size_t)16 * nt);
        for (long long t = 0; t < nt; ++t)
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b)
                    pairs.emplace_back((unsigned long long)tets[4 * t + b] * (unsigned long long)nn + (unsigned long long)tets[4 * t + a],
                                       (int)((t << 4) | (a << 2) | b));
        std::sort(pairs.begin(), pairs.end());
        std::vector<int> blk_ptr((size_t)nn + 1, 0), blk_row, inc_ptr, inc(pairs.size());
        blk_row.reserve(pairs.size() / 4);
        inc_ptr.reserve(pairs.size() / 4 + 1);
        for (size_t e = 0; e < pairs.size(); ++e)
        {
            if (e == 0 || pairs[e].first != pairs[e - 1].first)
            {
                const long long j = (long long)(pairs[e].first / (unsigned long long)nn), i = (long long)(pairs[e].first % (unsigned long long)nn);
                blk_row.push_back((int)i);
                inc_ptr.push_back((int)e);
                blk_ptr[j + 1]++;
            }
            inc[e] = pairs[e].second;
        }
        inc_ptr.push_back((int)pairs.size());
        for (long long j = 0; j < nn; ++j)
            blk_ptr[j + 1] += blk_ptr[j];
        h->nb = (long long)blk_row.size();
        if (9 * h->nb > 0x7fffffffLL - 1024)
            throw std::invalid_argument("psb200_nh: Hessian exceeds the int32 index range");
        // scalar CSC pattern
        h->outer.assign((size_t)3 * nn + 1, 0);
        h->inner.resize((size_t)9 * h->nb);
        for (long long j = 0; j < nn; ++j)
        {
            const int len = blk_ptr[j + 1] - blk_ptr[j];
            for (int c = 0; c < 3; ++c)
            {
                const int o = 9 * blk_ptr[j] + c * 3 * len;
                h->outer[3 * j + c] = o;
                for (int q = 0; q < len; ++q)
                    for (int r = 0; r < 3; ++r)
                        h->inner[(size_t)o + 3 * q + r] = 3 * blk_row[blk_ptr[j] + q] + r;
            }
        }
        h->outer[3 * nn] = (int32_t)(9 * h->nb);
        std::vector<unsigned char> fx((size_t)3 * nn, 0);
        if (fixed)
            for (size_t k = 0; k < fx.size(); ++k)
                fx[k] = fixed[k] ? 1 : 0;
        auto up_i = [&](psb::DevBuf<int> &d, const std::vector<int> &v) {
            d.alloc(std::max<size_t>(1, v.size()));
            if (!v.empty())
                PSB_CUDA(cudaMemcpyAsync(d.p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, h->st));
        };
        auto up_d = [&](psb::DevBuf<double> &d, const double *v, size_t n) {
            d.alloc(std::max<size_t>(1, n));
            PSB_CUDA(cudaMemcpyAsync(d.p, v, sizeof(double) * n, cudaMemcpyHostToDevice, h->st));
        };
        std::vector<int> tv(tets, tets + 4 * nt), ov(h->outer.begin(), h->outer.end());
        up_i(h->tets, tv);
        up_i(h->n2t_ptr, n2t_ptr);
        up_i(h->n2t, n2t);
        up_i(h->blk_ptr, blk_ptr);
        up_i(h->blk_row, blk_row);
        up_i(h->inc_ptr, inc_ptr);
        up_i(h->inc, inc);
        up_i(h->d_outer, ov);
        up_d(h->X, X, (size_t)3 * nn);
        up_d(h->G, G.data(), G.size());
        up_d(h->vol, vol.data(), vol.size());
        h->fixed.alloc((size_t)3 * nn);
        PSB_CUDA(cudaMemcpyAsync(h->fixed.p, fx.data(), fx.size(), cudaMemcpyHostToDevice, h->st));
        h->x.alloc((size_t)3 * nn, true);
        h->g.alloc((size_t)3 * nn, true);
        h->vals.alloc((size_t)9 * h->nb, true);
        PSB_CUDA(cudaStreamSynchronize(h->st)); // the staging vectors above are function-local
    }
    catch (const std::exception &e)
    {
        g_nh_create_error = e.what();
        delete h;
        return 2;
    }
    *out = h;
    return 0;
}

int psb200_nh_destroy(psb200_nh_handle h)
{
    if (!h)
        return 1;
    {
        psb::DeviceScope ds(h->device, true);
        if (h->st)
            cudaStreamSynchronize(h->st);
        for (auto *b : {&h->tets, &h->n2t_ptr, &h->n2t, &h->blk_ptr, &h->blk_row, &h->inc_ptr, &h->inc, &h->d_outer})
            b->release();
        for (auto *b : {&h->X, &h->G, &h->vol, &h->x, &h->g, &h->vals, &h->partial})
            b->release();
        h->fixed.release();
        if (h->st)
        {
            cudaStreamSynchronize(h->st);
            cudaStreamDestroy(h->st);
        }
    }
    delete h;
    return 0;
}

int psb200_nh_pattern(psb200_nh_handle h, int64_t *n, int64_t *nnz, const int32_t **outer, const int32_t **inner)
{
    if (!h)
        return 1;
    if (n)
        *n = 3 * h->nn;
    if (nnz)
        *nnz = 9 * h->nb;
    if (outer)
        *outer = h->outer.data();
    if (inner)
        *inner = h->inner.data();
    return 0;
}

int psb200_nh_value(psb200_nh_handle h, const double *x, double *value_out)
{
    return nh_guarded(h, [&](psb200_nh &s) {
        upload_x(s, x);
        evaluate_energy(s, value_out);
    });
}

int psb200_nh_gradient(psb200_nh_handle h, const double *x, double *grad_out)
{
    return nh_guarded(h, [&](psb200_nh &s) {
        upload_x(s, x);
        nh_gradient_kernel<<<nblk(s.nn, 128), 128, 0, s.st>>>(s.nn, s.tets.p, s.G.p, s.vol.p, s.n2t_ptr.p, s.n2t.p, s.x.p, s.fixed.p, s.mu, s.lam, s.g.p);
        PSB_CUDA(cudaGetLastError());
        PSB_CUDA(cudaMemcpyAsync(grad_out, s.g.p, sizeof(double) * 3 * (size_t)s.nn, cudaMemcpyDeviceToHost, s.st));
        PSB_CUDA(cudaStreamSynchronize(s.st));
    });
}

int psb200_nh_hessian_device(psb200_nh_handle h, const double *x, const double **d_vals_out)
{
    return nh_guarded(h, [&](psb200_nh &s) {
        upload_x(s, x);
        nh_hessian_kernel<<<nblk(s.nb, 128), 128, 0, s.st>>>(s.nn, s.tets.p, s.G.p, s.vol.p, s.blk_ptr.p, s.blk_row.p, s.inc_ptr.p, s.inc.p, s.d_outer.p,
                                                               s.x.p, s.fixed.p, s.mu, s.lam, s.nb, s.vals.p);
        PSB_CUDA(cudaGetLastError());
        PSB_CUDA(cudaStreamSynchronize(s.st)); // the consumer (the linear solver) works on its own stream
        if (d_vals_out)
            *d_vals_out = s.vals.p;
    });
}

int psb200_nh_hessian_host(psb200_nh_handle h, const double *x, double *vals_out)
{
    const double *d = nullptr;
    const int rc = psb200_nh_hessian_device(h, x, &d);
    if (rc)
        return rc;
    return nh_guarded(h, [&](psb200_nh &s) {
        PSB_CUDA(cudaMemcpyAsync(vals_out, d, sizeof(double) * 9 * (size_t)s.nb, cudaMemcpyDeviceToHost, s.st));
        PSB_CUDA(cudaStreamSynchronize(s.st));
    });
}

} // extern "C"
