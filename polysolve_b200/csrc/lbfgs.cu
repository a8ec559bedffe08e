// L-BFGS two-loop recursion on device vectors. Reference: the LBFGS descent strategy
// (src/polysolve/nonlinear/descent_strategies/LBFGS.cpp:22-61) keeps its memory in LBFGSpp::BFGSMat (un-vendored,
// cmake/recipes/LBFGSpp.cmake) and calls add_correction(x - x_prev, g - g_prev) followed by apply_Hv(g, -1, direction).
// Restated here from the published algorithm of BFGSMat:
//   add_correction(s, y): S[loc] = s, Y[loc] = y, ys[loc] = s.y, theta = y.y / s.y, loc = ptr mod m
//   apply_Hv(v, a):  res = a v;  newest -> oldest: alpha_j = S_j.res / ys_j, res -= alpha_j Y_j;  res /= theta;
//                    oldest -> newest: beta = Y_j.res / ys_j, res += (alpha_j - beta) S_j
// Every step is ONE fused pass: the vector update of step k and the dot product step k+1 needs, with the scalar
// (alpha, beta) produced on the device by the last CTA of the reduction -- 2 ncorr + 1 launches, no host round trip.
// The correction itself (s, y, s.y, y.y and the new x_prev / g_prev) is one more fused pass.
#include "../../include/psb200.h"
#include "../../include/psb200_nl.h"
#include "launch.cuh"

#include <string>

namespace psb {

// device scalars: ys[0..m), alpha[m..2m), theta = [2m], beta = [2m+1]
struct LbfgsScal
{
    double *p;
    int m;
    __device__ __forceinline__ double &ys(int j) const { return p[j]; }
    __device__ __forceinline__ double &alpha(int j) const { return p[m + j]; }
    __device__ __forceinline__ double &theta() const { return p[2 * m]; }
    __device__ __forceinline__ double &beta() const { return p[2 * m + 1]; }
};

// s = x - x_prev, y = g - g_prev, x_prev = x, g_prev = g;  s.y and y.y
struct OpLbfgsCorrection
{
    static constexpr int NV = 2;
    double *s, *y, *xp, *gp;
    const double *x, *g;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[2])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 xv = ld2(x, i), gv = ld2(g, i), xo = ld2(xp, i), go = ld2(gp, i);
            double2 sv, yv;
            sv.x = xv.x - xo.x;
            sv.y = xv.y - xo.y;
            yv.x = gv.x - go.x;
            yv.y = gv.y - go.y;
            st2(s, i, sv);
            st2(y, i, yv);
            st2(xp, i, xv);
            st2(gp, i, gv);
            acc[0] += sv.x * yv.x + sv.y * yv.y;
            acc[1] += yv.x * yv.x + yv.y * yv.y;
        }
    }
};
struct FinLbfgsCorrection
{
    LbfgsScal sc;
    int loc;
    __device__ __forceinline__ void operator()(const double *t) const
    {
        sc.ys(loc) = t[0];
        sc.theta() = t[1] / t[0];
    }
};
// res = a v; dot = S_j . res
struct OpLbfgsInit
{
    static constexpr int NV = 1;
    double *res;
    const double *v, *sj;
    double a;
    __device__ __forceinline__ void prologue() {}
    template <int U>
    __device__ __forceinline__ void apply(long long j, long long stride, double (&acc)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = j + u * stride;
            const double2 vv = ld2(v, i), sv = ld2(sj, i);
            double2 r;
            r.x = a * vv.x;
            r.y = a * vv.y;
            st2(res, i, r);
            acc[0] += sv.x * r.x + sv.y * r.y;
        }
    }
};
// loop 1: res -= alpha_j Y_j (then res /= theta on the last step); dot = next . res
struct OpLbfgsLoop1
{
    static constexpr int NV = 1;
    double *res;
    const double *yj, *next;
    LbfgsScal sc;
    int j;
    bool last;
    double alpha, theta;
    __device__ __forceinline__ void prologue()
    {
        alpha = sc.alpha(j);
        theta = sc.theta();
    }
    template <int U>
    __device__ __forceinline__ void apply(long long jj, long long stride, double (&acc)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = jj + u * stride;
            double2 r = ld2(res, i);
            const double2 yv = ld2(yj, i), nv = ld2(next, i);
            r.x -= alpha * yv.x;
            r.y -= alpha * yv.y;
            if (last)
            {
                r.x /= theta;
                r.y /= theta;
            }
            st2(res, i, r);
            acc[0] += nv.x * r.x + nv.y * r.y;
        }
    }
};
// loop 2: res += (alpha_j - beta) S_j; dot = Y_next . res
struct OpLbfgsLoop2
{
    static constexpr int NV = 1;
    double *res;
    const double *sj, *next; // next may alias sj on the final step (its dot is discarded)
    LbfgsScal sc;
    int j;
    double coef;
    __device__ __forceinline__ void prologue() { coef = sc.alpha(j) - sc.beta(); }
    template <int U>
    __device__ __forceinline__ void apply(long long jj, long long stride, double (&acc)[1])
    {
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const long long i = jj + u * stride;
            double2 r = ld2(res, i);
            const double2 sv = ld2(sj, i), nv = ld2(next, i);
            r.x += coef * sv.x;
            r.y += coef * sv.y;
            st2(res, i, r);
            acc[0] += nv.x * r.x + nv.y * r.y;
        }
    }
};
struct FinLbfgsAlpha
{
    LbfgsScal sc;
    int j;
    __device__ __forceinline__ void operator()(const double *t) const { sc.alpha(j) = t[0] / sc.ys(j); }
};
struct FinLbfgsBeta
{
    LbfgsScal sc;
    int j;
    __device__ __forceinline__ void operator()(const double *t) const { sc.beta() = t[0] / sc.ys(j); }
};

struct Lbfgs
{
    std::string err;
    Ctx ctx;
    long long n = 0, n_pad = 0;
    int device = 0;
    int m = 6, ncorr = 0, ptr = 0;
    bool has_prev = false;
    DevBuf<double> S, Y, xp, gp, res, x, g, scal;

    void init(long long n_, int m_, int device)
    {
        if (n_ < 0 || m_ < 1)
            throw std::invalid_argument("psb200_lbfgs_create: n < 0 or history_size < 1 (reference LBFGS.cpp:17-18)");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            throw CudaError("psb200: no CUDA device available; the CUDA backend has no CPU fallback");
        if (device >= 0)
            PSB_CUDA(cudaSetDevice(device));
        PSB_CUDA(cudaGetDevice(&this->device));
        n = n_;
        m = m_;
        n_pad = (n + 3) & ~3ll;
        ctx.init();
        AllocScope scope(ctx.stream);
        const size_t np = (size_t)std::max<long long>(n_pad, 4);
        S.alloc(np * m, true);
        Y.alloc(np * m, true);
        for (DevBuf<double> *b : {&xp, &gp, &res, &x, &g})
            b->alloc(np, true);
        scal.alloc((size_t)2 * m + 2, true);
    }
    double *col(DevBuf<double> &M, int j) { return M.p + (size_t)j * std::max<long long>(n_pad, 4); }
    void reset() // LBFGS::reset (LBFGS.cpp:22-28)
    {
        ncorr = 0;
        ptr = 0;
        has_prev = false;
    }
    // LBFGS::compute_update_direction (LBFGS.cpp:30-61); d_x, d_g, d_dir: device vectors of length n
    void direction_device(const double *d_x, const double *d_g, double *d_dir)
    {
        cudaStream_t st = ctx.stream;
        if (n == 0)
            return;
        // padded copies with zero tails (the fused kernels run over double2 lanes)
        PSB_CUDA(cudaMemcpyAsync(x.p, d_x, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        PSB_CUDA(cudaMemcpyAsync(g.p, d_g, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        run();
        PSB_CUDA(cudaMemcpyAsync(d_dir, res.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        PSB_CUDA(cudaStreamSynchronize(st));
    }
    void direction_host(const double *h_x, const double *h_g, double *h_dir)
    {
        cudaStream_t st = ctx.stream;
        if (n == 0)
            return;
        PSB_CUDA(cudaMemcpyAsync(x.p, h_x, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        PSB_CUDA(cudaMemcpyAsync(g.p, h_g, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        run();
        PSB_CUDA(cudaMemcpyAsync(h_dir, res.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
        PSB_CUDA(cudaStreamSynchronize(st));
    }
    void run()
    {
        cudaStream_t st = ctx.stream;
        LbfgsScal sc{scal.p, m};
        if (!has_prev)
        {
            // first iteration (or after a reset): gradient descent, direction = -grad (LBFGS.cpp:36-40)
            PSB_CUDA(cudaMemcpyAsync(xp.p, x.p, sizeof(double) * n_pad, cudaMemcpyDeviceToDevice, st));
            PSB_CUDA(cudaMemcpyAsync(gp.p, g.p, sizeof(double) * n_pad, cudaMemcpyDeviceToDevice, st));
            launch_vec(ctx, "lbfgs", n_pad, OpLbfgsInit{res.p, g.p, g.p, -1.0}, FinNone{});
            has_prev = true;
            return;
        }
        // m_bfgs.add_correction(x - m_prev_x, grad - m_prev_grad); m_prev_x = x; m_prev_grad = grad   (LBFGS.cpp:48-58)
        const int loc = ptr % m;
        launch_vec(ctx, "lbfgs", n_pad, OpLbfgsCorrection{col(S, loc), col(Y, loc), xp.p, gp.p, x.p, g.p}, FinLbfgsCorrection{sc, loc});
        if (ncorr < m)
            ++ncorr;
        ptr = loc + 1;
        // m_bfgs.apply_Hv(grad, -1, direction)   (LBFGS.cpp:51)
        int j = (ptr % m + m - 1) % m; // newest
        launch_vec(ctx, "lbfgs", n_pad, OpLbfgsInit{res.p, g.p, col(S, j), -1.0}, FinLbfgsAlpha{sc, j});
        for (int i = 0; i < ncorr; ++i)
        {
            const bool last = i + 1 == ncorr;
            const int jn = (j + m - 1) % m;
            if (!last)
                launch_vec(ctx, "lbfgs", n_pad, OpLbfgsLoop1{res.p, col(Y, j), col(S, jn), sc, j, false, 0, 0}, FinLbfgsAlpha{sc, jn});
            else // scale by 1/theta and start loop 2 at the same (oldest) index: beta = Y_j . res / ys_j
                launch_vec(ctx, "lbfgs", n_pad, OpLbfgsLoop1{res.p, col(Y, j), col(Y, j), sc, j, true, 0, 0}, FinLbfgsBeta{sc, j});
            if (!last)
                j = jn;
        }
        for (int i = 0; i < ncorr; ++i)
        {
            const bool last = i + 1 == ncorr;
            const int jn = (j + 1) % m;
            if (!last)
                launch_vec(ctx, "lbfgs", n_pad, OpLbfgsLoop2{res.p, col(S, j), col(Y, jn), sc, j, 0}, FinLbfgsBeta{sc, jn});
            else
                launch_vec(ctx, "lbfgs", n_pad, OpLbfgsLoop2{res.p, col(S, j), col(S, j), sc, j, 0}, FinNone{});
            j = jn;
        }
    }
};

} // namespace psb

struct psb200_lbfgs
{
    psb::Lbfgs l;
};

namespace {
thread_local std::string g_lbfgs_create_error;
template <class F>
int lbfgs_guarded(psb200_lbfgs_handle h, F &&f)
{
    if (!h)
        return PSB200_ERR_INVALID;
    try
    {
        h->l.err.clear();
        psb::DeviceScope device_scope(h->l.device, h->l.ctx.stream != nullptr);
        psb::AllocScope scope(h->l.ctx.stream);
        f(h->l);
        return PSB200_OK;
    }
    catch (const psb::CudaError &e)
    {
        h->l.err = e.what();
        cudaGetLastError();
        return PSB200_ERR_CUDA;
    }
    catch (const std::invalid_argument &e)
    {
        h->l.err = e.what();
        return PSB200_ERR_INVALID;
    }
    catch (const std::exception &e)
    {
        h->l.err = e.what();
        return PSB200_ERR_NUMERIC;
    }
}
} // namespace

extern "C" {

int psb200_lbfgs_create(psb200_lbfgs_handle *out, int64_t n, int history_size, int device)
{
    if (!out)
        return PSB200_ERR_INVALID;
    *out = nullptr;
    psb200_lbfgs *h = new psb200_lbfgs();
    const int rc = lbfgs_guarded(h, [&](psb::Lbfgs &l) { l.init(n, history_size, device); });
    if (rc != PSB200_OK)
    {
        g_lbfgs_create_error = h->l.err;
        delete h;
        return rc;
    }
    *out = h;
    return PSB200_OK;
}

int psb200_lbfgs_destroy(psb200_lbfgs_handle h)
{
    if (h)
    {
        psb::DeviceScope device_scope(h->l.device, h->l.ctx.stream != nullptr);
        delete h;
    }
    return PSB200_OK;
}

int psb200_lbfgs_reset(psb200_lbfgs_handle h)
{
    return lbfgs_guarded(h, [&](psb::Lbfgs &l) { l.reset(); });
}

int psb200_lbfgs_direction(psb200_lbfgs_handle h, const double *x, const double *grad, double *direction, int64_t n)
{
    return lbfgs_guarded(h, [&](psb::Lbfgs &l) {
        if (n != l.n || (n > 0 && (!x || !grad || !direction)))
            throw std::invalid_argument("psb200_lbfgs_direction: size mismatch or null vector");
        l.direction_host(x, grad, direction);
    });
}

int psb200_lbfgs_direction_device(psb200_lbfgs_handle h, const double *d_x, const double *d_grad, double *d_direction, int64_t n)
{
    return lbfgs_guarded(h, [&](psb::Lbfgs &l) {
        if (n != l.n || (n > 0 && (!d_x || !d_grad || !d_direction)))
            throw std::invalid_argument("psb200_lbfgs_direction_device: size mismatch or null vector");
        l.direction_device(d_x, d_grad, d_direction);
    });
}

const char *psb200_lbfgs_last_error(psb200_lbfgs_handle h) { return h ? h->l.err.c_str() : g_lbfgs_create_error.c_str(); }

} // extern "C"
