// TEST-ONLY host emulation of the device kernel bodies (nempc_generic.cuh / nempc_fast.cuh).
//
// The build container has no GPU, so the `-m "not gpu"` tests compile the very same per-step functions the
// CUDA kernels call for the host (one "thread", barriers are no-ops) and check their arithmetic and output
// indexing against the oracle.  This file is NOT part of the product: pyneuralempc_b200 never loads it, libnempc.so
// does not contain it, and it proves nothing about races or launch geometry -- the `-m gpu` tests do that.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../pyneuralempc_b200/csrc/nempc_fast.cuh"
#include "../../pyneuralempc_b200/csrc/nempc_generic.cuh"
#include "../../pyneuralempc_b200/csrc/nempc_layout.h"
#include "../../pyneuralempc_b200/csrc/nempc_solver.cuh"

namespace {

template <typename T> struct HostNet {
    std::vector<std::vector<T>> W, WT, b;
    NetView<T> view{};
};

template <typename T>
void build_net(HostNet<T>& hn, int d, int n_layers, const int* widths, int act, const double* wflat) {
    std::vector<int> dims{d};
    for (int l = 0; l < n_layers; ++l) dims.push_back(widths[l]);
    hn.W.resize(n_layers); hn.WT.resize(n_layers); hn.b.resize(n_layers);
    const double* p = wflat;
    NetView<T>& v = hn.view;
    v.L = n_layers; v.act = act;
    int off = 0, hm = 0;
    for (int l = 0; l <= n_layers; ++l) v.dims[l] = dims[l];
    for (int l = 0; l < n_layers; ++l) {
        const int fin = dims[l], fout = dims[l + 1];
        hn.W[l].resize((size_t)fin * fout); hn.WT[l].resize((size_t)fin * fout); hn.b[l].resize(fout);
        for (int i = 0; i < fin; ++i)
            for (int j = 0; j < fout; ++j) { hn.W[l][(size_t)i * fout + j] = (T)p[(size_t)i * fout + j]; hn.WT[l][(size_t)j * fin + i] = (T)p[(size_t)i * fout + j]; }
        p += (size_t)fin * fout;
        for (int j = 0; j < fout; ++j) hn.b[l][j] = (T)p[j];
        p += fout;
        v.W[l] = hn.W[l].data(); v.WT[l] = hn.WT[l].data(); v.b[l] = hn.b[l].data();
        if (l + 1 < n_layers) { v.hoff[l] = off; off += fout; hm = fout > hm ? fout : hm; }
    }
    v.sum_h = off; v.hmax = hm;
}

// exogenous model inputs of the NEXT hostsim_run call (generic kernel only); cleared by that call
struct Exo { int tvp_dim = 0, p_dim = 0; const double* tvp = nullptr; const double* p = nullptr; long long tvp_bstride = 0, p_bstride = 0; };
Exo g_exo;

template <typename T, int DMAX>
void run_generic(const HostNet<T>& hn, const StageTable<T>& st, const NlpLayout& L, const EvalArgs<double>& ar) {
    SlotLayout sl = make_slot_layout(L.x, L.d, hn.view.sum_h, hn.view.hmax, hn.view.tvp_dim + hn.view.p_dim);
    std::vector<T> ws(sl.total);
    for (long long s = 0; s < ar.nsteps; ++s) generic_step<T, double, DMAX>(hn.view, st, L, sl, ar, s, ws.data(), 0, 1, 0);
}

template <typename T>
void run_generic_d(const HostNet<T>& hn, const StageTable<T>& st, const NlpLayout& L, const EvalArgs<double>& ar) {
    if (L.d <= 4) run_generic<T, 4>(hn, st, L, ar);
    else if (L.d <= 8) run_generic<T, 8>(hn, st, L, ar);
    else run_generic<T, 16>(hn, st, L, ar);
}

template <int X, int U, int H1, int H2, int NCHUNK>
void run_fast(const double* wflat, const StageTable<float>& st, const NlpLayout& L, const EvalArgs<double>& ar, int mode) {
    typedef FastWeights<X, U, H1, H2, NCHUNK> FW;
    FW* fp = new FW();
    FW& f = *fp;
    constexpr int D = X + U;
    const double* W1 = wflat; const double* b1 = W1 + D * H1; const double* W2 = b1 + H1; const double* b2 = W2 + H1 * H2;
    const double* W3 = b2 + H2; const double* b3 = W3 + H2 * X;
    fill_fast_weights<X, U, H1, H2, NCHUNK>(f, W1, b1, W2, b2, W3, b3);
    std::vector<float> scr(FastScratch<X, U, H1, H2>::COUNT);
    for (long long s = 0; s < ar.nsteps; ++s) {
        if (mode == 0) fast_step<X, U, H1, H2, NCHUNK, 0, double>(f, st, L, ar, s, scr.data(), 1);
        else if (mode == 1) fast_step<X, U, H1, H2, NCHUNK, 1, double>(f, st, L, ar, s, scr.data(), 1);
        else fast_step<X, U, H1, H2, NCHUNK, 2, double>(f, st, L, ar, s, scr.data(), 1);
    }
    delete fp;
}

}  // namespace

extern "C" void hostsim_set_exo(int tvp_dim, int p_dim, const double* tvp, long long tvp_bstride, const double* p, long long p_bstride) {
    g_exo.tvp_dim = tvp_dim; g_exo.p_dim = p_dim; g_exo.tvp = tvp; g_exo.p = p; g_exo.tvp_bstride = tvp_bstride; g_exo.p_bstride = p_bstride;
}

// kernel: 0 generic, 1 fast.  what: 0 eval (resid/jac/hes), 1 blocks (pred/AB/Hblk), 2 model (zin -> f/jac/hes).
// Returns 0, or -1 when the combination is unsupported.  All arrays are double (io_dtype f64).
extern "C" int hostsim_run(int x, int u, int H, int n_layers, const int* widths, int act, int integ, double dt,
                           int compute_f64, int kernel, int what, const double* wflat, const double* quad,
                           long long B, const double* z, const double* x0, const double* lam, const double* sigma,
                           double sigma_scalar, double* out0, double* out1, double* out2) {
    const int d = x + u;
    std::vector<uint8_t> mask;
    if (quad) { mask.resize((size_t)H * d); for (size_t i = 0; i < mask.size(); ++i) mask[i] = quad[i] != 0.0; }
    NlpLayout L; nlp_layout_init(L, H, x, u, quad ? mask.data() : nullptr);
    EvalArgs<double> ar{};
    ar.z = z; ar.x0 = x0; ar.lam = lam; ar.sigma = sigma; ar.sigma_scalar = sigma_scalar; ar.quad = quad;
    const bool model = what == 2;
    ar.nsteps = model ? B : B * (long long)H;
    const int unity = integ == 1 ? NEMPC_UNITY : 0;
    if (what == 0) {
        ar.resid = out0; ar.jac = out1; ar.hes = out2;
        ar.flags = (out1 || out2 ? NEMPC_WANT_JAC : 0) | (out2 ? NEMPC_WANT_HES : 0) | unity;
    } else {
        ar.pred = out0; ar.AB = out1; ar.Hblk = out2;
        ar.flags = (model ? NEMPC_MODE_MODEL : NEMPC_MODE_BLOCKS) | (out1 || out2 ? NEMPC_WANT_JAC : 0) | (out2 ? NEMPC_WANT_HES : 0) | unity;
    }
    const bool rk4 = integ == 2 && !model;
    const Exo exo = g_exo;
    g_exo = Exo();
    ar.tvp = exo.tvp; ar.p = exo.p; ar.tvp_bstride = exo.tvp_bstride; ar.p_bstride = exo.p_bstride;
    const int n_ext = exo.tvp_dim + exo.p_dim;
    if (kernel == 1) {
        if (n_ext) return -1;
        if (what != 0 || compute_f64 || act != 0 || n_layers != 3 || x != 2 || u != 1) return -1;
        StageTable<float> st = make_stage_table<float>(rk4, dt);
        const int mode = out2 ? 2 : (out1 ? 1 : 0);
        if (widths[0] == 30 && widths[1] == 30) run_fast<2, 1, 30, 30, 3>(wflat, st, L, ar, mode);       // three register chunks of ten neurons: the production instantiation
        else if (widths[0] == 32 && widths[1] == 32) run_fast<2, 1, 32, 32, 2>(wflat, st, L, ar, mode);
        else if (widths[0] == 16 && widths[1] == 16) run_fast<2, 1, 16, 16, 1>(wflat, st, L, ar, mode);
        else return -1;
        return 0;
    }
    if (compute_f64) {
        HostNet<double> hn; build_net(hn, d + n_ext, n_layers, widths, act, wflat);
        hn.view.tvp_dim = exo.tvp_dim; hn.view.p_dim = exo.p_dim;
        run_generic_d<double>(hn, make_stage_table<double>(rk4, dt), L, ar);
    } else {
        HostNet<float> hn; build_net(hn, d + n_ext, n_layers, widths, act, wflat);
        hn.view.tvp_dim = exo.tvp_dim; hn.view.p_dim = exo.p_dim;
        run_generic_d<float>(hn, make_stage_table<float>(rk4, dt), L, ar);
    }
    return 0;
}

// Host emulation of nempc_solve: same per-problem bodies (nempc_solver.cuh), evaluations by the generic f64 kernel body.
// lb/ub: n doubles; z (B,n) in/out; lam (B,m) out; status/iters (B) out; kkt (B) out.  Returns outer iterations.
extern "C" int hostsim_solve(int x, int u, int H, int n_layers, const int* widths, int act, int integ, double dt,
                             const double* wflat, const double* lin, const double* quad, const double* ref, long long B,
                             const double* x0, const double* lb, const double* ub, double* z, int use_init, double* lam_out,
                             int* status, int* iters, double* kkt, int max_iter, double tol) {
    const int d = x + u;
    const size_t n = (size_t)H * d, m = (size_t)H * x;
    std::vector<uint8_t> mask(n);
    for (size_t i = 0; i < n; ++i) mask[i] = quad[i] != 0.0;
    NlpLayout L; nlp_layout_init(L, H, x, u, mask.data());
    HostNet<double> hn; build_net(hn, d, n_layers, widths, act, wflat);
    StageTable<double> st = make_stage_table<double>(integ == 2, dt);
    SolverOpts o = solver_defaults();
    o.max_iter = max_iter; o.tol = tol;
    const size_t nj = (size_t)L.nnz_jac, nh = (size_t)L.nnz_hes;
    auto vec = [&](size_t per) { return std::vector<double>(per * (size_t)B, 0.0); };
    std::vector<double> lam = vec(m), zL = vec(n), zU = vec(n), dz = vec(n), lamn = vec(m), dzL = vec(n), dzU = vec(n), grad = vec(n),
                        resid = vec(m), jac = vec(nj), hes = vec(nh), obj = vec(1), zt = vec(n), residt = vec(m), objt = vec(1),
                        K = vec((size_t)H * u * x), kf = vec((size_t)H * u), mu = vec(1), nu = vec(1), alpha = vec(1), alphaD = vec(1),
                        phi0 = vec(1), dphi = vec(1), err = vec(1);
    std::vector<int> stv(B), itv(B), acc(B);
    SolverWs w{};
    w.x0 = x0; w.lb = lb; w.ub = ub; w.z = z; w.lam = lam.data(); w.zL = zL.data(); w.zU = zU.data(); w.dz = dz.data(); w.lamn = lamn.data();
    w.dzL = dzL.data(); w.dzU = dzU.data(); w.grad = grad.data(); w.resid = resid.data(); w.jac = jac.data(); w.hes = hes.data(); w.obj = obj.data();
    w.zt = zt.data(); w.residt = residt.data(); w.objt = objt.data(); w.K = K.data(); w.kf = kf.data(); w.mu = mu.data(); w.nu = nu.data();
    w.alpha = alpha.data(); w.alphaD = alphaD.data(); w.phi0 = phi0.data(); w.dphi = dphi.data(); w.err = err.data();
    w.status = stv.data(); w.iters = itv.data(); w.accepted = acc.data();
    auto evaluate = [&](const double* zz, const double* lm, double* r, double* jv, double* hv, double* ob, double* gr) {
        EvalArgs<double> ar{};
        ar.z = zz; ar.x0 = x0; ar.lam = lm; ar.sigma = nullptr; ar.sigma_scalar = 1.0; ar.quad = quad;
        ar.resid = r; ar.jac = jv; ar.hes = hv; ar.nsteps = B * (long long)H;
        ar.flags = (jv || hv ? NEMPC_WANT_JAC : 0) | (hv ? NEMPC_WANT_HES : 0) | (integ == 1 ? NEMPC_UNITY : 0);
        run_generic_d<double>(hn, st, L, ar);
        for (long long b = 0; b < B; ++b) {
            double a = 0.0;
            for (size_t i = 0; i < n; ++i) {
                const double zi = zz[b * n + i], dd = zi - ref[i];
                a += lin[i] * zi + quad[i] * dd * dd;
                if (gr) gr[b * n + i] = lin[i] + 2.0 * quad[i] * dd;
            }
            ob[b] = a;
        }
    };
    for (long long b = 0; b < B; ++b) ipm_init_problem(L, w, b, o, use_init != 0);
    int it = 0;
    for (; it < o.max_iter; ++it) {
        evaluate(z, lam.data(), resid.data(), jac.data(), hes.data(), obj.data(), grad.data());
        int running = 0, pending = 0;
        for (long long b = 0; b < B; ++b) {
            if (x <= 4 && u <= 2) ipm_kkt_problem<4, 2>(L, w, b, o); else ipm_kkt_problem<NEMPC_SOLVER_XM, NEMPC_SOLVER_UM>(L, w, b, o);
            if (stv[b] == NEMPC_ST_RUNNING) { ++running; if (!acc[b]) ++pending; }
        }
        if (running == 0) break;
        for (int t = 0; t < o.max_backtrack && pending > 0; ++t) {
            evaluate(zt.data(), nullptr, residt.data(), nullptr, nullptr, objt.data(), nullptr);
            pending = 0;
            for (long long b = 0; b < B; ++b) { ipm_linesearch_problem(L, w, b, o); if (stv[b] == NEMPC_ST_RUNNING && !acc[b]) ++pending; }
        }
        for (long long b = 0; b < B; ++b) ipm_update_problem(L, w, b, o);
    }
    for (long long b = 0; b < B; ++b) {
        status[b] = stv[b] == NEMPC_ST_RUNNING ? NEMPC_ST_MAXITER : stv[b];
        iters[b] = itv[b]; kkt[b] = err[b];
    }
    if (lam_out) memcpy(lam_out, lam.data(), sizeof(double) * m * (size_t)B);
    return it;
}
