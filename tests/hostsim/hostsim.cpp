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

template <typename T, int DMAX>
void run_generic(const HostNet<T>& hn, const StageTable<T>& st, const NlpLayout& L, const EvalArgs<double>& ar) {
    SlotLayout sl = make_slot_layout(L.x, L.d, hn.view.sum_h, hn.view.hmax);
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
    if (kernel == 1) {
        if (what != 0 || compute_f64 || act != 0 || n_layers != 3 || x != 2 || u != 1) return -1;
        StageTable<float> st = make_stage_table<float>(rk4, dt);
        const int mode = out2 ? 2 : (out1 ? 1 : 0);
        if (widths[0] == 30 && widths[1] == 30) run_fast<2, 1, 30, 30, 2>(wflat, st, L, ar, mode);
        else if (widths[0] == 32 && widths[1] == 32) run_fast<2, 1, 32, 32, 2>(wflat, st, L, ar, mode);
        else if (widths[0] == 16 && widths[1] == 16) run_fast<2, 1, 16, 16, 1>(wflat, st, L, ar, mode);
        else return -1;
        return 0;
    }
    if (compute_f64) {
        HostNet<double> hn; build_net(hn, d, n_layers, widths, act, wflat);
        run_generic_d<double>(hn, make_stage_table<double>(rk4, dt), L, ar);
    } else {
        HostNet<float> hn; build_net(hn, d, n_layers, widths, act, wflat);
        run_generic_d<float>(hn, make_stage_table<float>(rk4, dt), L, ar);
    }
    return 0;
}
