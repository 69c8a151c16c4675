// Rolling-window (NARX) models: the network of step t sees the last `rolling_window` states and controls
// (reference model/tensorflow.py:112-340 KerasTFModelRollingInput, model/jax.py:93-259 DiffDiscretJaxModelRollingWindow; SURVEY 8f
// rank 2).  The Jacobian band and the Hessian blocks of the NLP widen from one step to `rolling_window` steps, everything else of the
// transcription stays (integrator/discret.py:13-81, unity.py:15-81 slice the model's dense arrays the same way for any band).
//
// Device pipeline per evaluation (host side: pyneuralempc_b200/rolling.py):
//   1. nempc_rolling_gather_kernel    window rows  zin[b, t, :] = [x_{t-w+1} .. x_t | u_{t-w+1} .. u_t]  from z, x0 and the history
//   2. nempc_model_eval               value / Jacobian / per-output Hessian of every row with respect to its dw = w (x + u) inputs
//                                     (the generic kernel's model mode)
//   3. nempc_rolling_assemble_kernel  residual, banded sparse Jacobian values and lambda-contracted sparse Lagrangian-Hessian values:
//                                     ONE THREAD PER OUTPUT SLOT gathers its contributions through index tables (a Hessian slot collects
//                                     up to w window blocks) -- no atomics, fixed summation order, so results are deterministic.
// Index tables are built once on the host from the same closed-form structure the Python mirror reports (bit-identical indices).
#pragma once
#include <cstdint>

// source code of a gathered scalar: >= 0 -> z[b, code];  < 0 -> aux[b, -1 - code]  (aux = [x0 | prev_x | prev_u], per problem)
template <typename TIO>
__global__ void nempc_rolling_gather_kernel(const TIO* __restrict__ z, const TIO* __restrict__ aux, const int32_t* __restrict__ gidx,
                                            TIO* __restrict__ zin, int n, int naux, int rows, long long B) {
    const long long total = B * (long long)rows;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / rows;
        const int code = gidx[i - b * rows];
        zin[i] = code >= 0 ? z[b * n + code] : aux[b * naux + (-1 - code)];
    }
}

struct RollingTables {
    const int32_t* resid_base;   // [m]   gather code of the "x_{t-1} +" term (discrete), INT32_MIN = none (unity)
    const int32_t* jac_src;      // [nnz_jac]  index into the problem's J[H][x][dw] block, -1 = constant entry
    const double* jac_add;       // [nnz_jac]  added constant (-1 on x_t, +1 on the newest window state for the discrete integrator)
    const int32_t* hes_ptr;      // [nnz_hes + 1]  CSR over the contributions of a Hessian slot
    const int32_t* hes_src;      // [nnz contributions]  t * dw * dw + a * dw + b  into the problem's Hs[H][x][dw][dw]
    const double* hes_obj;       // [nnz_hes]  objective Hessian entry of the slot (2 quad on the diagonal), scaled by obj_factor
};

template <typename TIO>
__global__ void nempc_rolling_assemble_kernel(const RollingTables tb, const TIO* __restrict__ z, const TIO* __restrict__ aux,
                                              const TIO* __restrict__ f, const TIO* __restrict__ J, const TIO* __restrict__ Hs,
                                              const TIO* __restrict__ lam, const TIO* __restrict__ sigma, double sigma_scalar,
                                              TIO* __restrict__ resid, TIO* __restrict__ jac, TIO* __restrict__ hes,
                                              int H, int x, int dw, int n, int naux, long long nnz_jac, long long nnz_hes, long long B) {
    const int m = H * x;
    const long long per = (resid ? m : 0) + (jac ? nnz_jac : 0) + (hes ? nnz_hes : 0);
    const long long total = B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / per;
        long long s = i - b * per;
        if (resid) {
            if (s < m) {
                const int code = tb.resid_base[s];
                double v = (double)f[b * m + s] - (double)z[b * n + s];
                if (code != INT32_MIN) v += (double)(code >= 0 ? z[b * n + code] : aux[b * naux + (-1 - code)]);
                resid[b * m + s] = (TIO)v;
                continue;
            }
            s -= m;
        }
        if (jac) {
            if (s < nnz_jac) {
                const int src = tb.jac_src[s];
                double v = tb.jac_add[s];
                if (src >= 0) v += (double)J[b * (long long)m * dw + src];
                jac[b * nnz_jac + s] = (TIO)v;
                continue;
            }
            s -= nnz_jac;
        }
        // Lagrangian Hessian slot: sum over the window blocks that hold it, each contracted with the step's multipliers
        const double sg = sigma ? (double)sigma[b] : sigma_scalar;
        double v = sg * tb.hes_obj[s];
        const long long hb = b * (long long)m * dw * dw;
        for (int e = tb.hes_ptr[s]; e < tb.hes_ptr[s + 1]; ++e) {
            const int code = tb.hes_src[e];
            const int t = code / (dw * dw), ab = code - t * dw * dw;
            for (int p = 0; p < x; ++p)
                v += (double)lam[b * m + t * x + p] * (double)Hs[hb + ((long long)(t * x + p) * dw * dw) + ab];
        }
        hes[b * nnz_hes + s] = (TIO)v;
    }
}
