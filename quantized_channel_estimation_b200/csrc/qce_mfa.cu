// Bussgang-MFA estimate kernel in Woodbury (low-rank + diagonal) form.
//
// For C_h,k = Lambda_k Lambda_k^H + diag(psi_k) (latent rank M), A = I and a multi-bit quantiser the Bussgang covariance
// C_r,k = beta^2 Lambda Lambda^H + Delta_k stays rank-M + diagonal (SURVEY.md section 2.1 "MFA Woodbury"; the reference
// builds the dense N x N matrix and pinvh's it, modules/mofa_cplx_bussgang.py:199-207).  With U = beta Lambda,
// S = I + U^H Delta^-1 U = L_S L_S^H:
//     x = r - m_r,k                 u = T_k x,  T_k = L_S^-1 U^H Delta^-1                       (M x N)
//     l_k = logc_k - ( sum_i |x_i|^2 / Delta_i - |u|^2 )                                         (mofa:370-381)
//     h_k = mu_k + e_k .* x + Y_k [V1_k x ; u]                                                   (mofa:215-216)
// where W_k = C_h B C_r^-1 = diag(e) + Lambda V1 - (e .* U) L_S^-H T, e = psi b / Delta, Y = [Lambda | -(e .* U) L_S^-H].
// Work per (pilot, component): M N complex MACs for the likelihood and 4 M N for the estimate instead of 2 N^2.
// Complex128 SIMT, same two-phase organisation and mode semantics as dense_fp64_kernel.
#include "qce_common.cuh"

namespace qce {

struct MfaArgs {
    int N, M, K;
    int64_t B;
    const double* inv_delta;   // [K][N]
    const double* evec;        // [K][N]
    const double2* D;          // [K][2M][N]   rows 0..M-1: V1, rows M..2M-1: T
    const double2* Y;          // [K][N][2M]
    const double2* m_r;        // [K][N]
    const double2* mu;         // [K][N]
    const double* logc;        // [K]
    const double2* r;
    double2* h_est;
    double* logp_out;
    const double2* h_true;
    double* acc;
    int mode, n_top, flags;
    double rho;
};

__device__ __forceinline__ void cfma2(double2& acc, const double2 a, const double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

template <int TS>
__global__ void __launch_bounds__(256) mfa_kernel(MfaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, M = a.M, K = a.K, M2 = 2 * a.M;
    double2* rT = reinterpret_cast<double2*>(smem_raw);                 // [N][TS] pilots (transposed)
    double2* xT = rT + (size_t)N * TS;                                  // [N][TS] residual r - m_r,k of the current component
    double2* tb = xT + (size_t)N * TS;                                  // [2M][TS] down-projected residual
    double* lp = reinterpret_cast<double*>(tb + (size_t)M2 * TS);       // [TS][K]
    double* q1 = lp + (size_t)TS * K;                                   // [TS]
    __shared__ double red[3];
    const int t = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * TS;
    const int nvalid = (int)((a.B - base) < TS ? (a.B - base) : TS);

    for (int idx = t; idx < TS * N; idx += 256) {
        const int s = idx / N, i = idx % N;
        rT[(size_t)i * TS + s] = (s < nvalid) ? a.r[(base + s) * N + i] : make_double2(0.0, 0.0);
    }
    if (t < 3) red[t] = 0.0;
    __syncthreads();

    // down-projection of the residual: tb[j][s] = sum_i D[j][i] x[i][s], rows j in [j0, j1)
    auto residual = [&](int k) {
        for (int idx = t; idx < TS * N; idx += 256) {
            const int i = idx / TS, s = idx % TS;
            const double2 m = a.m_r[(size_t)k * N + i], v = rT[idx];
            xT[idx] = make_double2(v.x - m.x, v.y - m.y);
            (void)s;
        }
    };
    auto project = [&](int k, int j0, int j1) {
        const int nj = j1 - j0;
        for (int o = t; o < nj * TS; o += 256) {
            const int j = j0 + o / TS, s = o % TS;
            const double2* __restrict__ Drow = a.D + ((size_t)k * M2 + j) * N;
            double2 accv = make_double2(0.0, 0.0);
            #pragma unroll 4
            for (int i = 0; i < N; ++i) cfma2(accv, __ldg(Drow + i), xT[(size_t)i * TS + s]);
            tb[(size_t)j * TS + s] = accv;
        }
    };

    // ---- phase 1: weighted log-probabilities
    for (int k = 0; k < K; ++k) {
        residual(k);
        __syncthreads();
        project(k, M, M2);                                               // u = T x only
        if (t < TS) {                                                    // sum_i |x_i|^2 / Delta_i
            double q = 0.0;
            for (int i = 0; i < N; ++i) {
                const double2 v = xT[(size_t)i * TS + t];
                q = fma(v.x * v.x + v.y * v.y, a.inv_delta[(size_t)k * N + i], q);
            }
            q1[t] = q;
        }
        __syncthreads();
        if (t < TS) {
            double un = 0.0;
            for (int j = M; j < M2; ++j) { const double2 v = tb[(size_t)j * TS + t]; un += v.x * v.x + v.y * v.y; }
            lp[(size_t)t * K + k] = a.logc[k] - (q1[t] - un);
        }
        __syncthreads();
    }
    if (a.logp_out) {
        for (int idx = t; idx < nvalid * K; idx += 256) a.logp_out[base * K + idx] = lp[idx];
        __syncthreads();
    }
    if (t < TS) weights_from_logp(lp + (size_t)t * K, K, a.mode, a.n_top, a.rho, a.flags);
    __syncthreads();

    // ---- phase 2: weighted combination of the component estimates
    if (a.h_est || a.acc) {
        // each thread owns (pilot s, antennas i = ig + IG*m): IG = 256 / TS antenna groups
        constexpr int IG = 256 / TS;
        constexpr int RM = 16;                                           // antennas per thread per chunk
        const int s = t % TS, ig = t / TS;
        double err = 0.0, pw = 0.0;
        for (int i0 = 0; i0 < N; i0 += IG * RM) {
            double2 accv[RM];
            #pragma unroll
            for (int m = 0; m < RM; ++m) accv[m] = make_double2(0.0, 0.0);
            for (int k = 0; k < K; ++k) {
                // skip components no pilot of the tile uses (block-uniform test keeps the barriers legal)
                bool used = false;
                for (int ss = 0; ss < TS; ++ss) used |= (lp[(size_t)ss * K + k] != 0.0);
                if (!used) continue;
                __syncthreads();
                residual(k);
                __syncthreads();
                project(k, 0, M2);
                __syncthreads();
                const double w = lp[(size_t)s * K + k];
                if (w != 0.0) {
                    #pragma unroll
                    for (int m = 0; m < RM; ++m) {
                        const int i = i0 + ig + IG * m;
                        if (i >= N) break;
                        const double2 x = xT[(size_t)i * TS + s];
                        const double e = a.evec[(size_t)k * N + i];
                        double2 hk = a.mu[(size_t)k * N + i];
                        hk.x = fma(e, x.x, hk.x); hk.y = fma(e, x.y, hk.y);
                        const double2* __restrict__ Yrow = a.Y + ((size_t)k * N + i) * M2;
                        #pragma unroll 4
                        for (int j = 0; j < M2; ++j) cfma2(hk, __ldg(Yrow + j), tb[(size_t)j * TS + s]);
                        accv[m].x += w * hk.x; accv[m].y += w * hk.y;
                    }
                }
            }
            #pragma unroll
            for (int m = 0; m < RM; ++m) {
                const int i = i0 + ig + IG * m;
                if (i >= N || s >= nvalid) continue;
                const int64_t o = (base + s) * N + i;
                if (a.h_est) a.h_est[o] = accv[m];
                if (a.acc && a.h_true) {
                    const double2 h = a.h_true[o];
                    const double dx = accv[m].x - h.x, dy = accv[m].y - h.y;
                    err += dx * dx + dy * dy;
                    pw += h.x * h.x + h.y * h.y;
                }
            }
        }
        if (a.acc) {
            #pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                err += __shfl_xor_sync(0xffffffffu, err, off);
                pw += __shfl_xor_sync(0xffffffffu, pw, off);
            }
            if ((t & 31) == 0) { atomicAdd(&red[0], err); atomicAdd(&red[1], pw); }
            __syncthreads();
            if (t == 0) { atomicAdd(a.acc + 0, red[0]); atomicAdd(a.acc + 1, red[1]); atomicAdd(a.acc + 2, (double)nvalid); }
        }
    }
}

template <int TS>
static qce_status launch_ts(const MfaArgs& a, cudaStream_t s, size_t smem) {
    static PerDeviceOnce once;
    if (once.first(current_device())) QCE_CUDA_TRY(cudaFuncSetAttribute(mfa_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    mfa_kernel<TS><<<(unsigned)((a.B + TS - 1) / TS), 256, smem, s>>>(a);
    QCE_CHECK_LAUNCH("mfa_kernel");
    return QCE_OK;
}

qce_status launch_mfa(const qce_mfa_model* m, cudaStream_t s, const double* r, int64_t B, int mode, int n_top, double rho,
                      double* h_est, double* logp_out, const double* h_true, double* acc) {
    if (B == 0) return QCE_OK;
    MfaArgs a;
    a.N = m->n_ant; a.M = m->latent; a.K = m->n_comp; a.B = B;
    a.inv_delta = m->inv_delta; a.evec = m->evec; a.D = (const double2*)m->D; a.Y = (const double2*)m->Y;
    a.m_r = (const double2*)m->m_r; a.mu = (const double2*)m->mu; a.logc = m->logc;
    a.r = (const double2*)r; a.h_est = (double2*)h_est; a.logp_out = logp_out; a.h_true = (const double2*)h_true; a.acc = acc;
    a.mode = mode; a.n_top = n_top; a.flags = m->flags; a.rho = rho;
    auto need = [&](int ts) {
        return (size_t)(2 * a.N + 2 * a.M) * ts * sizeof(double2) + (size_t)ts * a.K * sizeof(double) + (size_t)ts * sizeof(double);
    };
    const size_t cap = 220 * 1024;
    if (need(32) <= cap && B >= 32 * 148) return launch_ts<32>(a, s, need(32));
    if (need(16) <= cap && B >= 16 * 64) return launch_ts<16>(a, s, need(16));
    if (need(8) <= cap) return launch_ts<8>(a, s, need(8));
    set_error("MFA kernel: N=%d, M=%d, K=%d do not fit shared memory", a.N, a.M, a.K);
    return QCE_ERR_UNSUPPORTED;
}

}  // namespace qce
