// Circulant / block-circulant (BCCB) Bussgang-GMM estimate kernel.
//
// New algorithm (the reference densifies every covariance type, modules/gmm_cplx_bussgang.py:110-136, and only has a dense
// inference path); checked against the reference's dense estimate_from_y.  With C_h,k = F^H diag(c_k) F, F = F_n1 (x) F_n2
// unitary, A = I and zero means, every quantity of the Bussgang estimator is diagonalised by F:
//     rt = F r          l_k = logc_k - sum_i |rt_i|^2 / lambda_k,i          h = F^H [ (sum_k w_k(l) g_k) .* rt ]
// lambda_k = eigenvalues of C_r,k (arcsine law / beta model, computed on the host), g_k = b_k c_k / lambda_k.
// Per pilot: two small DFTs and two [N x K] real contractions -- no N x N matrix is ever formed.  All arithmetic is
// IEEE double (FP64 SIMT: the kernel is bounded by the FP64 pipe, 2 N (n1 + n2) complex MACs + 2 N K real FMAs per pilot;
// the 16 N bytes of I/O per pilot are far below the HBM roofline at that rate).
// One CTA owns TS pilots; the DFT-domain tile, |rt|^2, the [TS][K] log-probabilities and the per-bin gains live in shared
// memory.  cuFFT-free: the DFT is a direct twiddle-table product per dimension (n <= 256).
#include "qce_common.cuh"

namespace qce {

struct CircArgs {
    int n1, n2, N, K;
    int64_t B;
    const double* inv_lambda_t;   // [N][K]
    const double* gain;           // [K][N]
    const double* logc;           // [K]
    const double2* r;
    double2* h_est;
    double* logp_out;
    const double2* h_true;
    double* acc;
    int mode, n_top, flags;
    double rho;
    const int* rows;              // re-evaluation of listed pilots (null: pilots 0 .. B-1 in order): tile entry i is pilot rows[i],
    const int* n_rows_dev;        // i < *n_rows_dev
};

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// out[s][a][b'] = n^-1/2 sum_b in[s][a][b] w^(b b')   along the contiguous (n2) axis when AXIS2, else along n1
template <int TS, bool AXIS2>
__device__ void dft_pass(const double2* __restrict__ in, double2* __restrict__ out, const double2* __restrict__ tw, int n1, int n2, bool inverse) {
    const int N = n1 * n2, n = AXIS2 ? n2 : n1;
    const double scale = rsqrt((double)n);
    for (int o = threadIdx.x; o < TS * N; o += blockDim.x) {
        const int s = o / N, idx = o % N;
        const int a = idx / n2, b = idx % n2;
        const int f = AXIS2 ? b : a;                                  // output frequency along the transformed axis
        const double2* src = in + (size_t)s * N + (AXIS2 ? a * n2 : b);
        const int stride = AXIS2 ? 1 : n2;
        double2 accv = make_double2(0.0, 0.0);
        int ti = 0;
        for (int j = 0; j < n; ++j) {
            double2 w = tw[ti];
            if (inverse) w.y = -w.y;
            const double2 x = src[j * stride];
            accv.x = fma(x.x, w.x, accv.x); accv.x = fma(-x.y, w.y, accv.x);
            accv.y = fma(x.x, w.y, accv.y); accv.y = fma(x.y, w.x, accv.y);
            ti += f;
            if (ti >= n) ti -= n;
        }
        out[o] = make_double2(accv.x * scale, accv.y * scale);
    }
}

template <int TS>
__global__ void __launch_bounds__(256) circ_kernel(CircArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, K = a.K, n1 = a.n1, n2 = a.n2;
    double2* X = reinterpret_cast<double2*>(smem_raw);                // [TS][N]
    double2* Y = X + (size_t)TS * N;                                  // [TS][N] scratch (later: E, partial sums, G)
    double* lp = reinterpret_cast<double*>(Y + (size_t)TS * N);       // [TS][K]
    double* part = lp + (size_t)TS * K;                               // [256][TS] partial sums of the two contractions
    double2* tw1 = reinterpret_cast<double2*>(part + (size_t)256 * TS);   // [n1] e^{-2 pi i j / n1}
    double2* tw2 = tw1 + n1;                                          // [n2]
    __shared__ double red[3];
    __shared__ int64_t s_row[TS];
    const int t = threadIdx.x;
    for (int j = t; j < n1; j += 256) { double sn, cs; sincospi(-2.0 * j / n1, &sn, &cs); tw1[j] = make_double2(cs, sn); }
    for (int j = t; j < n2; j += 256) { double sn, cs; sincospi(-2.0 * j / n2, &sn, &cs); tw2[j] = make_double2(cs, sn); }
    const int64_t n_rows = a.n_rows_dev ? (int64_t)__ldg(a.n_rows_dev) : a.B;
    for (int64_t base = (int64_t)blockIdx.x * TS; base < n_rows; base += (int64_t)gridDim.x * TS) {
    const int nvalid = (int)((n_rows - base) < TS ? (n_rows - base) : TS);
    if (t < TS) s_row[t] = (t < nvalid) ? (a.rows ? (int64_t)a.rows[base + t] : base + t) : 0;
    if (t < 3) red[t] = 0.0;
    __syncthreads();
    for (int o = t; o < TS * N; o += 256) {
        const int s = o / N;
        X[o] = (s < nvalid) ? a.r[s_row[s] * N + (o % N)] : make_double2(0.0, 0.0);
    }
    __syncthreads();

    // ---- rt = F r  (n2 axis, then n1 axis)
    dft_pass<TS, true>(X, Y, tw2, n1, n2, false);
    __syncthreads();
    if (n1 > 1) { dft_pass<TS, false>(Y, X, tw1, n1, n2, false); }
    else { for (int o = t; o < TS * N; o += 256) X[o] = Y[o]; }
    __syncthreads();

    // ---- E = |rt|^2 (in Y), l = logc - E * inv_lambda^T
    double* E = reinterpret_cast<double*>(Y);                         // [TS][N]
    for (int o = t; o < TS * N; o += 256) { const double2 v = X[o]; E[o] = v.x * v.x + v.y * v.y; }
    __syncthreads();
    {
        const int Kc = K < 256 ? K : 256;                             // components per sweep
        const int ng = 256 / Kc;                                      // thread groups splitting the bin range
        for (int k0 = 0; k0 < K; k0 += Kc) {
            const int kl = t % Kc, g = t / Kc;
            const int k = k0 + kl;
            double accv[TS];
            #pragma unroll
            for (int s = 0; s < TS; ++s) accv[s] = 0.0;
            if (g < ng && k < K) {
                for (int i = g; i < N; i += ng) {
                    const double il = __ldg(a.inv_lambda_t + (size_t)i * K + k);
                    #pragma unroll
                    for (int s = 0; s < TS; ++s) accv[s] = fma(E[s * N + i], il, accv[s]);
                }
                #pragma unroll
                for (int s = 0; s < TS; ++s) part[((size_t)g * TS + s) * Kc + kl] = accv[s];
            }
            __syncthreads();
            for (int o = t; o < TS * Kc; o += 256) {
                const int s = o / Kc, kk = o % Kc;
                if (k0 + kk < K) {
                    double q = 0.0;
                    for (int gg = 0; gg < ng; ++gg) q += part[((size_t)gg * TS + s) * Kc + kk];
                    lp[(size_t)s * K + k0 + kk] = a.logc[k0 + kk] - q;
                }
            }
            __syncthreads();
        }
    }
    if (a.logp_out) {
        for (int o = t; o < nvalid * K; o += 256) a.logp_out[s_row[o / K] * K + (o % K)] = lp[o];
        __syncthreads();
    }
    if (t < TS) weights_from_logp(lp + (size_t)t * K, K, a.mode, a.n_top, a.rho, a.flags);
    __syncthreads();

    // ---- per-bin gain G = sum_k w_k g_k, rt <- G .* rt
    if (a.h_est || a.acc) {
        double* Gp = part;                                            // [ng2][TS][Nc] partial sums
        const int Nc = N < 256 ? N : 256;
        const int ng2 = 256 / Nc;
        for (int i0 = 0; i0 < N; i0 += Nc) {
            const int il = t % Nc, g = t / Nc;
            const int i = i0 + il;
            double accv[TS];
            #pragma unroll
            for (int s = 0; s < TS; ++s) accv[s] = 0.0;
            if (g < ng2 && i < N) {
                for (int k = g; k < K; k += ng2) {
                    const double gk = __ldg(a.gain + (size_t)k * N + i);
                    #pragma unroll
                    for (int s = 0; s < TS; ++s) accv[s] = fma(lp[(size_t)s * K + k], gk, accv[s]);
                }
                #pragma unroll
                for (int s = 0; s < TS; ++s) Gp[((size_t)g * TS + s) * Nc + il] = accv[s];
            }
            __syncthreads();
            for (int o = t; o < TS * Nc; o += 256) {
                const int s = o / Nc, ii = o % Nc;
                if (i0 + ii < N) {
                    double gsum = 0.0;
                    for (int gg = 0; gg < ng2; ++gg) gsum += Gp[((size_t)gg * TS + s) * Nc + ii];
                    double2& x = X[(size_t)s * N + i0 + ii];
                    x.x *= gsum; x.y *= gsum;
                }
            }
            __syncthreads();
        }
        // ---- h = F^H (.)
        if (n1 > 1) { dft_pass<TS, false>(X, Y, tw1, n1, n2, true); __syncthreads(); dft_pass<TS, true>(Y, X, tw2, n1, n2, true); }
        else { dft_pass<TS, true>(X, Y, tw2, n1, n2, true); __syncthreads(); for (int o = t; o < TS * N; o += 256) X[o] = Y[o]; }
        __syncthreads();
        double err = 0.0, pw = 0.0;
        for (int o = t; o < nvalid * N; o += 256) {
            const double2 v = X[o];
            const int64_t go = s_row[o / N] * N + (o % N);
            if (a.h_est) a.h_est[go] = v;
            if (a.acc && a.h_true) {
                const double2 h = a.h_true[go];
                const double dx = v.x - h.x, dy = v.y - h.y;
                err += dx * dx + dy * dy;
                pw += h.x * h.x + h.y * h.y;
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
    __syncthreads();      // the tile buffers are rewritten by the next tile of this CTA
    }
}

template <int TS>
static qce_status launch_ts(const CircArgs& a, cudaStream_t s, size_t smem, int64_t grid_cap = 0) {
    static PerDeviceOnce once;
    if (once.first(current_device())) QCE_CUDA_TRY(cudaFuncSetAttribute(circ_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int64_t grid = (a.B + TS - 1) / TS;
    if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
    circ_kernel<TS><<<(unsigned)grid, 256, smem, s>>>(a);
    QCE_CHECK_LAUNCH("circ_kernel");
    return QCE_OK;
}

qce_status launch_circ(const qce_circ_model* m, cudaStream_t s, const double* r, int64_t B, int mode, int n_top, double rho,
                       double* h_est, double* logp_out, const double* h_true, double* acc) {
    if (B == 0) return QCE_OK;
    CircArgs a;
    a.n1 = m->n1; a.n2 = m->n2; a.N = m->n_ant; a.K = m->n_comp; a.B = B;
    a.inv_lambda_t = m->inv_lambda_t; a.gain = m->gain; a.logc = m->logc;
    a.r = (const double2*)r; a.h_est = (double2*)h_est; a.logp_out = logp_out; a.h_true = (const double2*)h_true; a.acc = acc;
    a.mode = mode; a.n_top = n_top; a.flags = m->flags; a.rho = rho;
    a.rows = nullptr; a.n_rows_dev = nullptr;
    auto need = [&](int ts) {
        return (size_t)2 * ts * a.N * sizeof(double2) + (size_t)ts * a.K * sizeof(double) + (size_t)256 * ts * sizeof(double) +
               (size_t)(a.n1 + a.n2) * sizeof(double2);
    };
    const size_t cap = 220 * 1024;
    if (need(16) <= cap && B >= 16 * 148) return launch_ts<16>(a, s, need(16));
    if (need(8) <= cap && B >= 8 * 64) return launch_ts<8>(a, s, need(8));
    if (need(4) <= cap) return launch_ts<4>(a, s, need(4));
    set_error("circulant kernel: N=%d, K=%d do not fit shared memory", a.N, a.K);
    return QCE_ERR_UNSUPPORTED;
}

// complex128 re-evaluation of a device-side list of pilots (rows[0 .. *n_rows_dev)): small tiles, a grid that does not depend on
// the length of the list (CTAs without a tile exit at once; long lists are walked grid-stride)
qce_status launch_circ_rows(const qce_circ_model* m, cudaStream_t s, const double* r, const int* rows, const int* n_rows_dev, int64_t max_rows,
                            int mode, int n_top, double rho, double* h_est, const double* h_true, double* acc) {
    if (max_rows == 0) return QCE_OK;
    CircArgs a;
    a.n1 = m->n1; a.n2 = m->n2; a.N = m->n_ant; a.K = m->n_comp; a.B = max_rows;
    a.inv_lambda_t = m->inv_lambda_t; a.gain = m->gain; a.logc = m->logc;
    a.r = (const double2*)r; a.h_est = (double2*)h_est; a.logp_out = nullptr; a.h_true = (const double2*)h_true; a.acc = acc;
    a.mode = mode; a.n_top = n_top; a.flags = m->flags; a.rho = rho;
    a.rows = rows; a.n_rows_dev = n_rows_dev;
    const size_t need = (size_t)2 * 4 * a.N * sizeof(double2) + (size_t)4 * a.K * sizeof(double) + (size_t)256 * 4 * sizeof(double) +
                        (size_t)(a.n1 + a.n2) * sizeof(double2);
    if (need > (size_t)220 * 1024) { set_error("circulant kernel: N=%d, K=%d do not fit shared memory", a.N, a.K); return QCE_ERR_UNSUPPORTED; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, current_device());
    return launch_ts<4>(a, s, need, 2 * (int64_t)sms);
}

}  // namespace qce
