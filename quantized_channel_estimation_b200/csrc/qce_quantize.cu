// Quantiser prologue: utils.quant (modules/utils.py:189-203) and get_observation_nbit with A = I
// (modules/utils.py:241-251).  HBM-bound elementwise kernels: one complex sample (16 B) per thread
// per iteration, coalesced 128-bit loads/stores, grid-stride over a grid sized in multiples of the
// SM count.  All arithmetic is IEEE double with explicit round-to-nearest mul/add (no FMA
// contraction) so that the quantised pilots are bit-exact to numpy for identical noise draws.
#include "qce_common.cuh"

namespace qce {

template <bool OBSERVE, bool H_C64>
__global__ void __launch_bounds__(256) quantize_kernel(QuantTables t, bool have_q, const void* __restrict__ src,
                                                       const double2* __restrict__ noise, double noise_scale,
                                                       int64_t n, double2* __restrict__ y_out,
                                                       double2* __restrict__ r_out, uchar2* __restrict__ codes_out) {
    extern __shared__ double s_tab[];      // thr[n_thr] then labels[n_thr + 1] (b > 1 only)
    const int n_thr = t.n_thr;
    if (have_q && t.n_bits > 1) {
        for (int i = threadIdx.x; i < 2 * n_thr + 1; i += blockDim.x)
            s_tab[i] = (i < n_thr) ? t.thr[i] : t.labels[i - n_thr];
        __syncthreads();
    }
    const double* s_thr = s_tab;
    const double* s_lab = s_tab + n_thr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 y;
        if (OBSERVE) {
            double2 h;
            if (H_C64) {
                float2 hf = reinterpret_cast<const float2*>(src)[i];
                h = make_double2((double)hf.x, (double)hf.y);
            } else {
                h = reinterpret_cast<const double2*>(src)[i];
            }
            double2 w = noise[i];
            // y = h + s*n: two roundings (utils.py:247), never an FMA
            y.x = __dadd_rn(h.x, __dmul_rn(noise_scale, w.x));
            y.y = __dadd_rn(h.y, __dmul_rn(noise_scale, w.y));
            if (y_out) y_out[i] = y;
        } else {
            y = reinterpret_cast<const double2*>(src)[i];
        }
        if (!have_q) continue;
        uchar2 code;
        const double2 r = quantize_value(t.n_bits, n_thr, s_thr, s_lab, y, &code);
        if (r_out) r_out[i] = r;
        if (codes_out) codes_out[i] = code;
    }
}

static int quant_grid(int64_t n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t need = (n + 255) / 256;
    int64_t cap = (int64_t)sms * 8;        // 8 resident CTAs of 256 threads per SM
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

qce_status launch_quantize(const QuantTables* t, cudaStream_t s, const double* y, int64_t n, double* r_out,
                           uint8_t* codes_out) {
    if (n == 0) return QCE_OK;
    size_t smem = (t->n_bits > 1) ? (size_t)(2 * t->n_thr + 1) * sizeof(double) : 0;
    quantize_kernel<false, false><<<quant_grid(n), 256, smem, s>>>(*t, true, y, nullptr, 0.0, n, nullptr,
                                                                   (double2*)r_out, (uchar2*)codes_out);
    QCE_CHECK_LAUNCH("quantize_kernel");
    return QCE_OK;
}

qce_status launch_observe_quantize(const QuantTables* t, cudaStream_t s, const void* h, int h_is_c64,
                                   const double* noise, double noise_scale, int64_t n, double* y_out, double* r_out,
                                   uint8_t* codes_out) {
    if (n == 0) return QCE_OK;
    QuantTables tt{};
    bool have_q = (t != nullptr);
    if (have_q) tt = *t;
    size_t smem = (have_q && tt.n_bits > 1) ? (size_t)(2 * tt.n_thr + 1) * sizeof(double) : 0;
    if (h_is_c64)
        quantize_kernel<true, true><<<quant_grid(n), 256, smem, s>>>(tt, have_q, h, (const double2*)noise, noise_scale,
                                                                     n, (double2*)y_out, (double2*)r_out,
                                                                     (uchar2*)codes_out);
    else
        quantize_kernel<true, false><<<quant_grid(n), 256, smem, s>>>(tt, have_q, h, (const double2*)noise, noise_scale,
                                                                      n, (double2*)y_out, (double2*)r_out,
                                                                      (uchar2*)codes_out);
    QCE_CHECK_LAUNCH("observe_quantize_kernel");
    return QCE_OK;
}

// Level codes (the uint8 pairs quantize_kernel writes) back to the quantised pilots: the inverse of the code output, so that a host
// can ship 2 bytes per complex pilot entry instead of 16.  1 bit: 0 neg / 1 zero / 2 pos / 3 NaN; b bit: index into the labels.
__global__ void __launch_bounds__(256) decode_codes_kernel(QuantTables t, const uchar2* __restrict__ codes, int64_t n, double2* __restrict__ r_out) {
    extern __shared__ double s_lab[];
    if (t.n_bits > 1) {
        for (int i = threadIdx.x; i <= t.n_thr; i += blockDim.x) s_lab[i] = t.labels[i];
        __syncthreads();
    }
    const double nan = __longlong_as_double(0x7FF8000000000000LL);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uchar2 c = codes[i];
        double2 r;
        if (t.n_bits == 1) {
            if (c.x > 2 || c.y > 2) r = make_double2(nan, nan);
            else r = make_double2(__dmul_rn(inv_sqrt2(), (double)((int)c.x - 1)) + 0.0, __dmul_rn(inv_sqrt2(), (double)((int)c.y - 1)) + 0.0);
        } else {
            r = make_double2(c.x <= t.n_thr ? s_lab[c.x] : nan, c.y <= t.n_thr ? s_lab[c.y] : nan);
        }
        r_out[i] = r;
    }
}

__global__ void __launch_bounds__(256) c128_to_c64_kernel(const double2* __restrict__ in, int64_t n, float2* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        out[i] = make_float2((float)v.x, (float)v.y);
    }
}

qce_status launch_decode_codes(const QuantTables* t, cudaStream_t s, const uint8_t* codes, int64_t n_complex, double* r_out) {
    if (n_complex == 0) return QCE_OK;
    const size_t smem = (t->n_bits > 1) ? (size_t)(t->n_thr + 1) * sizeof(double) : 0;
    decode_codes_kernel<<<quant_grid(n_complex), 256, smem, s>>>(*t, (const uchar2*)codes, n_complex, (double2*)r_out);
    QCE_CHECK_LAUNCH("decode_codes_kernel");
    return QCE_OK;
}

qce_status launch_c128_to_c64(cudaStream_t s, const double* in, int64_t n_complex, float* out) {
    if (n_complex == 0) return QCE_OK;
    c128_to_c64_kernel<<<quant_grid(n_complex), 256, 0, s>>>((const double2*)in, n_complex, (float2*)out);
    QCE_CHECK_LAUNCH("c128_to_c64_kernel");
    return QCE_OK;
}

}  // namespace qce
