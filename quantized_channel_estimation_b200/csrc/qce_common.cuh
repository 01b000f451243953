// Shared declarations of the qce_b200 library (internal; the public ABI is include/qce_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "qce_b200.h"

namespace qce {

void set_error(const char* fmt, ...);
extern int64_t g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

#define QCE_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            qce::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                           __LINE__);                                                           \
            return QCE_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define QCE_CHECK_LAUNCH(name)                                                                  \
    do {                                                                                        \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            qce::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
            return QCE_ERR_CUDA;                                                                \
        }                                                                                       \
        qce::count_launch();                                                                    \
    } while (0)

struct QuantTables {            // device-resident quantiser tables
    int n_bits;
    int n_thr;                  // 2^b - 1
    const double* thr;          // device [n_thr]
    const double* labels;       // device [n_thr + 1]
};

}  // namespace qce

struct qce_quantizer {
    qce::QuantTables t;
    double* dev_buf;            // thr followed by labels
};

// Tensor-core parameter image of one model (see qce_dense_tc.cu)
struct TcParams {
    void* image = nullptr;      // packed FP16 hi/lo operand images, device
    size_t image_bytes = 0;
    void* image2 = nullptr;     // per-CTA half images of the SM-pair (cta_group::2) kernel
    size_t image2_bytes = 0;
    float* zoff = nullptr;      // [K][2*n_obs] whitened offsets (fp32)
    float* hoff = nullptr;      // [K][2*n_ant]
    float* zscale = nullptr;    // [K] power-of-two scale folded out of the Linv_k image
    float* hscale = nullptr;    // [K]
    void* logc2 = nullptr;      // [K] float2: logc_k as an FP32 (hi, lo) pair
    int* flags = nullptr;       // device scratch: [0] offsets non-zero, [1] some Linv_k not lower triangular
    bool has_offsets = false;
    bool triangular = false;
    bool ready = false;
};

struct qce_model {
    int n_obs, n_ant, n_comp, flags;
    // fp64 parameter copies (device)
    double* Linv = nullptr;     // c128 [K][No][No]
    double* W = nullptr;        // c128 [K][N][No]
    double* zoff = nullptr;     // c128 [K][No]
    double* hoff = nullptr;     // c128 [K][N]
    double* logc = nullptr;     // f64  [K]
    double data_scale = 0.0;
    bool params_set = false;
    TcParams tc;
    // scratch for the fused pipeline (quantised pilots of one chunk)
    void* pipe_r = nullptr;
    int64_t pipe_cap = 0;
};

namespace qce {
// qce_quantize.cu
qce_status launch_quantize(const QuantTables* t, cudaStream_t s, const double* y, int64_t n_complex, double* r_out,
                           uint8_t* codes_out);
qce_status launch_observe_quantize(const QuantTables* t, cudaStream_t s, const void* h, int h_is_c64,
                                   const double* noise, double noise_scale, int64_t n_complex, double* y_out,
                                   double* r_out, uint8_t* codes_out);
// qce_dense_fp64.cu
qce_status launch_dense_fp64(const qce_model* m, cudaStream_t s, const double* r, int64_t B, int mode, int n_top,
                             double rho, double* h_est, double* logp_out, const double* h_true, double* acc);
qce_status launch_dense_fp64_raw(const qce_model* m, cudaStream_t s, const double* r, int64_t B, int mode, int n_top,
                                 double rho, double* h_est, double* logp_out, const void* h_true, int h_true_c64,
                                 double* acc);
// qce_dense_tc.cu
qce_status tc_pack_params(qce_model* m, cudaStream_t s);
void tc_free(qce_model* m);
qce_status launch_dense_tc(qce_model* m, cudaStream_t s, const double* r, int64_t B, int mode, int n_top,
                           double rho, double* h_est, double* logp_out, const void* h_true, int h_true_c64,
                           double* acc);
qce_status tc_format(qce_model* m, cudaStream_t s, const double* r, int64_t B);
qce_status tc_estimate_formatted(qce_model* m, cudaStream_t s, int64_t B, double* h_est, const void* h_true, int h_true_c64,
                                 double* acc);
qce_status launch_pipeline_tc(qce_model* m, const QuantTables* qt, cudaStream_t s, const void* h, int h_is_c64,
                              const double* noise, double noise_scale, int64_t B, int mode, int n_top, double rho, double* h_est,
                              double* acc);
bool tc_supported(const qce_model* m, int mode);
}  // namespace qce
