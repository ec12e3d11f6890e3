// C-ABI entry points: library info, the host COO->CSR hand-off, SpMM / pooling wrappers and the
// Chebyshev convolution forward / backward compositions.  See include/mvb.h for the contract and
// the reference symbols (file:line) each entry point replaces.
#include <stdarg.h>
#include <atomic>
#include <vector>
#include "mvb_internal.cuh"

namespace mvb {

static thread_local char g_err[512] = {0};
char *err_buf() { return g_err; }
int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace mvb

using namespace mvb;

extern "C" int mvb_version(void) { return MVB_VERSION; }
extern "C" int mvb_sm_arch(void) { return 100; }
extern "C" const char *mvb_last_error(void) { return err_buf(); }

namespace mvb { long long launch_count(); }
extern "C" int64_t mvb_launch_count(void) { return (int64_t)mvb::launch_count(); }

extern "C" int mvb_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(MVB_ECUDA, "mvb_device_cc: %s", cudaGetErrorString(e));
    }
    return major * 10 + minor;
}

// ---------------------------------------------------------------------------------------------
// host: stable counting sort COO -> CSR (of P or of P^T)
// ---------------------------------------------------------------------------------------------
extern "C" int mvb_csr_from_coo_host(int64_t n_out_rows, int64_t n_out_cols, int64_t nnz,
                                     const int64_t *coo_row, const int64_t *coo_col,
                                     const float *coo_val, int transpose, int32_t *rowptr,
                                     int32_t *colidx, float *vals) {
    MVB_REQUIRE(n_out_rows >= 0 && n_out_cols >= 0 && nnz >= 0, "csr_from_coo: negative size");
    MVB_REQUIRE(n_out_rows < INT32_MAX && n_out_cols < INT32_MAX && nnz < INT32_MAX, "csr_from_coo: sizes exceed int32");
    MVB_REQUIRE(rowptr && (nnz == 0 || (coo_row && coo_col && coo_val && colidx && vals)), "csr_from_coo: null pointer");
    const int64_t *r = transpose ? coo_col : coo_row;
    const int64_t *c = transpose ? coo_row : coo_col;
    for (int64_t i = 0; i <= n_out_rows; ++i) rowptr[i] = 0;
    for (int64_t e = 0; e < nnz; ++e) {
        if (r[e] < 0 || r[e] >= n_out_rows || c[e] < 0 || c[e] >= n_out_cols)
            return set_err(MVB_EINVAL, "csr_from_coo: entry %lld = (%lld,%lld) outside [%lld,%lld]", (long long)e,
                           (long long)r[e], (long long)c[e], (long long)n_out_rows, (long long)n_out_cols);
        rowptr[r[e] + 1]++;
    }
    for (int64_t i = 0; i < n_out_rows; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int32_t> cursor(rowptr, rowptr + n_out_rows);
    for (int64_t e = 0; e < nnz; ++e) {
        const int32_t pos = cursor[r[e]]++;
        colidx[pos] = (int32_t)c[e];
        vals[pos] = coo_val[e];
    }
    return MVB_OK;
}

// ---------------------------------------------------------------------------------------------
// SpMM / pooling
// ---------------------------------------------------------------------------------------------
extern "C" int mvb_spmm(int n_rows, const int32_t *rowptr, const int32_t *colidx, const float *vals,
                        const float *x, float *y, const float *z, const float *w, float alpha,
                        float beta, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_rows >= 0 && ncols >= 0, "spmm: negative size");
    MVB_REQUIRE(rowptr && x && y, "spmm: null pointer");
    MVB_REQUIRE(x != y, "spmm: y must not alias x");
    return launch_spmm(n_rows, rowptr, colidx, vals, x, y, z, w, alpha, beta, ncols, (cudaStream_t)stream);
}

extern "C" int mvb_pool_fwd(int n_out_rows, const int32_t *rowptr, const int32_t *colidx,
                            const float *vals, const float *x, float *y, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_out_rows >= 0 && ncols >= 0 && rowptr && x && y, "pool_fwd: bad arguments");
    return launch_spmm(n_out_rows, rowptr, colidx, vals, x, y, nullptr, nullptr, 1.f, 0.f, ncols, (cudaStream_t)stream);
}

extern "C" int mvb_pool_bwd(int n_in_rows, const int32_t *rowptr_t, const int32_t *colidx_t,
                            const float *vals_t, const float *dy, float *dx, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_in_rows >= 0 && ncols >= 0 && rowptr_t && dy && dx, "pool_bwd: bad arguments");
    return launch_spmm(n_in_rows, rowptr_t, colidx_t, vals_t, dy, dx, nullptr, nullptr, 1.f, 0.f, ncols, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Chebyshev convolution
// ---------------------------------------------------------------------------------------------
extern "C" int mvb_cheb_fwd(int N, int B, int Fin, int Fout, int K, const int32_t *rowptr,
                            const int32_t *colidx, const float *vals, const float *x,
                            const float *weight, const float *bias, int relu, float *basis, float *y,
                            void *stream) {
    MVB_REQUIRE(N > 0 && B > 0 && Fin > 0 && Fout > 0 && K > 0, "cheb_fwd: bad sizes N=%d B=%d Fin=%d Fout=%d K=%d", N, B, Fin, Fout, K);
    MVB_REQUIRE(x && weight && y && (K == 1 || (basis && rowptr)), "cheb_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncols = (int64_t)B * Fin;
    const int64_t plane = (int64_t)N * ncols;
    // recurrence: T_1 = L x ; T_k = 2 L T_{k-1} - T_{k-2}     (nn/conv.py:564, 568-569)
    for (int k = 1; k < K; ++k) {
        float *tk = basis + (int64_t)(k - 1) * plane;
        const float *tkm1 = (k == 1) ? x : basis + (int64_t)(k - 2) * plane;
        const float *tkm2 = (k == 1) ? nullptr : (k == 2 ? x : basis + (int64_t)(k - 3) * plane);
        int rc = launch_spmm(N, rowptr, colidx, vals, tkm1, tk, tkm2, nullptr, k == 1 ? 1.f : 2.f, -1.f, ncols, st);
        if (rc) return rc;
    }
    ContractArgs a;
    a.rows = (int64_t)N * B;
    a.in_planes = K;
    a.in_w = Fin;
    a.in0 = x;
    a.in_rest = basis;
    a.mask = nullptr;
    a.wmat = weight;
    a.w_transposed = 0;
    a.bias = bias;
    a.relu = relu;
    a.out_planes = 1;
    a.out_w = Fout;
    a.out = y;
    return launch_contract(a, st);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t mvb_cheb_bwd_workspace_bytes(int N, int B, int Fin, int Fout, int K, int need_dx) {
    size_t bytes = align_up(wgrad_partial_bytes(K * Fin, Fout), 256);
    if (need_dx) bytes += align_up((size_t)K * N * B * Fin * sizeof(float), 256);
    return bytes;
}

extern "C" int mvb_cheb_bwd(int N, int B, int Fin, int Fout, int K, const int32_t *rowptr_t,
                            const int32_t *colidx_t, const float *vals_t, const float *x,
                            const float *basis, const float *weight, const float *y_for_relu,
                            const float *dy, float *dx, float *dweight, float *dbias,
                            void *workspace, size_t workspace_bytes, void *stream) {
    MVB_REQUIRE(N > 0 && B > 0 && Fin > 0 && Fout > 0 && K > 0, "cheb_bwd: bad sizes");
    MVB_REQUIRE(x && weight && dy && dweight && workspace && (K == 1 || basis), "cheb_bwd: null pointer");
    MVB_REQUIRE(!dx || K == 1 || rowptr_t, "cheb_bwd: dx requested without L^T");
    if (!aligned16(workspace)) return set_err(MVB_EALIGN, "cheb_bwd: workspace not 16-byte aligned");
    if (workspace_bytes < mvb_cheb_bwd_workspace_bytes(N, B, Fin, Fout, K, dx != nullptr))
        return set_err(MVB_EWORKSPACE, "cheb_bwd: workspace %zu < %zu", workspace_bytes,
                       mvb_cheb_bwd_workspace_bytes(N, B, Fin, Fout, K, dx != nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)N * B;
    const int64_t ncols = (int64_t)B * Fin;
    const int64_t plane = (int64_t)N * ncols;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t part_bytes = align_up(wgrad_partial_bytes(K * Fin, Fout), 256);

    // dW_k = T_k^T dY, db = 1^T dY
    WgradArgs wa;
    wa.rows = rows;
    wa.in_planes = K;
    wa.in_w = Fin;
    wa.in0 = x;
    wa.in_rest = basis;
    wa.dy = dy;
    wa.mask = y_for_relu;
    wa.n_out = Fout;
    wa.dweight = dweight;
    wa.dbias = dbias;
    wa.partials = reinterpret_cast<float *>(ws);
    wa.partial_bytes = part_bytes;
    int rc = launch_wgrad(wa, st);
    if (rc || !dx) return rc;

    // P_k = dY W_k^T for all k in one pass (plane k of the workspace); P_0 lands in a scratch plane
    float *P = reinterpret_cast<float *>(ws + part_bytes);
    ContractArgs a;
    a.rows = rows;
    a.in_planes = 1;
    a.in_w = Fout;
    a.in0 = dy;
    a.in_rest = nullptr;
    a.mask = y_for_relu;
    a.wmat = weight;          // [K*Fin, Fout] row-major == [Nn, M] -> transposed view
    a.w_transposed = 1;
    a.bias = nullptr;
    a.relu = 0;
    a.out_planes = K;
    a.out_w = Fin;
    a.out = P;
    rc = launch_contract(a, st);
    if (rc) return rc;

    // reverse recurrence, in place on the P planes (G_k overwrites P_k):
    //   G_{K-1} = P_{K-1};  G_k = P_k + 2 L^T G_{k+1} - G_{k+2}  (k >= 1);  dX = P_0 + L^T G_1 - G_2
    for (int k = K - 2; k >= 0; --k) {
        float *pk = P + (int64_t)k * plane;
        const float *gk1 = P + (int64_t)(k + 1) * plane;
        const float *gk2 = (k + 2 <= K - 1) ? P + (int64_t)(k + 2) * plane : nullptr;
        float *dst = (k == 0) ? dx : pk;
        rc = launch_spmm(N, rowptr_t, colidx_t, vals_t, gk1, dst, gk2, pk, k == 0 ? 1.f : 2.f, -1.f, ncols, st);
        if (rc) return rc;
    }
    if (K == 1) {
        cudaError_t e = cudaMemcpyAsync(dx, P, (size_t)plane * sizeof(float), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return set_err(MVB_ECUDA, "cheb_bwd: memcpy: %s", cudaGetErrorString(e));
    }
    return MVB_OK;
}
