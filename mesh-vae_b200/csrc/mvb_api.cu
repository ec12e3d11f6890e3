// C-ABI entry points: library info, the host COO->CSR hand-off, SpMM / pooling wrappers and the
// Chebyshev convolution forward / backward compositions.  See include/mvb.h for the contract and
// the reference symbols (file:line) each entry point replaces.
#include <stdarg.h>
#include <stdlib.h>
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
    static int cached[64] = {0};
    const int d = device_slot();
    if (cached[d] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) == cudaSuccess && n > 0)
            cached[d] = n;
        else {
            cudaGetLastError();
            cached[d] = 148;
        }
    }
    return cached[d];
}

}  // namespace mvb

using namespace mvb;

namespace mvb {
static int g_pdl = 1;
int pdl_enabled() { return g_pdl; }
void set_pdl(int v) { g_pdl = v ? 1 : 0; }
}  // namespace mvb

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
extern "C" int mvb_spmm(int n_rows, int n_src_rows, const int32_t *rowptr, const int32_t *colidx, const float *vals,
                        const float *x, float *y, const float *z, const float *w, float alpha,
                        float beta, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_rows >= 0 && ncols >= 0, "spmm: negative size");
    MVB_REQUIRE(rowptr && x && y, "spmm: null pointer");
    MVB_REQUIRE(x != y, "spmm: y must not alias x");
    return launch_spmm(n_rows, n_src_rows, rowptr, colidx, vals, x, y, z, w, alpha, beta, ncols, (cudaStream_t)stream);
}

extern "C" int mvb_pool_fwd(int n_out_rows, int n_in_rows, const int32_t *rowptr, const int32_t *colidx,
                            const float *vals, const float *x, float *y, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_out_rows >= 0 && ncols >= 0 && rowptr && x && y, "pool_fwd: bad arguments");
    return launch_spmm(n_out_rows, n_in_rows, rowptr, colidx, vals, x, y, nullptr, nullptr, 1.f, 0.f, ncols, (cudaStream_t)stream);
}

extern "C" int mvb_pool_bwd(int n_in_rows, int n_out_rows, const int32_t *rowptr_t, const int32_t *colidx_t,
                            const float *vals_t, const float *dy, float *dx, int64_t ncols, void *stream) {
    MVB_REQUIRE(n_in_rows >= 0 && ncols >= 0 && rowptr_t && dy && dx, "pool_bwd: bad arguments");
    return launch_spmm(n_in_rows, n_out_rows, rowptr_t, colidx_t, vals_t, dy, dx, nullptr, nullptr, 1.f, 0.f, ncols, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Chebyshev convolution
// ---------------------------------------------------------------------------------------------
// n_active: rows >= n_active of L have no entries and no entry references a column >= n_active
// (n_active == N for an ordinary operator).  For the empty rows the recurrence has the closed form
// T_k = c_k x, c_k = cos(k pi/2), so they collapse to ONE plane with the folded weight
// sum_k c_k W_k; only the first n_active vertices (a contiguous prefix in the vertex-major layout)
// run the SpMM recurrence.  This is what makes the reference's output layer - the 20-vertex
// operator applied to the 4998-vertex mesh, models/cheb_VAE.py:288 - a streaming pass.
static void fill_contract(ContractArgs &a) {
    memset(&a, 0, sizeof(a));
}

extern "C" int mvb_cheb_fwd(int N, int B, int Fin, int Fout, int K, int n_active, int nnz,
                            const int32_t *rowptr, const int32_t *colidx, const float *vals,
                            const float *x, const float *weight, const float *bias, int relu,
                            float *basis, float *y, void *stream) {
    MVB_REQUIRE(N > 0 && B > 0 && Fin > 0 && Fout > 0 && K > 0, "cheb_fwd: bad sizes N=%d B=%d Fin=%d Fout=%d K=%d", N, B, Fin, Fout, K);
    MVB_REQUIRE(n_active >= 0 && n_active <= N, "cheb_fwd: n_active=%d outside [0,%d]", n_active, N);
    MVB_REQUIRE(x && weight && y && (K == 1 || n_active == 0 || (basis && rowptr)), "cheb_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncols = (int64_t)B * Fin;
    const int64_t plane = (int64_t)n_active * ncols;
    int rc;
    if (n_active > 0) {
        // recurrence: T_1 = L x ; T_k = 2 L T_{k-1} - T_{k-2}     (nn/conv.py:564, 568-569)
        // one fused launch when the level fits shared memory, else one SpMM launch per step
        rc = (K > 1) ? launch_cheb_recur_fwd(n_active, nnz, K, rowptr, colidx, vals, x, basis, ncols, st) : 1;
        if (rc < 0) return rc;
        for (int k = 1; k < K && rc == 0; ++k) {
            float *tk = basis + (int64_t)(k - 1) * plane;
            const float *tkm1 = (k == 1) ? x : basis + (int64_t)(k - 2) * plane;
            const float *tkm2 = (k == 1) ? nullptr : (k == 2 ? x : basis + (int64_t)(k - 3) * plane);
            int rc2 = launch_spmm(n_active, n_active, rowptr, colidx, vals, tkm1, tk, tkm2, nullptr, k == 1 ? 1.f : 2.f, -1.f, ncols, st);
            if (rc2) return rc2;
        }
        ContractArgs a;
        fill_contract(a);
        a.rows = (int64_t)n_active * B;
        a.in_planes = K;
        a.in_w = Fin;
        a.in0 = x;
        a.in_rest = basis;
        a.wmat = weight;
        a.bias = bias;
        a.relu = relu;
        a.out_planes = 1;
        a.out_w = Fout;
        a.out = y;
        rc = launch_contract(a, st);
        if (rc) return rc;
    }
    if (n_active < N) {
        const int64_t off = (int64_t)n_active * B;
        rc = launch_fold_fwd((int64_t)(N - n_active) * B, K, Fin, Fout, x + off * Fin, weight, bias, relu, y + off * Fout, st);
        if (rc < 0) return rc;
        if (rc == 1) return MVB_OK;
        ContractArgs a;
        fill_contract(a);
        a.rows = (int64_t)(N - n_active) * B;
        a.in_planes = 1;
        a.in_w = Fin;
        a.in0 = x + off * Fin;
        a.wmat = weight;
        a.w_fold = K;
        a.bias = bias;
        a.relu = relu;
        a.out_planes = 1;
        a.out_w = Fout;
        a.out = y + off * Fout;
        rc = launch_contract(a, st);
        if (rc) return rc;
    }
    return MVB_OK;
}

static inline size_t align_up_c(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Chebyshev convolution followed by a row selection (the encoder loop body at the levels that do not fit a
// mesh-resident kernel: x = relu(cheb[i](x, L)); x = pool(x, D), models/cheb_VAE.py:264-265, D = one 1.0 per row,
// mesh_operations.py:72-85).  Only the rows D keeps are contracted and written (1/4 of them at every level), and the
// backward pass - the basis form for a layer whose input needs no gradient, i.e. the first encoder layer - reduces
// dW / db over those rows only: dY is zero everywhere else.  Tensor-core kernels only (4 / 16 / 32-wide planes).
// ---------------------------------------------------------------------------------------------
static bool sel_shape_ok(int N, int B, int Fin, int Fout, int K, int n_sel) {
    if (!tc_enabled() || N < 1 || B < 1 || K < 1 || n_sel < 1 || n_sel > N) return false;
    if (!((Fin == 4 && K <= 8) || (Fin == 16 && K <= 6) || (Fin == 32 && K <= 3))) return false;     // planar layouts of tc_rowgemm / tc_wgrad
    if (Fout % 4 || Fout > 32 || (int64_t)n_sel * B < 256) return false;
    if (K * Fin + 1 > 128) return false;
    return true;
}

extern "C" int mvb_cheb_sel_supported(int N, int B, int Fin, int Fout, int K, int n_sel) {
    return sel_shape_ok(N, B, Fin, Fout, K, n_sel) ? 1 : 0;
}

extern "C" int mvb_cheb_sel_fwd(int N, int B, int Fin, int Fout, int K, int nnz, const int32_t *rowptr, const int32_t *colidx,
                                const float *vals, const float *x, const float *weight, const float *bias, int relu, int n_sel,
                                const int32_t *sel, float *basis, float *y_sel, void *stream) {
    MVB_REQUIRE(sel_shape_ok(N, B, Fin, Fout, K, n_sel), "cheb_sel_fwd: shape N=%d B=%d Fin=%d Fout=%d K=%d n_sel=%d not supported", N, B, Fin,
                Fout, K, n_sel);
    MVB_REQUIRE(x && weight && y_sel && sel && (K == 1 || (basis && rowptr)), "cheb_sel_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ncols = (int64_t)B * Fin;
    const int64_t plane = (int64_t)N * ncols;
    int rc = (K > 1) ? launch_cheb_recur_fwd(N, nnz, K, rowptr, colidx, vals, x, basis, ncols, st) : 1;
    if (rc < 0) return rc;
    for (int k = 1; k < K && rc == 0; ++k) {
        float *tk = basis + (int64_t)(k - 1) * plane;
        const float *tkm1 = (k == 1) ? x : basis + (int64_t)(k - 2) * plane;
        const float *tkm2 = (k == 1) ? nullptr : (k == 2 ? x : basis + (int64_t)(k - 3) * plane);
        int rc2 = launch_spmm(N, N, rowptr, colidx, vals, tkm1, tk, tkm2, nullptr, k == 1 ? 1.f : 2.f, -1.f, ncols, st);
        if (rc2) return rc2;
    }
    ContractArgs a;
    fill_contract(a);
    a.rows = (int64_t)n_sel * B;
    a.in_planes = K;
    a.in_w = Fin;
    a.in0 = x;
    a.in_rest = basis;
    a.wmat = weight;
    a.bias = bias;
    a.relu = relu;
    a.out_planes = 1;
    a.out_w = Fout;
    a.out = y_sel;
    a.row_sel = sel;
    a.sel_group = B;
    a.plane_rows = (int64_t)N * B;
    rc = launch_contract_tc(a, st);
    if (rc < 0) return rc;
    if (rc == 0) return set_err(MVB_EINVAL, "cheb_sel_fwd: the tensor-core contraction does not cover this call (alignment?)");
    return MVB_OK;
}

extern "C" size_t mvb_cheb_sel_bwd_workspace_bytes(int Fin, int Fout, int K) {
    return align_up_c(wgrad_partial_bytes(K * Fin, Fout), 256);
}

extern "C" int mvb_cheb_sel_bwd(int N, int B, int Fin, int Fout, int K, const float *x, const float *basis,
                                const float *y_sel_for_relu, const float *dy_sel, int n_sel, const int32_t *sel, float *dweight,
                                float *dbias, void *workspace, size_t workspace_bytes, void *stream) {
    MVB_REQUIRE(sel_shape_ok(N, B, Fin, Fout, K, n_sel), "cheb_sel_bwd: shape not supported");
    MVB_REQUIRE(x && dy_sel && sel && dweight && workspace && (K == 1 || basis), "cheb_sel_bwd: null pointer");
    const size_t need = mvb_cheb_sel_bwd_workspace_bytes(Fin, Fout, K);
    if (workspace_bytes < need) return set_err(MVB_EWORKSPACE, "cheb_sel_bwd: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.rows = (int64_t)n_sel * B;
    wa.in_planes = K;
    wa.in_w = Fin;
    wa.in0 = x;
    wa.in_rest = basis;
    wa.dy = dy_sel;
    wa.mask = y_sel_for_relu;
    wa.n_out = Fout;
    wa.partials = reinterpret_cast<float *>(workspace);
    wa.partial_bytes = need;
    wa.row_sel = sel;
    wa.sel_group = B;
    wa.plane_rows = (int64_t)N * B;
    int nA = 0, m4A = 0;
    int rc = launch_wgrad_partials(wa, dbias != nullptr, &nA, &m4A, st);
    if (rc) return rc;
    return launch_wgrad_finalize(wa.partials, nA, m4A, nullptr, 0, 0, Fin, K * Fin, Fout, dweight, dbias, st);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// per-host-thread side stream + fork/join events for intra-call concurrency
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
static int g_overlap = 1;
static SideStream *side_stream() {
    static thread_local SideStream s;
    static thread_local bool tried = false;
    if (!tried) {
        tried = true;
        // creation is not a stream operation: legal while another stream is capturing only in relaxed
        // mode, so it is done during the (uncaptured) warm-up calls; if it fails we simply do not fork
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess)
            s.ok = true;
        else
            cudaGetLastError();
    }
    return s.ok ? &s : nullptr;
}

// Two algebraically equal forms of the backward pass (L acts on vertices, W on features: they commute):
//   basis form  : dW_k = T_k(L x)^T G,  dX = sum_k T_k(L^T) (G W_k^T)      - needs the forward basis T_k,
//                 K planes P_k = G W_k^T and a reverse recurrence on Fin-wide planes;
//   adjoint form: S_k = T_k(L^T) G,  dW_k = x^T S_k,  dX = sum_k S_k W_k^T - the FORWARD kernels run on G
//                 (recurrence on Fout-wide planes, one contraction), nothing saved by the forward pass.
// Plane traffic is ~(32 Fin + 2 Fout) vs ~(27 Fout + 2 Fin) plane-widths: the adjoint form is used when
// the input gradient is wanted and Fout <= 1.2 Fin (every 16->16 layer of the reference model; the
// first encoder layer, which needs no dX, keeps the basis form: its basis is only 3-4 columns wide).
static bool adjoint_form(int Fin, int Fout, int need_dx) { return need_dx && 25 * Fout <= 30 * Fin && Fout % 4 == 0; }

// fork the per-thread side stream from `st` (NULL when overlap is off or the fork fails)
static SideStream *fork_side(cudaStream_t st) {
    if (!g_overlap) return nullptr;
    SideStream *side = side_stream();
    if (!side) return nullptr;
    if (cudaEventRecord(side->fork, st) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return side;
}
struct JoinGuard {   // the side stream must rejoin on every exit path (a captured graph must not end forked)
    SideStream *s;
    cudaStream_t main;
    ~JoinGuard() {
        if (s) {
            cudaEventRecord(s->join, s->stream);
            cudaStreamWaitEvent(main, s->join, 0);
        }
    }
};

// the same fork / join for other translation units (mvb_layer.cu): returns the side stream or NULL
namespace mvb {
cudaStream_t side_fork(cudaStream_t st) {
    SideStream *s = fork_side(st);
    return s ? s->stream : nullptr;
}
void side_join(cudaStream_t side, cudaStream_t st) {
    SideStream *s = side_stream();
    if (side && s && s->stream == side) {
        cudaEventRecord(s->join, s->stream);
        cudaStreamWaitEvent(st, s->join, 0);
    }
}
}  // namespace mvb

// ---------------------------------------------------------------------------------------------
// Deferred side chains.  The weight-gradient reduction of a mesh-resident layer (tc_wgrad + finalize kernels, 10-20 us)
// is not on the critical path of the backward pass: nothing reads dW / db before the optimizer.  In deferred mode
// (mvb_tune "defer_wgrad=1", switched on by the step engine around its backward pass) such a chain is forked onto a
// per-device side stream (the chains of successive layers queue up there, each started by an event on the caller's
// stream once its producer kernel has been enqueued) and joined back only by mvb_side_join - so they fill the idle
// SMs / memory bandwidth under the following layers' latency-bound kernels; in a captured CUDA graph they form a
// parallel branch.  Off (default): the chain is joined before the call returns.
// The state is per device and process-wide (autograd runs backward functions on its own thread).
// ---------------------------------------------------------------------------------------------
#include <mutex>
namespace mvb {
struct LazySide {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool tried = false, ok = false, pending = false;
};
constexpr int LAZY_SLOTS = 8;       // slot 1 = lane 1 (dense-layer weight gradients); slots 0, 2..7 = lane 0 (convolution weight gradients:
                                    // successive chains go round robin over g_conv_lanes of them, so that chains can overlap each other)
static LazySide g_lazy[64][LAZY_SLOTS];
static int g_conv_lanes = 3;        // tuning: conv_lanes=1..7; same-box A/B (bench.py, 300 steps, second pass): 1 lane 0.9197 ms per step,
                                    // 2: 0.9152, 3: 0.9116, 4: 0.9109, 5: 0.9111, 7: 0.9109 - the chains queued behind each other in the tail of the
                                    // backward pass (the last layers' reductions) now run side by side
static int g_conv_next[64] = {0};
void set_conv_lanes(int v) { if (v >= 1 && v <= 7) g_conv_lanes = v; }
static const int kConvSlot[7] = {0, 2, 3, 4, 5, 6, 7};
static std::mutex g_lazy_mu;
static int g_defer_wgrad = 0;
void set_defer_wgrad(int v) { g_defer_wgrad = v ? 1 : 0; }
static int g_background_div = 0;        // deferred level-0 weight-gradient reduction: SMs per CTA, 0 = the usual two CTAs per SM
                                        // (tuning: background_div; measured at 64 meshes: 0 -> 921 us per step, 1 -> 948, 2 -> 988:
                                        // the slower reduction delays every chain queued behind it more than it spares the main stream)
void set_background_div(int v) { g_background_div = v < 0 ? 0 : v; }

static LazySide *lazy_side(int slot) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    LazySide &s = g_lazy[dev][slot & (LAZY_SLOTS - 1)];
    if (!s.tried) {
        s.tried = true;
        // LOWEST priority: when a chain and a kernel of the caller's stream are ready together (the input-gradient
        // contraction next to the weight-gradient reduction of the same layer), the block scheduler serves the caller's
        // kernel first - the chain takes what is left (a captured graph keeps the priority as a kernel-node attribute)
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { least = 0; cudaGetLastError(); }
        if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, least) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess)
            s.ok = true;
        else
            cudaGetLastError();
    }
    return s.ok ? &s : nullptr;
}

// make `st` wait for a pending chain (if any)
static void lazy_join_locked(LazySide *s, cudaStream_t st) {
    if (s && s->pending) {
        cudaStreamWaitEvent(st, s->join, 0);
        s->pending = false;
    }
}

// Call AFTER the producer kernel of the chain has been enqueued on `st`.  Makes the side stream wait for that point of
// `st`; returns the stream to enqueue the chain on, or NULL (not deferred / unavailable): run it on `st`.
cudaStream_t lazy_fork(cudaStream_t st, int lane) {
    if (!g_defer_wgrad) return nullptr;
    std::lock_guard<std::mutex> lk(g_lazy_mu);
    int slot = 1;
    if (lane == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
        slot = kConvSlot[g_conv_next[dev] % g_conv_lanes];
        g_conv_next[dev] = (g_conv_next[dev] + 1) % g_conv_lanes;
    }
    LazySide *s = lazy_side(slot);
    if (!s) return nullptr;
    if (cudaEventRecord(s->fork, st) != cudaSuccess || cudaStreamWaitEvent(s->stream, s->fork, 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return s->stream;
}

// the chain has been enqueued on `side`: record its end (the side stream is in order: the latest record covers every
// earlier chain); joined by mvb_side_join
void lazy_done(cudaStream_t side, cudaStream_t st) {
    if (!side) return;
    std::lock_guard<std::mutex> lk(g_lazy_mu);
    LazySide *s = nullptr;
    for (int slot = 0; slot < LAZY_SLOTS; ++slot) {
        LazySide *c = lazy_side(slot);
        if (c && c->stream == side) { s = c; break; }
    }
    if (!s) return;
    cudaEventRecord(s->join, s->stream);
    s->pending = true;
    if (!g_defer_wgrad) lazy_join_locked(s, st);
}
}  // namespace mvb

extern "C" int mvb_side_join(void *stream) {
    std::lock_guard<std::mutex> lk(mvb::g_lazy_mu);
    for (int slot = 0; slot < mvb::LAZY_SLOTS; ++slot) mvb::lazy_join_locked(mvb::lazy_side(slot), (cudaStream_t)stream);
    return MVB_OK;
}

// join only the chains of one lane (1 = the dense layers' weight gradients: the data-parallel engine reduces their
// bucket while the convolution chains of lane 0 are still running)
extern "C" int mvb_side_join_lane(void *stream, int lane) {
    std::lock_guard<std::mutex> lk(mvb::g_lazy_mu);
    if (lane == 1) {
        mvb::lazy_join_locked(mvb::lazy_side(1), (cudaStream_t)stream);
    } else {
        for (int c = 0; c < 7; ++c) mvb::lazy_join_locked(mvb::lazy_side(mvb::kConvSlot[c]), (cudaStream_t)stream);
    }
    return MVB_OK;
}

extern "C" int mvb_cheb_bwd_uses_basis(int Fin, int Fout, int need_dx) { return adjoint_form(Fin, Fout, need_dx) ? 0 : 1; }

extern "C" size_t mvb_cheb_bwd_workspace_bytes(int N, int B, int Fin, int Fout, int K, int n_active, int need_dx) {
    if (adjoint_form(Fin, Fout, need_dx)) {
        size_t bytes = align_up(wgrad_partial_bytes(K * Fout, Fin), 256);
        if (n_active < N) bytes += align_up(wgrad_partial_bytes(Fin, Fout), 256);
        bytes += align_up((size_t)mask_colsum_blocks((int64_t)N * B) * Fout * sizeof(float), 256);
        bytes += align_up((size_t)N * B * Fout * sizeof(float), 256);                       // G
        bytes += align_up((size_t)(K > 1 ? K - 1 : 0) * n_active * B * Fout * sizeof(float), 256);   // S_1..S_{K-1}
        return bytes;
    }
    size_t bytes = align_up(wgrad_partial_bytes(K * Fin, Fout), 256);
    if (n_active < N) bytes += align_up(wgrad_partial_bytes(Fin, Fout), 256);
    if (need_dx) bytes += align_up((size_t)K * n_active * B * Fin * sizeof(float), 256);
    return bytes;
}

static int cheb_bwd_adjoint(int N, int B, int Fin, int Fout, int K, int n_active, int nnz, const int32_t *rowptr_t,
                            const int32_t *colidx_t, const float *vals_t, const float *x, const float *weight,
                            const float *y_for_relu, const float *dy, float *dx, float *dweight, float *dbias, char *ws,
                            cudaStream_t st);

extern "C" int mvb_cheb_bwd(int N, int B, int Fin, int Fout, int K, int n_active, int nnz,
                            const int32_t *rowptr_t, const int32_t *colidx_t, const float *vals_t,
                            const float *x, const float *basis, const float *weight,
                            const float *y_for_relu, const float *dy, float *dx, float *dweight,
                            float *dbias, void *workspace, size_t workspace_bytes, void *stream) {
    MVB_REQUIRE(N > 0 && B > 0 && Fin > 0 && Fout > 0 && K > 0, "cheb_bwd: bad sizes");
    MVB_REQUIRE(n_active >= 0 && n_active <= N, "cheb_bwd: n_active=%d outside [0,%d]", n_active, N);
    const bool adj = adjoint_form(Fin, Fout, dx != nullptr);
    MVB_REQUIRE(x && weight && dy && dweight && workspace && (adj || K == 1 || n_active == 0 || basis), "cheb_bwd: null pointer");
    MVB_REQUIRE(!dx || K == 1 || n_active == 0 || rowptr_t, "cheb_bwd: dx requested without L^T");
    if (!aligned16(workspace)) return set_err(MVB_EALIGN, "cheb_bwd: workspace not 16-byte aligned");
    const size_t need = mvb_cheb_bwd_workspace_bytes(N, B, Fin, Fout, K, n_active, dx != nullptr);
    if (workspace_bytes < need) return set_err(MVB_EWORKSPACE, "cheb_bwd: workspace %zu < %zu", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    if (adj) {
        if (!aligned16(dy) || (y_for_relu && !aligned16(y_for_relu)))
            return set_err(MVB_EALIGN, "cheb_bwd: dy / y must be 16-byte aligned");
        return cheb_bwd_adjoint(N, B, Fin, Fout, K, n_active, nnz, rowptr_t, colidx_t, vals_t, x, weight, y_for_relu, dy, dx,
                                dweight, dbias, reinterpret_cast<char *>(workspace), st);
    }
    const int64_t rows_act = (int64_t)n_active * B;
    const int64_t rows_in = (int64_t)(N - n_active) * B;
    const int64_t ncols = (int64_t)B * Fin;
    const int64_t plane = (int64_t)n_active * ncols;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t partA_bytes = align_up(wgrad_partial_bytes(K * Fin, Fout), 256);
    const size_t partB_bytes = (n_active < N) ? align_up(wgrad_partial_bytes(Fin, Fout), 256) : 0;
    float *partA = reinterpret_cast<float *>(ws);
    float *partB = reinterpret_cast<float *>(ws + partA_bytes);
    float *P = reinterpret_cast<float *>(ws + partA_bytes + partB_bytes);
    const int has_bias = dbias != nullptr;
    int rc, nA = 0, m4A = 0, nB = 0, m4B = 0;

    // The weight-gradient branch (partials + ordered finalize) and the input-gradient branch
    // (P_k, reverse recurrence) are independent: fork the former onto a side stream so that the two
    // overlap (and become parallel branches of a captured CUDA graph), join before returning.
    cudaStream_t wst = st;
    SideStream *side = dx ? fork_side(st) : nullptr;
    if (side) wst = side->stream;
    JoinGuard join_guard{side, st};

    // dW_k = T_k^T dY, db = 1^T dY over the active prefix ...
    WgradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.rows = rows_act;
    wa.in_planes = K;
    wa.in_w = Fin;
    wa.in0 = x;
    wa.in_rest = basis;
    wa.dy = dy;
    wa.mask = y_for_relu;
    wa.n_out = Fout;
    wa.partials = partA;
    wa.partial_bytes = partA_bytes;
    rc = launch_wgrad_partials(wa, has_bias, &nA, &m4A, wst);
    if (rc) return rc;
    // ... plus S = x^T dY over the empty rows (dW_k += c_k S)
    if (rows_in > 0) {
        WgradArgs wb;
        memset(&wb, 0, sizeof(wb));
        wb.rows = rows_in;
        wb.in_planes = 1;
        wb.in_w = Fin;
        wb.in0 = x + rows_act * Fin;
        wb.dy = dy + rows_act * Fout;
        wb.mask = y_for_relu ? y_for_relu + rows_act * Fout : nullptr;
        wb.n_out = Fout;
        wb.partials = partB;
        wb.partial_bytes = partB_bytes;
        rc = launch_wgrad_partials(wb, has_bias, &nB, &m4B, wst);
        if (rc) return rc;
    }
    rc = launch_wgrad_finalize(partA, nA, m4A, partB, nB, m4B, Fin, K * Fin, Fout, dweight, dbias, wst);
    if (rc || !dx) return rc;

    if (rows_act > 0) {
        // P_k = dY W_k^T for all k in one pass (plane k of the workspace)
        ContractArgs a;
        fill_contract(a);
        a.rows = rows_act;
        a.in_planes = 1;
        a.in_w = Fout;
        a.in0 = dy;
        a.mask = y_for_relu;
        a.wmat = weight;          // [K*Fin, Fout] row-major == [Nn, M] -> transposed view
        a.w_transposed = 1;
        a.out_planes = K;
        a.out_w = Fin;
        a.out = P;
        rc = launch_contract(a, st);
        if (rc) return rc;
        // reverse recurrence, in place on the P planes (G_k overwrites P_k):
        //   G_{K-1} = P_{K-1};  G_k = P_k + 2 L^T G_{k+1} - G_{k+2}  (k >= 1);  dX = P_0 + L^T G_1 - G_2
        int fused = (K > 1) ? launch_cheb_recur_bwd(n_active, nnz, K, rowptr_t, colidx_t, vals_t, P, dx, ncols, st) : 0;
        if (fused < 0) return fused;
        for (int k = K - 2; k >= 0 && fused == 0; --k) {
            float *pk = P + (int64_t)k * plane;
            const float *gk1 = P + (int64_t)(k + 1) * plane;
            const float *gk2 = (k + 2 <= K - 1) ? P + (int64_t)(k + 2) * plane : nullptr;
            float *dst = (k == 0) ? dx : pk;
            rc = launch_spmm(n_active, n_active, rowptr_t, colidx_t, vals_t, gk1, dst, gk2, pk, k == 0 ? 1.f : 2.f, -1.f, ncols, st);
            if (rc) return rc;
        }
        if (K == 1) {
            cudaError_t e = cudaMemcpyAsync(dx, P, (size_t)plane * sizeof(float), cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return set_err(MVB_ECUDA, "cheb_bwd: memcpy: %s", cudaGetErrorString(e));
        }
    }
    if (rows_in > 0) {
        // empty rows: dX = dY (sum_k c_k W_k)^T
        ContractArgs a;
        fill_contract(a);
        a.rows = rows_in;
        a.in_planes = 1;
        a.in_w = Fout;
        a.in0 = dy + rows_act * Fout;
        a.mask = y_for_relu ? y_for_relu + rows_act * Fout : nullptr;
        a.wmat = weight;
        a.w_transposed = 1;
        a.w_fold = K;
        a.out_planes = 1;
        a.out_w = Fin;
        a.out = dx + rows_act * Fin;
        rc = launch_contract(a, st);
        if (rc) return rc;
    }
    return MVB_OK;
}

static int cheb_bwd_adjoint(int N, int B, int Fin, int Fout, int K, int n_active, int nnz, const int32_t *rowptr_t,
                            const int32_t *colidx_t, const float *vals_t, const float *x, const float *weight,
                            const float *y_for_relu, const float *dy, float *dx, float *dweight, float *dbias, char *ws,
                            cudaStream_t st) {
    const int64_t rows_act = (int64_t)n_active * B, rows_in = (int64_t)(N - n_active) * B, rows_all = (int64_t)N * B;
    const int64_t ncols = (int64_t)B * Fout;
    const int64_t plane = (int64_t)n_active * ncols;
    const size_t partA_bytes = align_up(wgrad_partial_bytes(K * Fout, Fin), 256);
    const size_t partB_bytes = (n_active < N) ? align_up(wgrad_partial_bytes(Fin, Fout), 256) : 0;
    const size_t col_bytes = align_up((size_t)mask_colsum_blocks(rows_all) * Fout * sizeof(float), 256);
    const size_t g_bytes = align_up((size_t)rows_all * Fout * sizeof(float), 256);
    float *partA = reinterpret_cast<float *>(ws);
    float *partB = reinterpret_cast<float *>(ws + partA_bytes);
    float *colpart = reinterpret_cast<float *>(ws + partA_bytes + partB_bytes);
    float *Gbuf = reinterpret_cast<float *>(ws + partA_bytes + partB_bytes + col_bytes);
    float *S = reinterpret_cast<float *>(ws + partA_bytes + partB_bytes + col_bytes + g_bytes);
    // G = dY * [y > 0] over ALL rows, db = column sums of G
    const float *G = y_for_relu ? Gbuf : dy;
    int rc = launch_mask_colsum(rows_all, Fout, dy, y_for_relu, y_for_relu ? Gbuf : nullptr, dbias, colpart, st);
    if (rc) return rc;
    // S_k = T_k(L^T) G on the active prefix: the forward recurrence kernels with CSR(L^T)
    if (n_active > 0 && K > 1) {
        rc = launch_cheb_recur_fwd(n_active, nnz, K, rowptr_t, colidx_t, vals_t, G, S, ncols, st);
        if (rc < 0) return rc;
        for (int k = 1; k < K && rc == 0; ++k) {
            float *sk = S + (int64_t)(k - 1) * plane;
            const float *skm1 = (k == 1) ? G : S + (int64_t)(k - 2) * plane;
            const float *skm2 = (k == 1) ? nullptr : (k == 2 ? G : S + (int64_t)(k - 3) * plane);
            int rc2 = launch_spmm(n_active, n_active, rowptr_t, colidx_t, vals_t, skm1, sk, skm2, nullptr, k == 1 ? 1.f : 2.f, -1.f, ncols, st);
            if (rc2) return rc2;
        }
    }
    // weight-gradient branch on a side stream, input-gradient contraction on the main stream.  Deferred mode (step
    // engine): the branch is joined lazily - nothing reads dW before the optimizer - so the contraction that the next
    // layer waits for no longer shares the memory system with it; otherwise forked and joined within this call.
    cudaStream_t wst = st;
    cudaStream_t lazy = lazy_fork(st);
    SideStream *side = lazy ? nullptr : fork_side(st);
    if (lazy) wst = lazy;
    else if (side) wst = side->stream;
    JoinGuard join_guard{side, st};
    struct LazyGuard {
        cudaStream_t s, main;
        ~LazyGuard() { lazy_done(s, main); }
    } lazy_guard{lazy, st};
    // the input-gradient contraction is launched FIRST (the next layer waits for it; in deferred mode the
    // weight-gradient chain behind it then only takes the block slots it leaves free)
    if (rows_act > 0) {      // dX = sum_k S_k W_k^T
        ContractArgs a;
        fill_contract(a);
        a.rows = rows_act;
        a.in_planes = K;
        a.in_w = Fout;
        a.in0 = G;
        a.in_rest = S;
        a.wmat = weight;
        a.w_transposed = 2;
        a.out_planes = 1;
        a.out_w = Fin;
        a.out = dx;
        rc = launch_contract(a, st);
        if (rc) return rc;
    }
    if (rows_in > 0) {       // empty rows: dX = G (sum_k c_k W_k)^T
        rc = launch_fold_dx(rows_in, K, Fin, Fout, G + rows_act * Fout, weight, dx + rows_act * Fin, st);
        if (rc < 0) return rc;
        if (rc == 0) {
        ContractArgs a;
        fill_contract(a);
        a.rows = rows_in;
        a.in_planes = 1;
        a.in_w = Fout;
        a.in0 = G + rows_act * Fout;
        a.wmat = weight;
        a.w_transposed = 1;
        a.w_fold = K;
        a.out_planes = 1;
        a.out_w = Fin;
        a.out = dx + rows_act * Fin;
        rc = launch_contract(a, st);
        if (rc) return rc;
        }
    }
    int nA = 0, m4A = 0, nB = 0, m4B = 0;
    if (rows_act > 0) {      // A = [S_0|..|S_{K-1}]^T x   ([K*Fout, Fin]; dW_k[i][o] = A[k*Fout + o][i])
        WgradArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.rows = rows_act;
        wa.in_planes = K;
        wa.in_w = Fout;
        wa.in0 = G;
        wa.in_rest = S;
        wa.dy = x;
        wa.n_out = Fin;
        wa.partials = partA;
        wa.partial_bytes = partA_bytes;
        // (tuning hook) deferred chain of a level too large for L2: fewer CTAs, so that the HBM-bound kernels of the
        // critical path that follow keep more of the memory system - measured a net loss, off by default
        wa.background = (lazy && rows_act * (int64_t)K * Fout * 4 > (int64_t)64 << 20) ? g_background_div : 0;
        rc = launch_wgrad_partials(wa, 0, &nA, &m4A, wst);
        if (rc) return rc;
    }
    if (rows_in > 0) {       // empty rows of the operator: S = x^T G, dW_k += cos(k pi/2) S
        rc = launch_fold_wgrad(rows_in, Fin, Fout, x + rows_act * Fin, G + rows_act * Fout, partB, partB_bytes, &nB, &m4B, wst);
        if (rc < 0) return rc;
    }
    if (rows_in > 0 && nB == 0) {
        WgradArgs wb;
        memset(&wb, 0, sizeof(wb));
        wb.rows = rows_in;
        wb.in_planes = 1;
        wb.in_w = Fin;
        wb.in0 = x + rows_act * Fin;
        wb.dy = G + rows_act * Fout;
        wb.n_out = Fout;
        wb.partials = partB;
        wb.partial_bytes = partB_bytes;
        rc = launch_wgrad_partials(wb, 0, &nB, &m4B, wst);
        if (rc) return rc;
    }
    rc = launch_wgrad_finalize(partA, nA, m4A, partB, nB, m4B, Fin, K * Fin, Fout, dweight, nullptr, wst, 1);
    if (rc) return rc;
    return MVB_OK;
}

// Make `stream` wait for the LATEST record of `event` at execution time.  With cudaEventWaitExternal the call is
// legal while `stream` is being captured and becomes an external event-wait node: every replay of the graph then
// waits for whatever the host recorded on the event last (the step engine's ground-truth H2D copy).
extern "C" int mvb_stream_wait_external_event(void *stream, void *event) {
    MVB_REQUIRE(event != nullptr, "stream_wait_external_event: null event");
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, cudaEventWaitExternal);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(MVB_ECUDA, "stream_wait_external_event: %s", cudaGetErrorString(e));
    }
    return MVB_OK;
}

// ---------------------------------------------------------------------------------------------
// tuning hooks for A/B runs (scripts/, tests): NOT part of the operator API.  spec = "key=a[,b];key=..."
//   tc_tuning=pg,cap   planes staged at a time by the 6-plane row GEMM (1,2,3,6) / resident CTAs per SM (1..4)
//   tc_balance=0|1     equal tile counts per row-GEMM block
//   layer_tuning=p,c   FFMA mesh-layer backward: blocks per SM, concurrent halves
//   fused_recurrence=v 0 off, 1 on, >= 64 on with that many threads
//   spmm_shape=tx,chunk / spmm_mode=m   SpMM block shape / size (0 = automatic)
//   overlap=0|1        side-stream fork inside mvb_cheb_bwd
//   mesh_tc=e,c        tensor-core mesh layers on/off, CTAs per mesh (0 = automatic)
//   defer_wgrad=0|1    weight-gradient chains of the mesh layers joined lazily (engine only; see mvb_side_join)
//   stream_tc=e,sw     row-streaming fused level-0 layers (mvb_cheb_stream_*) on/off, meshes per slab (8 / 16, 0 = least work per CTA)
//   stream_nt=n        threads per CTA of that kernel (1024 / 768)
//   tc_tma=e,ring      TMA-fed row GEMM for the 16-wide planes on/off, hi tiles in the ring (3 / 4)
//   conv_lanes=n       side streams the deferred convolution weight-gradient chains rotate over (1..7, default 3)
//   pdl=0|1            programmatic dependent launch of the step's kernels (default on; the engine switches it off for N > 1)
//   wgrad_perm=0|1     conflict-free lane order of the staging stores in tc_wgrad (default off)
//   background_div=d   grid of the deferred level-0 weight-gradient reduction: one CTA per d SMs (0 = two per SM)
//   mesh_dbg=bits      timing probes of the tensor-core mesh forward kernel (1 no MMAs, 2 no recurrence, 4 no epilogue):
//                      results are then wrong - scripts/mesh_tc_probe.py only
// ---------------------------------------------------------------------------------------------
extern "C" int mvb_tune(const char *spec) {
    MVB_REQUIRE(spec != nullptr, "mvb_tune: null spec");
    const char *p = spec;
    while (*p) {
        char key[32];
        int n = 0, v[2] = {0, 0}, nv = 0;
        while (*p && *p != '=' && *p != ';' && n < 31) key[n++] = *p++;
        key[n] = 0;
        if (*p == '=') {
            ++p;
            while (nv < 2) {
                char *end = nullptr;
                v[nv] = (int)strtol(p, &end, 10);
                if (end == p) break;
                ++nv;
                p = end;
                if (*p == ',') ++p; else break;
            }
        }
        while (*p && *p != ';') ++p;
        if (*p == ';') ++p;
        if (!strcmp(key, "tc_tuning")) { if (v[0]) set_tc_pg6(v[0]); if (v[1]) set_tc_cap(v[1]); }
        else if (!strcmp(key, "tc_balance")) set_tc_balance(v[0]);
        else if (!strcmp(key, "tc_tma")) set_tc_tma(v[0], v[1]);
        else if (!strcmp(key, "layer_tuning")) set_layer_tuning(v[0], v[1]);
        else if (!strcmp(key, "fused_recurrence")) set_recur_fused(v[0]);
        else if (!strcmp(key, "spmm_shape")) set_spmm_shape(v[0], v[1]);
        else if (!strcmp(key, "spmm_mode")) set_spmm_mode(v[0] & 15);
        else if (!strcmp(key, "overlap")) g_overlap = v[0] ? 1 : 0;
        else if (!strcmp(key, "mesh_tc")) set_mesh_tc(v[0], v[1]);
        else if (!strcmp(key, "mesh_dbg")) set_mesh_dbg(v[0]);
        else if (!strcmp(key, "stream_tc")) set_stream_tc(v[0], nv > 1 ? v[1] : -1);
        else if (!strcmp(key, "stream_nt")) set_stream_nt(v[0]);
        else if (!strcmp(key, "wgrad_perm")) set_wgrad_perm(v[0]);
        else if (!strcmp(key, "conv_lanes")) set_conv_lanes(v[0]);
        else if (!strcmp(key, "pdl")) set_pdl(v[0]);
        else if (!strcmp(key, "defer_wgrad")) set_defer_wgrad(v[0]);
        else if (!strcmp(key, "background_div")) set_background_div(v[0]);
        else if (key[0]) return set_err(MVB_EINVAL, "mvb_tune: unknown key '%s'", key);
    }
    return MVB_OK;
}
