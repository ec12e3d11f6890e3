// CSR SpMM over vertex-major activations: the B200 replacement of the reference's
// gather / scale / scatter_add propagate (nn/conv.py:199-200, :579-581, :363-364) and of the
// recurrence arithmetic `2 * prop - Tx_0` (nn/conv.py:569).
//
//   y[r,:] = alpha * sum_j vals[j] * x[colidx[j],:] + beta * z[r,:] + w[r,:]
//
// One thread owns one 16-byte column quad of one output row; consecutive threads walk the
// B*F columns of a row first, so every neighbour gather is a run of fully coalesced 512-byte
// warp requests.  Nothing is materialised per edge (the reference materialises [E,B,F]).
// Roofline: HBM/L2 bandwidth; algorithmic bytes = (2 or 3)*N*ncols*4 + CSR.
#include "mvb_internal.cuh"

namespace mvb {

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

__device__ __forceinline__ void fma4(float4 &acc, float v, const float4 &x) {
    acc.x = fmaf(v, x.x, acc.x);
    acc.y = fmaf(v, x.y, acc.y);
    acc.z = fmaf(v, x.z, acc.z);
    acc.w = fmaf(v, x.w, acc.w);
}

template <bool HAS_Z, bool HAS_W>
__global__ void __launch_bounds__(256)
spmm_v4_kernel(int n_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
               const float *__restrict__ vals, const float4 *__restrict__ x, float4 *y,
               const float4 *z, const float4 *w, float alpha, float beta, int nc4) {
    const int64_t total = (int64_t)n_rows * nc4;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / nc4);
        const int c = (int)(idx - (int64_t)r * nc4);
        const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = s;
        // 4 independent gathers in flight per thread (mean degree is 6)
        for (; j + 4 <= e; j += 4) {
            const int c0 = __ldg(colidx + j), c1 = __ldg(colidx + j + 1);
            const int c2 = __ldg(colidx + j + 2), c3 = __ldg(colidx + j + 3);
            const float v0 = __ldg(vals + j), v1 = __ldg(vals + j + 1);
            const float v2 = __ldg(vals + j + 2), v3 = __ldg(vals + j + 3);
            const float4 x0 = ldg4(x + (int64_t)c0 * nc4 + c);
            const float4 x1 = ldg4(x + (int64_t)c1 * nc4 + c);
            const float4 x2 = ldg4(x + (int64_t)c2 * nc4 + c);
            const float4 x3 = ldg4(x + (int64_t)c3 * nc4 + c);
            fma4(acc, v0, x0);
            fma4(acc, v1, x1);
            fma4(acc, v2, x2);
            fma4(acc, v3, x3);
        }
        for (; j < e; ++j) {
            const int c0 = __ldg(colidx + j);
            const float v0 = __ldg(vals + j);
            fma4(acc, v0, ldg4(x + (int64_t)c0 * nc4 + c));
        }
        float4 o = make_float4(alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w);
        if (HAS_Z) {
            const float4 zz = z[idx];
            o.x = fmaf(beta, zz.x, o.x);
            o.y = fmaf(beta, zz.y, o.y);
            o.z = fmaf(beta, zz.z, o.z);
            o.w = fmaf(beta, zz.w, o.w);
        }
        if (HAS_W) {
            const float4 ww = w[idx];
            o.x += ww.x;
            o.y += ww.y;
            o.z += ww.z;
            o.w += ww.w;
        }
        y[idx] = o;
    }
}

__global__ void __launch_bounds__(256)
spmm_scalar_kernel(int n_rows, const int32_t *__restrict__ rowptr,
                   const int32_t *__restrict__ colidx, const float *__restrict__ vals,
                   const float *__restrict__ x, float *y, const float *z, const float *w,
                   float alpha, float beta, int64_t ncols) {
    const int64_t total = (int64_t)n_rows * ncols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / ncols);
        const int64_t c = idx - (int64_t)r * ncols;
        const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
        float acc = 0.f;
        for (int j = s; j < e; ++j)
            acc = fmaf(__ldg(vals + j), __ldg(x + (int64_t)__ldg(colidx + j) * ncols + c), acc);
        float o = alpha * acc;
        if (z) o = fmaf(beta, z[idx], o);
        if (w) o += w[idx];
        y[idx] = o;
    }
}

int launch_spmm(int n_rows, const int32_t *rowptr, const int32_t *colidx, const float *vals,
                const float *x, float *y, const float *z, const float *w, float alpha, float beta,
                int64_t ncols, cudaStream_t st) {
    if (n_rows == 0 || ncols == 0) return MVB_OK;
    const bool vec = (ncols % 4 == 0) && aligned16(x) && aligned16(y) && (!z || aligned16(z)) &&
                     (!w || aligned16(w));
    const int threads = 256;
    // enough CTAs for every SM to hold its full complement of resident warps; grid-stride beyond
    const int64_t max_blocks = (int64_t)num_sms() * 8 * 4;
    if (vec) {
        const int nc4 = (int)(ncols / 4);
        int64_t blocks = ((int64_t)n_rows * nc4 + threads - 1) / threads;
        if (blocks > max_blocks) blocks = max_blocks;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        float4 *y4 = reinterpret_cast<float4 *>(y);
        const float4 *z4 = reinterpret_cast<const float4 *>(z);
        const float4 *w4 = reinterpret_cast<const float4 *>(w);
        if (z && w)
            spmm_v4_kernel<true, true><<<(unsigned)blocks, threads, 0, st>>>(n_rows, rowptr, colidx, vals, x4, y4, z4, w4, alpha, beta, nc4);
        else if (z)
            spmm_v4_kernel<true, false><<<(unsigned)blocks, threads, 0, st>>>(n_rows, rowptr, colidx, vals, x4, y4, z4, w4, alpha, beta, nc4);
        else if (w)
            spmm_v4_kernel<false, true><<<(unsigned)blocks, threads, 0, st>>>(n_rows, rowptr, colidx, vals, x4, y4, z4, w4, alpha, beta, nc4);
        else
            spmm_v4_kernel<false, false><<<(unsigned)blocks, threads, 0, st>>>(n_rows, rowptr, colidx, vals, x4, y4, z4, w4, alpha, beta, nc4);
    } else {
        int64_t blocks = ((int64_t)n_rows * ncols + threads - 1) / threads;
        if (blocks > max_blocks) blocks = max_blocks;
        spmm_scalar_kernel<<<(unsigned)blocks, threads, 0, st>>>(n_rows, rowptr, colidx, vals, x, y, z, w, alpha, beta, ncols);
    }
    return check_launch("mvb_spmm");
}

}  // namespace mvb
