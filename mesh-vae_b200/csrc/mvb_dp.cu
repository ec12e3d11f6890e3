// Data-parallel gradient exchange fused with the optimizer (new - the reference is single device, main.py:194-195, :251, :81).
//
// Every rank keeps its flat gradient buffer in memory that its peers have mapped (NVLink / NVSwitch peer access:
// torch symmetric memory or CUDA IPC - the pointers are handed in, the library does not care which).  ONE launch at the
// end of the backward pass then replaces "all-reduce, wait, Adam":
//
//   1. ready:   CTA 0 tells every peer "my gradient of epoch e is complete" (release store into the peer's signal pad);
//               every CTA waits until all peers have said so (acquire loads from the LOCAL pad);
//   2. reduce:  g[i] = sum_r grads[r][i] in rank order r = 0..world-1 - the same order on every rank, so all ranks compute
//               bit-identical sums (no reduction tree, no atomics) - read straight from the peers' buffers, 16 bytes per
//               thread per peer; Adam (torch.optim.Adam's arithmetic, as adam_hp_kernel) on the sum in the same pass:
//               every rank updates its full replica of p / m / v, nothing is sent back;
//   3. done:    the last CTA tells every peer "I have finished reading your buffer".  The wait for THAT signal is at the
//               start of the next step (mvb_dp_begin, one tiny launch at the head of the step graph): by then - a whole
//               forward pass later - it has long arrived, so the second synchronisation of an all-reduce costs nothing.
//
// Epochs instead of flags that are reset: signal words only ever grow, so a captured CUDA graph can be replayed with
// identical arguments.  2.85 MB of gradients for cheb_VAE: one-shot reads (world-1) x 2.85 MB per rank, ~5 us at 2 GPUs
// and ~30 us at 8 - against two NCCL all-reduce launches plus the Adam launch before.
#include "mvb_internal.cuh"

namespace mvb {

constexpr int DP_MAX_WORLD = 16;
constexpr int DP_CHANNELS = 2;       // exchanges that may be in flight within one step (gradient buckets)
// uint32 index inside a 256-byte signal pad: channel c, ready[r] at c * 32 + r, done[r] at c * 32 + 16 + r
__host__ __device__ constexpr int dp_ready(int c) { return c * 32; }
__host__ __device__ constexpr int dp_done(int c) { return c * 32 + 16; }

struct DpPtrs {
    const float *grads[DP_MAX_WORLD];
    unsigned int *pads[DP_MAX_WORLD];
};
struct DpState {                     // local device memory, zero at creation
    unsigned int epoch;
    unsigned int timeouts;           // waits given up after DP_TIMEOUT_CYCLES (a peer died or never launched): results are then invalid
    unsigned int cta_done[DP_CHANNELS];
};
constexpr long long DP_TIMEOUT_CYCLES = 60000000000ll;       // ~30 s at 2 GHz

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float *p) {        // never from a stale L1 line: the peers rewrite it every step
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// spin until *flag has reached `target` (epoch numbers: wrap-safe signed difference); bounded, so that a lost peer turns
// into an error the host can read (DpState::timeouts) instead of a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned int *flag, unsigned int target, DpState *state) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - target) < 0) {
        if (clock64() - t0 > DP_TIMEOUT_CYCLES) {
            atomicAdd(&state->timeouts, 1u);
            break;
        }
    }
}

__global__ void dp_tick_kernel(int64_t *step) { *step += 1; }

// head of a step: a new epoch, once every peer has finished reading this rank's gradient buffer of the previous one
__global__ void dp_begin_kernel(int world, int rank, int channels, DpPtrs ptrs, DpState *state) {
    const unsigned int e = state->epoch + 1;
    const int r = threadIdx.x;
    if (r < world) {
        for (int c = 0; c < channels; ++c) wait_flag(ptrs.pads[rank] + dp_done(c) + r, e - 1, state);
    }
    __syncthreads();
    if (r == 0) state->epoch = e;
}

__device__ __forceinline__ float adam_one(float pi, float gsum, float &mi, float &vi, float beta1, float beta2, float eps, float wd,
                                          float gscale, float step_size, float inv_sqrt_bc2) {
    const float gi = fmaf(wd, pi, gsum * gscale);
    mi = fmaf(beta1, mi, (1.f - beta1) * gi);
    vi = fmaf(beta2, vi, (1.f - beta2) * gi * gi);
    return pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
}

// elements [off, off + n) of the flat buffers; off and n multiples of 4 (the buffers are padded to 32 elements per parameter)
__global__ void __launch_bounds__(256)
dp_reduce_adam_kernel(int world, int rank, int channel, int64_t off, int64_t n, float *__restrict__ p, float *__restrict__ m,
                      float *__restrict__ v, float *gsum_out, const int64_t *__restrict__ step, const float *__restrict__ hp,
                      DpPtrs ptrs, DpState *state) {
    const unsigned int e = state->epoch;
    const int tid = threadIdx.x;
    // ---- 1. ready ----
    if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();
        st_release_sys(ptrs.pads[tid] + dp_ready(channel) + rank, e);
    }
    if (tid < world) wait_flag(ptrs.pads[rank] + dp_ready(channel) + tid, e, state);
    __syncthreads();
    // ---- 2. reduce in rank order + Adam ----
    const float lr = hp[0], beta1 = hp[1], beta2 = hp[2], eps = hp[3], wd = hp[4], gscale = hp[5];
    const float t = (float)(*step);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2 = 1.f - powf(beta2, t);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const int64_t i0 = off >> 2, n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e4 = 4 * (i0 + i);
        // all peers' quads in flight at once (the reads cross NVLink: latency, not arithmetic, paces this loop), then the
        // sum in rank order
        // eight peers at a time (keeps the kernel at ~64 registers: it runs next to the encoder backward)
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < world; r0 += 8) {
            float4 gr[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < world) gr[r] = ld_peer4(ptrs.grads[r0 + r] + e4);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r0 + r < world) {
                    if (r0 + r == 0) g = gr[0];
                    else { g.x += gr[r].x; g.y += gr[r].y; g.z += gr[r].z; g.w += gr[r].w; }
                }
        }
        float4 pi = *reinterpret_cast<const float4 *>(p + e4);
        float4 mi = *reinterpret_cast<const float4 *>(m + e4);
        float4 vi = *reinterpret_cast<const float4 *>(v + e4);
        pi.x = adam_one(pi.x, g.x, mi.x, vi.x, beta1, beta2, eps, wd, gscale, step_size, inv_sqrt_bc2);
        pi.y = adam_one(pi.y, g.y, mi.y, vi.y, beta1, beta2, eps, wd, gscale, step_size, inv_sqrt_bc2);
        pi.z = adam_one(pi.z, g.z, mi.z, vi.z, beta1, beta2, eps, wd, gscale, step_size, inv_sqrt_bc2);
        pi.w = adam_one(pi.w, g.w, mi.w, vi.w, beta1, beta2, eps, wd, gscale, step_size, inv_sqrt_bc2);
        *reinterpret_cast<float4 *>(p + e4) = pi;
        *reinterpret_cast<float4 *>(m + e4) = mi;
        *reinterpret_cast<float4 *>(v + e4) = vi;
        if (gsum_out) *reinterpret_cast<float4 *>(gsum_out + e4) = g;
    }
    // ---- 3. done: the last CTA of this rank releases every peer's buffer ----
    __syncthreads();
    __shared__ unsigned int last;
    if (tid == 0) {
        __threadfence();
        last = atomicAdd(&state->cta_done[channel], 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (last) {
        if (tid == 0) state->cta_done[channel] = 0;
        if (tid < world) {
            __threadfence_system();
            st_release_sys(ptrs.pads[tid] + dp_done(channel) + rank, e);
        }
    }
}

static int fill_ptrs(DpPtrs &pp, int world, void *const *grads, void *const *pads) {
    memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < world; ++r) {
        if ((grads && !grads[r]) || !pads[r]) return set_err(MVB_EINVAL, "dp: null peer pointer for rank %d", r);
        if (grads) pp.grads[r] = reinterpret_cast<const float *>(grads[r]);
        pp.pads[r] = reinterpret_cast<unsigned int *>(pads[r]);
    }
    return MVB_OK;
}

}  // namespace mvb

using namespace mvb;

extern "C" int mvb_dp_max_world(void) { return DP_MAX_WORLD; }
extern "C" size_t mvb_dp_pad_bytes(void) { return 256; }
extern "C" size_t mvb_dp_state_bytes(void) { return sizeof(DpState); }      // uint32 words: [0] epoch, [1] timeouts, (internal)

extern "C" int mvb_dp_begin(int world, int rank, int channels, void *const *pads, void *state, void *stream) {
    MVB_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world && pads && state, "dp_begin: bad arguments");
    MVB_REQUIRE(channels >= 1 && channels <= DP_CHANNELS, "dp_begin: channels=%d outside [1,%d]", channels, DP_CHANNELS);
    DpPtrs pp;
    int rc = fill_ptrs(pp, world, nullptr, pads);
    if (rc) return rc;
    dp_begin_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(world, rank, channels, pp, reinterpret_cast<DpState *>(state));
    return check_launch("mvb_dp_begin");
}

extern "C" int mvb_dp_reduce_adam(int world, int rank, int channel, int64_t offset, int64_t n, float *p, void *const *grads, float *m,
                                  float *v, float *grad_sum_out, int64_t *step, int tick, const float *hyper, void *const *pads,
                                  void *state, int max_ctas, void *stream) {
    MVB_REQUIRE(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "dp_reduce_adam: world=%d rank=%d", world, rank);
    MVB_REQUIRE(channel >= 0 && channel < DP_CHANNELS, "dp_reduce_adam: channel=%d outside [0,%d)", channel, DP_CHANNELS);
    MVB_REQUIRE(offset >= 0 && offset % 4 == 0 && n >= 0 && n % 4 == 0 && p && grads && m && v && step && hyper && pads && state,
                "dp_reduce_adam: bad arguments");
    DpPtrs pp;
    int rc = fill_ptrs(pp, world, grads, pads);
    if (rc) return rc;
    for (int r = 0; r < world; ++r)
        if (!aligned16(grads[r])) return set_err(MVB_EALIGN, "dp_reduce_adam: gradient buffer of rank %d is not 16-byte aligned", r);
    if (!aligned16(p) || !aligned16(m) || !aligned16(v) || (grad_sum_out && !aligned16(grad_sum_out)))
        return set_err(MVB_EALIGN, "dp_reduce_adam: p / m / v must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (tick) {
        dp_tick_kernel<<<1, 1, 0, st>>>(step);
        rc = check_launch("mvb_dp_reduce_adam tick");
        if (rc) return rc;
    }
    // every CTA waits for the peers' signals before it works: at most one wave of them
    int64_t blocks = (n / 4 + 255) / 256;
    int64_t cap = (int64_t)num_sms() * 4;
    if (max_ctas > 0 && max_ctas < cap) cap = max_ctas;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dp_reduce_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(world, rank, channel, offset, n, p, m, v, grad_sum_out, step, hyper, pp,
                                                          reinterpret_cast<DpState *>(state));
    return check_launch("mvb_dp_reduce_adam");
}
