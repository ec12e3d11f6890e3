// Fused Adam over one flat fp32 parameter buffer (main.py:251, :81).
// One streaming pass: reads p, g, m, v and writes p, m, v (28 B/parameter) - HBM-bound; the step
// counter lives on the device so the launch is CUDA-graph capturable.
#include "mvb_internal.cuh"

namespace mvb {

__global__ void adam_tick_kernel(int64_t *step) { *step += 1; }

__global__ void __launch_bounds__(256)
adam_kernel(int64_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, const int64_t *__restrict__ step, float lr, float beta1,
            float beta2, float eps, float wd, float gscale) {
    const float t = (float)(*step);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2 = 1.f - powf(beta2, t);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(wd, pi, g[i] * gscale);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

// the same update with the hyper-parameters read from DEVICE memory: hp = {lr, beta1, beta2, eps, weight_decay, grad_scale}.
// A captured CUDA graph then follows a learning-rate schedule (main.py:266-269 rewrites param_groups[...]['lr'] per
// epoch) or a restored optimizer state without being re-captured.
__global__ void __launch_bounds__(256)
adam_hp_kernel(int64_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
               float *__restrict__ v, const int64_t *__restrict__ step, const float *__restrict__ hp) {
    pdl_wait();          // (no trigger: nothing may start early next to the kernel that rewrites the weights)
    const float lr = hp[0], beta1 = hp[1], beta2 = hp[2], eps = hp[3], wd = hp[4], gscale = hp[5];
    const float t = (float)(*step);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2 = 1.f - powf(beta2, t);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(wd, pi, g[i] * gscale);
        const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
        const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

}  // namespace mvb

using namespace mvb;

extern "C" int mvb_adam_step_hp(int64_t n, float *p, const float *g, float *m, float *v, int64_t *step,
                                const float *hyper, void *stream) {
    MVB_REQUIRE(n >= 0 && p && g && m && v && step && hyper, "adam_step_hp: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(step);
    int rc = check_launch("mvb_adam_step_hp tick");
    if (rc || n == 0) return rc;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(adam_hp_kernel, dim3((unsigned)blocks), dim3(256), 0, st, n, p, g, m, v, step, hyper);
    return check_launch("mvb_adam_step_hp");
}


extern "C" int mvb_adam_step(int64_t n, float *p, const float *g, float *m, float *v, int64_t *step,
                             float lr, float beta1, float beta2, float eps, float weight_decay,
                             float grad_scale, void *stream) {
    MVB_REQUIRE(n >= 0 && p && g && m && v && step, "adam_step: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(step);
    int rc = check_launch("mvb_adam_step tick");
    if (rc || n == 0) return rc;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(n, p, g, m, v, step, lr, beta1, beta2, eps, weight_decay, grad_scale);
    return check_launch("mvb_adam_step");
}
