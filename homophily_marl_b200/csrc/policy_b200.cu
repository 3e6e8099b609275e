// policy_b200.cu -- sm_100a kernels on the policy side of the rollout loop, behind the same C ABI.
//
//   ssd_select_actions : EpsilonGreedyActionSelector.select_action (src/components/action_selectors.py:44-68)
//
// Reference citations are relative to drdh/Homophily-MARL.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "ssd_b200.h"

namespace {

thread_local int g_policy_cuda_error = 0;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// One thread per row of Q-values (n_actions <= 32: 9 / 8 env actions, 3 incentive actions).
//   masked_q[avail == 0] = -inf; greedy = FIRST index of the maximum (torch.max(dim=-1)[1]);
//   pick_random = u_pick < epsilon (fp32 compare, as torch compares a float32 tensor with a Python scalar);
//   random action = the k-th available action, k = floor(u_act * #available)  (th.multinomial over the 0/1 mask is a
//   uniform choice among the available actions; the position-indexed draw makes it injectable).
__global__ void select_actions_kernel(const float* __restrict__ q, const int32_t* __restrict__ avail, long long rows, int A,
                                      float eps, const float* __restrict__ u_pick, const float* __restrict__ u_act,
                                      uint32_t seed_lo, uint32_t seed_hi, uint32_t ctr_lo, uint32_t ctr_hi,
                                      long long* __restrict__ picked) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* qr = q + row * A;
    const int32_t* ar = avail ? avail + row * A : nullptr;
    float best = -INFINITY;
    int arg = 0, n_av = 0;
    uint32_t mask = 0;
    for (int a = 0; a < A; ++a) {
        const bool ok = !ar || ar[a] != 0;
        const float v = ok ? qr[a] : -INFINITY;
        if (ok) { mask |= 1u << a; ++n_av; }
        if (v > best) { best = v; arg = a; }
    }
    float up, ua;
    if (u_pick) { up = u_pick[row]; ua = u_act[row]; }
    else {
        const uint4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), ctr_lo, ctr_hi, seed_lo, seed_hi);
        up = (float)(r.x >> 8) * 5.9604644775390625e-8f;       // 24-bit uniform in [0, 1)
        ua = (float)(r.y >> 8) * 5.9604644775390625e-8f;
    }
    int out = arg;
    if (up < eps && n_av > 0) {
        int k = (int)(ua * (float)n_av);
        if (k >= n_av) k = n_av - 1;
        uint32_t m = mask;
        for (int i = 0; i < k; ++i) m &= m - 1;                // drop the k lowest available actions
        out = __ffs(m) - 1;
    }
    picked[row] = out;
}

}  // namespace

extern "C" {

int ssd_select_actions(const float* q, const int32_t* avail, int64_t rows, int32_t n_actions, float epsilon,
                       const float* u_pick, const float* u_act, uint64_t seed, uint64_t counter,
                       int64_t* picked, void* stream) {
    if (!q || !picked || rows < 0 || n_actions < 1 || n_actions > 32) return SSD_ERR_INVALID;
    if ((u_pick == nullptr) != (u_act == nullptr)) return SSD_ERR_INVALID;
    if (rows == 0) return SSD_OK;
    const int threads = 128;
    select_actions_kernel<<<(unsigned)((rows + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        q, avail, rows, n_actions, epsilon, u_pick, u_act, (uint32_t)seed, (uint32_t)(seed >> 32),
        (uint32_t)counter, (uint32_t)(counter >> 32), reinterpret_cast<long long*>(picked));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_policy_cuda_error = (int)e; return SSD_ERR_CUDA; }
    return SSD_OK;
}

int ssd_policy_last_cuda_error(void) { return g_policy_cuda_error; }

}  // extern "C"
