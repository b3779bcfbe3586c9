// Hypothesis assignment on the device.
//
// The reference's callers draw the expert of every hypothesis on the host from the gating distribution and build the
// per-expert histogram that decides which experts run and scales the gating gradient:
//   util.clamp_probs(gating_probs[0], maxexperts)                       util.py:38-48, train_esac.py:130
//   e_hyps = torch.multinomial(gating_probs[0], hypotheses, True)       train_esac.py:133-137, test_esac.py:169-174
//   e_hyps_hist = torch.histc(e_hyps.float(), bins=E, min=0, max=E-1)   train_esac.py:140,  test_esac.py:177
// Here one CTA per image does the three steps without leaving the GPU, so a batch of gating outputs turns into the
// [B, M] assignment that esacb200_forward_batch / esacb200_backward_batch consume.  The draws come from the same
// counter-based generator as the minimal sets (esac_rng.cuh), keyed (seed, image, hypothesis): the oracle
// (oracle/esac_oracle.py: assign_hypotheses) reproduces them bit for bit.  torch.multinomial's own stream is not
// reproduced (it is a property of torch's Philox/mt19937 state, not of the algorithm); the distribution is.
#include "esac_internal.h"
#include "esac_rng.cuh"

namespace esacb200 {

namespace {

constexpr int kMaxExperts = 1024;

// weights [B, E] (>= 0, need not sum to 1: multinomial normalises), out_assign [B, M] int64, out_hist [B, E] float.
// keep_top < 0: no clamping; else all but the keep_top largest weights are zeroed first (ties: the later index wins a place,
// as a stable ascending sort would leave it nearer the top).  single != 0: one draw per image, repeated M times
// (the "expertselection" mode, train_esac.py:133-135).
__global__ void __launch_bounds__(256) assign_kernel(const float* __restrict__ weights, int E, int M, int keep_top, int single,
                                                     uint64_t seed, int64_t* __restrict__ out_assign,
                                                     float* __restrict__ out_hist, int* __restrict__ flags) {
    __shared__ float w[kMaxExperts];
    __shared__ double cdf[kMaxExperts];
    __shared__ int hist[kMaxExperts];
    __shared__ int last_pos;
    const int b = blockIdx.x;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        w[e] = weights[(size_t)b * E + e];
        hist[e] = 0;
    }
    __syncthreads();
    if (keep_top >= 0 && keep_top < E) {
        // rank of entry e in the stable ascending order = #{j: w[j] < w[e]} + #{j < e: w[j] == w[e]}
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const float we = w[e];
            int rank = 0;
            for (int j = 0; j < E; ++j) rank += (w[j] < we) || (w[j] == we && j < e);
            if (rank < E - keep_top) cdf[e] = 0.;  // remember the verdict; w is still being read by other threads
            else cdf[e] = 1.;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < E; e += blockDim.x)
            if (cdf[e] == 0.) w[e] = 0.f;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double acc = 0.;
        int lp = -1;
        bool bad = false;
        for (int e = 0; e < E; ++e) {
            const float v = w[e];
            if (!(v >= 0.f) || isinf(v)) bad = true;  // torch.multinomial: "probability tensor contains either inf, nan or element < 0"
            if (v > 0.f) { acc += (double)v; lp = e; }
            cdf[e] = acc;
        }
        if (bad || lp < 0) atomicOr(flags, bad ? 1 : 2);  // 2: "invalid multinomial distribution (sum of probabilities <= 0)"
        last_pos = lp;
    }
    __syncthreads();
    const int lp = last_pos;
    const double total = lp >= 0 ? cdf[E - 1] : 0.;
    for (int h = threadIdx.x; h < M; h += blockDim.x) {
        const uint32_t k = single ? 0u : (uint32_t)h;
        const uint64_t r = try_state(seed, (uint32_t)b, k);
        const double u = (double)(r >> 11) * 0x1.0p-53 * total;
        // first expert whose cumulative weight exceeds u (experts of zero weight can never be it)
        int lo = 0, hi = E - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid;
            else lo = mid + 1;
        }
        int e = lo;
        if (lp >= 0 && !(cdf[e] > u)) e = lp;  // u rounded up to the total
        if (lp < 0) e = 0;
        out_assign[(size_t)b * M + h] = e;
        atomicAdd(&hist[e], 1);
    }
    __syncthreads();
    if (out_hist)
        for (int e = threadIdx.x; e < E; e += blockDim.x) out_hist[(size_t)b * E + e] = (float)hist[e];
}

}  // namespace

int assign_max_experts() { return kMaxExperts; }

void launch_assign(const float* weights, int B, int E, int M, int keep_top, int single, uint64_t seed, int64_t* out_assign,
                   float* out_hist, int* flags, cudaStream_t stream) {
    assign_kernel<<<B, 256, 0, stream>>>(weights, E, M, keep_top, single, seed, out_assign, out_hist, flags);
}

}  // namespace esacb200
