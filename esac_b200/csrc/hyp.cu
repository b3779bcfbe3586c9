// Hypothesis sampling: minimal sets -> P3P -> 4-point gate, one CTA per hypothesis.
//
// Replaces sampleHypotheses (esac_util.h:129-225, called from esac.cpp:112 / 276).  The reference loops
// tries sequentially per hypothesis under `omp parallel for`; the counter-based stream (esac_rng.cuh) makes
// every try addressable, so a CTA evaluates a contiguous block of tries at once and keeps the LOWEST passing
// try -- exactly the try the sequential loop would have stopped at.  Maps whose experts are wrong need
// ~1/P(4th point lands within tau) ~ 1e3 tries per hypothesis (the dominant cost of a step), so each try
// first goes through a float prefilter (p3p_may_pass) that discards the >98% of tries whose every P3P root
// misses the 4th point by more than 4 tau; only the survivors are compacted and run through the exact
// fp64 path (p3p_pose + minimal_set_gate) whose verdict is the only one that counts.
#include "esac_internal.h"
#include "esac_rng.cuh"

namespace esacb200 {

constexpr int kSampleThreads = 128;
constexpr int kSampleMaxK = 8;  // tries per thread per super-round

__device__ __forceinline__ void load_try(const float* __restrict__ pl, const Problem& P, const int* injected, int inj_T,
                                         uint64_t seed, int h, int t, int cx[4], int cy[4], float obj[4][3], float img[4][2]) {
    if (injected) {
        const int* c = injected + ((size_t)h * inj_T + t) * 8;
        for (int j = 0; j < 4; ++j) { cx[j] = c[2 * j]; cy[j] = c[2 * j + 1]; }
    } else {
        draw_minimal_set(seed, (uint32_t)h, (uint32_t)t, P.W, P.H, cx, cy);
    }
    for (int j = 0; j < 4; ++j) {
        const int p = cy[j] * P.W + cx[j];
        obj[j][0] = pl[p]; obj[j][1] = pl[P.N + p]; obj[j][2] = pl[2 * (size_t)P.N + p];
        img[j][0] = (float)(cx[j] * P.sub + P.sub / 2 - P.shiftX);
        img[j][1] = (float)(cy[j] * P.sub + P.sub / 2 - P.shiftY);
    }
}

__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const float* __restrict__ coords, const int* __restrict__ assign32,
                                                                Problem P, uint64_t seed, int max_tries,
                                                                const int* __restrict__ injected, int inj_T, Pose* poses,
                                                                int* cells, int* tries) {
    const int h = blockIdx.x;
    const int tid = threadIdx.x;
    const int e = assign32[h];
    const float* pl = coords + (size_t)e * 3 * P.N;
    const int limit = injected ? min(max_tries, inj_T) : max_tries;
    const double f = (double)P.f, ppx = (double)P.ppx, ppy = (double)P.ppy;
    __shared__ int s_list[kSampleThreads * kSampleMaxK];
    __shared__ int s_n, s_best;
    if (tid == 0) { s_best = 0x7fffffff; }
    int base = 0;
    int K = 1;  // first super-round: one try per thread (easy maps finish here), then kSampleMaxK
    while (base < limit) {
        if (tid == 0) s_n = 0;
        __syncthreads();
        const int span = min(limit - base, K * kSampleThreads);
        // ---- float prefilter over [base, base + span) ----
        for (int t = base + tid; t < base + span; t += kSampleThreads) {
            int cx[4], cy[4];
            float obj[4][3], img[4][2];
            load_try(pl, P, injected, inj_T, seed, h, t, cx, cy, obj, img);
            if (p3p_may_pass(obj, img, P.f, P.ppx, P.ppy, P.tau)) s_list[atomicAdd(&s_n, 1)] = t;
        }
        __syncthreads();
        // ---- exact path on the survivors ----
        const int n = s_n;
        for (int i = tid; i < n; i += kSampleThreads) {
            const int t = s_list[i];
            int cx[4], cy[4];
            float obj[4][3], img[4][2];
            load_try(pl, P, injected, inj_T, seed, h, t, cx, cy, obj, img);
            Pose pose;
            if (p3p_pose(obj, img, f, ppx, ppy, pose) && minimal_set_gate(obj, img, pose, f, ppx, ppy, P.tau)) atomicMin(&s_best, t);
        }
        __syncthreads();
        if (s_best != 0x7fffffff) break;
        base += span;
        K = kSampleMaxK;
    }
    // one thread re-derives the winning try (or, when exhausted, the last try whose state survives in the reference)
    if (tid == 0) {
        const bool found = s_best != 0x7fffffff;
        const int t = found ? s_best : limit - 1;
        int cx[4], cy[4];
        float obj[4][3], img[4][2];
        load_try(pl, P, injected, inj_T, seed, h, t, cx, cy, obj, img);
        Pose pose;
        if (!p3p_pose(obj, img, f, ppx, ppy, pose)) {
            for (int c = 0; c < 3; ++c) { pose.r[c] = 0; pose.t[c] = 0; }  // safeSolvePnP failure state
        }
        poses[h] = pose;
        for (int j = 0; j < 4; ++j) { cells[h * 8 + 2 * j] = cx[j]; cells[h * 8 + 2 * j + 1] = cy[j]; }
        tries[h] = t + 1;
    }
}

void launch_sample(const float* coords, const int* assign32, const Problem& P, uint64_t seed, int max_tries,
                   const int* injected, int inj_T, Pose* poses, int* cells, int* tries, cudaStream_t st) {
    sample_kernel<<<P.M, kSampleThreads, 0, st>>>(coords, assign32, P, seed, max_tries, injected, inj_T, poses, cells, tries);
}

}  // namespace esacb200
