// Hypothesis sampling: minimal sets -> P3P -> 4-point gate, one warp per hypothesis.
//
// Replaces sampleHypotheses (esac_util.h:129-225, called from esac.cpp:112 / 276).  The reference
// loops tries sequentially per hypothesis under `omp parallel for`; here lane l of the warp evaluates
// try 32*round + l of the counter-based stream (esac_rng.cuh), the warp votes, and the lowest passing
// try wins -- the same "first try that passes" the sequential loop returns.
#include "esac_internal.h"
#include "esac_rng.cuh"

namespace esacb200 {

__global__ void __launch_bounds__(128) sample_kernel(const float* __restrict__ coords, const int* __restrict__ assign32,
                                                     Problem P, uint64_t seed, int max_tries,
                                                     const int* __restrict__ injected, int inj_T, Pose* poses, int* cells,
                                                     int* tries) {
    const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (h >= P.M) return;
    const int e = assign32[h];
    const float* pl = coords + (size_t)e * 3 * P.N;
    const int limit = injected ? min(max_tries, inj_T) : max_tries;
    const double f = (double)P.f, ppx = (double)P.ppx, ppy = (double)P.ppy;

    Pose pose;
    int cx[4], cy[4];
    bool found = false;
    for (int base = 0; base < limit && !found; base += 32) {
        const int t = base + lane;
        bool ok = false;
        bool solved = false;
        if (t < limit) {
            if (injected) {
                const int* c = injected + ((size_t)h * inj_T + t) * 8;
                for (int j = 0; j < 4; ++j) { cx[j] = c[2 * j]; cy[j] = c[2 * j + 1]; }
            } else {
                draw_minimal_set(seed, (uint32_t)h, (uint32_t)t, P.W, P.H, cx, cy);
            }
            float obj[4][3], img[4][2];
            for (int j = 0; j < 4; ++j) {
                const int p = cy[j] * P.W + cx[j];
                obj[j][0] = pl[p]; obj[j][1] = pl[P.N + p]; obj[j][2] = pl[2 * (size_t)P.N + p];
                img[j][0] = (float)(cx[j] * P.sub + P.sub / 2 - P.shiftX);
                img[j][1] = (float)(cy[j] * P.sub + P.sub / 2 - P.shiftY);
            }
            solved = p3p_pose(obj, img, f, ppx, ppy, pose);
            if (solved) ok = minimal_set_gate(obj, img, pose, f, ppx, ppy, P.tau);
            if (!solved) { for (int c = 0; c < 3; ++c) { pose.r[c] = 0; pose.t[c] = 0; } }  // safeSolvePnP failure state
        }
        const unsigned vote = __ballot_sync(0xffffffffu, ok);
        int src = -1;
        if (vote) { src = __ffs(vote) - 1; found = true; }
        else if (base + 32 >= limit) src = (limit - 1) - base;  // exhausted: the state of the last try survives
        if (src >= 0 && lane == src) {
            poses[h] = pose;
            for (int j = 0; j < 4; ++j) { cells[h * 8 + 2 * j] = cx[j]; cells[h * 8 + 2 * j + 1] = cy[j]; }
            tries[h] = t + 1;
        }
    }
}

void launch_sample(const float* coords, const int* assign32, const Problem& P, uint64_t seed, int max_tries,
                   const int* injected, int inj_T, Pose* poses, int* cells, int* tries, cudaStream_t st) {
    const int warps_per_block = 4;
    sample_kernel<<<(P.M + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        coords, assign32, P, seed, max_tries, injected, inj_T, poses, cells, tries);
}

}  // namespace esacb200
