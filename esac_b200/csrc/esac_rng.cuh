// Counter-based minimal-set sampling stream.
//
// Replaces the reference's per-OpenMP-thread std::mt19937 (thread_rand.cpp:13-30, 68-71), whose
// stream depends on the OpenMP schedule and cannot be re-seeded from Python.  A try is addressed
// as (seed, hypothesis, try) so the 32 lanes of a warp can evaluate 32 tries of one hypothesis
// at once and still agree, bit for bit, with the sequential oracle (oracle/esac_oracle.py:
// mix64 / try_state / cell_draw).  The distribution is the reference's: x in [0, W-2],
// y in [0, H-2] (irand(0, imW-1) -> uniform_int(0, imW-2), esac_util.h:167-168), 4 distinct
// cells, duplicates re-drawn (esac_util.h:170-174).
#pragma once
#include <stdint.h>

#ifndef ESAC_HD
#ifdef __CUDACC__
#define ESAC_HD __host__ __device__ __forceinline__
#else
#define ESAC_HD inline
#endif
#endif

namespace esacb200 {

constexpr uint64_t kGold = 0x9E3779B97F4A7C15ull;

ESAC_HD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

ESAC_HD uint64_t try_state(uint64_t seed, uint32_t h, uint32_t t) {
    uint64_t s = mix64(seed + kGold * (uint64_t)(h + 1u));
    return mix64(s + kGold * (uint64_t)(t + 1u));
}

ESAC_HD void cell_draw(uint64_t state, uint32_t k, int W, int H, int& x, int& y) {
    uint64_t r = mix64(state + kGold * (uint64_t)(k + 1u));
    x = (int)(((r & 0xFFFFFFFFull) * (uint64_t)(W - 1)) >> 32);
    y = (int)(((r >> 32) * (uint64_t)(H - 1)) >> 32);
}

// 4 distinct cells of try (seed, h, t).  cx/cy receive the cell coordinates.
ESAC_HD void draw_minimal_set(uint64_t seed, uint32_t h, uint32_t t, int W, int H, int cx[4], int cy[4]) {
    uint64_t st = try_state(seed, h, t);
    int n = 0;
    uint32_t k = 0;
    while (n < 4) {
        int x, y;
        cell_draw(st, k++, W, H, x, y);
        bool dup = false;
        for (int i = 0; i < n; ++i) dup = dup || (cx[i] == x && cy[i] == y);
        if (dup) continue;
        cx[n] = x;
        cy[n] = y;
        ++n;
    }
}

}  // namespace esacb200
