// Pose refinement: iterated inlier re-selection + least-squares PnP.
//
// Replaces refineHyp (esac_util.h:378-454; forward: winner only, esac.cpp:167; backward: every
// hypothesis with p >= PROB_THRESH, esac.cpp:328-347) including the cv::solvePnP(SOLVEPNP_ITERATIVE,
// useExtrinsicGuess=true) it calls (esac_util.h:426-436), which minimises the plain squared
// reprojection error of the current inlier set.
//
// * The inlier test reproduces getReproErrs' arithmetic exactly (fp64 transform, float-rounded
//   projection, float difference, double norm -> float, clamp, `< tau`), because which cells pass is
//   what makes two implementations agree or not.
// * The least-squares solve is Levenberg-Marquardt on the local rotation parametrisation
//   R <- exp([w]x) R (no Rodrigues Jacobian per point); it is run to convergence (1e-12), so the
//   minimiser -- which does not depend on the parametrisation -- is what is handed back as (rvec, tvec).
// * A job (one hypothesis) is worked on by a group of `group` CTAs; reductions over cells go
//   warp shuffle -> shared memory -> (if group > 1) per-CTA slots in global memory + a counting
//   barrier, every CTA summing the slots in the same order so all take identical decisions.
#include <cooperative_groups.h>

#include "esac_internal.h"

namespace esacb200 {

constexpr int kRefThreads = 512;
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kRedN = 28;   // 21 (J^T J upper) + 6 (J^T r) + 1 (cost)
constexpr int kSlot = 32;   // doubles per CTA slot

struct RefShared {
    double red[kRefWarps][kRedN];
    double tot[kRedN];
    double R[9];
    double t[3];
    double Rc[9];
    double tc[3];
    double cur[kRedN];
    double lambda;
    int flag;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Block (and group) all-reduce of `nv` doubles held per thread in v[]; result in sh.tot[0..nv).
template <int NV>
__device__ __forceinline__ void all_reduce(double (&v)[NV], RefShared& sh, const RefineArgs& a, int grp, int cta,
                                           unsigned& epoch) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (NV == 1) {
        double x = v[0];
#pragma unroll
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh.red[warp][0] = x;
    } else {
        const double x = warp_reduce_scatter<NV>(v);
        if (lane < NV) sh.red[warp][lane] = x;
    }
    __syncthreads();
    if (tid < NV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kRefWarps; ++w) s += sh.red[w][tid];
        sh.tot[tid] = s;
    }
    __syncthreads();
    if (a.group > 1) {
        double* slots = a.scratch + (size_t)grp * a.group * 2 * kSlot;
        const int buf = epoch & 1;
        if (tid < NV) slots[((size_t)cta * 2 + buf) * kSlot + tid] = sh.tot[tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&a.barrier[grp], 1u);
            const unsigned target = (unsigned)a.group * (epoch + 1u);
            while (ld_acquire(&a.barrier[grp]) < target) { }
        }
        __syncthreads();
        if (tid < NV) {
            double s = 0;
            for (int c = 0; c < a.group; ++c) s += __ldcg(&slots[((size_t)c * 2 + buf) * kSlot + tid]);
            sh.tot[tid] = s;
        }
        __syncthreads();
        ++epoch;
    }
}

// J^T J, J^T r and cost of the reprojection residuals over the masked cells, pose (R, t) in shared memory.
// Coordinates are taken relative to the plane centre c (t here is R*c + t of the true pose): the same least-squares
// problem, but rotation updates pivot inside the scene, which keeps J^T J well conditioned for world-scale maps.
__device__ __forceinline__ void lm_accumulate(const float* __restrict__ pl, const Problem& P, const double* R, const double* t,
                                              const double* c, const uint32_t* mask, int w0, int w1, double (&acc)[kRedN]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kRedN; ++i) acc[i] = 0;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy;
    for (int w = w0 + warp; w < w1; w += kRefWarps) {
        const uint32_t bits = mask[w];
        if (!((bits >> lane) & 1u)) continue;
        const int p = w * 32 + lane;
        const int yy = p / P.W, xx = p - yy * P.W;
        const double px = (double)(xx * P.sub + P.sub / 2 - P.shiftX);
        const double py = (double)(yy * P.sub + P.sub / 2 - P.shiftY);
        const double X = (double)pl[p] - c[0], Y = (double)pl[P.N + p] - c[1], Z = (double)pl[2 * (size_t)P.N + p] - c[2];
        const double qx = R[0] * X + R[1] * Y + R[2] * Z;
        const double qy = R[3] * X + R[4] * Y + R[5] * Z;
        const double qz = R[6] * X + R[7] * Y + R[8] * Z;
        double zc = qz + t[2];
        const double iz = zc != 0. ? 1. / zc : 1.;
        const double xn = (qx + t[0]) * iz, yn = (qy + t[1]) * iz;
        const double ru = xn * f + cx - px, rv = yn * f + cy - py;
        const double a_ = f * iz, c_ = -f * xn * iz, d_ = -f * yn * iz;
        const double Ju[6] = {c_ * qy, a_ * qz - c_ * qx, -a_ * qy, a_, 0., c_};
        const double Jv[6] = {-a_ * qz + d_ * qy, -d_ * qx, a_ * qx, 0., a_, d_};
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) acc[k++] += Ju[i] * Ju[j] + Jv[i] * Jv[j];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[21 + i] += Ju[i] * ru + Jv[i] * rv;
        acc[27] += ru * ru + rv * rv;
    }
}

__device__ __forceinline__ void mat3mul(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

__global__ void __launch_bounds__(kRefThreads, 1) refine_kernel(const __grid_constant__ RefineArgs a) {
    __shared__ RefShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_groups = gridDim.x / a.group;
    const int grp = blockIdx.x / a.group, cta = blockIdx.x - grp * a.group;
    const int n_jobs = a.n_jobs ? *a.n_jobs : a.n_jobs_host;
    const Problem& P = a.P;
    const int words = (P.N + 31) / 32;
    const int wpc = (words + a.group - 1) / a.group;
    const int w0 = min(words, cta * wpc), w1 = min(words, w0 + wpc);
    unsigned epoch = 0;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy;

    for (int job = grp; job < n_jobs; job += n_groups) {
        const int h = a.jobs[job];
        const int e = a.assign32[h];
        const float* pl = a.coords + (size_t)e * 3 * P.N;
        uint32_t* mbase = a.masks + (size_t)job * 2 * a.mask_words;
        const double cen[3] = {(double)a.centres[e * 3], (double)a.centres[e * 3 + 1], (double)a.centres[e * 3 + 2]};
        Pose pose = a.poses_in[h];
        double best = 4;
        int rounds = 0, sel = 0;
        for (int step = 0; step < a.max_ref_steps; ++step) {
            // ---- inliers of the current pose (esac_util.h:403-415) ----
            double R0[9];
            rodrigues_v2m(pose.r, R0, nullptr);
            uint32_t* mtent = mbase + (size_t)(1 - sel) * a.mask_words;
            double cnt[1] = {0};
            for (int w = w0 + warp; w < w1; w += kRefWarps) {
                const int p = w * 32 + lane;
                bool inl = false;
                if (p < P.N) {
                    const int yy = p / P.W, xx = p - yy * P.W;
                    const float px = (float)(xx * P.sub + P.sub / 2 - P.shiftX);
                    const float py = (float)(yy * P.sub + P.sub / 2 - P.shiftY);
                    float err = repro_err_f(R0, pose.t, f, cx, cy, pl[p], pl[P.N + p], pl[2 * (size_t)P.N + p], px, py);
                    err = (P.max_reproj < err) ? P.max_reproj : err;  // std::min(err, maxReproj): NaN stays NaN
                    inl = err < P.tau;
                }
                const uint32_t bits = __ballot_sync(0xffffffffu, inl);
                if (lane == 0) { mtent[w] = bits; cnt[0] += (double)__popc(bits); }
            }
            all_reduce<1>(cnt, sh, a, grp, cta, epoch);   // also publishes the mask words to the block
            const double n_in = sh.tot[0];
            __syncthreads();
            if (!(n_in > best)) break;  // converged (esac_util.h:417-418)
            best = n_in;
            // ---- least-squares PnP on that set, started at the current pose ----
            if (tid == 0) {
                for (int i = 0; i < 9; ++i) sh.R[i] = R0[i];
                for (int i = 0; i < 3; ++i) sh.t[i] = R0[i * 3] * cen[0] + R0[i * 3 + 1] * cen[1] + R0[i * 3 + 2] * cen[2] + pose.t[i];
            }
            __syncthreads();
            double acc[kRedN];
            lm_accumulate(pl, P, sh.R, sh.t, cen, mtent, w0, w1, acc);
            all_reduce<kRedN>(acc, sh, a, grp, cta, epoch);
            if (tid < kRedN) sh.cur[tid] = sh.tot[tid];
            if (tid == 0) sh.lambda = 1e-3;
            __syncthreads();
            bool failed = false;
            for (int it = 0; it < 60; ++it) {
                // thread 0 solves the damped normal equations and publishes the candidate
                if (tid == 0) {
                    double H[36], g[6], d[6];
                    const double* cur = sh.cur;
                    double lambda = sh.lambda;
                    int k = 0;
                    for (int i = 0; i < 6; ++i)
                        for (int j = i; j < 6; ++j) { H[i * 6 + j] = cur[k]; H[j * 6 + i] = cur[k]; ++k; }
                    for (int i = 0; i < 6; ++i) g[i] = -cur[21 + i];
                    int ok = 0;
                    for (int tr = 0; tr < 40 && !ok; ++tr) {
                        double Hd[36];
                        for (int i = 0; i < 36; ++i) Hd[i] = H[i];
                        for (int i = 0; i < 6; ++i) Hd[i * 6 + i] = H[i * 6 + i] * (1. + lambda);
                        ok = chol_solve6(Hd, g, d) ? 1 : 0;
                        if (!ok) lambda *= 10.;
                    }
                    bool fin = ok;
                    for (int i = 0; i < 6; ++i) fin = fin && (d[i] == d[i]) && fabs(d[i]) < 1e300;
                    if (fin) {
                        double dR[9];
                        rodrigues_v2m(d, dR, nullptr);
                        mat3mul(dR, sh.R, sh.Rc);
                        for (int i = 0; i < 3; ++i) sh.tc[i] = sh.t[i] + d[3 + i];
                        double sw = fmax(fabs(d[0]), fmax(fabs(d[1]), fabs(d[2])));
                        double st = fmax(fabs(d[3]), fmax(fabs(d[4]), fabs(d[5])));
                        double tn = fmax(fabs(sh.t[0]), fmax(fabs(sh.t[1]), fabs(sh.t[2])));
                        sh.flag = (sw < 1e-12 && st < 1e-12 * (1. + tn)) ? 2 : 1;
                    } else {
                        sh.flag = 0;
                    }
                    sh.lambda = lambda;
                }
                __syncthreads();
                const int flag = sh.flag;
                if (flag == 0) { failed = true; break; }
                lm_accumulate(pl, P, sh.Rc, sh.tc, cen, mtent, w0, w1, acc);
                all_reduce<kRedN>(acc, sh, a, grp, cta, epoch);
                const bool accept = sh.tot[27] <= sh.cur[27];   // false for NaN
                const double lambda = sh.lambda;
                __syncthreads();
                if (accept) {
                    if (tid < kRedN) sh.cur[tid] = sh.tot[tid];
                    if (tid == 32) {
                        for (int i = 0; i < 9; ++i) sh.R[i] = sh.Rc[i];
                        for (int i = 0; i < 3; ++i) sh.t[i] = sh.tc[i];
                        sh.lambda = fmax(lambda * 0.1, 1e-15);
                    }
                    __syncthreads();
                    if (flag == 2) break;
                } else {
                    if (flag == 2) break;  // step below resolution and no decrease: at the minimum
                    if (tid == 0) sh.lambda = lambda * 10.;
                    __syncthreads();
                    if (lambda * 10. > 1e12) break;
                }
            }
            if (failed) break;  // "abort if PnP fails" (esac_util.h:426-437): previous pose and map stay
            Pose np_;
            rodrigues_m2v(sh.R, np_.r);
            for (int i = 0; i < 3; ++i) np_.t[i] = sh.t[i] - (sh.R[i * 3] * cen[0] + sh.R[i * 3 + 1] * cen[1] + sh.R[i * 3 + 2] * cen[2]);
            bool bad = false;
            for (int i = 0; i < 3; ++i) bad = bad || !(np_.r[i] == np_.r[i]) || !(np_.t[i] == np_.t[i]);
            __syncthreads();
            if (bad) break;
            pose = np_;
            sel = 1 - sel;
            ++rounds;
        }
        if (cta == 0 && tid == 0) {
            a.poses_out[h] = pose;
            a.rounds[job * 2] = rounds;
            a.rounds[job * 2 + 1] = sel;
        }
    }
}

void launch_refine(const RefineArgs& a, int n_groups, cudaStream_t st) {
    dim3 grid(n_groups * a.group), block(kRefThreads);
    if (a.group > 1) {
        void* params[] = {(void*)&a};
        cudaLaunchCooperativeKernel((const void*)refine_kernel, grid, block, params, 0, st);
    } else {
        refine_kernel<<<grid, block, 0, st>>>(a);
    }
}

int refine_max_coresident_blocks(int sm_count) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, refine_kernel, kRefThreads, 0) != cudaSuccess) {
        cudaGetLastError();
        nb = 1;
    }
    return nb * sm_count;
}

__global__ void finish_forward_kernel(const Pose* poses, const int* winner, const int* assign32, float* out17) {
    if (threadIdx.x == 0) {
        const int w = *winner;
        double T[16];
        pose2trans(poses[w], T);
        for (int i = 0; i < 16; ++i) out17[i] = (float)T[i];
        out17[16] = (float)assign32[w];
    }
}

void launch_finish_forward(const Pose* poses, const int* winner, const int* assign32, float* out17, cudaStream_t st) {
    finish_forward_kernel<<<1, 32, 0, st>>>(poses, winner, assign32, out17);
}

}  // namespace esacb200
