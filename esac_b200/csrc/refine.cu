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
// * The least-squares solve reproduces OpenCV's own iteration (cvFindExtrinsicCameraParams2 + CvLevMarq as
//   observed on cv2 4.13: parameters (rvec, tvec), J^T J with its diagonal scaled by 1 + 10^k, k from -3,
//   step accepted when the error norm does not grow, at most 20 iterations, stop when the relative parameter
//   change drops below FLT_EPSILON).  This matters: on world-scale maps (|t| ~ 1e3) that criterion stops long
//   before the minimiser, so only the same iteration gives the same pose and the same next inlier set.
//   Per cell the Jacobian is taken in a cheap, well-conditioned local frame (rotation increment about the
//   plane centre); the reduced 6x6 sums are then mapped to the (rvec, tvec) frame by one 6x6 change of
//   variables per evaluation.
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
    double red[kRefWarps][32];
    double tot[32];
    double R[9];
    double t[3];
    double cur[kRedN];   // (rvec, tvec)-frame sums at the last accepted parameters
    double cand[kRedN];  // same at the candidate
    double par[6], prev[6];
    double Rc[3], dR[27], T[9], G[36], H1[36];
    double prev_err, err_norm;
    int lamlg, iters, flag;
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
        // every CTA sums the group's slots in the same fixed order: warp w owns values w, w + 16; its lanes stride
        // over the CTAs (independent L2 reads in flight), then a shuffle tree
        for (int v_ = warp; v_ < NV; v_ += kRefWarps) {
            double s = 0;
            for (int c = lane; c < a.group; c += 32) s += __ldcg(&slots[((size_t)c * 2 + buf) * kSlot + v_]);
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sh.tot[v_] = s;
        }
        __syncthreads();
        ++epoch;
    }
}

// J^T J, J^T r and cost of the reprojection residuals over the masked cells, pose (R, t) in shared memory.
// Coordinates are taken relative to the plane centre c (t here is R*c + t of the true pose): the same least-squares
// problem, but rotation updates pivot inside the scene, which keeps J^T J well conditioned for world-scale maps.
// BUILD_MASK: the pass also decides, for every cell, whether it is an inlier of the true pose (R0, t0) -- exactly
// getReproErrs' arithmetic --, writes the bit mask and counts (acc[28]); otherwise the mask is read.
template <bool BUILD_MASK>
__device__ __forceinline__ void lm_accumulate(const float* __restrict__ pl, const Problem& P, const double* R, const double* t,
                                              const double* c, uint32_t* mask, int w0, int w1, const double* R0, const double* t0,
                                              double (&acc)[kRedN + 1]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kRedN + 1; ++i) acc[i] = 0;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy;
    for (int w = w0 + warp; w < w1; w += kRefWarps) {
        const int p = w * 32 + lane;
        const int yy = p / P.W, xx = p - yy * P.W;
        const int ipx = xx * P.sub + P.sub / 2 - P.shiftX, ipy = yy * P.sub + P.sub / 2 - P.shiftY;
        float Xf = 0.f, Yf = 0.f, Zf = 0.f;
        bool inl;
        if (BUILD_MASK) {
            inl = false;
            if (p < P.N) {
                Xf = pl[p]; Yf = pl[P.N + p]; Zf = pl[2 * (size_t)P.N + p];
                float err = repro_err_f(R0, t0, f, cx, cy, Xf, Yf, Zf, (float)ipx, (float)ipy);
                err = (P.max_reproj < err) ? P.max_reproj : err;  // std::min(err, maxReproj): NaN stays NaN
                inl = err < P.tau;                                  // esac_util.h:406
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, inl);
            if (lane == 0) { mask[w] = bits; acc[kRedN] += (double)__popc(bits); }
        } else {
            inl = (mask[w] >> lane) & 1u;
            if (inl) { Xf = pl[p]; Yf = pl[P.N + p]; Zf = pl[2 * (size_t)P.N + p]; }
        }
        if (!inl) continue;
        const double px = (double)ipx, py = (double)ipy;
        const double X = (double)Xf - c[0], Y = (double)Yf - c[1], Z = (double)Zf - c[2];
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
            // ---- inliers of the current pose (esac_util.h:403-415), fused with the first Jacobian evaluation ----
            double R0[9];
            rodrigues_v2m(pose.r, R0, nullptr);
            uint32_t* mtent = mbase + (size_t)(1 - sel) * a.mask_words;
            // ---- solvePnP(SOLVEPNP_ITERATIVE, useExtrinsicGuess) on that set, started at the current pose ----
            // evaluate(): sums of the cell Jacobians / residuals at sh.par, mapped to the (rvec, tvec) frame -> sh.cand
            auto evaluate = [&](bool first) {
                if (tid == 0) {
                    double R[9];
                    rodrigues_v2m(sh.par, R, nullptr);
                    for (int i = 0; i < 3; ++i) {
                        sh.Rc[i] = R[i * 3] * cen[0] + R[i * 3 + 1] * cen[1] + R[i * 3 + 2] * cen[2];
                        sh.t[i] = sh.Rc[i] + sh.par[3 + i];
                    }
                    for (int i = 0; i < 9; ++i) sh.R[i] = R[i];
                }
                __syncthreads();
                double acc[kRedN + 1];
                if (first) lm_accumulate<true>(pl, P, sh.R, sh.t, cen, mtent, w0, w1, R0, pose.t, acc);
                else lm_accumulate<false>(pl, P, sh.R, sh.t, cen, mtent, w0, w1, R0, pose.t, acc);
                all_reduce<kRedN + 1>(acc, sh, a, grp, cta, epoch);
                // Change of variables local frame -> (rvec, tvec), spread over the lanes of warp 0:
                //   local (w', t') = Q (w, t), Q = [[I, 0], [-[Rc]x, I]];  (w, t) = P (r, t), P = blkdiag(T, I),
                //   T[:, i] = vee((dR/dr_i) R^T);  G = Q P;  JtJ = G^T H G, JtErr = G^T g.
                if (warp == 0) {
                    const double rx0 = sh.par[0], ry0 = sh.par[1], rz0 = sh.par[2];
                    const double theta = sqrt(rx0 * rx0 + ry0 * ry0 + rz0 * rz0);
                    if (lane < 27) {  // dR[i*9 + e] = d R[e] / d r_i  (cv::Rodrigues' Jacobian, esac_geom.cuh)
                        const int i = lane / 9, e = lane - 9 * i, ra = e / 3, cb = e - 3 * ra;
                        auto eps3 = [](int p_, int q_, int r_) { return (double)((p_ - q_) * (q_ - r_) * (r_ - p_)) * 0.5; };
                        double v;
                        if (theta < DBL_EPSILON) {
                            v = -eps3(ra, cb, i);
                        } else {
                            const double c = cos(theta), sn = sin(theta), c1 = 1. - c, it = 1. / theta;
                            const double rr[3] = {rx0 * it, ry0 * it, rz0 * it};
                            const double ri = rr[i];
                            const double a0 = -sn * ri, a1 = (sn - 2 * c1 * it) * ri, a2 = c1 * it, a3 = (c - sn * it) * ri, a4 = sn * it;
                            const double I_e = ra == cb ? 1. : 0.;
                            const double rrt = rr[ra] * rr[cb];
                            const double drrt = (ra == i ? rr[cb] : 0.) + (cb == i ? rr[ra] : 0.);
                            double rxm = 0;  // [r]x entry (ra, cb) = -eps(ra, cb, k) r_k
                            for (int k = 0; k < 3; ++k) rxm -= eps3(ra, cb, k) * rr[k];
                            const double drx = -eps3(ra, cb, i);
                            v = a0 * I_e + a1 * rrt + a2 * drrt + a3 * rxm + a4 * drx;
                        }
                        sh.dR[lane] = v;
                    }
                    __syncwarp();
                    if (lane < 9) {  // T[row][i]: rows (2,1), (0,2), (1,0) of (dR_i R^T)
                        const int row = lane / 3, i = lane - 3 * row;
                        const int p_ = row == 0 ? 2 : (row == 1 ? 0 : 1), q_ = row == 0 ? 1 : (row == 1 ? 2 : 0);
                        const double* d = sh.dR + i * 9 + p_ * 3;
                        sh.T[lane] = d[0] * sh.R[q_ * 3] + d[1] * sh.R[q_ * 3 + 1] + d[2] * sh.R[q_ * 3 + 2];
                    }
                    __syncwarp();
                    for (int idx = lane; idx < 36; idx += 32) {
                        const int row = idx / 6, col = idx - 6 * row;
                        double v = 0;
                        if (row < 3) {
                            v = col < 3 ? sh.T[row * 3 + col] : 0.;
                        } else if (col < 3) {
                            const int r = row - 3;  // -[Rc]x row r
                            const double K0 = r == 0 ? 0. : (r == 1 ? -sh.Rc[2] : sh.Rc[1]);
                            const double K1 = r == 0 ? sh.Rc[2] : (r == 1 ? 0. : -sh.Rc[0]);
                            const double K2 = r == 0 ? -sh.Rc[1] : (r == 1 ? sh.Rc[0] : 0.);
                            v = K0 * sh.T[col] + K1 * sh.T[3 + col] + K2 * sh.T[6 + col];
                        } else {
                            v = (col - 3 == row - 3) ? 1. : 0.;
                        }
                        sh.G[idx] = v;
                    }
                    __syncwarp();
                    auto sym = [&](int i, int j) {  // upper-triangle storage of the reduced local J^T J
                        const int lo = i < j ? i : j, hi = i < j ? j : i;
                        return sh.tot[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
                    };
                    for (int idx = lane; idx < 36; idx += 32) {
                        const int i = idx / 6, j = idx - 6 * i;
                        double v = 0;
                        for (int q = 0; q < 6; ++q) v += sym(i, q) * sh.G[q * 6 + j];
                        sh.H1[idx] = v;
                    }
                    __syncwarp();
                    if (lane < 21) {
                        int i = 0, rem = lane;
                        while (rem >= 6 - i) { rem -= 6 - i; ++i; }
                        const int j = i + rem;
                        double v = 0;
                        for (int q = 0; q < 6; ++q) v += sh.G[q * 6 + i] * sh.H1[q * 6 + j];
                        sh.cand[lane] = v;
                    } else if (lane < 27) {
                        const int i = lane - 21;
                        double v = 0;
                        for (int q = 0; q < 6; ++q) v += sh.G[q * 6 + i] * sh.tot[21 + q];
                        sh.cand[lane] = v;
                    } else if (lane == 27) {
                        sh.cand[27] = sh.tot[27];
                    }
                }
                __syncthreads();
            };
            // step(): param = prevParam - solve(JtJ with diag * (1 + 10^lamlg), JtErr)   (CvLevMarq::step)
            auto lm_step = [&]() {
                double A[36], Ai[36];
                int k = 0;
                for (int i = 0; i < 6; ++i)
                    for (int j = i; j < 6; ++j) { A[i * 6 + j] = sh.cur[k]; A[j * 6 + i] = sh.cur[k]; ++k; }
                const double lambda = exp((double)sh.lamlg * log(10.));
                for (int i = 0; i < 6; ++i) A[i * 6 + i] *= 1. + lambda;
                // solve(JtJN, JtErr, DECOMP_SVD): Cholesky when the damped matrix is positive definite (the solution is
                // the same up to rounding, which the iteration is insensitive to), SVD-style pseudo-inverse otherwise
                double dlt[6];
                if (!chol_solve6(A, &sh.cur[21], dlt)) {
                    pinv_sym6(A, Ai);
                    for (int i = 0; i < 6; ++i) {
                        double d = 0;
                        for (int j = 0; j < 6; ++j) d += Ai[i * 6 + j] * sh.cur[21 + j];
                        dlt[i] = d;
                    }
                }
                for (int i = 0; i < 6; ++i) sh.par[i] = sh.prev[i] - dlt[i];
            };
            if (tid == 0) {
                for (int i = 0; i < 3; ++i) { sh.par[i] = pose.r[i]; sh.par[3 + i] = pose.t[i]; }
                sh.lamlg = -3;
                sh.iters = 0;
            }
            __syncthreads();
            evaluate(true);
            const double n_in = sh.tot[kRedN];
            __syncthreads();
            if (!(n_in > best)) break;  // converged (esac_util.h:417-418)
            best = n_in;
            if (tid == 0) {
                for (int i = 0; i < kRedN; ++i) sh.cur[i] = sh.cand[i];
                sh.prev_err = sqrt(sh.cand[27]);  // iters == 0: prevErrNorm = ||err(param0)||
            }
            __syncthreads();
            for (int outer = 0; outer < 20; ++outer) {
                if (tid == 0) {
                    for (int i = 0; i < 6; ++i) sh.prev[i] = sh.par[i];
                    lm_step();
                }
                __syncthreads();
                for (;;) {  // CHECK_ERR (the evaluation also yields the Jacobian sums reused if the step is kept)
                    evaluate(false);
                    if (tid == 0) {
                        sh.err_norm = sqrt(sh.cand[27]);
                        if (sh.err_norm > sh.prev_err && ++sh.lamlg <= 16) { lm_step(); sh.flag = 1; }
                        else sh.flag = 0;
                    }
                    __syncthreads();
                    const int retry = sh.flag;
                    __syncthreads();  // every thread has read the flag before thread 0 may rewrite it below
                    if (retry == 0) break;
                }
                if (tid == 0) {
                    sh.lamlg = sh.lamlg - 1 > -16 ? sh.lamlg - 1 : -16;
                    double dn = 0, pn = 0;
                    for (int i = 0; i < 6; ++i) { const double d = sh.par[i] - sh.prev[i]; dn += d * d; pn += sh.prev[i] * sh.prev[i]; }
                    const bool done = (++sh.iters >= 20) || (sqrt(dn) / sqrt(pn) < (double)FLT_EPSILON);
                    sh.flag = done ? 1 : 0;
                    if (!done) {
                        sh.prev_err = sh.err_norm;
                        for (int i = 0; i < kRedN; ++i) sh.cur[i] = sh.cand[i];
                    }
                }
                __syncthreads();
                const int done = sh.flag;
                __syncthreads();
                if (done) break;
            }
            Pose np_;
            for (int i = 0; i < 3; ++i) { np_.r[i] = sh.par[i]; np_.t[i] = sh.par[3 + i]; }
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

__global__ void finish_forward_kernel(const Pose* poses, const int* winner, const int* assign32, const int* flags, float* out) {
    if (threadIdx.x == 0) {
        const int w = *winner;
        double T[16];
        pose2trans(poses[w], T);
        for (int i = 0; i < 16; ++i) out[i] = (float)T[i];
        out[16] = (float)assign32[w];
        out[17] = (float)flags[0];
        out[18] = (float)w;
    }
}

// [scores (M) | camera pose (16) | global expert id or -1 on a bad assignment | local winner] as doubles: the record one
// shard contributes to the all-gather of the sharded forward
__global__ void pack_forward_kernel(const double* scores, const float* out20, int M, int expert_offset, double* pack) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M + 18; i += gridDim.x * blockDim.x) {
        double v;
        if (i < M) v = scores[i];
        else if (i < M + 16) v = (double)out20[i - M];
        else if (i == M + 16) v = out20[17] != 0.f ? -1. : (double)out20[16] + (double)expert_offset;
        else v = (double)out20[18];
        pack[i] = v;
    }
}

void launch_pack_forward(const double* scores, const float* out20, int M, int expert_offset, double* pack, cudaStream_t st) {
    pack_forward_kernel<<<(M + 18 + 255) / 256, 256, 0, st>>>(scores, out20, M, expert_offset, pack);
}

void launch_finish_forward(const Pose* poses, const int* winner, const int* assign32, const int* flags, float* out20,
                           cudaStream_t st) {
    finish_forward_kernel<<<1, 32, 0, st>>>(poses, winner, assign32, flags, out20);
}

}  // namespace esacb200
