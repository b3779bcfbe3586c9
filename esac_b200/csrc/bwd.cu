// Backward pass: expected pose loss and its gradient w.r.t. the scene coordinates, matrix-free.
//
// Replaces esac_backward's tail (esac.cpp:353-510) and esac_derivative.h / esac_loss.h:
//   losses + expectation              esac.cpp:354-362, esac_loss.h:66-83
//   path I  (refinement, implicit)    esac.cpp:373-463   J_R = -(J^T J)^-1 J^T, clamp > 10, dLoss * dHyp_dObjs
//   path II (score)                   esac_derivative.h:205-324 (dScore), 347-420 (dSMScore), 128-185 (dPNP)
//   assembly                          esac.cpp:491-508   out[e][c][y][x] += p_h * gradI + gradII   (float += double)
// The reference materialises a 6 x 3N matrix and an N x 6 Jacobian per hypothesis (11 GB at 480x640 x 256);
// here per hypothesis only 27 reduced numbers exist between the passes:
//   A = J^T J at the refined pose over the final inlier set (21), s = sum_cells w * jacobeanHyp row (6);
//   v = -dLoss * pinv(A);   gradI(cell)  = (v . J_cell) * dProjectdObj_refined(cell)
//   gradII(cell) = w(cell) * dProjectdObj_initial(cell) [+ (s * dPNP) for the 3 minimal-set cells]
#include "esac_internal.h"

namespace esacb200 {

constexpr int kBwdThreads = 256;
constexpr int kBwdPix = 4;                       // cells per thread in the reduction passes
constexpr int kBwdTile = kBwdThreads * kBwdPix;  // 1024
constexpr int kRed = 27;

struct HypGrad {
    int h, expert, flagI, pad;
    double p, g;
    double Ri[9], ti[3], dRi[27];
    double Rr[9], tr[3], dRr[27];
    double dl[6], v[6], inv[36], support[12];
    unsigned long long maxjr_bits;
    int cells[8];
};

size_t bwd_hypgrad_bytes() { return sizeof(HypGrad); }
int bwd_red_vals() { return kRed; }
int bwd_tiles(int N) { return (N + kBwdTile - 1) / kBwdTile; }

// trans2pose for the (float) ground truth: general affine inverse like cv::Mat::inv, then Rodrigues.
__device__ void gt_trans2pose(const float* gt, Pose& p, double T[16]) {
    for (int i = 0; i < 16; ++i) T[i] = (double)gt[i];
    double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    double B[9];
    adj3(R, B);
    double d = R[0] * B[0] + R[1] * B[3] + R[2] * B[6];
    double Ri[9];
    for (int i = 0; i < 9; ++i) Ri[i] = B[i] / d;
    for (int r = 0; r < 3; ++r) p.t[r] = -(Ri[r * 3] * T[3] + Ri[r * 3 + 1] * T[7] + Ri[r * 3 + 2] * T[11]);
    // polar factor via Newton: X <- (X + X^-T)/2
    double X[9];
    for (int i = 0; i < 9; ++i) X[i] = Ri[i];
    for (int it = 0; it < 8; ++it) {
        double C[9];
        adj3(X, C);
        double dd = X[0] * C[0] + X[1] * C[3] + X[2] * C[6];
        if (!(fabs(dd) > 0)) break;
        double Y[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Y[r * 3 + c] = 0.5 * (X[r * 3 + c] + C[c * 3 + r] / dd);  // X^-T[r][c] = adj[c][r]/det
        for (int i = 0; i < 9; ++i) X[i] = Y[i];
    }
    rodrigues_m2v(X, p.r);
}

struct BwdAux {
    HypGrad* hg;
    double* red;        // [job][tiles][kRed]
    int* expert_njobs;  // [E]
    int tiles;
};

// ---------------------------------------------------------------------------------------------
// B1: losses, expectation, score gradients, per-job records
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) bwd_loss_kernel(const __grid_constant__ BwdArgs a, BwdAux x) {
    const Problem& P = a.P;
    __shared__ Pose gtp;
    __shared__ double gtT[16];
    __shared__ double expected;
    const int tid = threadIdx.x;
    if (tid == 0) gt_trans2pose(a.gt, gtp, gtT);
    for (int e = tid; e < P.E; e += blockDim.x) x.expert_njobs[e] = 0;
    __syncthreads();
    for (int h = tid; h < P.M; h += blockDim.x) {
        double T[16];
        pose2trans(a.ref[h], T);
        a.losses[h] = pose_loss(T, gtT, (double)a.wRot, (double)a.wTrans, (double)a.cut);
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        for (int h = 0; h < P.M; ++h) s += a.probs[h] * a.losses[h];  // esac.cpp:357-362 order
        *a.out_loss = s;                                              // local (this rank's) part of the expectation
        expected = a.expected_override ? *a.expected_override : s;    // multi-GPU: sum over all ranks
    }
    __syncthreads();
    const int nc = *a.n_contrib;
    for (int j = tid; j < nc; j += blockDim.x) {
        const int h = a.contrib[j];
        HypGrad& g = x.hg[j];
        g.h = h;
        g.expert = a.assign32[h];
        g.flagI = a.rounds[2 * j] > 0 ? 1 : 0;
        g.pad = a.rounds[2 * j + 1];
        g.p = a.probs[h];
        g.g = a.probs[h] * a.losses[h] - a.probs[h] * expected;  // esac_derivative.h:372-374
        rodrigues_v2m(a.init[h].r, g.Ri, g.dRi);
        rodrigues_v2m(a.ref[h].r, g.Rr, g.dRr);
        for (int i = 0; i < 3; ++i) { g.ti[i] = a.init[h].t[i]; g.tr[i] = a.ref[h].t[i]; }
        pose_dloss(a.ref[h], gtp, (double)a.wRot, (double)a.wTrans, (double)a.cut, g.dl);
        for (int i = 0; i < 8; ++i) g.cells[i] = a.cells[h * 8 + i];
        g.maxjr_bits = 0ull;
        atomicAdd(&x.expert_njobs[g.expert], 1);
    }
}

// jacobeanR / jacobeanHyp row of one cell (esac_util.h:339-351, esac.cpp:419-431): zero row when err > maxReproj.
__device__ __forceinline__ bool jac_row(const double* R, const double* t, const double* dRdr, double f, double cx, double cy,
                                        float X, float Y, float Z, float px, float py, double max_reproj, double row[6]) {
    float uf, vf;
    project_point_f(R, t, f, cx, cy, X, Y, Z, uf, vf);
    const float dxf = uf - px, dyf = vf - py;
    double err = sqrt((double)dxf * (double)dxf + (double)dyf * (double)dyf);
    err = fmax(err, kEps);
    if (!(err <= max_reproj)) return false;  // `err > maxReproj` -> skipped; NaN rows are dropped here
    double u, v, Ju[6], Jv[6];
    project_point_jac(R, t, dRdr, f, cx, cy, (double)X, (double)Y, (double)Z, u, v, Ju, Jv);
    const double a_ = 1. / err * (double)dxf, b_ = 1. / err * (double)dyf;
#pragma unroll
    for (int i = 0; i < 6; ++i) row[i] = a_ * Ju[i] + b_ * Jv[i];
    return true;
}

// d score / d reprojection error of one cell (esac_derivative.h:261-266)
__device__ __forceinline__ double score_weight(float err_clamped, const Problem& P, double g, double fac) {
    const float stf = P.beta * (err_clamped - P.tau);
    double st = (double)stf;
    st = 1 / (1 + exp(-st));
    return (-st * (1 - st) * (double)P.beta * g) * fac;
}

__device__ __forceinline__ void block_reduce_store(double (&v)[kRed], double* dst) {
    __shared__ double sred[kBwdThreads / 32][kRed];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const double s = warp_reduce_scatter<kRed>(v);
        if (lane < kRed) sred[warp][lane] = s;
    }
    __syncthreads();
    if (threadIdx.x < kRed) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kBwdThreads / 32; ++w) s += sred[w][threadIdx.x];
        dst[threadIdx.x] = s;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// B2: per (job, tile) partial sums of J^T J (refined pose, final inliers) and sum w * jacobeanHyp (initial pose)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdThreads) bwd_reduce_kernel(const __grid_constant__ BwdArgs a, BwdAux x) {
    const Problem& P = a.P;
    const int nc = *a.n_contrib;
    const int n_items = nc * x.tiles;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy, mr = (double)P.max_reproj;
    const float facf = P.alpha / (float)P.W / (float)P.H;
    __shared__ HypGrad sg;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int j = item / x.tiles, tile = item - j * x.tiles;
        __syncthreads();
        for (int i = threadIdx.x; i < (int)(sizeof(HypGrad) / 8); i += blockDim.x) ((double*)&sg)[i] = ((const double*)&x.hg[j])[i];
        __syncthreads();
        const float* pl = a.coords + (size_t)sg.expert * 3 * P.N;
        const uint32_t* mask = a.masks + ((size_t)j * 2 + sg.pad) * a.mask_words;
        double acc[kRed];
#pragma unroll
        for (int i = 0; i < kRed; ++i) acc[i] = 0;
        for (int k = 0; k < kBwdPix; ++k) {
            const int p = tile * kBwdTile + k * kBwdThreads + threadIdx.x;
            if (p >= P.N) continue;
            const int yy = p / P.W, xx = p - yy * P.W;
            const float px = (float)(xx * P.sub + P.sub / 2 - P.shiftX);
            const float py = (float)(yy * P.sub + P.sub / 2 - P.shiftY);
            const float X = pl[p], Y = pl[P.N + p], Z = pl[2 * (size_t)P.N + p];
            double row[6];
            if (sg.flagI && ((mask[p >> 5] >> (p & 31)) & 1u)) {
                if (jac_row(sg.Rr, sg.tr, sg.dRr, f, cx, cy, X, Y, Z, px, py, mr, row)) {
                    int q = 0;
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int l = i; l < 6; ++l) acc[q++] += row[i] * row[l];
                }
            }
            // path II weight at the initial pose
            float err = repro_err_f(sg.Ri, sg.ti, f, cx, cy, X, Y, Z, px, py);
            err = (P.max_reproj < err) ? P.max_reproj : err;
            const double w = score_weight(err, P, sg.g, (double)facf);
            if (jac_row(sg.Ri, sg.ti, sg.dRi, f, cx, cy, X, Y, Z, px, py, mr, row)) {
#pragma unroll
                for (int i = 0; i < 6; ++i) acc[21 + i] += w * row[i];
            }
        }
        block_reduce_store(acc, x.red + ((size_t)j * x.tiles + tile) * kRed);
    }
}

// ---------------------------------------------------------------------------------------------
// B3: per job small algebra: pinv(J^T J), v, dPNP by central differences (18 P3P solves on 18 lanes), support
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) bwd_solve_kernel(const __grid_constant__ BwdArgs a, BwdAux x) {
    const Problem& P = a.P;
    const int nc = *a.n_contrib;
    const int lane = threadIdx.x;
    __shared__ double tot[kRed];
    __shared__ double fb[18][6];
    __shared__ int okf[18];
    __shared__ double dH[6][12];
    for (int j = blockIdx.x; j < nc; j += gridDim.x) {
        HypGrad& g = x.hg[j];
        __syncwarp();
        if (lane < kRed) {
            double s = 0;
            for (int t = 0; t < x.tiles; ++t) s += x.red[((size_t)j * x.tiles + t) * kRed + lane];
            tot[lane] = s;
        }
        __syncwarp();
        // dPNP (esac_derivative.h:128-185): lane = (i*3 + jj)*2 + dir
        const int h = g.h;
        const float* pl = a.coords + (size_t)g.expert * 3 * P.N;
        if (lane < 18) {
            float obj[4][3], img[4][2];
            for (int q = 0; q < 4; ++q) {
                const int cxq = g.cells[2 * q], cyq = g.cells[2 * q + 1];
                const int p = cyq * P.W + cxq;
                obj[q][0] = pl[p]; obj[q][1] = pl[P.N + p]; obj[q][2] = pl[2 * (size_t)P.N + p];
                img[q][0] = (float)(cxq * P.sub + P.sub / 2 - P.shiftX);
                img[q][1] = (float)(cyq * P.sub + P.sub / 2 - P.shiftY);
            }
            const int col = lane >> 1, dir = lane & 1, pi = col / 3, pj = col % 3;
            const float eps = 0.001f;
            // float arithmetic of the reference (esac_derivative.h:147-171): x += eps (forward solve);
            // x -= 2*eps (backward solve); x += eps (restore -- not always bit-exact, and the restored value
            // is what the later columns see)
            for (int c = 0; c < col; ++c) {
                float r = obj[c / 3][c % 3] + eps;
                r = r - 2 * eps;
                obj[c / 3][c % 3] = r + eps;
            }
            float vfw = obj[pi][pj] + eps;
            float vbw = vfw - 2 * eps;
            obj[pi][pj] = dir == 0 ? vfw : vbw;
            Pose ps;
            const bool ok = p3p_pose(obj, img, (double)P.f, (double)P.ppx, (double)P.ppy, ps);
            okf[lane] = ok ? 1 : 0;
            for (int q = 0; q < 3; ++q) { fb[lane][q] = ps.r[q]; fb[lane][3 + q] = ps.t[q]; }
        }
        __syncwarp();
        if (lane == 0) {
            bool good = true;
            for (int q = 0; q < 18; ++q) good = good && okf[q];
            const double two_eps = (double)(2 * 0.001f);
            double mx = -1;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 12; ++c) dH[r][c] = 0;
            if (good) {
                for (int col = 0; col < 9 && good; ++col)
                    for (int r = 0; r < 6; ++r) {
                        const double val = (fb[2 * col][r] - fb[2 * col + 1][r]) / two_eps;
                        if (!(val == val)) good = false;
                        dH[r][col] = val;
                    }
            }
            if (!good)
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 12; ++c) dH[r][c] = 0;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 12; ++c) { const double v_ = fabs(dH[r][c]); if (mx < 0 || v_ > mx) mx = v_; }
            if (mx > 10)
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 12; ++c) dH[r][c] = 0;
            for (int c = 0; c < 12; ++c) {
                double s = 0;
                for (int r = 0; r < 6; ++r) s += tot[21 + r] * dH[r][c];
                g.support[c] = s;
            }
            // path I algebra
            double A[36];
            int q = 0;
            for (int i = 0; i < 6; ++i)
                for (int l = i; l < 6; ++l) { A[i * 6 + l] = tot[q]; A[l * 6 + i] = tot[q]; ++q; }
            pinv_sym6(A, g.inv);
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int l = 0; l < 6; ++l) s += g.dl[l] * g.inv[l * 6 + i];
                g.v[i] = -s;
            }
        }
        (void)h;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// B4: max |J_R| = max_cells,k |(pinv(A) J_cell^T)_k|  (esac.cpp:434-437 clamp)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdThreads) bwd_maxjr_kernel(const __grid_constant__ BwdArgs a, BwdAux x) {
    const Problem& P = a.P;
    const int nc = *a.n_contrib;
    const int n_items = nc * x.tiles;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy, mr = (double)P.max_reproj;
    __shared__ HypGrad sg;
    __shared__ double smax[kBwdThreads / 32];
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int j = item / x.tiles, tile = item - j * x.tiles;
        __syncthreads();
        for (int i = threadIdx.x; i < (int)(sizeof(HypGrad) / 8); i += blockDim.x) ((double*)&sg)[i] = ((const double*)&x.hg[j])[i];
        __syncthreads();
        if (!sg.flagI) continue;
        const float* pl = a.coords + (size_t)sg.expert * 3 * P.N;
        const uint32_t* mask = a.masks + ((size_t)j * 2 + sg.pad) * a.mask_words;
        double mx = 0;
        for (int k = 0; k < kBwdPix; ++k) {
            const int p = tile * kBwdTile + k * kBwdThreads + threadIdx.x;
            if (p >= P.N) continue;
            if (!((mask[p >> 5] >> (p & 31)) & 1u)) continue;
            const int yy = p / P.W, xx = p - yy * P.W;
            const float px = (float)(xx * P.sub + P.sub / 2 - P.shiftX);
            const float py = (float)(yy * P.sub + P.sub / 2 - P.shiftY);
            double row[6];
            if (!jac_row(sg.Rr, sg.tr, sg.dRr, f, cx, cy, pl[p], pl[P.N + p], pl[2 * (size_t)P.N + p], px, py, mr, row)) continue;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double s = 0;
#pragma unroll
                for (int l = 0; l < 6; ++l) s += sg.inv[i * 6 + l] * row[l];
                mx = fmax(mx, fabs(s));
            }
        }
        for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kBwdThreads / 32; ++w) mx = fmax(mx, smax[w]);
            atomicMax(&x.hg[j].maxjr_bits, (unsigned long long)__double_as_longlong(mx));  // order-independent
        }
    }
}

// ---------------------------------------------------------------------------------------------
// B5: assembly, one thread per cell, hypotheses of the cell's expert in ascending order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBwdThreads) bwd_assemble_kernel(const __grid_constant__ BwdArgs a, BwdAux x) {
    const Problem& P = a.P;
    const int e = blockIdx.y;
    if (x.expert_njobs[e] == 0) return;
    const int nc = *a.n_contrib;
    const double f = (double)P.f, cx = (double)P.ppx, cy = (double)P.ppy, mr = (double)P.max_reproj;
    const float facf = P.alpha / (float)P.W / (float)P.H;
    const int p = blockIdx.x * kBwdThreads + threadIdx.x;
    const bool live = p < P.N;
    const float* pl = a.coords + (size_t)e * 3 * P.N;
    float* gr = a.grads + (size_t)e * 3 * P.N;
    float X = 0, Y = 0, Z = 0, px = 0, py = 0;
    int xx = 0, yy = 0;
    float o0 = 0, o1 = 0, o2 = 0;
    if (live) {
        yy = p / P.W; xx = p - yy * P.W;
        px = (float)(xx * P.sub + P.sub / 2 - P.shiftX);
        py = (float)(yy * P.sub + P.sub / 2 - P.shiftY);
        X = pl[p]; Y = pl[P.N + p]; Z = pl[2 * (size_t)P.N + p];
        o0 = gr[p]; o1 = gr[P.N + p]; o2 = gr[2 * (size_t)P.N + p];
    }
    __shared__ HypGrad sg;
    for (int j = 0; j < nc; ++j) {
        if (x.hg[j].expert != e) continue;  // uniform
        __syncthreads();
        for (int i = threadIdx.x; i < (int)(sizeof(HypGrad) / 8); i += blockDim.x) ((double*)&sg)[i] = ((const double*)&x.hg[j])[i];
        __syncthreads();
        if (!live) continue;
        double g0, g1, g2;
        {   // path II: direct influence on the score at the initial pose (esac_derivative.h:303-306)
            float err = repro_err_f(sg.Ri, sg.ti, f, cx, cy, X, Y, Z, px, py);
            err = (P.max_reproj < err) ? P.max_reproj : err;
            const double w = score_weight(err, P, sg.g, (double)facf);
            double d[3];
            d_project_d_obj(px, py, X, Y, Z, sg.Ri, sg.ti, f, cx, cy, mr, d);
            g0 = d[0] * w; g1 = d[1] * w; g2 = d[2] * w;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (sg.cells[2 * q] == xx && sg.cells[2 * q + 1] == yy) {
                    g0 += sg.support[3 * q]; g1 += sg.support[3 * q + 1]; g2 += sg.support[3 * q + 2];
                }
        }
        double h0 = 0, h1 = 0, h2 = 0;
        if (sg.flagI) {  // path I: through the refined pose (esac.cpp:442-449, 461-462)
            const uint32_t* mask = a.masks + ((size_t)j * 2 + sg.pad) * a.mask_words;
            const double maxjr = __longlong_as_double((long long)sg.maxjr_bits);
            if (((mask[p >> 5] >> (p & 31)) & 1u) && !(maxjr > 10)) {
                double row[6];
                double s = 0;
                if (jac_row(sg.Rr, sg.tr, sg.dRr, f, cx, cy, X, Y, Z, px, py, mr, row)) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) s += sg.v[i] * row[i];
                }
                double d[3];
                d_project_d_obj(px, py, X, Y, Z, sg.Rr, sg.tr, f, cx, cy, mr, d);
                h0 = s * d[0]; h1 = s * d[1]; h2 = s * d[2];
            }
        }
        // outGradients += hypProbs*gradients + dLoss_dScore : float += double (esac.cpp:501-506)
        o0 = (float)((double)o0 + (sg.p * h0 + g0));
        o1 = (float)((double)o1 + (sg.p * h1 + g1));
        o2 = (float)((double)o2 + (sg.p * h2 + g2));
    }
    if (live) { gr[p] = o0; gr[P.N + p] = o1; gr[2 * (size_t)P.N + p] = o2; }
}

void launch_backward_losses(const BwdArgs& a, cudaStream_t st) {
    BwdAux x;
    x.hg = (HypGrad*)a.hyp_grad;
    x.red = a.red;
    x.expert_njobs = (int*)a.job_of;
    x.tiles = bwd_tiles(a.P.N);
    bwd_loss_kernel<<<1, 1024, 0, st>>>(a, x);
}

void launch_backward(const BwdArgs& a, int max_jobs, cudaStream_t st) {
    BwdAux x;
    x.hg = (HypGrad*)a.hyp_grad;
    x.red = a.red;
    x.expert_njobs = (int*)a.job_of;  // [E] ints, buffer provided by the caller
    x.tiles = bwd_tiles(a.P.N);
    bwd_loss_kernel<<<1, 1024, 0, st>>>(a, x);
    long long items = (long long)max_jobs * x.tiles;
    int grid = (int)(items < 148 * 8 ? items : 148 * 8);
    if (grid < 1) grid = 1;
    bwd_reduce_kernel<<<grid, kBwdThreads, 0, st>>>(a, x);
    bwd_solve_kernel<<<max_jobs < 592 ? max_jobs : 592, 32, 0, st>>>(a, x);
    bwd_maxjr_kernel<<<grid, kBwdThreads, 0, st>>>(a, x);
    dim3 ga((a.P.N + kBwdThreads - 1) / kBwdThreads, a.P.E);
    bwd_assemble_kernel<<<ga, kBwdThreads, 0, st>>>(a, x);
}

// dst += src (the rank-summed gradient of a hypothesis-major sharded backward joins the caller's tensor: esac.cpp:501-506 is +=)
__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = reinterpret_cast<float4*>(dst)[i];
        const float4 b = reinterpret_cast<const float4*>(src)[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(dst)[i] = a;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

__global__ void add_inplace_scalar_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// flags[e] = 1 when expert e has a contributing hypothesis (p >= PROB_THRESH) on this rank: only those planes receive gradient
__global__ void expert_flags_kernel(const int* contrib, const int* n_contrib, const int* assign32, int E, int* flags) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) flags[e] = 0;
    __syncthreads();
    const int n = *n_contrib;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int e = assign32[contrib[i]];
        if (e >= 0 && e < E) flags[e] = 1;
    }
}
void launch_expert_flags(const int* contrib, const int* n_contrib, const int* assign32, int E, int* flags, cudaStream_t st) {
    expert_flags_kernel<<<1, 256, 0, st>>>(contrib, n_contrib, assign32, E, flags);
}

void launch_add_inplace(float* dst, const float* src, size_t n, cudaStream_t st) {
    if ((((uintptr_t)dst) | ((uintptr_t)src)) & 15) add_inplace_scalar_kernel<<<1184, 256, 0, st>>>(dst, src, n);  // unaligned views
    else add_inplace_kernel<<<1184, 256, 0, st>>>(dst, src, n);
}

}  // namespace esacb200
