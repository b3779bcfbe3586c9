// Reprojection loss of the expert refinement stage, forward and backward in one pass.
//
// ref_expert.py:103-146 projects every predicted scene coordinate with the ground-truth pose, takes the pixel distance to
// the cell centre, clamps it to [0, 100] px, applies an L1 / square-root robust loss and lets autograd walk back through
// six elementwise tensors.  It is the scoring kernel's projection with a different epilogue, and purely HBM-bound:
// 12 B read + 12 B written per cell.  One launch covers a batch of maps; each thread handles 4 consecutive cells with
// 128-bit loads/stores; the per-image loss is summed in fp64 in a fixed order (block partials, last block adds them up),
// so the result does not depend on scheduling.
//
//   eye = gtPose^-1[0:3,:] * [X;1]                                   ref_expert.py:127-132
//   px  = K * eye;  px[2].clamp_(min=0.1);  px = px[0:2] / px[2]     :135-137   (numerator keeps the unclamped depth)
//   err = ||px - (cell centre - pad)||, clamp(0, 100)                :140-142
//   loss = (sum_{err<=cut} err + sum_{err>cut} sqrt(cut*err)) / N    :144-148
#include "esac_internal.h"

namespace esacb200 {

namespace {

constexpr int kThreads = 256;
constexpr int kCellsPerThread = 4;

struct CellOut { float gx, gy, gz, loss; };

// Arithmetic in fp32 like the torch ops it replaces.  m = rows of the 3x4 world->camera matrix.
__device__ __forceinline__ CellOut reproj_cell(float X, float Y, float Z, const float* __restrict__ m, float f, float cx,
                                               float cy, float tx, float ty, float cut, float max_err, float min_depth,
                                               float inv_n) {
    const float xc = fmaf(m[0], X, fmaf(m[1], Y, fmaf(m[2], Z, m[3])));
    const float yc = fmaf(m[4], X, fmaf(m[5], Y, fmaf(m[6], Z, m[7])));
    const float zc = fmaf(m[8], X, fmaf(m[9], Y, fmaf(m[10], Z, m[11])));
    const float nu = fmaf(f, xc, cx * zc);
    const float nv = fmaf(f, yc, cy * zc);
    const bool open = zc >= min_depth;            // clamp_ passes the gradient where it did not clamp
    const float zz = open ? zc : min_depth;
    const float iz = 1.f / zz;
    const float u = nu * iz, v = nv * iz;
    const float du = u - tx, dv = v - ty;
    const float err = sqrtf(du * du + dv * dv);
    CellOut o;
    const float e = fminf(err, max_err);          // err >= 0 always
    const bool l1 = e <= cut;
    o.loss = l1 ? e : sqrtf(cut * e);
    // d loss / d err; the clamp's gradient is 1 on [0, max_err] (bounds included) and 0 outside
    float g = (err <= max_err) ? (l1 ? 1.f : 0.5f * cut / sqrtf(cut * e)) : 0.f;
    g *= inv_n;
    // norm backward: x / ||x||, defined as 0 at the origin
    const float in = err > 0.f ? 1.f / err : 0.f;
    const float gu = g * du * in, gv = g * dv * in;
    // u = nu / zz:  du/dxc = f/zz, du/dzc = cx/zz - [open] nu/zz^2   (same for v)
    const float gxc = gu * f * iz;
    const float gyc = gv * f * iz;
    float gzc = (gu * cx + gv * cy) * iz;
    if (open) gzc -= (gu * nu + gv * nv) * iz * iz;
    o.gx = fmaf(m[0], gxc, fmaf(m[4], gyc, m[8] * gzc));
    o.gy = fmaf(m[1], gxc, fmaf(m[5], gyc, m[9] * gzc));
    o.gz = fmaf(m[2], gxc, fmaf(m[6], gyc, m[10] * gzc));
    if (!(err == err)) o.loss = 0.f;              // NaN: in neither branch of the masked sums (its gradient stays NaN, as in torch)
    return o;
}

// grid = (blocks_per_image, B).  img[b] = 12 matrix entries, padX, padY, 2 unused.
template <bool VEC>
__global__ void __launch_bounds__(kThreads) reproj_kernel(const float* __restrict__ coords, float* __restrict__ grads,
                                                          const float* __restrict__ img, int N, int W, float sub, float f,
                                                          float cx, float cy, float cut, float max_err, float min_depth,
                                                          double* __restrict__ partial, unsigned* __restrict__ tickets,
                                                          double* __restrict__ losses) {
    __shared__ float m[16];
    __shared__ double warp_sum[kThreads / 32];
    __shared__ bool last;
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x] = img[b * 16 + threadIdx.x];
    __syncthreads();
    const float* px = coords + (size_t)b * 3 * N;
    const float* py = px + N;
    const float* pz = py + N;
    float* gx = grads ? grads + (size_t)b * 3 * N : nullptr;
    const float inv_n = 1.f / (float)N;
    const float half = sub * 0.5f;
    const float padX = m[12], padY = m[13];
    double acc = 0.;
    const int per_block = kThreads * kCellsPerThread;
    for (int base = blockIdx.x * per_block; base < N; base += gridDim.x * per_block) {
        const int p0 = base + threadIdx.x * kCellsPerThread;
        if (p0 >= N) continue;
        float X[4], Y[4], Z[4];
        int n = min(kCellsPerThread, N - p0);
        if (VEC) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(px + p0));
            const float4 c = __ldcs(reinterpret_cast<const float4*>(py + p0));
            const float4 d = __ldcs(reinterpret_cast<const float4*>(pz + p0));
            X[0] = a.x; X[1] = a.y; X[2] = a.z; X[3] = a.w;
            Y[0] = c.x; Y[1] = c.y; Y[2] = c.z; Y[3] = c.w;
            Z[0] = d.x; Z[1] = d.y; Z[2] = d.z; Z[3] = d.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = i < n;
                X[i] = ok ? px[p0 + i] : 0.f;
                Y[i] = ok ? py[p0 + i] : 0.f;
                Z[i] = ok ? pz[p0 + i] : 1.f;
            }
        }
        int y = p0 / W, x = p0 - y * W;
        float ox[4], oy[4], oz[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float tx = fmaf((float)x, sub, half) - padX;
            const float ty = fmaf((float)y, sub, half) - padY;
            const CellOut o = reproj_cell(X[i], Y[i], Z[i], m, f, cx, cy, tx, ty, cut, max_err, min_depth, inv_n);
            if (i < n) acc += (double)o.loss;
            ox[i] = o.gx; oy[i] = o.gy; oz[i] = o.gz;
            if (++x == W) { x = 0; ++y; }
        }
        if (gx) {
            if (VEC) {
                __stcs(reinterpret_cast<float4*>(gx + p0), make_float4(ox[0], ox[1], ox[2], ox[3]));
                __stcs(reinterpret_cast<float4*>(gx + N + p0), make_float4(oy[0], oy[1], oy[2], oy[3]));
                __stcs(reinterpret_cast<float4*>(gx + 2 * (size_t)N + p0), make_float4(oz[0], oz[1], oz[2], oz[3]));
            } else {
                for (int i = 0; i < n; ++i) {
                    gx[p0 + i] = ox[i];
                    gx[N + p0 + i] = oy[i];
                    gx[2 * (size_t)N + p0 + i] = oz[i];
                }
            }
        }
    }
    // block sum in a fixed order, then the last block of the image adds the partials in block order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.;
        for (int w = 0; w < kThreads / 32; ++w) s += warp_sum[w];
        partial[(size_t)b * gridDim.x + blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(&tickets[b], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.;
        for (unsigned k = 0; k < gridDim.x; ++k) s += __ldcg(&partial[(size_t)b * gridDim.x + k]);
        losses[b] = s / (double)N;
        tickets[b] = 0;  // ready for the next launch
    }
}

}  // namespace

int reproj_blocks_per_image(int N, int B, int sm_count) {
    const int per_block = kThreads * kCellsPerThread;
    int need = (N + per_block - 1) / per_block;
    // enough CTAs in flight to cover the HBM latency (8 resident CTAs per SM), never more than the work
    int want = (sm_count * 8 + B - 1) / B;
    if (want < 1) want = 1;
    return need < want ? need : want;
}

void launch_reproj(const float* coords, float* grads, const float* img, int B, int N, int W, float sub, float f, float cx,
                   float cy, float cut, float max_err, float min_depth, int blocks_per_image, double* partial,
                   unsigned* tickets, double* losses, cudaStream_t stream) {
    const dim3 grid(blocks_per_image, B);
    const bool vec = (N % 4 == 0) && ((uintptr_t)coords % 16 == 0) && (!grads || (uintptr_t)grads % 16 == 0);
    if (vec)
        reproj_kernel<true><<<grid, kThreads, 0, stream>>>(coords, grads, img, N, W, sub, f, cx, cy, cut, max_err, min_depth,
                                                           partial, tickets, losses);
    else
        reproj_kernel<false><<<grid, kThreads, 0, stream>>>(coords, grads, img, N, W, sub, f, cx, cy, cut, max_err, min_depth,
                                                            partial, tickets, losses);
}

}  // namespace esacb200
