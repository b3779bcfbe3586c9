// Pipe-throughput microbenchmark (B200): FFMA vs FFMA2 (f32x2) vs MUFU.{EX2,RSQ,RCP} vs DFMA.
// Sizes the scoring kernel's ceiling: ops per clock per SM, measured with clock64 inside the kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, float seed) {
    float a[ILP];
    float2 b[ILP];
    double d[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x; b[i] = make_float2(a[i], a[i] * 0.5f); d[i] = a[i]; }
    const float m = 0.999f, c = 0.001f;
    const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], m, c);
            if (OP == 1) b[i] = __ffma2_rn(b[i], m2, c2);
            if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 5) d[i] = fma(d[i], 0.999, 0.001);
            if (OP == 6) a[i] = fminf(a[i], m) + c;  // FMNMX + FADD mix
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i].x + b[i].y + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int sms, int threads, double per_instr) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    k<OP><<<sms, threads>>>(out, cyc, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<sms, threads>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    double thread_instr = (double)ITERS * ILP * threads;
    printf("%-8s threads/SM=%4d  %.2f thread-instr/clk/SM  (%.2f ops/clk/SM)  kernel %.3f ms  ~%.0f MHz\n", name, threads,
           thread_instr / avg, thread_instr * per_instr / avg, ms, avg / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s  SMs=%d\n", p.name, p.multiProcessorCount);
    int sms = p.multiProcessorCount;
    for (int threads : {256, 1024}) {
        run<0>("FFMA", sms, threads, 1);
        run<1>("FFMA2", sms, threads, 2);
        run<2>("EX2", sms, threads, 1);
        run<3>("RSQ", sms, threads, 1);
        run<4>("RCP", sms, threads, 1);
        run<5>("DFMA", sms, threads, 1);
        run<6>("FMNMX+ADD", sms, threads, 2);
    }
    return 0;
}
