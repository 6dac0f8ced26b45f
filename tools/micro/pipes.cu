// Throughput of the pipes the G2P kernel leans on: FFMA, DFMA, f32<->f64 conversions, MUFU.RCP (per SM and clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
    double da[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x, da[i] = a[i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
            if (MODE == 1) da[i] = fma(da[i], 1.0001, 0.5);
            if (MODE == 2) a[i] = (float)((double)a[i] + 1.0) ; // F2F.F64.F32 + DADD + F2F.F32.F64
            if (MODE == 3) a[i] = __frcp_rn(a[i]) ;
            if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 5) a[i] = rintf(a[i] * 1.5f);
            if (MODE == 6) a[i] = (float)(int)(a[i] * 1.5f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + (float)da[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_iter) {
    int dev_sms = 0, khz = 0;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* out;
    const int blocks = dev_sms * 8, threads = 256, iters = 4096;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double thread_ops = (double)blocks * threads * iters * 8 * ops_per_iter;
    printf("%-28s %8.3f ms  %7.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, ms, thread_ops / (ms * 1e-3) / (khz * 1e3) / dev_sms, khz / 1000);
    cudaFree(out);
}
int main() {
    run<0>("FFMA", 1);
    run<1>("DFMA", 1);
    run<2>("F2F+DADD+F2F", 1);
    run<3>("__frcp_rn", 1);
    run<4>("rcp.approx (MUFU.RCP)", 1);
    run<5>("FMUL+FRND", 1);
    run<6>("FMUL+F2I+I2F", 1);
    return 0;
}
