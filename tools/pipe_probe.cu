// Issue-rate probe for the sm_100a pipes the compute-bound kernels lean on (MUFU ex2, packed FP32, FMNMX):
// cycles per warp-instruction and SM sub-partition, for 1 / 2 / 4 warps per sub-partition, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && tools/pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(float *out, long long *cyc, int iters) {
    float v[8];
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = threadIdx.x * 1e-3f + i; p[i] = make_float2(v[i], v[i] + 0.5f); }
    const float2 c1 = make_float2(1.0001f + out[0], 0.9999f), c2 = make_float2(1e-3f, 2e-3f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) p[i] = __ffma2_rn(p[i], c1, c2);                       // three distinct register operands
            if (OP == 3) p[i] = __ffma2_rn(p[i], make_float2(1.0001f, 1.0001f), c2);  // one immediate
            if (OP == 4) p[i] = __fadd2_rn(p[i], make_float2(0.5f, 0.5f));
            if (OP == 5) v[i] = fminf(v[i], 29.f + it);
            if (OP == 6) v[i] = fmaf(v[i], c1.x, c2.x);
            if (OP == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); p[i] = __ffma2_rn(p[i], c1, c2); p[(i + 1) & 7] = __fmul2_rn(p[(i + 1) & 7], c1); }
            if (OP == 8) p[i] = __fmul2_rn(p[i], c1);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i] + p[i].x + p[i].y;
    out[1 + blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int per_iter, float *out, long long *cyc) {
    const int iters = 4096;
    printf("%-34s", name);
    for (int warps : {4, 8, 16, 32}) {
        probe<OP><<<148, warps * 32>>>(out, cyc, iters);
        probe<OP><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        // warp-instructions issued per sub-partition = (warps / 4) * iters * per_iter
        printf("  %2d w/SMSP: %5.2f clk/inst", warps / 4, (double)c / ((warps / 4.0) * iters * per_iter));
    }
    printf("\n");
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * (1 + 148 * 1024)); cudaMemset(out, 0, sizeof(float) * (1 + 148 * 1024));
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("MUFU.EX2", 8, out, cyc);
    run<2>("FFMA2 (3 register operands)", 8, out, cyc);
    run<3>("FFMA2 (immediate multiplier)", 8, out, cyc);
    run<4>("FADD2 (immediate)", 8, out, cyc);
    run<8>("FMUL2 (2 register operands)", 8, out, cyc);
    run<5>("FMNMX", 8, out, cyc);
    run<6>("FFMA (scalar)", 8, out, cyc);
    run<7>("EX2 + FFMA2 + FMUL2 (per triple)", 8, out, cyc);
    return 0;
}
