// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the
// compress kernel leans on.  One CTA of 128 threads per SM (1 warp per scheduler) or 512 (4 per scheduler);
// every thread runs ITER x 8 independent chains of the op.   nvcc -arch=sm_100a -O3 -o ubench ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int OP> __device__ __forceinline__ void body(float (&f)[8], double (&d)[8], unsigned (&u)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (OP == 0) { f[k] = __double2float_rn(d[k]); d[k] += (double)f[k]; }            // F2F + DADD (see OP 8 for DADD alone)
        if (OP == 1) { u[k] = __popc(u[k]) + 0x9e3779b9u * u[k]; }                         // POPC + IMAD
        if (OP == 2) { unsigned r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(u[k])); u[k] = r + 0x9e3779b9u * u[k]; }
        if (OP == 3) { u[k] = 0x9e3779b9u * u[k] + 12345u; }                               // IMAD alone
        if (OP == 4) { f[k] = __uint2float_rz(u[k]); u[k] = __float_as_uint(f[k]) * 0x9e3779b9u; }   // I2FP + IMAD
        if (OP == 8) { d[k] += 1.0; }
    }
}
template <int OP> __device__ __forceinline__ void body2(float2 (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (OP == 5) v[k] = __fadd2_rn(v[k], v[(k + 1) & 7]);
        if (OP == 6) v[k] = __fmul2_rn(v[k], make_float2(0.999f, 1.001f));
        if (OP == 7) v[k] = __ffma2_rn(v[k], make_float2(0.999f, 1.001f), v[(k + 1) & 7]);
        if (OP == 9) { v[k].x = fmaxf(fmaxf(v[k].x, v[(k + 1) & 7].x), v[(k + 2) & 7].y); }
        if (OP == 10) { v[k].x = __fadd_rn(v[k].x, v[(k + 1) & 7].x); }
    }
}
template <int OP> __global__ void k(float* out, long long* cyc, unsigned seed) {
    float f[8]; double d[8]; unsigned u[8]; float2 v[8];
    for (int k = 0; k < 8; ++k) { f[k] = k; d[k] = 1.0 + threadIdx.x * 1e-3 + k; u[k] = seed + threadIdx.x * 77 + k; v[k] = make_float2(1.f + k, 2.f + threadIdx.x); }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) { if (OP < 5 || OP == 8) body<OP>(f, d, u); else body2<OP>(v); }
    long long t1 = clock64();
    float acc = 0; for (int k = 0; k < 8; ++k) acc += f[k] + (float)d[k] + (float)u[k] + v[k].x + v[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int nt) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, nt>>>(out, cyc, 1u); k<OP><<<148, nt>>>(out, cyc, 2u); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    int warps_per_smsp = nt / 128;
    printf("%-28s nt=%4d  cycles per (8 ops x 1 warp-slot) iteration: %8.2f  -> per op per SMSP: %6.2f\n", name, nt, avg / ITER,
           avg / ITER / 8.0 / warps_per_smsp);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int nt : {128, 512, 1024}) {
        run<0>("F2F.F32.F64 + DADD", nt); run<8>("DADD", nt); run<1>("POPC + IMAD", nt); run<2>("FLO + IMAD", nt);
        run<3>("IMAD", nt); run<4>("I2FP + IMAD", nt); run<5>("FADD2", nt); run<6>("FMUL2", nt); run<7>("FFMA2", nt);
        run<9>("FMNMX3", nt); run<10>("FADD", nt);
    }
    return 0;
}
