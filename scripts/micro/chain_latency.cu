// chain_latency.cu -- microbenchmarks behind the solver design:
// dependent-issue latency of FADD / FMUL / FMNMX chains and the cycles one Gauss-Seidel sweep of
// the car's four revolute joints takes for a single warp (no contention) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I multi_car_racing_b200/csrc -o /tmp/chain scripts/micro/chain_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "solver.cuh"

__global__ void chain_kernel(float* out, long long* cyc, float a, float b, int n) {
    float x = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) x = x + b;           // FADD chain
    }
    long long t1 = clock64();
    float y = a;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 32; ++k) y = y * b;           // FMUL chain
    }
    long long t2 = clock64();
    float z = a;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { z = z * b; z = z + a; }   // FMUL, FADD alternating (no contraction)
    }
    long long t3 = clock64();
    float w = a;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { w = fminf(w + b, a); w = fmaxf(w, -a); }   // FADD, FMNMX, FMNMX
    }
    long long t4 = clock64();
    if (threadIdx.x == 0) { out[0] = x + y + z + w; cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}

template <int PAT>
__global__ void sweep_bench(float* out, long long* cyc, const float* in, int nsweeps) {
    VelState s; JointC J[4]; Masses m;
    const float* p = in;
    for (int i = 0; i < 5; ++i) { s.vx[i] = *p++; s.vy[i] = *p++; s.w[i] = *p++; }
    for (int k = 0; k < 4; ++k) {
        s.jix[k] = *p++; s.jiy[k] = *p++; s.jiz[k] = *p++; s.jmot[k] = *p++;
        JointC& j = J[k];
        j.rAx = *p++; j.rAy = *p++; j.k11 = *p++; j.k12 = *p++; j.k22 = *p++; j.ezx = *p++; j.ezy = *p++; j.ezz = *p++;
        j.det22 = *p++; j.cfx = *p++; j.cfy = *p++; j.cfz = *p++; j.det33 = *p++; j.motorMass = *p++; j.motorSpeed = *p++;
        j.limit = PAT >= 0 ? (((PAT >> k) & 1) ? LIM_LOWER : LIM_INACTIVE) : (int)*p; p++;
    }
    m.mA = *p++; m.iA = *p++; m.mB = *p++; m.iB = *p++; m.maxMotorImpulse = *p++;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < nsweeps; ++it) sweep<PAT>(s, J, m);
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 5; ++i) acc += s.vx[i] + s.vy[i] + s.w[i];
    for (int k = 0; k < 4; ++k) acc += s.jix[k] + s.jiy[k] + s.jiz[k] + s.jmot[k];
    if (threadIdx.x == 0) { out[blockIdx.x] = acc; cyc[blockIdx.x] = t1 - t0; }
}

int main() {
    float* out; long long* cyc; float* in;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 4096 * 8); cudaMalloc(&in, 1024);
    float h[128]; int n = 0;
    for (int i = 0; i < 5; ++i) { h[n++] = 10.0f + i; h[n++] = -3.0f + 0.1f * i; h[n++] = 0.2f * i; }
    const float mA = 0.141643f, iA = 0.054908f, mB = 16.5344f, iB = 134.0626f;
    const float ax[4] = {-1.1f, 1.1f, -1.1f, 1.1f}, ay[4] = {1.6825f, 1.6825f, -1.5575f, -1.5575f};
    for (int k = 0; k < 4; ++k) {
        h[n++] = 0.1f; h[n++] = -0.2f; h[n++] = 0.0f; h[n++] = 0.01f;
        float rAx = ax[k], rAy = ay[k];
        float k11 = mA + mB + rAy * rAy * iA, k12 = -rAy * rAx * iA, k22 = mA + mB + rAx * rAx * iA;
        float ezx = -rAy * iA, ezy = rAx * iA, ezz = iA + iB;
        float det = 1.0f / (k11 * k22 - k12 * k12);
        float cfx = k22 * ezz - ezy * ezy, cfy = ezy * ezx - k12 * ezz, cfz = k12 * ezy - k22 * ezx;
        float det33 = 1.0f / (k11 * cfx + k12 * cfy + ezx * cfz);
        float v[15] = {rAx, rAy, k11, k12, k22, ezx, ezy, ezz, det, cfx, cfy, cfz, det33, 1.0f / (iA + iB), 0.5f};
        for (int q = 0; q < 15; ++q) h[n++] = v[q];
        h[n++] = 0.0f;
    }
    h[n++] = mA; h[n++] = iA; h[n++] = mB; h[n++] = iB; h[n++] = 1.296f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    long long hc[4096];
    for (int rep = 0; rep < 2; ++rep) {
        chain_kernel<<<1, 32>>>(out, cyc, 1.0f, 1.0000001f, 1000);
        cudaMemcpy(hc, cyc, 32, cudaMemcpyDeviceToHost);
    }
    printf("dependent latency (cycles/op): FADD %.2f  FMUL %.2f  FMUL+FADD %.2f  FADD+FMNMX+FMNMX %.2f\n",
           hc[0] / 32000.0, hc[1] / 32000.0, hc[2] / 32000.0, hc[3] / 48000.0);
    const int grids[4] = {1, 148, 592, 2048};
    for (int g = 0; g < 4; ++g) {
        for (int rep = 0; rep < 2; ++rep) { sweep_bench<0><<<grids[g], 32>>>(out, cyc, in, 180); cudaDeviceSynchronize(); }
        cudaMemcpy(hc, cyc, grids[g] * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; double av = 0; for (int i = 0; i < grids[g]; ++i) { mx = hc[i] > mx ? hc[i] : mx; av += hc[i]; }
        printf("sweep<0>  warps=%4d (1 warp/CTA): %.1f cycles/sweep avg, %.1f max\n", grids[g], av / grids[g] / 180.0, mx / 180.0);
    }
    for (int rep = 0; rep < 2; ++rep) { sweep_bench<3><<<148, 32>>>(out, cyc, in, 180); cudaDeviceSynchronize(); }
    cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("sweep<3>  (front limits active): %.1f cycles/sweep\n", hc[0] / 180.0);
    for (int rep = 0; rep < 2; ++rep) { sweep_bench<-1><<<148, 32>>>(out, cyc, in, 180); cudaDeviceSynchronize(); }
    cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("sweep<-1> (generic, limits inactive): %.1f cycles/sweep\n", hc[0] / 180.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
