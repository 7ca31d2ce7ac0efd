// int_peak.cu — measures the integer-pipe ceilings the rooflines in DESIGN.md use ("the build must
// measure it on the box with an IMAD-only microbenchmark"):
//   1. IMAD.WIDE.U32 issue rate (the instruction the Montgomery multiply compiles to)
//   2. 32-bit IMAD issue rate
//   3. Fr Montgomery multiplications per second with 1/2/4 independent chains per thread
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I delay-encryption-in-halo2_b200/csrc -I tools tools/int_peak.cu -o tools/int_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
#include "mul29_experiment.cuh"
using namespace de;

template <int ILP>
__global__ void k_imad_wide(unsigned long long* out, unsigned int a, unsigned int b, int iters) {
    unsigned long long acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"((unsigned int)acc[i] ^ b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_imad32(unsigned int* out, unsigned int a, unsigned int b, int iters) {
    unsigned int acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++)
                asm volatile("mad.lo.u32 %0, %1, %0, %2;" : "+r"(acc[i]) : "r"(a + i), "r"(b));
    }
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 pipe: DFMA issue rate, alone and interleaved with IMAD.WIDE in the same thread (do the two pipes overlap?)
template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_mixed(double* out, double a, double b, unsigned int ua, int iters) {
    double acc[ILP];
    unsigned long long wacc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        acc[i] = threadIdx.x + i;
        wacc[i] = threadIdx.x + 7 * i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(wacc[i]) : "r"((unsigned int)wacc[i]), "r"(ua));  // operand changes: not hoistable
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + (double)wacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// IMAD.WIDE.U32 with a 64-bit accumulator and NO carry flag; one multiplicand changes every step (by an ALU add, on the other
// pipe) so that the product cannot be hoisted out of the loop
template <int ILP>
__global__ void k_imad_wide_clean(unsigned long long* out, unsigned int a, int iters) {
    unsigned long long acc[ILP];
    unsigned int x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        acc[i] = threadIdx.x + i;
        x[i] = a + 3 * i + threadIdx.x;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(a));
                x[i] += 0x9e3779b9u;
            }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the multiplier's own instruction pattern: carry chains of mad.lo.cc / madc.hi.cc pairs (ptxas fuses each pair into one
// IMAD.WIDE.U32.X with predicate carries); ops counted as fused wide multiply-adds
template <int ILP>
__global__ void k_carry_chain(unsigned int* out, unsigned int a0, unsigned int b, int iters) {
    unsigned int acc[ILP][8];
    unsigned int a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = a0 + 77u * j + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) de::cmad_n(acc[i], a, b + i + r);  // 4 fused wide mads per call
    }
    unsigned int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_frmul(Fr* out, const Fr* in, int iters) {
    Fr x[ILP];
    Fr y = load(&in[threadIdx.x & 31]);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = load(&in[(threadIdx.x + i) & 63]);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = mul(x[i], y);
    }
    Fr s = x[0];
#pragma unroll
    for (int i = 1; i < ILP; i++) s = add(s, x[i]);
    store(&out[blockIdx.x * blockDim.x + threadIdx.x], s);
}

template <int ILP>
__global__ void k_frmul29(Fr* out, const Fr* in, int iters) {
    Fr x[ILP];
    Fr y = load(&in[threadIdx.x & 31]);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = load(&in[(threadIdx.x + i) & 63]);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = mul29(x[i], y);
    }
    Fr s = x[0];
#pragma unroll
    for (int i = 1; i < ILP; i++) s = add(s, x[i]);
    store(&out[blockIdx.x * blockDim.x + threadIdx.x], s);
}

template <class F>
static float time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    void* buf; cudaMalloc(&buf, (size_t)sms * 8 * 1024 * 32);
    Fr* in; cudaMalloc((void**)&in, 64 * sizeof(Fr));
    Fr h[64];
    for (int i = 0; i < 64; i++) for (int k = 0; k < 8; k++) h[i].l[k] = (k == 7) ? (0x0fffffffu - i) : (0x9e3779b9u * (i * 8 + k + 1));
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 4096;
    for (int tpb : {256, 512, 1024}) {
        int blocks = sms * (2048 / tpb);
        double n_threads = (double)blocks * tpb;
        float ms = time_ms([&] { k_imad_wide<8><<<blocks, tpb>>>((unsigned long long*)buf, 12345u, 0x9e3779b9u, iters); });
        printf("{\"bench\": \"imad_wide_u32\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, n_threads * iters * 64 / ms / 1e6);
        ms = time_ms([&] { k_imad32<8><<<blocks, tpb>>>((unsigned int*)buf, 12345u, 0x9e3779b9u, iters); });
        printf("{\"bench\": \"imad_lo_u32\", \"threads_per_sm\": 2048, \"tpb\": %d, \"gops\": %.1f}\n", tpb, n_threads * iters * 64 / ms / 1e6);
    }
    {
        int tpb = 256, blocks = sms * 8;
        double n_threads = (double)blocks * tpb;
        float ms = time_ms([&] { k_dfma<8><<<blocks, tpb>>>((double*)buf, 1.0000001, 0.5, iters); });
        printf("{\"bench\": \"dfma_f64\", \"tpb\": %d, \"gops\": %.1f}\n", tpb, n_threads * iters * 64 / ms / 1e6);
        ms = time_ms([&] { k_imad_wide_clean<8><<<blocks, tpb>>>((unsigned long long*)buf, 12345u, iters); });
        printf("{\"bench\": \"imad_wide_u32_clean\", \"tpb\": %d, \"gops\": %.1f}\n", tpb, n_threads * iters * 64 / ms / 1e6);
        ms = time_ms([&] { k_mixed<8><<<blocks, tpb>>>((double*)buf, 1.0000001, 0.5, 12345u, iters); });
        printf("{\"bench\": \"dfma_plus_imad_wide_interleaved\", \"tpb\": %d, \"gops_each\": %.1f}\n", tpb, n_threads * iters * 64 / ms / 1e6);
    }
    {
        int tpb = 256, blocks = sms * 8;
        double n_threads = (double)blocks * tpb;
        float ms = time_ms([&] { k_carry_chain<4><<<blocks, tpb>>>((unsigned int*)buf, 12345u, 0x9e3779b9u, iters); });
        printf("{\"bench\": \"carry_chain_wide_mads\", \"tpb\": %d, \"gops\": %.1f}\n", tpb, n_threads * iters * 2 * 4 * 4 / ms / 1e6);
    }
    const int fiters = 512;
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int tpb = 128;
        int blocks = sms * warps_per_sm * 32 / tpb;
        double n_threads = (double)blocks * tpb;
        float ms = time_ms([&] { k_frmul<1><<<blocks, tpb>>>((Fr*)buf, in, fiters); });
        printf("{\"bench\": \"fr_mul\", \"ilp\": 1, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, n_threads * fiters * 1 / ms / 1e6);
        ms = time_ms([&] { k_frmul<2><<<blocks, tpb>>>((Fr*)buf, in, fiters); });
        printf("{\"bench\": \"fr_mul\", \"ilp\": 2, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, n_threads * fiters * 2 / ms / 1e6);
        ms = time_ms([&] { k_frmul<4><<<blocks, tpb>>>((Fr*)buf, in, fiters); });
        printf("{\"bench\": \"fr_mul\", \"ilp\": 4, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, n_threads * fiters * 4 / ms / 1e6);
        ms = time_ms([&] { k_frmul29<1><<<blocks, tpb>>>((Fr*)buf, in, fiters); });
        printf("{\"bench\": \"fr_mul29\", \"ilp\": 1, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, n_threads * fiters * 1 / ms / 1e6);
        ms = time_ms([&] { k_frmul29<2><<<blocks, tpb>>>((Fr*)buf, in, fiters); });
        printf("{\"bench\": \"fr_mul29\", \"ilp\": 2, \"warps_per_sm\": %d, \"gmul_s\": %.2f}\n", warps_per_sm, n_threads * fiters * 2 / ms / 1e6);
    }
    return 0;
}
