// Throughput of the instruction classes the sweep kernel leans on (B200, sm_100a): per-SM warp-instructions
// per clock for I2F, MUFU, FFMA, IMAD, LOP3, PRMT, shared loads / stores / atomics, SHFL.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, long long* clk) {
    extern __shared__ uint32_t sm[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    uint32_t x[UNROLL];
    float f[UNROLL];
    for (int j = 0; j < UNROLL; ++j) { x[j] = seed + threadIdx.x * 7 + j; f[j] = (float)(x[j] & 1023) + 1.5f; }
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + lane * 4u + warp * 128u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            if (OP == 0) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[j]) : "r"(x[j])); x[j] += __float_as_uint(f[j]) & 1; }
            if (OP == 1) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[j])); }
            if (OP == 2) { asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[j])); }
            if (OP == 3) { asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(x[j])); }
            if (OP == 4) { asm volatile("lop3.b32 %0, %0, %1, 0x4b000000, 0x96;" : "+r"(x[j]) : "r"(seed)); }
            if (OP == 5) { asm volatile("prmt.b32 %0, %0, %1, 0x4441;" : "+r"(x[j]) : "r"(seed)); }
            if (OP == 6) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + ((x[j] & 7u) << 12))); x[j] += v; }
            if (OP == 7) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(sbase + (j << 12)), "r"(x[j])); }
            if (OP == 8) { asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(sbase + (j << 12)), "r"(x[j])); }
            if (OP == 9) { asm volatile("red.shared.add.f32 [%0], %1;" :: "r"(sbase + (j << 12)), "f"(f[j])); }
            if (OP == 10) { asm volatile("shfl.sync.idx.b32 %0, %0, 3, 0x1f, 0xffffffff;" : "+r"(x[j])); }
            if (OP == 11) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sbase + ((x[j] & 7u) << 12) + (j & 3))); x[j] += v; }
            if (OP == 12) { asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(x[j]) : "f"(f[j])); f[j] += 1.0f; }
            if (OP == 13) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j])); }
            if (OP == 14) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(seed)); }
            if (OP == 15) { asm volatile("add.f32 %0, %0, %1;" : "+f"(f[j]) : "f"(1.0f)); }
            if (OP == 16) { asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(seed)); }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int j = 0; j < UNROLL; ++j) acc += x[j] + __float_as_uint(f[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    k<OP><<<148, threads, 160 * 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    k<OP><<<148, threads, 160 * 1024>>>(out, 12345u, clk);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double winstr = (double)ITERS * UNROLL * (threads / 32);
    printf("%-22s threads %4d: %8.3f warp-instr/clk/SM  (%6.2f clk per warp-instr per SM) %s\n", name, threads, winstr / avg, avg / winstr,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out); cudaFree(clk);
}

int main() {
    for (int threads : {1024, 512}) {
        run<0>("I2F.F32.S32 (+LOP,IADD)", threads);
        run<1>("MUFU.LG2", threads);
        run<13>("MUFU.EX2", threads);
        run<12>("F2I (+FADD)", threads);
        run<2>("FFMA", threads);
        run<15>("FADD", threads);
        run<3>("IMAD", threads);
        run<14>("IMAD.HI", threads);
        run<16>("IADD", threads);
        run<4>("LOP3", threads);
        run<5>("PRMT", threads);
        run<6>("LDS.32 (+LOP,SHF,IADD)", threads);
        run<11>("LDS.U8 (+LOP,SHF,IADD)", threads);
        run<7>("STS.32", threads);
        run<8>("RED.shared.add.s32", threads);
        run<9>("RED.shared.add.f32", threads);
        run<10>("SHFL.IDX", threads);
    }
    return 0;
}
