// ubench_pipes.cu -- issue-rate micro-benchmark of the epilogue's candidate instructions on B200.
// Each thread runs 16 independent dependent-chains of ONE instruction kind (inputs are
// loop-carried, so nothing can be hoisted); 1024 threads per SM, all SMs.
// Output: warp-instructions per clock per SM sub-partition (SMSP) at the max clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 ubench_pipes.cu -o ubench_pipes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Op { I2F_XU, I2F_F64, DFMA_OP, WIDE_CHAIN, MIX_XU3, MIX_XU4, MIX_XU3_WIDE2, MIX_F64, FFMA2_DENORM, FFMA_DENORM, I2FP, IADD3, IMAD, FADD, FADD2, FFMA, FFMA2, LOP3, IMADWIDE, MIX_I2FP_FFMA2, MIX_IMAD_FADD2_FFMA2, MIX_HALF };

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(int iters, const int *in, int *out)
{
    unsigned r[16];
    unsigned long long q[8];
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = in[(threadIdx.x + j) & 1023];
#pragma unroll
    for (int j = 0; j < 8; j++) q[j] = ((unsigned long long)r[2 * j] << 32) | r[2 * j + 1];
    if (OP == FFMA2_DENORM || OP == FFMA_DENORM) {
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = (r[j] & 0x3FFFF) | 0x400;        // denormal bit patterns (< 2^23)
#pragma unroll
        for (int j = 0; j < 8; j++) q[j] = ((unsigned long long)r[2 * j] << 32) | r[2 * j + 1];
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == I2FP) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[j]));
            if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(r[(j + 1) & 15]));
            if (OP == I2F_XU) asm volatile("cvt.rp.f32.s32 %0, %0;" : "+r"(r[j]));
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, 3, 0x4B400000;" : "+r"(r[j]));
            if (OP == FADD) asm volatile("add.rn.f32 %0, %0, 0fCB400000;" : "+r"(r[j]));
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r[j]));
            if (OP == FFMA_DENORM) asm volatile("fma.rn.f32 %0, %0, 0f3F800000, 0f00000000;" : "+r"(r[j]));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, 0xF0F0F0F0, %1, 0x6A;" : "+r"(r[j]) : "r"(i));
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q[j]));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(q[j]));
            if (OP == FFMA2_DENORM) {   // operands stay denormal: x = x * 1.0 + 0 with x = small integer bit pattern
                asm volatile("{.reg .b64 o, z; mov.b64 o, 0x3F8000003F800000; mov.b64 z, 0; fma.rn.f32x2 %0, %0, o, z;}" : "+l"(q[j]));
            }
            if (OP == I2F_F64) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=l"(q[j]) : "r"((unsigned)q[j]));
            if (OP == DFMA_OP) asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+l"(q[j]));
            if (OP == WIDE_CHAIN) {           // two dependent 64-bit accumulates per accumulator, like the exact epilogue
                asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(q[j]) : "r"(r[2 * j]));
                asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(q[j]) : "r"(r[2 * j + 1]));
            }
            if (OP == MIX_XU3 || OP == MIX_XU4 || OP == MIX_XU3_WIDE2) {   // 16 elements: some converted on the XU pipe (I2F.RP), rest I2FP
                const int nxu = OP == MIX_XU4 ? 4 : 3;
                const bool wide = OP == MIX_XU3_WIDE2 && j == 7;           // one pair of 8 squared as integers
                if (wide) {
                    asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(q[j]) : "r"(r[2 * j]));
                    asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(q[j]) : "r"(r[2 * j + 1]));
                } else {
                    if (j < nxu) asm volatile("cvt.rp.f32.s32 %0, %0;" : "+r"(r[2 * j]));
                    else         asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j]));
                    asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j + 1]));
                    asm volatile("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
                }
            }
            if (OP == MIX_F64) {              // 16 elements: 4 through I2F.F64 + DFMA, 12 through I2FP + FFMA2
                if (j < 2) {
                    unsigned long long d0, d1;
                    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=l"(d0) : "r"(r[2 * j]));
                    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=l"(d1) : "r"(r[2 * j + 1]));
                    asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+l"(q[j]) : "l"(d0));
                    asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+l"(q[j]) : "l"(d1));
                    r[2 * j] ^= (unsigned)q[j]; r[2 * j + 1] += i;
                } else {
                    asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j]));
                    asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j + 1]));
                    asm volatile("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
                }
            }
            if (OP == IMADWIDE) asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(q[j]) : "r"(r[j]));
            if (OP == MIX_I2FP_FFMA2) {      // the current epilogue: 2 I2FP + 1 FFMA2 per 2 elements
                asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j]));
                asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j + 1]));
                asm volatile("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
            }
            if (OP == MIX_IMAD_FADD2_FFMA2) { // integer bias add on the FMA pipe + packed float un-bias + packed FMA
                asm volatile("mad.lo.s32 %0, %0, 1, 0x4B400000;" : "+r"(r[2 * j]));
                asm volatile("mad.lo.s32 %0, %0, 1, 0x4B400000;" : "+r"(r[2 * j + 1]));
                asm volatile("{.reg .b64 t, m; mov.b64 t, {%1, %2}; mov.b64 m, 0xCB400000CB400000; add.rn.f32x2 t, t, m; fma.rn.f32x2 %0, t, t, %0;}"
                             : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
            }
            if (OP == MIX_HALF) {             // alternate pairs between the two conversions
                if (j & 1) {
                    asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j]));
                    asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[2 * j + 1]));
                    asm volatile("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
                } else {
                    asm volatile("mad.lo.s32 %0, %0, 1, 0x4B400000;" : "+r"(r[2 * j]));
                    asm volatile("mad.lo.s32 %0, %0, 1, 0x4B400000;" : "+r"(r[2 * j + 1]));
                    asm volatile("{.reg .b64 t, m; mov.b64 t, {%1, %2}; mov.b64 m, 0xCB400000CB400000; add.rn.f32x2 t, t, m; fma.rn.f32x2 %0, t, t, %0;}"
                                 : "+l"(q[j]) : "r"(r[2 * j]), "r"(r[2 * j + 1]));
                }
            }
        }
    }
    unsigned acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) acc ^= r[j];
#pragma unroll
    for (int j = 0; j < 8; j++) acc ^= (unsigned)q[j] ^ (unsigned)(q[j] >> 32);
    if (acc == 0x12345u) out[threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, double inst_per_iter, double elems_per_iter, int sms, int clk_khz, const int *in, int *out)
{
    const int iters = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<OP><<<sms, 1024>>>(iters, in, out);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k<OP><<<sms, 1024>>>(iters, in, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * clk_khz * 1e3;
    const double warp_inst = inst_per_iter * iters * 32.0;      // 32 warps per SM
    printf("%-22s %.3f ms  %.3f warp-inst/clk/SMSP", name, best, warp_inst / 4.0 / cycles);
    if (elems_per_iter > 0) printf("   %.1f elements/clk/SM", elems_per_iter * iters * 1024.0 / cycles);
    printf("\n");
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    int *in, *out; CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, 4096));
    int h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i * 7919) % 4001 - 2000;
    CK(cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice));
    printf("device %s, %d SMs, max clock %d MHz (rates assume the max clock)\n", prop.name, sms, clk_khz / 1000);
    run<I2F_XU>("I2F.RP (XU pipe)", 16, 0, sms, clk_khz, in, out);
    run<I2F_F64>("I2F.F64.S32", 8, 0, sms, clk_khz, in, out);
    run<DFMA_OP>("DFMA", 8, 0, sms, clk_khz, in, out);
    run<WIDE_CHAIN>("IMAD.WIDE chain x2", 16, 16, sms, clk_khz, in, out);
    run<MIX_XU3>("3 I2F.RP+13 I2FP+8 FFMA2", 24, 16, sms, clk_khz, in, out);
    run<MIX_XU4>("4 I2F.RP+12 I2FP+8 FFMA2", 24, 16, sms, clk_khz, in, out);
    run<MIX_XU3_WIDE2>("3 XU+11 I2FP+7 FFMA2+2 WIDE", 23, 16, sms, clk_khz, in, out);
    run<MIX_F64>("4 F64+12 I2FP+6 FFMA2", 26, 16, sms, clk_khz, in, out);
    run<FFMA_DENORM>("FFMA denormal operand", 16, 0, sms, clk_khz, in, out);
    run<FFMA2_DENORM>("FFMA2 denormal operand", 8, 0, sms, clk_khz, in, out);
    run<I2FP>("I2FP.F32.S32", 16, 0, sms, clk_khz, in, out);
    run<IADD3>("IADD (add.s32 imm)", 16, 0, sms, clk_khz, in, out);
    run<IMAD>("IMAD (mad.lo imm)", 16, 0, sms, clk_khz, in, out);
    run<LOP3>("LOP3", 16, 0, sms, clk_khz, in, out);
    run<FADD>("FADD", 16, 0, sms, clk_khz, in, out);
    run<FFMA>("FFMA", 16, 0, sms, clk_khz, in, out);
    run<FADD2>("FADD2 (add.f32x2)", 8, 0, sms, clk_khz, in, out);
    run<FFMA2>("FFMA2 (fma.f32x2)", 8, 0, sms, clk_khz, in, out);
    run<IMADWIDE>("IMAD.WIDE sq-acc", 8, 8, sms, clk_khz, in, out);
    run<MIX_I2FP_FFMA2>("2xI2FP+FFMA2", 24, 16, sms, clk_khz, in, out);
    run<MIX_IMAD_FADD2_FFMA2>("2xIMAD+FADD2+FFMA2", 32, 16, sms, clk_khz, in, out);
    run<MIX_HALF>("half/half mix", 28, 16, sms, clk_khz, in, out);
    return 0;
}
