// ubench.cu -- B200 micro-benchmarks that size the fused kernel's roofline:
//   1. bare tcgen05.mma kind::i8 issue rate (the int8 tensor peak that
//      MEASURED_PEAKS.json does not contain),
//   2. TMEM -> register read rate (tcgen05.ld 32x32b.x32),
//   3. the candidate epilogue instruction mixes (IMAD.WIDE square-accumulate,
//      I2FP+FFMA, FADD+FFMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../dsabeamformer_b200/csrc ubench.cu -o ubench
#include "sm100_ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace dsabf::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---- 1. int8 MMA issue rate ------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *a = sm;                      // 128 x 128 B
    uint8_t *b = sm + 16384;              // N x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x01010101u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_i8(128, N);
        const uint64_t ad = make_desc_k_sw128(smem_u32(a)), bd = make_desc_k_sw128(smem_u32(b));
        for (int i = 0; i < iters; i++) {
            const uint32_t d = tbase + (i & 1) * 256;
#pragma unroll
            for (int k = 0; k < 4; k++) mma_i8_ss(d, ad + 2 * k, bd + 2 * k, idesc, k > 0);
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0, 100);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// ---- 2. TMEM read rate -----------------------------------------------------
__global__ void __launch_bounds__(256, 1) tmem_ld_kernel(int iters, uint32_t *sink)
{
    __shared__ uint32_t tbase;
    if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    const uint32_t taddr = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    uint32_t v[32];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait(v);
            acc ^= v[0] ^ v[31];
        }
    }
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tbase, 512);
}

// ---- 3. epilogue instruction mixes ------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512, 1) alu_kernel(int iters, const int *in, float *sink)
{
    int v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = in[(threadIdx.x + j) & 1023];
    long long s64[4] = {0, 0, 0, 0};
    float sf[4] = {0, 0, 0, 0};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (MODE == 0) {
                asm volatile("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s64[j & 3]) : "r"(v[j]));
            } else if (MODE == 1) {
                float f;
                asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(v[j]));
                asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(sf[j & 3]) : "f"(f));
            } else if (MODE == 2) {
                float f;
                asm volatile("add.rn.f32 %0, %1, 0fCB400000;" : "=f"(f) : "f"(__int_as_float(v[j])));
                asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(sf[j & 3]) : "f"(f));
            } else if (MODE == 3) {
                asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(sf[j & 3]) : "f"(__int_as_float(v[j])));
            }
        }
    }
    float r = sf[0] + sf[1] + sf[2] + sf[3] + (float)(s64[0] + s64[1] + s64[2] + s64[3]);
    if (r == 1.2345f) sink[threadIdx.x] = r;
}

template <typename F>
float time_ms(F f, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s sms %d max_clock_mhz %d\n", prop.name, sms, clk_khz / 1000);

    {   // int8 MMA peak, N = 256 and N = 128
        const int iters = 20000;
        CK(cudaFuncSetAttribute(mma_peak_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
        CK(cudaFuncSetAttribute(mma_peak_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560));
        float ms = time_ms([&] { mma_peak_kernel<256><<<sms, 128, 16384 + 256 * 128 + 1024>>>(iters); });
        double ops = 2.0 * 128 * 256 * 128 * (double)iters * sms;
        printf("i8_mma_peak N=256: %.3f ms  %.1f TOPS  (%.1f cycles per M128N256K32 at max clock)\n", ms, ops / ms * 1e-9,
               ms * 1e-3 * clk_khz * 1e3 / (iters * 4.0));
        ms = time_ms([&] { mma_peak_kernel<128><<<sms, 128, 16384 + 128 * 128 + 1024>>>(iters); });
        ops = 2.0 * 128 * 128 * 128 * (double)iters * sms;
        printf("i8_mma_peak N=128: %.3f ms  %.1f TOPS\n", ms, ops / ms * 1e-9);
    }
    {   // TMEM read
        uint32_t *sink;
        CK(cudaMalloc(&sink, 4096));
        const int iters = 20000;
        float ms = time_ms([&] { tmem_ld_kernel<<<sms, 256>>>(iters, sink); });
        double bytes = (double)iters * 8 * 4096.0 * 8 * sms;   // 8 warps x 8 loads x 4 KB
        printf("tmem_ld x32: %.3f ms  %.1f TB/s chip  %.1f B/clk/SM at max clock  (%.1f accumulators/clk/SM)\n", ms, bytes / ms * 1e-9,
               bytes / sms / (ms * 1e-3 * clk_khz * 1e3), bytes / 4 / sms / (ms * 1e-3 * clk_khz * 1e3));
    }
    {   // ALU mixes
        int *in; float *sink;
        std::vector<int> h(1024);
        for (int i = 0; i < 1024; i++) h[i] = (i * 7919) % 4001 - 2000;
        CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&sink, 4096));
        CK(cudaMemcpy(in, h.data(), 4096, cudaMemcpyHostToDevice));
        const int iters = 4000;
        const char *names[4] = {"IMAD.WIDE sq-acc", "I2FP+FFMA", "FADD(magic)+FFMA", "FFMA only"};
        float ms[4];
        ms[0] = time_ms([&] { alu_kernel<0><<<sms, 512>>>(iters, in, sink); });
        ms[1] = time_ms([&] { alu_kernel<1><<<sms, 512>>>(iters, in, sink); });
        ms[2] = time_ms([&] { alu_kernel<2><<<sms, 512>>>(iters, in, sink); });
        ms[3] = time_ms([&] { alu_kernel<3><<<sms, 512>>>(iters, in, sink); });
        for (int m = 0; m < 4; m++) {
            double elems = (double)iters * 32 * 512 * sms;
            printf("alu %-18s: %.3f ms  %.1f elements/clk/SM at max clock\n", names[m], ms[m], elems / sms / (ms[m] * 1e-3 * clk_khz * 1e3));
        }
    }
    return 0;
}
