/* dsabf_compat.cuh -- source-level compatibility with the reference's CUDA header.
 *
 * A translation unit written against /root/reference/src/beamformer.cuh keeps compiling when it includes this file
 * instead: the same names with the same argument meaning and launch contract, implemented here from scratch.
 *
 *   gpuErrchk / gpuAssert          beamformer.cuh:19-29   print "GPUassert: <msg> <file> <line>" and exit(code)
 *   gpuBLASchk                     beamformer.cuh:32-36   the reference prints and carries on; this one prints and exits
 *   START_TIMER / STOP_RECORD_TIMER, globals start / stop   beamformer.cuh:43-58   CUDA-event stopwatch, milliseconds
 *   expand_input<<<g, b>>>(in, out, input_size)            beamformer.cuh:66-109  4+4-bit -> interleaved int8 (re, im);
 *                                  any grid: threads stride over the 32-bit words of the input
 *   detect_dimGrid, detect_dimBlock, detect_sum<<<>>>(in, n_avg, out)            beamformer.cuh:112-155
 *                                  grid = (outputs per chunk, channels), block = beams; out[o][f][b] = sum of n_avg |C|^2
 *   print_data, print_data_scalar  beamformer.cuh:158-169
 *   CUDA_select_GPU(name)          beamformer.cuh:171-192 first device whose name matches, else the current one stays
 *
 * These are the UNFUSED steps, for code that calls them one by one (and for checking them against each other); the
 * product path is the fused kernel behind dsabf_enqueue_block / dsabf_beamform_device (include/dsabf.h), which
 * replaces expand_input + cublasGemmStridedBatchedEx + detect_sum together.  Include after the shape macros
 * (N_BEAMS, N_FREQUENCIES, N_OUTPUTS_PER_GEMM of beamformer.hh) if the defaults below are not wanted. */
#ifndef DSABF_COMPAT_CUH
#define DSABF_COMPAT_CUH

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include <cuComplex.h>
#include <cuda_runtime.h>

#ifndef N_BEAMS
#define N_BEAMS 256
#endif
#ifndef N_FREQUENCIES
#define N_FREQUENCIES 256
#endif
#ifndef N_OUTPUTS_PER_GEMM
#define N_OUTPUTS_PER_GEMM 8
#endif

/* ---- error checking ---- */
#define gpuErrchk(ans) { gpuAssert((ans), __FILE__, __LINE__); }
inline void gpuAssert(cudaError_t code, const char *file, int line, bool abort = true)
{
    if (code == cudaSuccess) return;
    fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(code), file, line);
    if (abort) exit((int)code);
}
/* errval: a cublasStatus_t (0 = success).  Unlike the reference this does not continue after a failed call. */
inline void gpuBLASchk(int errval)
{
    if (errval == 0) return;
    std::cerr << "Failed BLAS call, error code " << errval << std::endl;
    exit(errval);
}

/* ---- stopwatch ---- */
static cudaEvent_t start;
static cudaEvent_t stop;
#define START_TIMER() {                       \
    gpuErrchk(cudaEventCreate(&start));       \
    gpuErrchk(cudaEventCreate(&stop));        \
    gpuErrchk(cudaEventRecord(start));        \
}
#define STOP_RECORD_TIMER(name) {                          \
    gpuErrchk(cudaEventRecord(stop));                      \
    gpuErrchk(cudaEventSynchronize(stop));                 \
    gpuErrchk(cudaEventElapsedTime(&name, start, stop));   \
    gpuErrchk(cudaEventDestroy(start));                    \
    gpuErrchk(cudaEventDestroy(stop));                     \
}

/* ---- unpack: byte (re<<4)|(im&15) -> int8 re, int8 im ---- */
__device__ __forceinline__ unsigned dsabf_compat_spread(unsigned two_bytes)
{
    /* two packed bytes b0, b1 (low half of the argument) -> re0, im0, re1, im1 as four int8 lanes */
    const int b0 = (int)(signed char)(two_bytes & 0xFFu), b1 = (int)(signed char)((two_bytes >> 8) & 0xFFu);
    const unsigned re0 = (unsigned)(b0 >> 4) & 0xFFu, im0 = (unsigned)((int)(signed char)(b0 << 4) >> 4) & 0xFFu;
    const unsigned re1 = (unsigned)(b1 >> 4) & 0xFFu, im1 = (unsigned)((int)(signed char)(b1 << 4) >> 4) & 0xFFu;
    return re0 | (im0 << 8) | (re1 << 16) | (im1 << 24);
}
__global__ void expand_input(char const *__restrict__ input, char *output, int input_size)
{
    const unsigned *in = reinterpret_cast<const unsigned *>(input);
    uint2 *out = reinterpret_cast<uint2 *>(output);
    const size_t n_words = (size_t)input_size / sizeof(unsigned);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned w = in[i];
        out[i] = make_uint2(dsabf_compat_spread(w & 0xFFFFu), dsabf_compat_spread(w >> 16));
    }
}

/* ---- detect + sum ---- */
static dim3 detect_dimGrid(N_OUTPUTS_PER_GEMM, N_FREQUENCIES, 1);
static dim3 detect_dimBlock(N_BEAMS, 1, 1);
__global__ void detect_sum(cuComplex const *__restrict__ input, int n_avg, float *__restrict__ output)
{
    const int n_beams = blockDim.x, n_freq = gridDim.y, n_out = gridDim.x;
    const int o = blockIdx.x, f = blockIdx.y, b = threadIdx.x;
    const cuComplex *p = input + ((size_t)f * n_out + o) * n_avg * n_beams + b;     /* [f][o*n_avg + i][b] */
    float s = 0.f;
    for (int i = 0; i < n_avg; i++, p += n_beams) {
        const cuComplex c = *p;
        s = __fadd_rn(__fmaf_rn(c.x, c.x, __fmul_rn(c.y, c.y)), s);   /* the order nvcc gives x*x + y*y accumulated in turn */
    }
    output[((size_t)o * n_freq + f) * n_beams + b] = s;
}

/* ---- debugging prints ---- */
__global__ void print_data(float *data)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    printf("data[%d] = %f\n", idx, data[idx]);
}
__global__ void print_data_scalar(float *data) { printf("*data = %f\n", *data); }

/* ---- device selection by marketing name ---- */
inline void CUDA_select_GPU(char *prefered_device_name)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) != cudaSuccess) continue;
        if (strcmp(prefered_device_name, prop.name) == 0) {
            std::cout << "Selected: " << prop.name << std::endl;
            gpuErrchk(cudaSetDevice(d));
            return;
        }
    }
}

#endif
