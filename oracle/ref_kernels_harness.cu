/*
 * ref_kernels_harness.cu -- replays the REFERENCE's GPU sequence for one chunk.
 *
 * TEST / BASELINE INFRASTRUCTURE.  Built only where /root/reference is mounted
 * (oracle/Makefile -> oracle/_ref/ref_kernels_{dbg,obs}); runs on the GPU box.
 * It #includes the reference's own kernels (expand_input, detect_sum:
 * src/beamformer.cuh:66-155) and issues the reference's own cuBLAS call with the
 * reference's arguments (src/beamformer.cu:464-481, 498-504).  The reference
 * binary never exposes d_C / d_out, so this is how the "reference rebuilt for
 * sm_100" is compared bit for bit and timed.
 *
 * usage: ref_kernels <weights.i8> <chunks.u8> <n_chunks> <out_prefix> [reps]
 *   weights.i8 : [256][64][256][2] int8  (the reference's fourier_coefficients)
 *   chunks.u8  : n_chunks * N_BYTES_PRE_EXPANSION_PER_GEMM bytes
 *   writes <out_prefix>.C.f32   (first chunk only: [f][col][beam][2] float)
 *          <out_prefix>.out.f32 ([n_chunks][8][256][256] float)
 *          <out_prefix>.ded.f32 ([n_chunks][256] float, gemv of output time 0)
 *   prints "ref_ms_per_chunk <ms>" (CUDA events over `reps` replays, 1 stream)
 *   and    "ref_ms_per_chunk_8streams <ms>" (the reference's 8-stream fan-out)
 */
#include "beamformer.hh"
#include "beamformer.cuh"
#include <cstdio>
#include <vector>
#include <string>

static std::vector<char> slurp(const char *name, size_t expect)
{
    FILE *f = fopen(name, "rb");
    if (!f) { perror(name); exit(2); }
    std::vector<char> v(expect);
    if (fread(v.data(), 1, expect, f) != expect) { fprintf(stderr, "%s: short read\n", name); exit(2); }
    fclose(f);
    return v;
}
static void dump(const std::string &name, const void *p, size_t n)
{
    FILE *f = fopen(name.c_str(), "wb");
    if (!f || fwrite(p, 1, n, f) != n) { perror(name.c_str()); exit(2); }
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s weights chunks n_chunks out_prefix [reps]\n", argv[0]); return 1; }
    int n_chunks = atoi(argv[3]);
    std::string prefix = argv[4];
    int reps = argc > 5 ? atoi(argv[5]) : 0;

    const size_t w_bytes = (size_t)N_BEAMS * N_ANTENNAS * N_FREQUENCIES * sizeof(CxInt8_t);
    std::vector<char> w = slurp(argv[1], w_bytes);
    std::vector<char> in = slurp(argv[2], (size_t)n_chunks * N_BYTES_PRE_EXPANSION_PER_GEMM);

    /* src/beamformer.cu:164-172 */
    int A_rows = N_BEAMS, A_cols = N_ANTENNAS, A_stride = A_rows * A_cols;
    int B_cols = N_TIMESTEPS_PER_GEMM, B_rows = A_cols, B_stride = B_rows * B_cols;
    int C_rows = A_rows, C_stride = C_rows * B_cols;

    CxInt8_t *d_A, *d_B; char *d_data; cuComplex *d_C; float *d_out, *d_ones, *d_one, *d_zero_f, *d_ded;
    cuComplex h_alpha, h_zero, *d_alpha, *d_zero;
    h_alpha.x = 1.0 / MAX_VAL; h_alpha.y = 0; h_zero.x = 0; h_zero.y = 0;      /* :191-194 */
    const int NS = N_STREAMS;
    gpuErrchk(cudaMalloc(&d_A, w_bytes));
    gpuErrchk(cudaMalloc(&d_B, (size_t)N_CX_IN_PER_GEMM * NS * sizeof(CxInt8_t)));
    gpuErrchk(cudaMalloc(&d_C, (size_t)N_CX_OUT_PER_GEMM * NS * sizeof(cuComplex)));
    gpuErrchk(cudaMalloc(&d_data, in.size()));
    gpuErrchk(cudaMalloc(&d_out, (size_t)N_F_PER_DETECT * NS * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_alpha, sizeof(cuComplex)));
    gpuErrchk(cudaMalloc(&d_zero, sizeof(cuComplex)));
    gpuErrchk(cudaMalloc(&d_ones, N_FREQUENCIES * sizeof(float)));
    gpuErrchk(cudaMalloc(&d_one, sizeof(float)));
    gpuErrchk(cudaMalloc(&d_zero_f, sizeof(float)));
    gpuErrchk(cudaMalloc(&d_ded, N_BEAMS * NS * sizeof(float)));
    std::vector<float> ones(N_FREQUENCIES, 1.0f);
    float f1 = 1.0f, f0 = 0.0f;
    gpuErrchk(cudaMemcpy(d_A, w.data(), w_bytes, cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_data, in.data(), in.size(), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_alpha, &h_alpha, sizeof(cuComplex), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_zero, &h_zero, sizeof(cuComplex), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_ones, ones.data(), N_FREQUENCIES * sizeof(float), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_one, &f1, sizeof(float), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_zero_f, &f0, sizeof(float), cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemset(d_B, 0, (size_t)N_CX_IN_PER_GEMM * NS * sizeof(CxInt8_t)));
    gpuErrchk(cudaMemset(d_C, 0, (size_t)N_CX_OUT_PER_GEMM * NS * sizeof(cuComplex)));

    cudaStream_t stream[NS]; cublasHandle_t handle[NS];
    for (int i = 0; i < NS; i++) {                                               /* :310-320 */
        gpuErrchk(cudaStreamCreate(&stream[i]));
        cublasStatus_t s = cublasCreate(&handle[i]);
        if (s != CUBLAS_STATUS_SUCCESS) { fprintf(stderr, "cublasCreate failed %d\n", (int)s); return 3; }
        cublasSetStream(handle[i], stream[i]);
        cublasSetPointerMode(handle[i], CUBLAS_POINTER_MODE_DEVICE);
        cublasSetMathMode(handle[i], CUBLAS_TENSOR_OP_MATH);
    }

    int gemm_status = 0;
    auto one_chunk = [&](int chunk, int st) {
        expand_input<<<10000, 32, 0, stream[st]>>>(&d_data[(size_t)N_BYTES_PRE_EXPANSION_PER_GEMM * chunk],
                                                   (char *)&d_B[(size_t)N_CX_IN_PER_GEMM * st], B_stride * N_FREQUENCIES);
        cublasStatus_t s = cublasGemmStridedBatchedEx(handle[st], CUBLAS_OP_N, CUBLAS_OP_N,
                              A_rows, B_cols, A_cols, d_alpha,
                              d_A, CUDA_C_8I, A_rows, A_stride,
                              &d_B[(size_t)N_CX_IN_PER_GEMM * st], CUDA_C_8I, B_rows, B_stride,
                              d_zero,
                              &d_C[(size_t)N_CX_OUT_PER_GEMM * st], CUDA_C_32F, C_rows, C_stride,
                              N_FREQUENCIES, CUDA_C_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP);
        if (s != CUBLAS_STATUS_SUCCESS) gemm_status = (int)s;
        detect_sum<<<detect_dimGrid, detect_dimBlock, 0, stream[st]>>>(&d_C[(size_t)N_CX_OUT_PER_GEMM * st], N_INPUTS_PER_OUTPUT,
                                                                      &d_out[(size_t)st * N_F_PER_DETECT]);
    };

    std::vector<float> out((size_t)n_chunks * N_F_PER_DETECT), ded((size_t)n_chunks * N_BEAMS);
    for (int c = 0; c < n_chunks; c++) {
        one_chunk(c, 0);
        cublasSgemv(handle[0], CUBLAS_OP_N, N_BEAMS, N_FREQUENCIES, d_one, d_out, N_BEAMS, d_ones, 1, d_zero_f, d_ded, 1);
        gpuErrchk(cudaMemcpyAsync(&out[(size_t)c * N_F_PER_DETECT], d_out, N_F_PER_DETECT * sizeof(float), cudaMemcpyDeviceToHost, stream[0]));
        gpuErrchk(cudaMemcpyAsync(&ded[(size_t)c * N_BEAMS], d_ded, N_BEAMS * sizeof(float), cudaMemcpyDeviceToHost, stream[0]));
        gpuErrchk(cudaStreamSynchronize(stream[0]));
        if (c == 0) {
            std::vector<cuComplex> C((size_t)N_CX_OUT_PER_GEMM);
            gpuErrchk(cudaMemcpy(C.data(), d_C, C.size() * sizeof(cuComplex), cudaMemcpyDeviceToHost));
            dump(prefix + ".C.f32", C.data(), C.size() * sizeof(cuComplex));
        }
    }
    printf("gemm_status %d\n", gemm_status);
    dump(prefix + ".out.f32", out.data(), out.size() * sizeof(float));
    dump(prefix + ".ded.f32", ded.data(), ded.size() * sizeof(float));

    if (reps > 0) {
        cudaEvent_t e0, e1; float ms;
        gpuErrchk(cudaEventCreate(&e0)); gpuErrchk(cudaEventCreate(&e1));
        for (int r = 0; r < 3; r++) one_chunk(r % n_chunks, 0);
        gpuErrchk(cudaDeviceSynchronize());
        gpuErrchk(cudaEventRecord(e0, stream[0]));
        for (int r = 0; r < reps; r++) one_chunk(r % n_chunks, 0);
        gpuErrchk(cudaEventRecord(e1, stream[0]));
        gpuErrchk(cudaEventSynchronize(e1));
        gpuErrchk(cudaEventElapsedTime(&ms, e0, e1));
        printf("ref_ms_per_chunk %.6f\n", ms / reps);
        /* the reference's own fan-out over N_STREAMS streams (src/beamformer.cu:454-521) */
        for (int r = 0; r < NS; r++) one_chunk(r % n_chunks, r);
        gpuErrchk(cudaDeviceSynchronize());
        gpuErrchk(cudaEventRecord(e0, 0));
        for (int r = 0; r < reps; r++) one_chunk(r % n_chunks, r % NS);
        gpuErrchk(cudaDeviceSynchronize());
        gpuErrchk(cudaEventRecord(e1, 0));
        gpuErrchk(cudaEventSynchronize(e1));
        gpuErrchk(cudaEventElapsedTime(&ms, e0, e1));
        printf("ref_ms_per_chunk_8streams %.6f\n", ms / reps);
    }
    printf("ref_kernels done: n_avg %d cols %d\n", N_AVERAGING, N_TIMESTEPS_PER_GEMM);
    return 0;
}
