// Test program for include/dsabf_compat.cuh: every name a reference-derived translation unit uses, with the
// reference's own launch configurations (beamformer.cu:464-466, :481).  usage: compat_check <in.u8> <n_bytes> <c.f32> <out_prefix>
#define N_BEAMS 256
#define N_FREQUENCIES 4
#define N_OUTPUTS_PER_GEMM 8
#include "dsabf_compat.cuh"
#include <vector>

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    char name[] = "NVIDIA B200";
    CUDA_select_GPU(name);
    const int n_bytes = atoi(argv[2]);
    const int n_avg = 32;
    const size_t n_c = (size_t)N_FREQUENCIES * N_OUTPUTS_PER_GEMM * n_avg * N_BEAMS;
    std::vector<char> h_in(n_bytes), h_out(2 * (size_t)n_bytes);
    std::vector<cuComplex> h_c(n_c);
    std::vector<float> h_p((size_t)N_OUTPUTS_PER_GEMM * N_FREQUENCIES * N_BEAMS);
    FILE *f = fopen(argv[1], "rb"); if (!f || fread(h_in.data(), 1, n_bytes, f) != (size_t)n_bytes) return 3; fclose(f);
    f = fopen(argv[3], "rb"); if (!f || fread(h_c.data(), sizeof(cuComplex), n_c, f) != n_c) return 3; fclose(f);
    char *d_in, *d_out; cuComplex *d_c; float *d_p;
    gpuErrchk(cudaMalloc(&d_in, n_bytes)); gpuErrchk(cudaMalloc(&d_out, 2 * (size_t)n_bytes));
    gpuErrchk(cudaMalloc(&d_c, n_c * sizeof(cuComplex))); gpuErrchk(cudaMalloc(&d_p, h_p.size() * sizeof(float)));
    gpuErrchk(cudaMemcpy(d_in, h_in.data(), n_bytes, cudaMemcpyHostToDevice));
    gpuErrchk(cudaMemcpy(d_c, h_c.data(), n_c * sizeof(cuComplex), cudaMemcpyHostToDevice));
    float ms = 0;
    START_TIMER();
    expand_input<<<10000, 32>>>(d_in, d_out, n_bytes);
    detect_sum<<<detect_dimGrid, detect_dimBlock>>>(d_c, n_avg, d_p);
    STOP_RECORD_TIMER(ms);
    gpuErrchk(cudaGetLastError());
    gpuBLASchk(0);
    gpuErrchk(cudaMemcpy(h_out.data(), d_out, h_out.size(), cudaMemcpyDeviceToHost));
    gpuErrchk(cudaMemcpy(h_p.data(), d_p, h_p.size() * sizeof(float), cudaMemcpyDeviceToHost));
    print_data_scalar<<<1, 1>>>(d_p);
    gpuErrchk(cudaDeviceSynchronize());
    std::string pre(argv[4]);
    f = fopen((pre + ".expanded.i8").c_str(), "wb"); fwrite(h_out.data(), 1, h_out.size(), f); fclose(f);
    f = fopen((pre + ".power.f32").c_str(), "wb"); fwrite(h_p.data(), sizeof(float), h_p.size(), f); fclose(f);
    printf("compat_check ok %.3f ms\n", ms);
    return 0;
}
