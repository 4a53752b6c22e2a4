/*
 * ref_host_harness.cu -- runs the REFERENCE's own host expressions on the CPU.
 *
 * TEST INFRASTRUCTURE.  Built only where /root/reference is mounted, by
 * oracle/Makefile, into oracle/_ref/ref_host.  It #includes the reference
 * headers where they lie (-I/root/reference/src) and a sed-extracted copy of
 * the steering-weight loop (src/beamformer.cu:232-241, written to the
 * git-ignored oracle/_ref/weights_snippet.inc at build time), so the numbers it
 * prints are the reference's, compiled by the same nvcc host pass the
 * reference's makefile uses.  No reference source is committed.
 *
 * usage: ref_host <positions> <directions> <sources|-> <gpu> <out_prefix> [n_batches]
 *   writes <out_prefix>.weights.i8   [256][64][256][2]
 *          <out_prefix>.data.b<k>.u8 [1024][256][16][64]  for each source batch k
 */
#include "beamformer.hh"   /* reference: constants, antenna, beam_direction, readers */
#include "beamformer.cuh"  /* reference: gpuErrchk (used by the generator)          */

static cudaError_t shim_host_alloc(void **p, size_t sz) { *p = malloc(sz); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static cudaError_t shim_host_free(void *p) { free(p); return cudaSuccess; }
#define cudaHostAlloc(p, sz, fl) shim_host_alloc((void **)(p), (sz))
#define cudaFreeHost(p) shim_host_free((void *)(p))

#include "test_data_generator.hh" /* reference: test_data_generator */

#include <cstdio>
#include <cstring>
#include <string>

static void dump(const std::string &name, const void *p, size_t n)
{
    FILE *f = fopen(name.c_str(), "wb");
    if (!f || fwrite(p, 1, n, f) != n) { perror(name.c_str()); exit(2); }
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s pos dir src|- gpu out_prefix [n_batches]\n", argv[0]); return 1; }
    int gpu = atoi(argv[4]);
    std::string prefix = argv[5];
    int want_batches = argc > 6 ? atoi(argv[6]) : -1;

    antenna *pos = new antenna[N_ANTENNAS]();
    beam_direction *dir = new beam_direction[N_BEAMS]();
    if (strcmp(argv[1], "-")) read_in_position_locations(argv[1], pos);
    else for (int i = 0; i < N_ANTENNAS; i++) pos[i].x = i * 500.0 / (N_ANTENNAS - 1) - 250.0;            /* beamformer.cu:135-140 */
    if (strcmp(argv[2], "-")) read_in_beam_directions(argv[2], N_BEAMS, dir);
    else for (int i = 0; i < N_BEAMS; i++) dir[i].theta = i * DEG2RAD(2 * HALF_FOV) / (N_BEAMS - 1) - DEG2RAD(HALF_FOV); /* :142-147 */

    /* names the extracted snippet expects (src/beamformer.cu:164-173, 230) */
    int fourier_coefficients_rows = N_BEAMS;
    int fourier_coefficients_cols = N_ANTENNAS;
    int fourier_coefficients_stride = fourier_coefficients_rows * fourier_coefficients_cols;
    float bw_per_channel = BW_PER_CHANNEL;
    CxInt8_t *fourier_coefficients = new CxInt8_t[fourier_coefficients_cols * fourier_coefficients_rows * N_FREQUENCIES];
#include "weights_snippet.inc"
    dump(prefix + ".weights.i8", fourier_coefficients, sizeof(CxInt8_t) * fourier_coefficients_stride * N_FREQUENCIES);

    test_data_generator gen;
    if (strcmp(argv[3], "-")) gen.read_in_source_directions(argv[3]);
    int n_src = gen.get_n_pt_sources();
    int n_batches = CEILING(n_src, N_SOURCES_PER_BATCH);
    if (want_batches >= 0 && want_batches < n_batches) n_batches = want_batches;
    for (int b = 0; b < n_batches; b++) {
        if (strcmp(argv[3], "-")) gen.generate_test_data(pos, gpu);
        dump(prefix + ".data.b" + std::to_string(b) + ".u8", gen.get_data(), (size_t)INPUT_DATA_SIZE);
    }
    printf("n_src %d n_batches %d\n", n_src, n_batches);
    return 0;
}
