/*
 * dsabf_oracle.c -- CPU restatement of the DSAbeamformer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The product path (dsabeamformer_b200/csrc) never links or calls it.
 *
 * Every function cites the reference file:line (under /root/reference) it
 * restates.  Parity status: PINNED -- against (i) the reference's nibble
 * known-answer vectors (sandbox/bitshift.cpp:5-28, sandbox/kernelTest.cu:127-148,
 * src/test_data_generator.hh:8), (ii) the reference's own host expressions for
 * weights and test data compiled from /root/reference by oracle/Makefile into
 * oracle/_ref/ref_host (hashes committed in tests/golden/), (iii) the
 * notebook scalar np.sum(A) (sandbox/2D Beamformer.ipynb cell 6), and (iv) on
 * the GPU box the reference's own kernels + cuBLAS replayed by
 * oracle/_ref/ref_kernels.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference's host code is built by nvcc's host
 * pass without -march, i.e. without FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

/* Constants: src/beamformer.hh:63-77 */
#define ORC_N_GPUS 8
#define ORC_TOT_CHANNELS 2048
#define ORC_START_F 1.28
#define ORC_END_F 1.53
#define ORC_ZERO_PT 0
#define ORC_BW_PER_CHANNEL ((ORC_END_F - ORC_START_F) / ORC_TOT_CHANNELS)
#define ORC_C_SPEED 299792458.0
#define ORC_PI 3.14159265358979
#define ORC_MAX_VAL 127
#define ORC_SIG_MAX_VAL 7

/* ------------------------------------------------------------------------- */
/* Channel frequency in GHz, the steering-weight flavour: bw is a *float*
 * (src/beamformer.cu:173 `float bw_per_channel = BW_PER_CHANNEL;`) and the
 * sub-band offset uses integer division by N_GPUS-1 (src/beamformer.cu:233). */
float orc_freq_weights(int gpu, int chan)
{
    float bw = (float)ORC_BW_PER_CHANNEL;
    int idx = ORC_ZERO_PT + gpu * ORC_TOT_CHANNELS / (ORC_N_GPUS - 1) + chan;
    float prod = (float)idx * bw;               /* int*float -> float */
    float freq = (float)(ORC_END_F - (double)prod);
    return freq;
}

/* The test-data flavour: bw stays the double macro
 * (src/test_data_generator.hh:72). */
float orc_freq_generator(int gpu, int chan)
{
    int idx = ORC_ZERO_PT + gpu * ORC_TOT_CHANNELS / (ORC_N_GPUS - 1) + chan;
    float freq = (float)(ORC_END_F - (double)idx * ORC_BW_PER_CHANNEL);
    return freq;
}

/* wavelength: src/beamformer.cu:234, src/test_data_generator.hh:74 */
static float orc_wavelength(float freq)
{
    return (float)(ORC_C_SPEED / (1E9 * (double)freq));
}

/* geometry term in *float* (sin(float) resolves to the float overload under
 * nvcc): pos.x*sin(theta) + pos.y*sin(phi), src/beamformer.cu:237-238 */
static float orc_geom(float x, float y, float theta, float phi)
{
    float a = x * sinf(theta);
    float b = y * sinf(phi);
    return a + b;
}

/* ------------------------------------------------------------------------- */
/* Steering weights, src/beamformer.cu:230-241.
 * pos: [n_ant][3] floats (x,y,z; z unused), dir: [n_beams][2] floats (theta,phi)
 * w:   [n_chan][n_ant][n_beams][2] int8 (re,im), beam fastest == the reference's
 *      column-major N_BEAMS x N_ANTENNAS CxInt8 with lda = N_BEAMS. */
void orc_weights(const float *pos, const float *dir, int n_ant, int n_beams,
                 int n_chan, int gpu, int8_t *w)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_chan; i++) {
        float wavelength = orc_wavelength(orc_freq_weights(gpu, i));
        for (int j = 0; j < n_ant; j++) {
            for (int k = 0; k < n_beams; k++) {
                float g = orc_geom(pos[3 * j], pos[3 * j + 1], dir[2 * k], dir[2 * k + 1]);
                double ph = -2 * ORC_PI * (double)g / (double)wavelength;
                size_t o = (((size_t)i * n_ant + j) * n_beams + k) * 2;
                w[o]     = (int8_t)round(ORC_MAX_VAL * cos(ph));
                w[o + 1] = (int8_t)round(ORC_MAX_VAL * sin(ph));
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* DEBUG test data, src/test_data_generator.hh:63-95.  One "source" per chunk,
 * all columns of a chunk identical.  data: [n_chunks][n_chan][cols][n_ant]
 * bytes; chunk c of this batch carries source (c + batch*n_chunks), zero bytes
 * past n_src (hh:78-86).  bogus fill 0x70 when src == NULL (hh:8,36). */
void orc_generate(const float *pos, const float *src, int n_src, int batch,
                  int n_chunks, int n_ant, int n_chan, int cols, int gpu,
                  uint8_t *data)
{
    size_t chunk_bytes = (size_t)n_chan * cols * n_ant;
    if (!src) {
        memset(data, 0x70, chunk_bytes * n_chunks);
        return;
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (int c = 0; c < n_chunks; c++) {
        int s = c + batch * n_chunks;
        uint8_t *d = data + chunk_bytes * c;
        if (s >= n_src) {
            memset(d, 0, chunk_bytes);
            continue;
        }
        for (int i = 0; i < n_chan; i++) {
            float wavelength = orc_wavelength(orc_freq_generator(gpu, i));
            uint8_t *row0 = d + (size_t)i * cols * n_ant;
            for (int k = 0; k < n_ant; k++) {
                float g = orc_geom(pos[3 * k], pos[3 * k + 1], src[2 * s], src[2 * s + 1]);
                double ph = 2 * ORC_PI * (double)g / (double)wavelength;
                signed char high = (signed char)round(ORC_SIG_MAX_VAL * cos(ph));
                signed char low  = (signed char)round(ORC_SIG_MAX_VAL * sin(ph));
                row0[k] = (uint8_t)((high << 4) | (0x0F & low));
            }
            for (int j = 1; j < cols; j++)
                memcpy(row0 + (size_t)j * n_ant, row0, n_ant);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* 4-bit -> 8-bit expansion, src/beamformer.cuh:92-103: high nibble = real,
 * low nibble = imaginary, two's complement. out[2*i]=re, out[2*i+1]=im. */
void orc_unpack(const uint8_t *in, int8_t *out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        int8_t t = (int8_t)in[i];
        int8_t high = (int8_t)(t >> 4);
        int8_t low = (int8_t)((int8_t)(t << 4) >> 4);
        out[2 * i] = high;
        out[2 * i + 1] = low;
    }
}

/* ------------------------------------------------------------------------- */
/* Exact integer beam voltages: the integer sum inside
 * cublasGemmStridedBatchedEx (src/beamformer.cu:470-477) before alpha.
 * packed: [n_chan][n_cols][n_ant]; w: [n_chan][n_ant][n_beams][2];
 * V: [n_chan][n_cols][n_beams][2] int32 (re,im) == d_C layout [f][col][beam]. */
void orc_beamform(const uint8_t *packed, const int8_t *w, int n_chan, int n_cols,
                  int n_ant, int n_beams, int32_t *V)
{
#pragma omp parallel
    {
        int16_t *wre = (int16_t *)malloc(sizeof(int16_t) * n_ant * n_beams);
        int16_t *wim = (int16_t *)malloc(sizeof(int16_t) * n_ant * n_beams);
        int32_t *are = (int32_t *)malloc(sizeof(int32_t) * n_beams);
        int32_t *aim = (int32_t *)malloc(sizeof(int32_t) * n_beams);
#pragma omp for schedule(dynamic, 1)
        for (int f = 0; f < n_chan; f++) {
            const int8_t *wf = w + (size_t)f * n_ant * n_beams * 2;
            for (int i = 0; i < n_ant * n_beams; i++) {
                wre[i] = wf[2 * i];
                wim[i] = wf[2 * i + 1];
            }
            for (int c = 0; c < n_cols; c++) {
                const uint8_t *x = packed + ((size_t)f * n_cols + c) * n_ant;
                for (int b = 0; b < n_beams; b++) { are[b] = 0; aim[b] = 0; }
                for (int a = 0; a < n_ant; a++) {
                    int8_t t = (int8_t)x[a];
                    int xr = (int8_t)(t >> 4);
                    int xi = (int8_t)((int8_t)(t << 4) >> 4);
                    if (xr == 0 && xi == 0) continue;
                    const int16_t *r = wre + (size_t)a * n_beams;
                    const int16_t *m = wim + (size_t)a * n_beams;
                    for (int b = 0; b < n_beams; b++) {
                        are[b] += r[b] * xr - m[b] * xi;
                        aim[b] += r[b] * xi + m[b] * xr;
                    }
                }
                int32_t *v = V + ((size_t)f * n_cols + c) * n_beams * 2;
                for (int b = 0; b < n_beams; b++) {
                    v[2 * b] = are[b];
                    v[2 * b + 1] = aim[b];
                }
            }
        }
        free(wre); free(wim); free(are); free(aim);
    }
}

/* Exact integer power: P[o][f][b] = sum_{i<n_avg} |V[f][o*n_avg+i][b]|^2
 * (the integer content of detect_sum, src/beamformer.cuh:141-154, without
 * alpha^2).  n_out = n_cols / n_avg. */
void orc_power_int(const int32_t *V, int n_chan, int n_cols, int n_beams,
                   int n_avg, int64_t *P)
{
    int n_out = n_cols / n_avg;
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < n_out; o++)
        for (int f = 0; f < n_chan; f++)
            for (int b = 0; b < n_beams; b++) {
                int64_t s = 0;
                for (int i = 0; i < n_avg; i++) {
                    const int32_t *v = V + (((size_t)f * n_cols + o * n_avg + i) * n_beams + b) * 2;
                    s += (int64_t)v[0] * v[0] + (int64_t)v[1] * v[1];
                }
                P[((size_t)o * n_chan + f) * n_beams + b] = s;
            }
}

/* The reference's float path, emulated operation by operation:
 *   C = fl32(alpha * V) with alpha = (float)(1.0/127)  (src/beamformer.cu:191,
 *       470-477; the integer sum is exact in fp32 because |V| < 2^17),
 *   detect_sum (src/beamformer.cuh:141-154) as nvcc 12.9 -O3 compiles it for
 *   sm_100 (SASS read back with cuobjdump): t = y*y; u = fma(x,x,t); s = u + s,
 *   sequentially over the n_avg columns.
 * out: [n_out][n_chan][n_beams] float == d_out. */
void orc_detect_ref_float(const int32_t *V, int n_chan, int n_cols, int n_beams,
                          int n_avg, float *out)
{
    const float alpha = (float)(1.0 / ORC_MAX_VAL);
    int n_out = n_cols / n_avg;
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < n_out; o++)
        for (int f = 0; f < n_chan; f++)
            for (int b = 0; b < n_beams; b++) {
                float s = 0.0f;
                for (int i = 0; i < n_avg; i++) {
                    const int32_t *v = V + (((size_t)f * n_cols + o * n_avg + i) * n_beams + b) * 2;
                    float x = alpha * (float)v[0];
                    float y = alpha * (float)v[1];
                    float t = y * y;
                    float u = fmaf(x, x, t);
                    s = u + s;
                }
                out[((size_t)o * n_chan + f) * n_beams + b] = s;
            }
}

/* DM-0 "dedispersion": cublasSgemv with a ones vector over the channel axis of
 * output time `o` (the reference uses o = 0 only, src/beamformer.cu:498-504).
 * Summation order inside cuBLAS is unspecified; we sum in double and round once
 * so that callers can compare with a 256*2^-24 relative tolerance. */
void orc_dm0(const float *out, int n_chan, int n_beams, int o, float *ded)
{
    for (int b = 0; b < n_beams; b++) {
        double s = 0;
        for (int f = 0; f < n_chan; f++)
            s += out[((size_t)o * n_chan + f) * n_beams + b];
        ded[b] = (float)s;
    }
}

/* ------------------------------------------------------------------------- */
/* Incoherent dedispersion for DM trials (generalises the DM-0 gemv).  The only
 * spec in the reference is sandbox/Dispersion Theory.ipynb: trial list by the
 * recurrence in cell 2, delay in output samples by the formula of cell 5:
 *   int(4.15 * DM * (f^-2 - f1^-2) / (0.131*16)), f in GHz, f1 = 1.53.
 * power: [n_t][n_chan][n_beams] float; delay: [n_dm][n_chan] int32 (>= 0);
 * ded: [n_dm][n_t_out][n_beams] with
 *   ded[d][t][b] = sum_f power[t + delay[d][f]][f][b]   (t + delay < n_t)
 * accumulated in double, rounded once. */
void orc_dedisperse(const float *power, int n_t, int n_chan, int n_beams,
                    const int32_t *delay, int n_dm, int n_t_out, float *ded)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < n_dm; d++)
        for (int t = 0; t < n_t_out; t++)
            for (int b = 0; b < n_beams; b++) {
                double s = 0;
                for (int f = 0; f < n_chan; f++) {
                    int tt = t + delay[(size_t)d * n_chan + f];
                    if (tt < n_t)
                        s += power[((size_t)tt * n_chan + f) * n_beams + b];
                }
                ded[((size_t)d * n_t_out + t) * n_beams + b] = (float)s;
            }
}

/* DM trial list, sandbox/Dispersion Theory.ipynb cells 1-2 (all double). */
int orc_dm_trials(double dm_max, int nchan, double *dms, int cap)
{
    double epsilon = 1.25;
    double alpha = 1.0 / (16 + (double)nchan * nchan);
    double nu = (1.28 + 1.53) / 2;
    double B = (1.53 - 1.28) / nchan * 1000;
    double ti = 40.0, tscat = 0, tsamp = 131;
    double beta = ti * ti + tscat * tscat + tsamp * tsamp;
    double dm_prev = 0;
    int n = 0;
    if (n < cap) dms[n] = dm_prev;
    n++;
    while (dm_prev < dm_max) {
        double n2a = (double)nchan * nchan * alpha;
        double q = (nu * nu * nu) / (8.3 * B);
        dm_prev = n2a * dm_prev +
                  4 * sqrt(alpha * (epsilon * epsilon - n2a) * dm_prev * dm_prev +
                           alpha * beta * (epsilon * epsilon - 1) * q * q);
        if (n < cap) dms[n] = dm_prev;
        n++;
    }
    return n;
}

/* Delay table, sandbox/Dispersion Theory.ipynb cell 5.  freq_ghz[f] is the
 * centre frequency of channel f; int() truncates toward zero. */
void orc_dm_delays(const double *dms, int n_dm, const double *freq_ghz, int n_chan,
                   double f1_ghz, double tsamp_ms_times_navg, int32_t *delay)
{
    for (int d = 0; d < n_dm; d++)
        for (int f = 0; f < n_chan; f++) {
            double fr = freq_ghz[f];
            double v = 4.15 * dms[d] * (-1.0 / (f1_ghz * f1_ghz) + 1.0 / (fr * fr)) / tsamp_ms_times_navg;
            delay[(size_t)d * n_chan + f] = (int32_t)v;
        }
}
