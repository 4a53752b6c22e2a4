// dsabf_api.cu -- the C ABI of include/dsabf.h.
//
// Absorbs the set-up, stream/event plumbing and teardown that the reference does
// inline in main() around its hot loop (/root/reference/src/beamformer.cu:249-320,
// 389-396, 461-526, 560-618; src/observation_loop.hh:71-119).  Differences by
// design: one fused kernel per block instead of 32 x (expand, GEMM, detect); three
// streams (H2D, compute, D2H) with device-side event dependencies instead of 8
// compute streams fenced only on the last one; a ring slot is never overwritten
// before the kernel that reads it has finished.
#include "../../include/dsabf.h"
#include "fused_beamform.cuh"

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *what)
{
    char buf[512];
    if (code > 0 && code < 10000)
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString((cudaError_t)code));
    else
        snprintf(buf, sizeof buf, "%s (code %d)", what, code);
    g_last_error = buf;
    return code;
}

#define CK(expr)                                              \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return fail((int)e__, #expr); \
    } while (0)

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

constexpr int kTicketRing = 256;

}  // namespace

struct dsabf_handle {
    dsabf_config cfg{};
    int device = 0;
    int num_sms = 0;
    int group = 0;            // columns per output = n_pol * n_averaging
    int cols_per_chunk = 0;
    size_t chunk_bytes = 0, block_bytes = 0, chunk_out_floats = 0;
    float scale = 0.f;        // power_scale
    bool weights_set = false;

    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    uint8_t *d_ring = nullptr;     // n_ring_slots * block_bytes
    float *d_out = nullptr;        // n_ring_slots * n_chunks_per_block * chunk_out_floats
    float *d_ded = nullptr;        // n_ring_slots * n_chunks_per_block * n_beams
    int8_t *d_w = nullptr;         // reference-layout weights
    uint8_t *d_wprep = nullptr;    // tensor-core layout
    int *d_bad = nullptr;
    long long *d_trace = nullptr;  // debug: clock64 trace buffer (dsabf_debug_trace)
    float *d_gen_par = nullptr;    // parameters of the GPU test-data generator
    int32_t *d_delay = nullptr; size_t delay_cap = 0;            // dedispersion delay table (grown on demand)
    cudaEvent_t ev_delay = nullptr; bool delay_in_use = false;   // last kernel that read it

    std::vector<cudaEvent_t> ev_slot_h2d, ev_slot_comp, ev_slot_d2h;   // per ring slot
    std::vector<char> slot_has_h2d, slot_has_comp, slot_has_d2h;
    cudaEvent_t ev_xfer[kTicketRing]{}, ev_ana[kTicketRing]{};
    uint64_t n_xfer = 0, n_ana = 0;

    // scratch for the *_host conveniences
    uint8_t *d_tmp_in = nullptr; size_t tmp_in_bytes = 0;
    float *d_tmp_out = nullptr;  size_t tmp_out_bytes = 0;

    uint64_t launches = 0;
    bool timing = false;
    std::vector<cudaEvent_t> t_start, t_stop;   // pool of timing events
    size_t t_used = 0;
};

namespace {

int ensure_tmp(dsabf_handle *h, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > h->tmp_in_bytes) {
        if (h->d_tmp_in) cudaFree(h->d_tmp_in);
        h->d_tmp_in = nullptr; h->tmp_in_bytes = 0;
        CK(cudaMalloc(&h->d_tmp_in, in_bytes));
        h->tmp_in_bytes = in_bytes;
    }
    if (out_bytes > h->tmp_out_bytes) {
        if (h->d_tmp_out) cudaFree(h->d_tmp_out);
        h->d_tmp_out = nullptr; h->tmp_out_bytes = 0;
        CK(cudaMalloc(&h->d_tmp_out, out_bytes));
        h->tmp_out_bytes = out_bytes;
    }
    return 0;
}

// One fused launch over n_chunks chunks.
int run_fused(dsabf_handle *h, const uint8_t *d_packed, int n_chunks, float *d_out, int32_t *vdump, long long *pdump, bool exact,
              cudaStream_t st)
{
    if (!h->weights_set) return fail(DSABF_E_STATE, "weights not set (call dsabf_set_weights first)");
    if (n_chunks <= 0) return 0;
    dsabf::FusedParams p{};
    p.packed = d_packed;
    p.wprep = h->d_wprep;
    p.out = d_out;
    p.vdump = vdump;
    p.pdump = pdump;
    p.trace = h->d_trace;
    p.n_chan = h->cfg.n_channels;
    p.n_chunks = n_chunks;
    p.cols_per_chunk = h->cols_per_chunk;
    p.group = h->group;
    p.n_ant = h->cfg.n_antennas;
    p.n_beams = h->cfg.n_beams;
    const long long cols = (long long)n_chunks * h->cols_per_chunk;
    p.n_tiles = (int)((cols + 255) / 256);
    p.scale = h->scale;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        if (h->t_used == h->t_start.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            h->t_start.push_back(a);
            h->t_stop.push_back(b);
        }
        e0 = h->t_start[h->t_used];
        e1 = h->t_stop[h->t_used];
        h->t_used++;
        CK(cudaEventRecord(e0, st));
    }
    // 64 antennas x <= 256 beams: weights resident per channel; anything larger: streamed-weights kernel
    static const bool force_big = getenv("DSABF_FORCE_STREAMED") != nullptr;     // testing hook: route every shape through the streamed-weights kernel
    if (h->cfg.n_antennas == 64 && h->cfg.n_beams <= 256 && !force_big) CK(dsabf::launch_fused_beamform(p, exact, h->num_sms, st));
    else CK(dsabf::launch_fused_beamform_big(p, exact, h->num_sms, st));
    if (e1) CK(cudaEventRecord(e1, st));
    h->launches++;
    return 0;
}

// Stream-ordered upload of a delay table into the handle's buffer: no host synchronisation, no allocation in the
// steady state.  The buffer is reused, so the copy first waits (on the device) for the last kernel that read it.
int upload_delays(dsabf_handle *h, const int32_t *h_delay, size_t n, cudaStream_t st)
{
    if (!h->ev_delay) CK(cudaEventCreateWithFlags(&h->ev_delay, cudaEventDisableTiming));
    if (n > h->delay_cap) {
        if (h->delay_in_use) CK(cudaEventSynchronize(h->ev_delay));
        if (h->d_delay) CK(cudaFree(h->d_delay));
        h->d_delay = nullptr; h->delay_cap = 0;
        CK(cudaMalloc(&h->d_delay, sizeof(int32_t) * n));
        h->delay_cap = n;
        h->delay_in_use = false;
    }
    if (h->delay_in_use) CK(cudaStreamWaitEvent(st, h->ev_delay, 0));
    CK(cudaMemcpyAsync(h->d_delay, h_delay, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace

extern "C" {

int dsabf_default_config(dsabf_config *cfg, int debug)
{
    if (!cfg) return fail(DSABF_E_INVALID, "cfg is NULL");
    cfg->n_antennas = 64;
    cfg->n_beams = 256;
    cfg->n_channels = 256;
    cfg->n_pol = 2;
    cfg->n_averaging = debug ? 1 : 16;
    cfg->n_outputs_per_chunk = 8;
    cfg->n_chunks_per_block = 32;
    cfg->n_ring_slots = 8;
    cfg->device = -1;
    cfg->exact_power = 0;
    cfg->power_scale = 0.f;
    return 0;
}

int dsabf_columns_per_chunk(const dsabf_config *c) { return c->n_outputs_per_chunk * c->n_pol * c->n_averaging; }
size_t dsabf_chunk_bytes(const dsabf_config *c) { return (size_t)c->n_antennas * c->n_channels * dsabf_columns_per_chunk(c); }
size_t dsabf_block_bytes(const dsabf_config *c) { return dsabf_chunk_bytes(c) * c->n_chunks_per_block; }
size_t dsabf_chunk_out_floats(const dsabf_config *c) { return (size_t)c->n_outputs_per_chunk * c->n_channels * c->n_beams; }

const char *dsabf_last_error_string(void) { return g_last_error.c_str(); }

int dsabf_create(const dsabf_config *cfg, dsabf_handle **out)
{
    if (!cfg || !out) return fail(DSABF_E_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->n_antennas < 64 || cfg->n_antennas > 256 || cfg->n_antennas % 64) return fail(DSABF_E_INVALID, "n_antennas must be 64, 128, 192 or 256");
    if (cfg->n_beams < 64 || cfg->n_beams > 4096 || cfg->n_beams % 64) return fail(DSABF_E_INVALID, "n_beams must be a multiple of 64 in [64, 4096]");
    if (cfg->n_channels < 1) return fail(DSABF_E_INVALID, "n_channels must be >= 1");
    if (cfg->n_pol < 1 || cfg->n_averaging < 1) return fail(DSABF_E_INVALID, "n_pol, n_averaging must be >= 1");
    const int group = cfg->n_pol * cfg->n_averaging;
    if (!is_pow2(group) || group < 2 || group > 64) return fail(DSABF_E_INVALID, "n_pol*n_averaging must be a power of two in [2,64]");
    const int cpc = cfg->n_outputs_per_chunk * group;
    if (cfg->n_outputs_per_chunk < 1 || !((cpc >= 256 && cpc % 256 == 0) || 256 % cpc == 0))
        return fail(DSABF_E_INVALID, "columns per chunk must divide 256 or be a multiple of 256");
    if (cfg->n_chunks_per_block < 1 || cfg->n_ring_slots < 1) return fail(DSABF_E_INVALID, "n_chunks_per_block, n_ring_slots must be >= 1");

    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail((int)cudaErrorNoDevice, "no CUDA device: the beamformer has no CPU fallback");
    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(DSABF_E_INVALID, "device ordinal out of range");
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(DSABF_E_INVALID, "this library contains sm_100a code only (needs a B200-class GPU)");

    dsabf_handle *h = new (std::nothrow) dsabf_handle;
    if (!h) return fail(DSABF_E_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->device = dev;
    h->num_sms = prop.multiProcessorCount;
    h->group = group;
    h->cols_per_chunk = cpc;
    h->chunk_bytes = dsabf_chunk_bytes(cfg);
    h->block_bytes = dsabf_block_bytes(cfg);
    h->chunk_out_floats = dsabf_chunk_out_floats(cfg);
    // reference: alpha = (float)(1.0/127) applied to both components (beamformer.cu:191), squared by detect_sum
    const float alpha = (float)(1.0 / 127);
    const float ps = cfg->power_scale != 0.f ? cfg->power_scale : (float)((double)alpha * (double)alpha);
    h->scale = ps / 256.0f;      // voltages are carried x16 through the tensor core

    int rc = 0;
    auto ck = [&](cudaError_t e, const char *w) { if (e != cudaSuccess && !rc) rc = fail((int)e, w); };
    ck(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking), "create h2d stream");
    ck(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking), "create compute stream");
    ck(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking), "create d2h stream");
    const size_t slots = cfg->n_ring_slots, cpb = cfg->n_chunks_per_block;
    ck(cudaMalloc(&h->d_ring, slots * h->block_bytes), "alloc input ring");
    ck(cudaMalloc(&h->d_out, slots * cpb * h->chunk_out_floats * sizeof(float)), "alloc power ring");
    ck(cudaMalloc(&h->d_ded, slots * cpb * cfg->n_beams * sizeof(float)), "alloc dm0 ring");
    const size_t wbytes = (size_t)cfg->n_channels * cfg->n_antennas * cfg->n_beams * 2;
    ck(cudaMalloc(&h->d_w, wbytes), "alloc weights");
    ck(cudaMalloc(&h->d_wprep, 2 * wbytes), "alloc tensor-core weights");   // realified: 2 rows of 2*n_ant bytes per beam
    ck(cudaMalloc(&h->d_bad, sizeof(int)), "alloc flag");
    if (!rc) {
        ck(cudaMemset(h->d_ring, 0, slots * h->block_bytes), "memset ring");           // beamformer.cu:291-298
        ck(cudaMemset(h->d_out, 0, slots * cpb * h->chunk_out_floats * sizeof(float)), "memset out");
        ck(cudaMemset(h->d_ded, 0, slots * cpb * cfg->n_beams * sizeof(float)), "memset ded");
    }
    h->ev_slot_h2d.resize(slots); h->ev_slot_comp.resize(slots); h->ev_slot_d2h.resize(slots);
    h->slot_has_h2d.assign(slots, 0); h->slot_has_comp.assign(slots, 0); h->slot_has_d2h.assign(slots, 0);
    for (size_t i = 0; i < slots && !rc; i++) {
        ck(cudaEventCreateWithFlags(&h->ev_slot_h2d[i], cudaEventDisableTiming), "event");
        ck(cudaEventCreateWithFlags(&h->ev_slot_comp[i], cudaEventDisableTiming), "event");
        ck(cudaEventCreateWithFlags(&h->ev_slot_d2h[i], cudaEventDisableTiming), "event");
    }
    for (int i = 0; i < kTicketRing && !rc; i++) {
        ck(cudaEventCreateWithFlags(&h->ev_xfer[i], cudaEventDisableTiming), "event");
        ck(cudaEventCreateWithFlags(&h->ev_ana[i], cudaEventDisableTiming), "event");
    }
    if (rc) { dsabf_destroy(h); return rc; }
    *out = h;
    return 0;
}

int dsabf_destroy(dsabf_handle *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto e : h->ev_slot_h2d) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_slot_comp) if (e) cudaEventDestroy(e);
    for (auto e : h->ev_slot_d2h) if (e) cudaEventDestroy(e);
    for (int i = 0; i < kTicketRing; i++) { if (h->ev_xfer[i]) cudaEventDestroy(h->ev_xfer[i]); if (h->ev_ana[i]) cudaEventDestroy(h->ev_ana[i]); }
    for (auto e : h->t_start) cudaEventDestroy(e);
    for (auto e : h->t_stop) cudaEventDestroy(e);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_comp) cudaStreamDestroy(h->s_comp);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    cudaFree(h->d_ring); cudaFree(h->d_out); cudaFree(h->d_ded); cudaFree(h->d_w); cudaFree(h->d_wprep); cudaFree(h->d_bad);
    cudaFree(h->d_delay);
    cudaFree(h->d_gen_par);
    if (h->ev_delay) cudaEventDestroy(h->ev_delay);
    cudaFree(h->d_tmp_in); cudaFree(h->d_tmp_out);
    delete h;
    return 0;
}

// ---- steering weights: the host loop of beamformer.cu:230-241 -------------
// Arithmetic types follow the reference expression by expression: the channel
// frequency uses a *float* bandwidth and an integer division by N_GPUS-1
// (beamformer.cu:173,233); the geometry term is float (sin(float) is the float
// overload under nvcc); the phase, cos/sin and round are double.
int dsabf_compute_weights(const float *pos, const float *dir, int n_ant, int n_beams, int n_chan, int gpu, int8_t *w)
{
    if (!pos || !dir || !w || n_ant < 1 || n_beams < 1 || n_chan < 1) return fail(DSABF_E_INVALID, "bad argument");
    const double kEndF = 1.53, kStartF = 1.28, kC = 299792458.0, kPi = 3.14159265358979;   // beamformer.hh:65-72
    const int kTotChannels = 2048, kNGpus = 8, kZeroPt = 0, kMaxVal = 127;
    const float bw = (float)((kEndF - kStartF) / kTotChannels);
    std::vector<float> sth(n_beams), sph(n_beams);
    for (int k = 0; k < n_beams; k++) { sth[k] = sinf(dir[2 * k]); sph[k] = sinf(dir[2 * k + 1]); }
    for (int i = 0; i < n_chan; i++) {
        const int idx = kZeroPt + gpu * kTotChannels / (kNGpus - 1) + i;
        const float freq = (float)(kEndF - (double)((float)idx * bw));
        const float wavelength = (float)(kC / (1E9 * (double)freq));
        for (int j = 0; j < n_ant; j++) {
            const float x = pos[3 * j], y = pos[3 * j + 1];
            int8_t *row = w + ((size_t)i * n_ant + j) * n_beams * 2;
            for (int k = 0; k < n_beams; k++) {
                const float gx = x * sth[k];
                const float gy = y * sph[k];
                const float g = gx + gy;
                const double ph = -2 * kPi * (double)g / (double)wavelength;
                row[2 * k] = (int8_t)round(kMaxVal * cos(ph));
                row[2 * k + 1] = (int8_t)round(kMaxVal * sin(ph));
            }
        }
    }
    return 0;
}

int dsabf_set_weights(dsabf_handle *h, const int8_t *w)
{
    if (!h || !w) return fail(DSABF_E_INVALID, "NULL argument");
    CK(cudaSetDevice(h->device));
    const dsabf_config &c = h->cfg;
    const size_t wbytes = (size_t)c.n_channels * c.n_antennas * c.n_beams * 2;
    CK(cudaStreamSynchronize(h->s_comp));
    CK(cudaMemcpyAsync(h->d_w, w, wbytes, cudaMemcpyHostToDevice, h->s_comp));
    CK(cudaMemsetAsync(h->d_bad, 0, sizeof(int), h->s_comp));
    CK(dsabf::launch_prep_weights(h->d_w, h->d_wprep, c.n_channels, c.n_antennas, c.n_beams, h->d_bad, h->s_comp));
    h->launches++;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, h->d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->s_comp));
    CK(cudaStreamSynchronize(h->s_comp));
    if (bad) { h->weights_set = false; return fail(DSABF_E_WEIGHT, "weight component -128 is not representable after negation"); }
    h->weights_set = true;
    return 0;
}

// ---- ingest ----------------------------------------------------------------
void *dsabf_ring_slot_ptr(dsabf_handle *h, int slot)
{
    if (!h || slot < 0 || slot >= h->cfg.n_ring_slots) return nullptr;
    return h->d_ring + (size_t)slot * h->block_bytes;
}

int dsabf_enqueue_h2d(dsabf_handle *h, int slot, const void *host_src, size_t bytes, uint64_t *ticket)
{
    if (!h || !host_src || slot < 0 || slot >= h->cfg.n_ring_slots || bytes > h->block_bytes) return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    // never overwrite a slot the fused kernel may still be reading
    if (h->slot_has_comp[slot]) CK(cudaStreamWaitEvent(h->s_h2d, h->ev_slot_comp[slot], 0));
    CK(cudaMemcpyAsync(h->d_ring + (size_t)slot * h->block_bytes, host_src, bytes, cudaMemcpyHostToDevice, h->s_h2d));
    CK(cudaEventRecord(h->ev_slot_h2d[slot], h->s_h2d));
    h->slot_has_h2d[slot] = 1;
    const uint64_t t = h->n_xfer++;
    CK(cudaEventRecord(h->ev_xfer[t % kTicketRing], h->s_h2d));
    if (ticket) *ticket = t;
    return 0;
}

static int query_ticket(dsabf_handle *h, const cudaEvent_t *ring, uint64_t issued, uint64_t ticket)
{
    if (!h || ticket >= issued) { fail(DSABF_E_INVALID, "unknown ticket"); return -1; }
    if (issued - ticket > (uint64_t)kTicketRing) return 1;   // long since recycled => done
    cudaError_t e = cudaEventQuery(ring[ticket % kTicketRing]);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) return 0;
    fail((int)e, "cudaEventQuery");
    return -1;
}
int dsabf_query_transfer(dsabf_handle *h, uint64_t ticket) { return query_ticket(h, h ? h->ev_xfer : nullptr, h ? h->n_xfer : 0, ticket); }
int dsabf_query_analysis(dsabf_handle *h, uint64_t ticket) { return query_ticket(h, h ? h->ev_ana : nullptr, h ? h->n_ana : 0, ticket); }

// ---- the hot path ------------------------------------------------------------
int dsabf_enqueue_block(dsabf_handle *h, int slot, int n_chunks, float *beam_out_host, float *dedispersed_host, int n_valid_dedisp,
                        uint64_t *ticket)
{
    if (!h || slot < 0 || slot >= h->cfg.n_ring_slots || n_chunks < 0 || n_chunks > h->cfg.n_chunks_per_block)
        return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    const dsabf_config &c = h->cfg;
    const size_t cpb = c.n_chunks_per_block;
    float *d_out = h->d_out + (size_t)slot * cpb * h->chunk_out_floats;
    float *d_ded = h->d_ded + (size_t)slot * cpb * c.n_beams;
    if (h->slot_has_h2d[slot]) CK(cudaStreamWaitEvent(h->s_comp, h->ev_slot_h2d[slot], 0));
    if (h->slot_has_d2h[slot]) CK(cudaStreamWaitEvent(h->s_comp, h->ev_slot_d2h[slot], 0));   // previous power of this slot copied out
    int rc = run_fused(h, h->d_ring + (size_t)slot * h->block_bytes, n_chunks, d_out, nullptr, nullptr, c.exact_power != 0, h->s_comp);
    if (rc) return rc;
    const int n_ded = dedispersed_host ? (n_valid_dedisp < n_chunks ? (n_valid_dedisp < 0 ? 0 : n_valid_dedisp) : n_chunks) : 0;
    if (n_ded > 0) {
        CK(dsabf::launch_dm0(d_out, d_ded, n_ded, c.n_outputs_per_chunk, c.n_channels, c.n_beams, h->s_comp));
        h->launches++;
    }
    CK(cudaEventRecord(h->ev_slot_comp[slot], h->s_comp));
    h->slot_has_comp[slot] = 1;
    CK(cudaStreamWaitEvent(h->s_d2h, h->ev_slot_comp[slot], 0));
    if (beam_out_host && n_chunks > 0)
        CK(cudaMemcpyAsync(beam_out_host, d_out, (size_t)n_chunks * h->chunk_out_floats * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
    if (n_ded > 0)
        CK(cudaMemcpyAsync(dedispersed_host, d_ded, (size_t)n_ded * c.n_beams * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
    CK(cudaEventRecord(h->ev_slot_d2h[slot], h->s_d2h));
    h->slot_has_d2h[slot] = 1;
    const uint64_t t = h->n_ana++;
    CK(cudaEventRecord(h->ev_ana[t % kTicketRing], h->s_d2h));
    if (ticket) *ticket = t;
    return 0;
}

float *dsabf_block_out_ptr(dsabf_handle *h, int slot)
{
    if (!h || slot < 0 || slot >= h->cfg.n_ring_slots) return nullptr;
    return h->d_out + (size_t)slot * h->cfg.n_chunks_per_block * h->chunk_out_floats;
}

int dsabf_sync(dsabf_handle *h)
{
    if (!h) return fail(DSABF_E_INVALID, "NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->s_h2d));
    CK(cudaStreamSynchronize(h->s_comp));
    CK(cudaStreamSynchronize(h->s_d2h));
    return 0;
}

int dsabf_beamform_device(dsabf_handle *h, const void *d_packed, int n_chunks, float *d_out, void *stream)
{
    if (!h || !d_packed || !d_out || n_chunks < 0) return fail(DSABF_E_INVALID, "bad argument");
    // bulk copies (resident kernel) and 16-byte vector loads (streamed kernel) read d_packed; float stores are scalar
    if (((uintptr_t)d_packed & 15) || ((uintptr_t)d_out & 3)) return fail(DSABF_E_INVALID, "d_packed must be 16-byte aligned, d_out 4-byte aligned");
    CK(cudaSetDevice(h->device));
    return run_fused(h, (const uint8_t *)d_packed, n_chunks, d_out, nullptr, nullptr, h->cfg.exact_power != 0,
                     stream ? (cudaStream_t)stream : h->s_comp);
}

int dsabf_beamform_host(dsabf_handle *h, const void *packed, int n_chunks, float *out)
{
    if (!h || !packed || !out || n_chunks < 0) return fail(DSABF_E_INVALID, "bad argument");
    if (n_chunks == 0) return 0;
    CK(cudaSetDevice(h->device));
    const size_t in_b = (size_t)n_chunks * h->chunk_bytes, out_b = (size_t)n_chunks * h->chunk_out_floats * sizeof(float);
    int rc = ensure_tmp(h, in_b, out_b);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_tmp_in, packed, in_b, cudaMemcpyHostToDevice, h->s_comp));
    rc = run_fused(h, h->d_tmp_in, n_chunks, h->d_tmp_out, nullptr, nullptr, h->cfg.exact_power != 0, h->s_comp);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->d_tmp_out, out_b, cudaMemcpyDeviceToHost, h->s_comp));
    CK(cudaStreamSynchronize(h->s_comp));
    return 0;
}

int dsabf_debug_voltages_host(dsabf_handle *h, const void *packed, int n_chunks, int32_t *V, int64_t *P)
{
    if (!h || !packed || !V || n_chunks <= 0) return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    const dsabf_config &c = h->cfg;
    const size_t in_b = (size_t)n_chunks * h->chunk_bytes, out_f = (size_t)n_chunks * h->chunk_out_floats;
    const size_t v_b = (size_t)c.n_channels * n_chunks * h->cols_per_chunk * c.n_beams * 2 * sizeof(int32_t);
    int rc = ensure_tmp(h, in_b, out_f * sizeof(float));
    if (rc) return rc;
    int32_t *d_v = nullptr;
    long long *d_p = nullptr;
    CK(cudaMalloc(&d_v, v_b));
    if (P) { cudaError_t e = cudaMalloc(&d_p, out_f * sizeof(long long)); if (e != cudaSuccess) { cudaFree(d_v); return fail((int)e, "alloc P"); } }
    cudaError_t e = cudaMemcpyAsync(h->d_tmp_in, packed, in_b, cudaMemcpyHostToDevice, h->s_comp);
    if (e == cudaSuccess) {
        rc = run_fused(h, h->d_tmp_in, n_chunks, h->d_tmp_out, d_v, d_p, true, h->s_comp);
        if (!rc) e = cudaMemcpyAsync(V, d_v, v_b, cudaMemcpyDeviceToHost, h->s_comp);
        if (!rc && e == cudaSuccess && P) e = cudaMemcpyAsync(P, d_p, out_f * sizeof(long long), cudaMemcpyDeviceToHost, h->s_comp);
        if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(h->s_comp);
    }
    cudaFree(d_v);
    if (d_p) cudaFree(d_p);
    if (rc) return rc;
    if (e != cudaSuccess) return fail((int)e, "debug voltages");
    return 0;
}

// debug: run one fused launch on device-resident data with the in-kernel clock trace on; trace_out: 2048 int64
int dsabf_debug_trace(dsabf_handle *h, const void *d_packed, int n_chunks, float *d_out, long long *trace_out)
{
    if (!h || !d_packed || !d_out || !trace_out) return fail(DSABF_E_INVALID, "bad argument");
#ifndef DSABF_TRACE
    return fail(DSABF_E_STATE, "library built without -DDSABF_TRACE (make EXTRA=-DDSABF_TRACE)");
#else
    CK(cudaSetDevice(h->device));
    CK(cudaMalloc(&h->d_trace, 2048 * sizeof(long long)));
    CK(cudaMemset(h->d_trace, 0, 2048 * sizeof(long long)));
    int rc = run_fused(h, (const uint8_t *)d_packed, n_chunks, d_out, nullptr, nullptr, h->cfg.exact_power != 0, h->s_comp);
    cudaError_t e = cudaStreamSynchronize(h->s_comp);
    if (!rc && e == cudaSuccess) e = cudaMemcpy(trace_out, h->d_trace, 2048 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(h->d_trace);
    h->d_trace = nullptr;
    if (rc) return rc;
    if (e != cudaSuccess) return fail((int)e, "trace");
    return 0;
#endif
}

// ---- multi-GPU exchange ----------------------------------------------------------
int dsabf_transpose_pull_device(dsabf_handle *h, const void *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams,
                                float *d_full, void *stream)
{
    if (!h || !peer_power || !d_full || world < 1 || rank < 0 || rank >= world || n_t < 0 || c_local < 0 || n_beams < 1)
        return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->s_comp;
    CK(dsabf::launch_transpose_pull(reinterpret_cast<const float *const *>(peer_power), world, rank, n_t, c_local, n_beams, d_full, h->num_sms, st));
    h->launches++;
    return 0;
}

// ---- dedispersion -------------------------------------------------------------
int dsabf_dedisperse_device(dsabf_handle *h, const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *h_delay, int n_dm,
                            int n_t_out, float *d_out, void *stream)
{
    if (!h || !d_power || !h_delay || !d_out || n_t < 1 || n_chan < 1 || n_beams < 2 || n_dm < 1 || n_t_out < 1)
        return fail(DSABF_E_INVALID, "bad argument");
    for (size_t i = 0; i < (size_t)n_dm * n_chan; i++)
        if (h_delay[i] < 0) return fail(DSABF_E_INVALID, "negative delay");
    // with n_beams a multiple of 4 the staged kernel moves 16-byte pieces of the rows with cp.async (other beam counts
    // take its scalar path; the single-trial kernel checks its own alignment and falls back to the staged one)
    if ((n_beams & 3) == 0 && ((uintptr_t)d_power & 15)) return fail(DSABF_E_INVALID, "d_power must be 16-byte aligned");
    if (((uintptr_t)d_power & 3) || ((uintptr_t)d_out & 3)) return fail(DSABF_E_INVALID, "d_power and d_out must be 4-byte aligned");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->s_comp;
    int rc = upload_delays(h, h_delay, (size_t)n_dm * n_chan, st);
    if (rc) return rc;
    CK(dsabf::launch_dedisperse(d_power, n_t, n_chan, n_beams, h->d_delay, n_dm, n_t_out, d_out, st));
    h->launches++;
    CK(cudaEventRecord(h->ev_delay, st));
    h->delay_in_use = true;
    return 0;
}

int dsabf_dedisperse_single_peers_device(dsabf_handle *h, const void *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams,
                                         const int32_t *h_delay, int n_t_out, float *d_out, void *stream)
{
    if (!h || !peer_power || !h_delay || !d_out || world < 1 || rank < 0 || rank >= world || n_t < 1 || c_local < 1 || n_beams < 4 || n_t_out < 1)
        return fail(DSABF_E_INVALID, "bad argument");
    const size_t n = (size_t)world * c_local;
    for (size_t i = 0; i < n; i++)
        if (h_delay[i] < 0) return fail(DSABF_E_INVALID, "negative delay");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->s_comp;
    int rc = upload_delays(h, h_delay, n, st);
    if (rc) return rc;
    CK(dsabf::launch_dedisperse_single_peers(reinterpret_cast<const float *const *>(peer_power), world, rank, n_t, c_local, n_beams, h->d_delay,
                                             n_t_out, d_out, st));
    h->launches++;
    CK(cudaEventRecord(h->ev_delay, st));
    h->delay_in_use = true;
    return 0;
}

int dsabf_dedisperse_host(dsabf_handle *h, const float *power, int n_t, int n_chan, int n_beams, const int32_t *delay, int n_dm,
                          int n_t_out, float *out)
{
    if (!h || !power || !out) return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    const size_t in_b = sizeof(float) * (size_t)n_t * n_chan * n_beams, out_b = sizeof(float) * (size_t)n_dm * n_t_out * n_beams;
    float *d_in = nullptr, *d_o = nullptr;
    CK(cudaMalloc(&d_in, in_b));
    cudaError_t e = cudaMalloc(&d_o, out_b);
    if (e != cudaSuccess) { cudaFree(d_in); return fail((int)e, "alloc"); }
    e = cudaMemcpy(d_in, power, in_b, cudaMemcpyHostToDevice);
    int rc = 0;
    if (e == cudaSuccess) rc = dsabf_dedisperse_device(h, d_in, n_t, n_chan, n_beams, delay, n_dm, n_t_out, d_o, nullptr);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(h->s_comp);      // dsabf_dedisperse_device is stream-ordered, not blocking
    if (e == cudaSuccess && !rc) e = cudaMemcpy(out, d_o, out_b, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_o);
    if (rc) return rc;
    if (e != cudaSuccess) return fail((int)e, "dedisperse_host");
    return 0;
}

// ---- DEBUG-mode synthetic input on the GPU (test_data_generator.hh:63-95) -----------------
// The float library calls of the reference's host loop (sinf of the source angles; the channel frequency from the
// *double* bandwidth macro and the integer sub-band offset gpu*2048/7, :72; the float wavelength, :74) are evaluated
// here on the host, exactly as the CPU generator does, and only the per-sample arithmetic runs in the kernel.
int dsabf_generate_test_data_device(dsabf_handle *h, const float *pos, const float *sources, int n_sources, int batch, int gpu, void *d_dst,
                                    void *stream)
{
    if (!h || !pos || (n_sources > 0 && !sources) || !d_dst || batch < 0 || n_sources < 0) return fail(DSABF_E_INVALID, "bad argument");
    if ((uintptr_t)d_dst & 15) return fail(DSABF_E_INVALID, "d_dst must be 16-byte aligned");
    CK(cudaSetDevice(h->device));
    const dsabf_config &c = h->cfg;
    const int per_batch = 1024;                                  // N_SOURCES_PER_BATCH
    const long first = (long)batch * per_batch;
    const int n_valid = (int)(n_sources - first < 0 ? 0 : (n_sources - first > per_batch ? per_batch : n_sources - first));
    std::vector<float> host((size_t)2 * c.n_antennas + 2 * per_batch + c.n_channels, 0.f);
    float *pxy = host.data(), *ssin = pxy + 2 * c.n_antennas, *wl = ssin + 2 * per_batch;
    for (int k = 0; k < c.n_antennas; k++) { pxy[2 * k] = pos[3 * k]; pxy[2 * k + 1] = pos[3 * k + 1]; }
    for (int s = 0; s < n_valid; s++) { ssin[2 * s] = sinf(sources[2 * (first + s)]); ssin[2 * s + 1] = sinf(sources[2 * (first + s) + 1]); }
    const double bw = (1.53 - 1.28) / 2048;
    for (int i = 0; i < c.n_channels; i++) {
        const float freq = (float)(1.53 - (0 + gpu * 2048 / (8 - 1) + i) * bw);
        wl[i] = (float)(299792458.0 / (1E9 * (double)freq));
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : h->s_comp;
    if (!h->d_gen_par) CK(cudaMalloc(&h->d_gen_par, sizeof(float) * (2 * 256 + 2 * per_batch + 65536)));
    if (host.size() > (size_t)(2 * 256 + 2 * per_batch + 65536)) return fail(DSABF_E_INVALID, "more than 65536 channels");
    CK(cudaStreamSynchronize(st));                               // the parameter block is reused: previous generator launch must be done
    CK(cudaMemcpyAsync(h->d_gen_par, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                               // `host` goes out of scope
    CK(dsabf::launch_generate_test_data((uint8_t *)d_dst, h->d_gen_par, h->d_gen_par + 2 * c.n_antennas, h->d_gen_par + 2 * c.n_antennas + 2 * per_batch,
                                        per_batch, n_valid, c.n_antennas, c.n_channels, h->cols_per_chunk, st));
    h->launches++;
    return 0;
}

int dsabf_generate_test_data_host(dsabf_handle *h, const float *pos, const float *sources, int n_sources, int batch, int gpu, void *dst)
{
    if (!h || !dst) return fail(DSABF_E_INVALID, "bad argument");
    CK(cudaSetDevice(h->device));
    const size_t bytes = h->chunk_bytes * 1024;
    int rc = ensure_tmp(h, bytes, 0);
    if (rc) return rc;
    rc = dsabf_generate_test_data_device(h, pos, sources, n_sources, batch, gpu, h->d_tmp_in, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dst, h->d_tmp_in, bytes, cudaMemcpyDeviceToHost, h->s_comp));
    CK(cudaStreamSynchronize(h->s_comp));
    return 0;
}

// ---- dedispersion plan (host): DM trial list and delay table -------------------------
// Trial spacing of sandbox/Dispersion Theory.ipynb cells 1-2: each next trial is the DM at which the smearing of a
// pulse dedispersed at the previous trial has grown by the tolerance epsilon (channel width B MHz at nu GHz, intrinsic
// width 40 us, sampling 131 us); 2048 channels to DM 2000 gives 1627 trials.
int dsabf_dm_trials(double dm_max, int n_chan_total, double *dms, int capacity)
{
    if (n_chan_total < 1 || (capacity > 0 && !dms)) { fail(DSABF_E_INVALID, "bad argument"); return -1; }
    const double eps2 = 1.25 * 1.25, nc2 = (double)n_chan_total * n_chan_total;
    const double alpha = 1.0 / (16 + nc2), centre_ghz = (1.28 + 1.53) / 2, chan_mhz = (1.53 - 1.28) / n_chan_total * 1000;
    const double width2_us = 40.0 * 40.0 + 0.0 * 0.0 + 131.0 * 131.0;
    const double smear = centre_ghz * centre_ghz * centre_ghz / (8.3 * chan_mhz);
    const double floor_term = alpha * width2_us * (eps2 - 1) * smear * smear;
    int n = 0;
    for (double dm = 0;; dm = nc2 * alpha * dm + 4 * sqrt(alpha * (eps2 - nc2 * alpha) * dm * dm + floor_term)) {
        if (n < capacity) dms[n] = dm;
        n++;
        if (!(dm < dm_max)) break;
    }
    return n;
}

// Delay of channel f (centre freq_ghz[f]) relative to the top of the band, in output samples of 16 x 131 us
// (cell 5 of the same notebook): int(4.15 DM (f^-2 - 1.53^-2) / (0.131 * 16)), truncated toward zero.
int dsabf_dm_delays(const double *dms, int n_dm, const double *freq_ghz, int n_chan, int32_t *delay)
{
    if (!dms || !freq_ghz || !delay || n_dm < 1 || n_chan < 1) return fail(DSABF_E_INVALID, "bad argument");
    const double top = 1.53, sample_ms = 0.131 * 16;
    for (int d = 0; d < n_dm; d++)
        for (int f = 0; f < n_chan; f++)
            delay[(size_t)d * n_chan + f] = (int32_t)(4.15 * dms[d] * (-1.0 / (top * top) + 1.0 / (freq_ghz[f] * freq_ghz[f])) / sample_ms);
    return 0;
}

// ---- introspection ------------------------------------------------------------
uint64_t dsabf_kernel_launches(dsabf_handle *h) { return h ? h->launches : 0; }

int dsabf_enable_kernel_timing(dsabf_handle *h, int on)
{
    if (!h) return fail(DSABF_E_INVALID, "NULL handle");
    h->timing = on != 0;
    return 0;
}

int dsabf_fused_kernel_ms(dsabf_handle *h, double *avg_ms, int *n_launches, int reset)
{
    if (!h || !avg_ms) return fail(DSABF_E_INVALID, "NULL argument");
    CK(cudaSetDevice(h->device));
    double tot = 0;
    for (size_t i = 0; i < h->t_used; i++) {
        CK(cudaEventSynchronize(h->t_stop[i]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, h->t_start[i], h->t_stop[i]));
        tot += ms;
    }
    *avg_ms = h->t_used ? tot / h->t_used : 0.0;
    if (n_launches) *n_launches = (int)h->t_used;
    if (reset) h->t_used = 0;
    return 0;
}

}  // extern "C"
