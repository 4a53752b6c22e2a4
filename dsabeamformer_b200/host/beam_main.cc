// beam_main.cc -- the `beam` program: command line, set-up, observation loop, reports.
//
// Drop-in for the reference's main() (/root/reference/src/beamformer.cu:12-628) with the GPU
// work behind the C ABI of include/dsabf.h: same options (-c -k -g -p -d [-s] -h), same config
// files, same default geometry (:135-147), same loop pacing (:364-534 via ObservationLoopState),
// same stdout milestones and timing report (:539-557), same bin/data.py (:568-571).
// Built twice: -DBEAM_DEBUG=1 -> bin/beam_debug (the reference's `make debug` / `fast_debug`),
// otherwise -> bin/beam (`make all` / `make verbose`, reads the PSRDADA-style ring).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

#include <cuda_runtime.h>

#include "beam_config.hh"
#include "observation_loop.hh"
#include "dada_hdu.h"          // PSRDADA API: $(PSRDADA_INC) = host/psrdada_compat + the shim library, or the real thing
#include "ascii_header.h"
#if BEAM_DEBUG
#include "test_data_generator.hh"
#else
#include "dada_handler.hh"
#endif

#ifndef BEAM_VERBOSE
#define BEAM_VERBOSE 0
#endif

using namespace beam;

// gpuErrchk convention of the reference (beamformer.cuh:19-29): print and exit with the code
static void check(int rc, const char *file, int line)
{
    if (rc != 0) {
        fprintf(stderr, "GPUassert: %s %s %d\n", dsabf_last_error_string(), file, line);
        exit(rc);
    }
}
#define CHK(x) check((x), __FILE__, __LINE__)
static void cuda_check(cudaError_t e, const char *file, int line)
{
    if (e != cudaSuccess) { fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(e), file, line); exit((int)e); }
}
#define CUCHK(x) cuda_check((x), __FILE__, __LINE__)

struct Stopwatch {   // START_TIMER / STOP_RECORD_TIMER (beamformer.cuh:43-58): device-complete wall time
    dsabf_handle *h;
    std::chrono::steady_clock::time_point t0;
    void start() { t0 = std::chrono::steady_clock::now(); }
    float stop_ms() { CHK(dsabf_sync(h)); return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

int main(int argc, char *argv[])
{
    const bool debug = BEAM_DEBUG;
    const bool verbose = BEAM_VERBOSE;
    dsabf_config cfg;
    CHK(dsabf_default_config(&cfg, debug));
    if (const char *dev = getenv("DSABF_DEVICE")) cfg.device = atoi(dev);   // CUDA ordinal; -g stays the sub-band index
    // The reference fixes the array shape at compile time (beamformer.hh:47-53); here it is run-time.  These
    // overrides select e.g. the scaled array (256 antennas x 2048 beams) without a rebuild.
    if (const char *v = getenv("DSABF_N_ANTENNAS")) cfg.n_antennas = atoi(v);
    if (const char *v = getenv("DSABF_N_BEAMS")) cfg.n_beams = atoi(v);
    if (const char *v = getenv("DSABF_N_CHANNELS")) cfg.n_channels = atoi(v);

    std::vector<Antenna> pos(cfg.n_antennas);
    std::vector<Direction> dir(cfg.n_beams);
    bool pos_set = false, dir_set = false;
    int gpu = 0;
#if BEAM_DEBUG
    const char *options = "s:g:p:d:h";
    std::string source_file;
#else
    const char *options = "c:k:g:p:d:h";
    int core = 0;
    unsigned in_key = 0x0000dada;
#endif
    int arg;
    while ((arg = getopt(argc, argv, options)) != -1) {
        switch (arg) {
#if BEAM_DEBUG
        case 's': source_file = optarg; break;
#else
        case 'c': core = atoi(optarg); break;
        case 'k':
            if (sscanf(optarg, "%x", &in_key) != 1) { fprintf(stderr, "beam: could not parse key from %s\n", optarg); return EXIT_FAILURE; }
            break;
#endif
        case 'g': gpu = atoi(optarg); break;
        case 'p':
            if (read_table(optarg, cfg.n_antennas, pos, "antennas", "N_ANTENNAS") < 0) { fprintf(stderr, "beam: could not parse antenna location file from %s\n", optarg); return EXIT_FAILURE; }
            pos_set = true;
            break;
        case 'd':
            if (read_table(optarg, cfg.n_beams, dir, "beams", "expected") < 0) { fprintf(stderr, "beam: could not parse direction file from %s\n", optarg); return EXIT_FAILURE; }
            dir_set = true;
            break;
        case 'h': usage(debug); return EXIT_SUCCESS;
        default: usage(debug); return EXIT_FAILURE;
        }
    }
    if (!pos_set)   // beamformer.cu:135-140
        for (int i = 0; i < cfg.n_antennas; i++) pos[i].x = i * 500.0 / (cfg.n_antennas - 1) - 250.0;
    if (!dir_set)   // beamformer.cu:142-147
        for (int i = 0; i < cfg.n_beams; i++) dir[i].theta = i * deg2rad(2 * kHalfFovDeg) / (cfg.n_beams - 1) - deg2rad(kHalfFovDeg);
    if (verbose) print_config(cfg, debug);

    dsabf_handle *h = nullptr;
    CHK(dsabf_create(&cfg, &h));
    const size_t chunk_bytes = dsabf_chunk_bytes(&cfg), block_bytes = dsabf_block_bytes(&cfg), chunk_out = dsabf_chunk_out_floats(&cfg);
    const int cpb = cfg.n_chunks_per_block;

#if BEAM_DEBUG
    TestDataGenerator generator(cfg);
    generator.use_gpu(h);
    if (!source_file.empty()) generator.read_in_source_directions(source_file.c_str(), verbose);
    else std::cout << "Using BOGUS DATA " << std::endl;
    const int n_pt_sources = generator.n_pt_sources();
    float *dedispersed_out = nullptr;
    CUCHK(cudaHostAlloc((void **)&dedispersed_out, sizeof(float) * (size_t)cfg.n_beams * n_pt_sources, 0));
    memset(dedispersed_out, 0, sizeof(float) * (size_t)cfg.n_beams * n_pt_sources);
#else
    DadaHandler dada("beam", core, (key_t)in_key, block_bytes, verbose);
    if (dada.ring_blocks() < 2) { fprintf(stderr, "beam: the input ring needs at least 2 blocks\n"); return EXIT_FAILURE; }
#endif

    // steering weights: beamformer.cu:230-241 behind dsabf_compute_weights
    {
        std::vector<float> p3(3 * (size_t)cfg.n_antennas), d2(2 * (size_t)cfg.n_beams);
        for (int i = 0; i < cfg.n_antennas; i++) { p3[3 * i] = pos[i].x; p3[3 * i + 1] = pos[i].y; p3[3 * i + 2] = pos[i].z; }
        for (int i = 0; i < cfg.n_beams; i++) { d2[2 * i] = dir[i].theta; d2[2 * i + 1] = dir[i].phi; }
        std::vector<int8_t> w((size_t)cfg.n_channels * cfg.n_antennas * cfg.n_beams * 2);
        CHK(dsabf_compute_weights(p3.data(), d2.data(), cfg.n_antennas, cfg.n_beams, cfg.n_channels, gpu, w.data()));
        CHK(dsabf_set_weights(h, w.data()));
    }

    // final data product buffer (pinned, caller-owned): kOutBufs blocks of power used in turn.  A block's buffer is
    // reused by block k + kOutBufs, which cannot be queued before block k was seen complete (TQ - A < MAX_TOTAL_SEP).
    constexpr int kOutBufs = 4;
    static_assert(kOutBufs >= kMaxTotalSep, "output buffers must cover the blocks in flight");
    const size_t out_block_floats = chunk_out * cpb;
    float *beam_out = nullptr;
    CUCHK(cudaHostAlloc((void **)&beam_out, sizeof(float) * out_block_floats * kOutBufs, 0));

    // Output sink (the reference drops beam_out, beamformer.cu:485-488): with DSABF_OUT_KEY=<hex>
    // every finished block of power, float [out_time][channel][beam], is appended to that PSRDADA-style ring for the
    // next stage (dedispersion / search) to read.
    dada_hdu_t *sink = nullptr;
    multilog_t *sink_log = nullptr;
    uint64_t blocks_sunk = 0;
    if (const char *k = getenv("DSABF_OUT_KEY")) {
        unsigned out_key = 0;
        sink_log = multilog_open("beam_out", 0);
        multilog_add(sink_log, stderr);
        sink = dada_hdu_create(sink_log);
        if (sscanf(k, "%x", &out_key) == 1) dada_hdu_set_key(sink, (key_t)out_key);
        if (sscanf(k, "%x", &out_key) != 1 || dada_hdu_connect(sink) < 0 || dada_hdu_lock_write(sink) < 0) {
            fprintf(stderr, "beam: could not connect to output ring %s\n", k);
            return EXIT_FAILURE;
        }
        if (ipcbuf_get_bufsz((ipcbuf_t *)sink->data_block) < sizeof(float) * out_block_floats) {
            fprintf(stderr, "beam: output ring blocks are smaller than %zu bytes\n", sizeof(float) * out_block_floats);
            return EXIT_FAILURE;
        }
        const uint64_t hsz = ipcbuf_get_bufsz(sink->header_block);
        char *hdr = ipcbuf_get_next_write(sink->header_block);
        memset(hdr, 0, hsz);
        ascii_header_set(hdr, "HDR_VERSION", "%s", "1.0");
        ascii_header_set(hdr, "INSTRUMENT", "%s", "DSAX");
        ascii_header_set(hdr, "DTYPE", "%s", "float32");
        ascii_header_set(hdr, "ORDER", "%s", "TFB");
        ascii_header_set(hdr, "NBEAM", "%d", cfg.n_beams);
        ascii_header_set(hdr, "NCHAN", "%d", cfg.n_channels);
        ascii_header_set(hdr, "NTIME", "%d", cfg.n_outputs_per_chunk * cpb);
        ascii_header_set(hdr, "SUBBAND", "%d", gpu);
        ipcbuf_mark_filled(sink->header_block, hsz);
    }
    const bool have_sink = sink != nullptr;

    ObservationLoopState obs(h, kMaxTransferSep, kMaxTotalSep, cpb, cfg.n_ring_slots, debug);
#if BEAM_DEBUG
    obs.set_n_pt_sources(n_pt_sources);
#else
    dada.read_headers();
    if (dada.get_block_size() != block_bytes) {      // the H2D copies read block_bytes out of every ring block
        fprintf(stderr, "beam: ring blocks are %llu bytes, this array shape needs %zu (dada_db -b %zu)\n",
                (unsigned long long)dada.get_block_size(), block_bytes, block_bytes);
        return EXIT_FAILURE;
    }
#endif
    if (verbose) {
        std::cout << "Executing beamformer (B200 fused path)\n";
        std::cout << "MAX_TOTAL_SEP: " << kMaxTotalSep << "\n";
        std::cout << "MAX_TRANSFER_SEP: " << kMaxTransferSep << std::endl;
    }
#if !BEAM_DEBUG
    std::cout << "Burning IN" << std::endl;
    for (int i = 0; i < kBurnIn; i++) { if (!dada.read() || dada.get_bytes_read() == 0) { dada.close(); break; } dada.close(); }
    std::cout << "Done Burn in" << std::endl;
#endif

    Stopwatch sw{h, {}};
    sw.start();
    float observation_time_ms = 0;
#if !BEAM_DEBUG
    uint64_t blocks_read = 0;   // data blocks taken from the ring after the burn-in
#endif

    while (!obs.observations_complete()) {
        if (verbose) { std::cout << "##########################################" << std::endl; std::cout << obs << std::endl; }

        // ---- copy data to the GPU (beamformer.cu:378-440) ----
        if (obs.ready_for_transfer()) {
#if !BEAM_DEBUG
            dada.retire(h);
            if (dada.in_flight() == 0) {            // PSRDADA: one open block per reader; the previous one is still being copied otherwise
                if (verbose) std::cout << "READING FROM PSRDADA" << std::endl;
                char *block = dada.read();
                if (block && !dada.check_transfers_complete()) {
                    uint64_t ticket = 0;
                    CHK(dsabf_enqueue_h2d(h, obs.next_transfer_slot(), block, block_bytes, &ticket));
                    obs.transfer_enqueued(ticket);
                    dada.close_when_copied(h, ticket);
                    blocks_read++;
                } else {
                    obs.set_transfers_complete(true);
                    dada.close();
                }
            }
#else
            if (generator.need_to_generate_more(obs.blocks_transferred())) {
                std::cout << "Generating new source data" << std::endl;
                observation_time_ms += sw.stop_ms();
                generator.generate_test_data(pos, gpu);
                sw.start();
                std::cout << "done generating test data" << std::endl;
            }
            if (generator.data_ready_for_transfer(obs.blocks_transfer_queue())) {
                const size_t off = (block_bytes * obs.blocks_transfer_queue()) % generator.bytes();
                uint64_t ticket = 0;
                CHK(dsabf_enqueue_h2d(h, obs.next_transfer_slot(), generator.data() + off, block_bytes, &ticket));
                obs.transfer_enqueued(ticket);
            }
            obs.check_transfers_complete_debug();
#endif
        }
        obs.check_transfer_events(verbose);

        // ---- beamform one block: the whole expand/GEMM/detect/D2H fan-out of beamformer.cu:452-526 ----
        if (obs.ready_for_analysis()) {
            const uint64_t first_chunk = obs.current_analysis_chunk(0);
            if (verbose) std::cout << "Queueing Beamforming. Start Dir = " << first_chunk << std::endl;
            float *out_host = beam_out + (obs.blocks_analysis_queue() % kOutBufs) * out_block_floats;
            uint64_t ticket = 0;
#if BEAM_DEBUG
            long n_valid = (long)n_pt_sources - (long)first_chunk;     // `current_gemm < n_pt_sources` guard (:492)
            if (n_valid > cpb) n_valid = cpb;
            if (n_valid < 0) n_valid = 0;
            if (verbose && n_valid > 0) std::cout << "Current GEMM: " << first_chunk << std::endl;
            CHK(dsabf_enqueue_block(h, obs.next_analysis_slot(), cpb, out_host,
                                    n_valid > 0 ? dedispersed_out + (size_t)first_chunk * cfg.n_beams : nullptr, (int)n_valid, &ticket));
#else
            CHK(dsabf_enqueue_block(h, obs.next_analysis_slot(), cpb, out_host, nullptr, 0, &ticket));
#endif
            obs.current_analysis_chunk(cpb - 1);
            obs.analysis_enqueued(ticket);
        }
        obs.check_analysis_events(verbose);
        while (have_sink && blocks_sunk < obs.blocks_analyzed()) {       // hand finished blocks to the next stage, in order
            uint64_t id = 0;
            char *dst = ipcio_open_block_write(sink->data_block, &id);
            if (!dst) { fprintf(stderr, "beam: could not open an output block\n"); return EXIT_FAILURE; }
            memcpy(dst, beam_out + (blocks_sunk % kOutBufs) * out_block_floats, sizeof(float) * out_block_floats);
            // always the ring's full block size: a short block would end the data (PSRDADA semantics)
            ipcio_close_block_write(sink->data_block, ipcbuf_get_bufsz((ipcbuf_t *)sink->data_block));
            blocks_sunk++;
        }
#if !BEAM_DEBUG
        dada.retire(h);
#endif
    }
    if (have_sink) { dada_hdu_unlock_write(sink); dada_hdu_destroy(sink); multilog_close(sink_log); }   // raises end of data

    // ---- report (beamformer.cu:539-557) ----
    observation_time_ms += sw.stop_ms();
#if BEAM_DEBUG
    std::cout << "Observation ran in " << observation_time_ms << "milliseconds.\n";
    std::cout << "Code produced outputs for " << n_pt_sources * cfg.n_outputs_per_chunk << " data chunks.\n";
    std::cout << "Time per data chunk: " << observation_time_ms / (n_pt_sources * cfg.n_outputs_per_chunk) << " milliseconds.\n";
    std::cout << "Approximate datarate: " << (double)chunk_bytes * n_pt_sources / observation_time_ms / 1e6 << "GB/s" << std::endl;
#else
    if (verbose || getenv("DSABF_REPORT")) {
        std::cout << "Observation ran in " << observation_time_ms << "milliseconds.\n";
        std::cout << "Code produced outputs for " << obs.current_transfer_chunk() * cfg.n_outputs_per_chunk << " data chunks.\n";
        std::cout << "Time per data chunk: " << observation_time_ms / (obs.current_transfer_chunk() * cfg.n_outputs_per_chunk) << " milliseconds.\n";
        std::cout << "Approximate datarate: " << (double)dada.get_block_size() * obs.blocks_transfer_queue() / observation_time_ms / 1e6 << "GB/s" << std::endl;
        std::cout << "Blocks read from the ring after the burn-in: " << blocks_read << std::endl;
    }
#endif
    CHK(dsabf_sync(h));
#if !BEAM_DEBUG
    dada.drain(h);
#endif
    std::cout << "Synchronized" << std::endl;

#if BEAM_DEBUG
    write_python_array(dedispersed_out, n_pt_sources, cfg.n_beams, "bin/data.py");
#endif
    std::cout << "Freeing CUDA Structures" << std::endl;
    const unsigned long long launches = dsabf_kernel_launches(h);
    CHK(dsabf_destroy(h));
    std::cout << "Freed cuda streams and handles" << std::endl;
    std::cout << "Freed GPU memory" << std::endl;
    CUCHK(cudaFreeHost(beam_out));
#if BEAM_DEBUG
    CUCHK(cudaFreeHost(dedispersed_out));
#endif
    std::cout << "Freed CPU memory" << std::endl;
    if (verbose) std::cout << "kernels launched: " << launches << std::endl;
    return 0;
}
