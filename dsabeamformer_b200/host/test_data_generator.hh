// test_data_generator.hh -- DEBUG-mode synthetic input: one point source per chunk.
//
// Same bytes as the reference's test_data_generator (/root/reference/src/test_data_generator.hh:
// 34-108): a pinned buffer of kSourcesPerBatch chunks, pre-filled with 0x70 ("bogus" data: re 7,
// im 0); with a source catalogue, batch k holds sources [1024k, 1024k+1024), every column of a
// chunk identical, samples round(7 cos) / round(7 sin) of 2 pi (x sin(theta) + y sin(phi)) / lambda
// packed as (re << 4) | (im & 0xF), zero bytes past the last source.  Arithmetic types are the
// reference's: channel frequency from the *double* bandwidth macro and the integer sub-band offset
// gpu*2048/7 (:72), wavelength in float (:74), geometry in float, phase / cos / sin / round in double.
// Unlike the reference the per-chunk row is computed once and replicated over the columns, and
// OpenMP is always on (`make fast_debug` semantics, README.md:195).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "beam_config.hh"

namespace beam {

class TestDataGenerator {
  public:
    TestDataGenerator(const dsabf_config &cfg) : cfg_(cfg), chunk_bytes_(dsabf_chunk_bytes(&cfg)), bytes_(chunk_bytes_ * kSourcesPerBatch)
    {
        if (cudaHostAlloc((void **)&data_, bytes_, 0) != cudaSuccess) { fprintf(stderr, "test_data_generator: cudaHostAlloc failed\n"); exit(1); }
        memset(data_, kBogusData, bytes_);
    }
    ~TestDataGenerator() { cudaFreeHost(data_); }

    int n_pt_sources() const { return n_pt_sources_; }
    char *data() const { return data_; }
    size_t bytes() const { return bytes_; }

    void read_in_source_directions(const char *file_name, bool verbose)
    {
        if (use_catalog_) return;
        const int n = read_sources(file_name, sources_);
        if (n < 0) { std::cout << "Could not open " << file_name << std::endl; return; }
        n_pt_sources_ = n;
        use_catalog_ = true;
        if (verbose) std::cout << "Read in " << n_pt_sources_ << " source directions" << std::endl;
    }

    void generate_test_data(const std::vector<Antenna> &pos, int gpu)
    {
        const int n_ant = cfg_.n_antennas, n_chan = cfg_.n_channels, cols = dsabf_columns_per_chunk(&cfg_);
        const double bw = (1.53 - 1.28) / 2048;
#pragma omp parallel for schedule(dynamic, 4)
        for (long c = 0; c < kSourcesPerBatch; c++) {
            const long s = c + (long)batch_counter_ * kSourcesPerBatch;
            char *d = data_ + chunk_bytes_ * c;
            if (s >= n_pt_sources_) { memset(d, 0, chunk_bytes_); continue; }
            const float sth = sinf(sources_[s].theta), sph = sinf(sources_[s].phi);
            for (int i = 0; i < n_chan; i++) {
                const float freq = (float)(1.53 - (0 + gpu * 2048 / (8 - 1) + i) * bw);
                const float wavelength = (float)(299792458.0 / (1E9 * (double)freq));
                char *row0 = d + (size_t)i * cols * n_ant;
                for (int k = 0; k < n_ant; k++) {
                    const float gx = pos[k].x * sth, gy = pos[k].y * sph;
                    const float g = gx + gy;
                    const double ph = 2 * kPi * (double)g / (double)wavelength;
                    const signed char high = (signed char)round(7 * cos(ph));
                    const signed char low = (signed char)round(7 * sin(ph));
                    row0[k] = (char)((high << 4) | (0x0F & low));
                }
                for (int j = 1; j < cols; j++) memcpy(row0 + (size_t)j * n_ant, row0, n_ant);
            }
        }
        batch_counter_++;
    }

    // test_data_generator.hh:98-108
    bool need_to_generate_more(uint64_t blocks_transferred) const
    {
        return use_catalog_ && blocks_transferred == (uint64_t)batch_counter_ * kSourcesPerBatch / cfg_.n_chunks_per_block;
    }
    bool data_ready_for_transfer(uint64_t blocks_transfer_queue)
    {
        if (!use_catalog_ && batch_counter_ == 0) batch_counter_ = 1;
        return blocks_transfer_queue < (uint64_t)batch_counter_ * kSourcesPerBatch / cfg_.n_chunks_per_block;
    }

  private:
    dsabf_config cfg_;
    size_t chunk_bytes_, bytes_;
    char *data_ = nullptr;
    int n_pt_sources_ = 1024;
    int batch_counter_ = 0;
    bool use_catalog_ = false;
    std::vector<Direction> sources_;
};

}  // namespace beam
