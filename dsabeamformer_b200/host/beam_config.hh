// beam_config.hh -- host-side constants, array/beam geometry types, config-file readers and
// the data.py writer of the `beam` program.
//
// Mirrors, behaviour for behaviour, the non-CUDA header of the reference
// (/root/reference/src/beamformer.hh): the shape constants (:47-60,108-147), the antenna /
// beam_direction records and their stream extractors (:170-213), usage() (:222-243), the
// readers (:250-284) and write_array_to_disk_as_python_file (:287-311).  Shapes are run-time
// here (they come from dsabf_config), the defaults are the reference's.
#pragma once
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/dsabf.h"

namespace beam {

constexpr int kBurnIn = 25;               // BURNIN, beamformer.hh:45
constexpr double kHalfFovDeg = 3.5;       // HALF_FOV
constexpr double kPi = 3.14159265358979;  // PI as the reference spells it (beamformer.hh:72)
constexpr int kSourcesPerBatch = 1024;    // N_SOURCES_PER_BATCH
constexpr int kMaxTransferSep = 2;        // MAX_TRANSFER_SEP
constexpr int kMaxTotalSep = 4;           // MAX_TOTAL_SEP
constexpr unsigned char kBogusData = 0x70;  // test_data_generator.hh:8

inline double deg2rad(double x) { return x * kPi / 180.0; }

struct Antenna { float x = 0, y = 0, z = 0; };        // metres; z is read but never used
struct Direction { float theta = 0, phi = 0; };       // radians

inline std::istream &operator>>(std::istream &in, Antenna &a) { return in >> a.x >> a.y >> a.z; }
inline std::istream &operator>>(std::istream &in, Direction &d) { return in >> d.theta >> d.phi; }

// "count, then one record per line"; a count mismatch only warns, missing records stay zero
// (beamformer.hh:250-284).  Returns the count found in the file, or -1 if it cannot be opened.
template <class T>
int read_table(const char *file_name, int expected, std::vector<T> &out, const char *what, const char *expected_name)
{
    std::ifstream f(file_name);
    out.assign(expected, T());
    if (!f) {
        std::cout << "Could not open " << file_name << std::endl;
        return -1;
    }
    int n = 0;
    f >> n;
    if (n != expected) {
        std::cout << "Number of " << what << " in file (" << n << ") does not match " << expected_name << " (" << expected << ")" << std::endl;
        std::cout << "Excess " << what << " will be ignored, missing " << what << " will be set to 0." << std::endl;
    }
    for (int i = 0; i < expected; i++) {
        T rec;
        if (f >> rec) out[i] = rec;
        else break;
    }
    return n;
}

// A source catalogue keeps every record in the file (test_data_generator.hh:43-60).
inline int read_sources(const char *file_name, std::vector<Direction> &out)
{
    std::ifstream f(file_name);
    out.clear();
    if (!f) return -1;
    int n = 0;
    f >> n;
    out.assign(n > 0 ? n : 0, Direction());
    for (int i = 0; i < n; i++) {
        Direction d;
        if (f >> d) out[i] = d;
        else break;
    }
    return n;
}

inline void usage(bool debug)
{
    if (debug)
        fprintf(stdout,
                "dsaX_beamformer_DEBUG_MODE [options]\n"
                " -g gpu                  select a predefined frequency range\n"
                " -p position_filename    file where the antenna positions are stored\n"
                " -d direction_filename   file where the beam directions are stored\n"
                " -s source_filename      file where the source directions are stored\n"
                " -h                      print usage\n");
    else
        fprintf(stdout,
                "dsaX_beamformer [options]\n"
                " -c core                 bind process to CPU core\n"
                " -k key                  [default dada]\n"
                " -g gpu                  select a predefined frequency range\n"
                " -p position_filename    file where the antenna positions are stored\n"
                " -d direction_filename   file where the beam directions are stored\n"
                " -h                      print usage\n");
}

// bin/data.py: `A = [[v,v,...],\n[...]]`, default ostream float formatting, importable from python
// (beamformer.hh:287-311; util/view_data.py reads data.A).
inline void write_python_array(const float *data, int rows, int cols, const char *file_name)
{
    std::ofstream f(file_name);
    f << "A = [[";
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            f << data[(size_t)r * cols + c];
            if (c != cols - 1) f << ",";
        }
        if (r != rows - 1) f << "],\n[";
        else f << "]]" << std::endl;
    }
}

inline void print_config(const dsabf_config &c, bool debug)
{
    const int n_in_per_out = c.n_pol * c.n_averaging;
    std::cout << "N_BEAMS: " << c.n_beams << "\nN_ANTENNAS: " << c.n_antennas << "\nN_FREQUENCIES: " << c.n_channels
              << "\nN_AVERAGING: " << c.n_averaging << "\nN_POL: " << c.n_pol << "\nN_OUTPUTS_PER_GEMM: " << c.n_outputs_per_chunk
              << "\nN_GEMMS_PER_BLOCK: " << c.n_chunks_per_block << "\nN_INPUTS_PER_OUTPUT: " << n_in_per_out
              << "\nN_TIMESTEPS_PER_GEMM: " << dsabf_columns_per_chunk(&c) << "\nN_BLOCKS_ON_GPU: " << c.n_ring_slots
              << "\nN_BYTES_PRE_EXPANSION_PER_GEMM: " << dsabf_chunk_bytes(&c) << "\nN_BYTES_PRE_EXPANSION_PER_BLOCK: " << dsabf_block_bytes(&c)
              << "\nN_GPUS: 8\nTOT_CHANNELS: 2048\nSTART_F: 1.28\nEND_F: 1.53\n";
    if (debug) std::cout << "N_SOURCES_PER_BATCH: " << kSourcesPerBatch << "\n";
    std::cout << std::endl;
}

}  // namespace beam
