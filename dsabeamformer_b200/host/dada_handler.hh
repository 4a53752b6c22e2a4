// dada_handler.hh -- block reader for the observation build: connect to the ring, pin its
// blocks for DMA, hand out full blocks.
//
// Mirrors the reference's dada_handler (/root/reference/src/dada_handler.hh:1-176): constructor
// connects + locks + cudaHostRegister()s every ring block and optionally binds the thread to a
// core (:25-60,127-155); read_headers() consumes the header block and prints the block size
// (:66-90); read() returns the next full block (:92-94); check_transfers_complete() reports a
// short block = end of observation (:100-116).  The transport underneath is DadaRing.
#pragma once
#include <deque>
#include <iostream>

#include <cuda_runtime.h>
#include <sched.h>

#include "../../include/dsabf.h"
#include "dada_ring.hh"

namespace beam {

class DadaHandler {
  public:
    DadaHandler(const char *name, int core, unsigned key, uint64_t expected_block_bytes, bool verbose)
        : expected_(expected_block_bytes), verbose_(verbose)
    {
        fprintf(stderr, "[%s] creating hdu\n", name);
        if (!ring_.connect(key)) { printf("Error: could not connect to dada buffer\n"); exit(-1); }
        for (uint64_t i = 0; i < ring_.nbufs(); i++) {
            if (cudaHostRegister(ring_.buffer(i), ring_.bufsz(), 0) != cudaSuccess) { printf("Error: could not pin dada buffer\n"); exit(-1); }
        }
        if (core >= 0) {
            printf("binding to core %d\n", core);
            cpu_set_t set; CPU_ZERO(&set); CPU_SET(core, &set);
            if (sched_setaffinity(0, sizeof set, &set) < 0) printf("failed to bind to core %d\n", core);
        }
        if (verbose_) fprintf(stderr, "[%s] Done setting up buffer\n", name);
    }
    ~DadaHandler()
    {
        for (uint64_t i = 0; i < ring_.nbufs(); i++) cudaHostUnregister(ring_.buffer(i));
        ring_.disconnect();
    }

    void read_headers()
    {
        uint64_t hs = 0;
        if (!ring_.read_header(&hs)) { fprintf(stderr, "main: could not read next header\n"); exit(-1); }
        block_size_ = ring_.bufsz();
        std::cout << "block size is: " << block_size_ << std::endl;
    }
    char *read() { return ring_.open_block_read(&bytes_read_, &block_id_); }
    void close() { ring_.close_block_read(); }
    // Release ring blocks only once the H2D copy that reads them has completed.
    void close_when_copied(dsabf_handle *h, uint64_t ticket)
    {
        pending_.push_back(ticket);
        while (!pending_.empty() && dsabf_query_transfer(h, pending_.front()) == 1) { pending_.pop_front(); ring_.close_block_read(); }
    }
    // retire ring blocks whose copies have completed (non-blocking)
    void retire(dsabf_handle *h)
    {
        while (!pending_.empty() && dsabf_query_transfer(h, pending_.front()) == 1) { pending_.pop_front(); ring_.close_block_read(); }
    }
    void drain(dsabf_handle *h)
    {
        while (!pending_.empty()) { if (dsabf_query_transfer(h, pending_.front()) == 1) { pending_.pop_front(); ring_.close_block_read(); } }
    }
    size_t in_flight() const { return pending_.size(); }
    uint64_t ring_blocks() const { return ring_.nbufs(); }

    bool check_transfers_complete()
    {
        if (bytes_read_ != expected_) std::cout << "ERROR: Async, Bytes Read: " << bytes_read_ << ", Should also be " << expected_ << std::endl;
        if (bytes_read_ < block_size_) {
            if (verbose_) std::cout << "bytes_read < block_size, ending transfers" << std::endl;
            return true;
        }
        return false;
    }
    uint64_t get_block_size() const { return block_size_; }
    uint64_t get_bytes_read() const { return bytes_read_; }

  private:
    DadaRing ring_;
    uint64_t expected_, block_size_ = 0, bytes_read_ = 0, block_id_ = 0;
    bool verbose_;
    std::deque<uint64_t> pending_;
};

}  // namespace beam
