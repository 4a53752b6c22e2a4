// observation_loop.hh -- the four-counter state machine that paces ingest against analysis.
//
// Same contract as the reference's observation_loop_state
// (/root/reference/src/observation_loop.hh:1-177): blocks move through
//   TQ (transfer queued) -> T (transferred) -> AQ (analysis queued) -> A (analysed),
// a transfer may be queued while TQ-A < max_total_sep and TQ-T < max_transfer_sep, and an
// analysis while AQ < T.  Differences: completion is polled through the C ABI's tickets
// (dsabf_query_transfer / dsabf_query_analysis) instead of raw cudaEvents that are destroyed
// and re-created after every use (observation_loop.hh:95-96,112-113), and an analysis ticket
// covers ALL streams of the block, not only the last one (beamformer.cu:525).
#pragma once
#include <cstdint>
#include <deque>
#include <iostream>

#include "../../include/dsabf.h"

namespace beam {

class ObservationLoopState {
  public:
    ObservationLoopState(dsabf_handle *h, uint64_t max_transfer_sep, uint64_t max_total_sep, int chunks_per_block, int ring_slots, bool debug)
        : h_(h), max_transfer_sep_(max_transfer_sep), max_total_sep_(max_total_sep), chunks_per_block_(chunks_per_block),
          ring_slots_(ring_slots), debug_(debug) {}

    // a transfer / an analysis has been enqueued; remember its ticket
    void transfer_enqueued(uint64_t ticket) { transfer_tickets_.push_back(ticket); blocks_transfer_queue_++; }
    void analysis_enqueued(uint64_t ticket) { analysis_tickets_.push_back(ticket); blocks_analysis_queue_++; }

    // retire completed work in order; stop at the first pending ticket (observation_loop.hh:83-119)
    void check_transfer_events(bool verbose)
    {
        while (!transfer_tickets_.empty() && dsabf_query_transfer(h_, transfer_tickets_.front()) == 1) {
            if (verbose) std::cout << "Block " << blocks_transferred_ << " transfered to GPU" << std::endl;
            transfer_tickets_.pop_front();
            blocks_transferred_++;
        }
    }
    void check_analysis_events(bool verbose)
    {
        while (!analysis_tickets_.empty() && dsabf_query_analysis(h_, analysis_tickets_.front()) == 1) {
            if (verbose) std::cout << "Block " << blocks_analyzed_ << " Analyzed" << std::endl;
            analysis_tickets_.pop_front();
            blocks_analyzed_++;
        }
    }

    uint64_t blocks_analyzed() const { return blocks_analyzed_; }
    uint64_t blocks_transferred() const { return blocks_transferred_; }
    uint64_t blocks_analysis_queue() const { return blocks_analysis_queue_; }
    uint64_t blocks_transfer_queue() const { return blocks_transfer_queue_; }
    int next_transfer_slot() const { return (int)(blocks_transfer_queue_ % ring_slots_); }
    int next_analysis_slot() const { return (int)(blocks_analysis_queue_ % ring_slots_); }
    // first chunk ("GEMM") index of the block about to be analysed (observation_loop.hh:121-124)
    uint64_t current_analysis_chunk(int chunk_in_block)
    {
        most_recent_chunk_ = blocks_analysis_queue_ * chunks_per_block_ + chunk_in_block;
        return most_recent_chunk_;
    }
    uint64_t current_transfer_chunk() const { return blocks_transfer_queue_ * chunks_per_block_; }

    void set_transfers_complete(bool v) { transfers_complete_ = v; }
    bool transfers_complete() const { return transfers_complete_; }
    void set_n_pt_sources(int n) { n_pt_sources_ = n; }

    bool ready_for_transfer() const
    {
        return (blocks_transfer_queue_ - blocks_analyzed_ < max_total_sep_) && (blocks_transfer_queue_ - blocks_transferred_ < max_transfer_sep_) &&
               !transfers_complete_;
    }
    bool ready_for_analysis() const { return blocks_analysis_queue_ < blocks_transferred_; }

    // DEBUG: stop queueing once every source has a chunk (observation_loop.hh:161-169)
    bool check_transfers_complete_debug()
    {
        if ((int64_t)(blocks_transfer_queue_ * chunks_per_block_) >= n_pt_sources_) { transfers_complete_ = true; return true; }
        return false;
    }
    bool observations_complete()
    {
        const bool drained = (blocks_analyzed_ == blocks_transfer_queue_) && transfers_complete_;
        const bool done = debug_ ? (drained && (int64_t)most_recent_chunk_ >= n_pt_sources_ - 1) : drained;
        if (done) std::cout << "obs Complete" << std::endl;
        return done;
    }

    friend std::ostream &operator<<(std::ostream &out, const ObservationLoopState &a)
    {
        return out << "A: " << a.blocks_analyzed_ << ", AQ: " << a.blocks_analysis_queue_ << ", T: " << a.blocks_transferred_
                   << ", TQ: " << a.blocks_transfer_queue_ << "\n"
                   << "current_gemm: " << a.most_recent_chunk_ << ", transfers_complete: " << a.transfers_complete_;
    }

  private:
    dsabf_handle *h_;
    uint64_t blocks_analyzed_ = 0, blocks_transferred_ = 0, blocks_analysis_queue_ = 0, blocks_transfer_queue_ = 0;
    uint64_t max_transfer_sep_, max_total_sep_;
    int chunks_per_block_, ring_slots_;
    bool debug_;
    bool transfers_complete_ = false;
    uint64_t most_recent_chunk_ = 0;
    int64_t n_pt_sources_ = 0;
    std::deque<uint64_t> transfer_tickets_, analysis_tickets_;
};

}  // namespace beam
