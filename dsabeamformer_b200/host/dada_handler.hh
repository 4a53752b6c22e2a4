// dada_handler.hh -- block reader for the observation build: connect to the PSRDADA ring, pin its
// blocks for DMA, hand out full blocks.
//
// Mirrors the reference's dada_handler (/root/reference/src/dada_handler.hh:1-176) call for call, on the
// PSRDADA C API itself (dada_hdu_* / ipcbuf_* / ipcio_* / multilog / dada_bind_thread_to_core): the
// constructor connects, locks for reading, cudaHostRegister()s every ring block and optionally binds the
// thread to a core (:25-60,127-155); read_headers() consumes the header block and prints the block size
// (:66-90); read() opens the next full block (:92-94); check_transfers_complete() reports a short block =
// end of observation (:100-116).  The prototypes come from $(PSRDADA_INC): host/psrdada_compat/ + the shim
// library where PSRDADA is not installed, the real headers + -lpsrdada where it is -- no source change.
//
// One difference by design: PSRDADA allows ONE open block per reader, and the reference closes it right
// after *queueing* the asynchronous H2D copy (beamformer.cu:389-401), i.e. hands the block back to the
// writer while the DMA may still be reading it.  Here the block stays open until the copy's ticket has
// completed (close_when_copied / retire), and the next block is opened only after that.
#pragma once
#include <iostream>

#include <cuda_runtime.h>

#include "dada_hdu.h"
#include "ipcio.h"
#include "ipcbuf.h"
#include "multilog.h"
#include "dada_affinity.h"

#include "../../include/dsabf.h"

namespace beam {

class DadaHandler {
  public:
    DadaHandler(const char *name, int core, key_t in_key, uint64_t expected_block_bytes, bool verbose)
        : expected_(expected_block_bytes), verbose_(verbose)
    {
        log_ = multilog_open(name, 0);
        multilog_add(log_, stderr);
        multilog(log_, LOG_INFO, "creating hdu\n");
        hdu_in_ = dada_hdu_create(log_);
        dada_hdu_set_key(hdu_in_, in_key);
        if (dada_hdu_connect(hdu_in_) < 0) { printf("Error: could not connect to dada buffer\n"); exit(-1); }
        if (dada_hdu_lock_read(hdu_in_) < 0) {
            printf("Error: could not lock to dada buffer (try relaxing memlock limits in /etc/security/limits.conf)\n");
            exit(-1);
        }
        if (dbregister() < 0) { printf("Error: could not pin dada buffer\n"); exit(-1); }
        if (core >= 0) {
            printf("binding to core %d\n", core);
            if (dada_bind_thread_to_core(core) < 0) printf("failed to bind to core %d\n", core);
        }
        if (verbose_) multilog(log_, LOG_INFO, "Done setting up buffer\n");
    }
    ~DadaHandler()
    {
        dbunregister();
        if (dada_hdu_unlock_read(hdu_in_) < 0) multilog(log_, LOG_ERR, "could not unlock read on hdu_in\n");
        dada_hdu_destroy(hdu_in_);
        multilog_close(log_);
    }

    void read_headers()
    {
        uint64_t header_size = 0;
        char *header_in = ipcbuf_get_next_read(hdu_in_->header_block, &header_size);
        if (!header_in) { multilog(log_, LOG_ERR, "main: could not read next header\n"); exit(-1); }
        if (ipcbuf_mark_cleared(hdu_in_->header_block) < 0) { multilog(log_, LOG_ERR, "could not mark header block cleared\n"); exit(-1); }
        block_size_ = ipcbuf_get_bufsz((ipcbuf_t *)hdu_in_->data_block);
        if (verbose_) multilog(log_, LOG_INFO, "Done setting up header \n");
        std::cout << "block size is: " << block_size_ << std::endl;
    }

    // Next full block, or nullptr once the writer has ended the data (then bytes_read = 0).  Requires that the
    // previous block has been closed: close() / retire() / drain().
    char *read()
    {
        bytes_read_ = 0;
        char *b = ipcio_open_block_read(hdu_in_->data_block, &bytes_read_, &block_id_);
        if (!b) bytes_read_ = 0;
        open_ = b != nullptr;
        return b;
    }
    void close()
    {
        if (open_) ipcio_close_block_read(hdu_in_->data_block, bytes_read_);
        open_ = false;
        pending_ = false;
    }
    // Keep the open block until the H2D copy that reads it (ticket) has completed.
    void close_when_copied(dsabf_handle *h, uint64_t ticket)
    {
        pending_ = true;
        ticket_ = ticket;
        retire(h);
    }
    // close the block if its copy has completed (non-blocking)
    void retire(dsabf_handle *h)
    {
        if (pending_ && dsabf_query_transfer(h, ticket_) == 1) close();
    }
    void drain(dsabf_handle *h)
    {
        while (pending_) {
            const int q = dsabf_query_transfer(h, ticket_);
            if (q != 0) close();            // done, or a CUDA error (reported by the next API call): either way let go
        }
    }
    size_t in_flight() const { return pending_ ? 1 : 0; }
    uint64_t ring_blocks() const { return ipcbuf_get_nbufs((ipcbuf_t *)hdu_in_->data_block); }

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
    // dada_cuda_dbregister / dbunregister of the reference (dada_handler.hh:127-176)
    int dbregister()
    {
        ipcbuf_t *db = (ipcbuf_t *)hdu_in_->data_block;
        if (ipcbuf_lock(db) < 0) { perror("dada_dbregister: ipcbuf_lock failed\n"); return -1; }
        if (ipcbuf_get_device(db) >= 0) return 0;      // buffers that already live on a device are not registered
        const size_t bufsz = ipcbuf_get_bufsz(db);
        for (uint64_t i = 0; i < ipcbuf_get_nbufs(db); i++)
            if (cudaHostRegister((void *)db->buffer[i], bufsz, 0) != cudaSuccess) { perror("dada_dbregister:  cudaHostRegister failed\n"); return -1; }
        return 0;
    }
    void dbunregister()
    {
        ipcbuf_t *db = (ipcbuf_t *)hdu_in_->data_block;
        if (ipcbuf_get_device(db) >= 0) return;
        for (uint64_t i = 0; i < ipcbuf_get_nbufs(db); i++) cudaHostUnregister((void *)db->buffer[i]);
    }

    multilog_t *log_ = nullptr;
    dada_hdu_t *hdu_in_ = nullptr;
    uint64_t expected_, block_size_ = 0, bytes_read_ = 0, block_id_ = 0, ticket_ = 0;
    bool verbose_, open_ = false, pending_ = false;
};

}  // namespace beam
