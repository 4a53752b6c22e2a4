// dada_ring.hh -- a small shared-memory ring of fixed-size blocks: the stand-in for the
// PSRDADA data block when PSRDADA itself is not installed.
//
// THIS IS A TEST DOUBLE for the un-vendored PSRDADA library the reference links against
// (/root/reference/makefile:5-6,13; used through dada_hdu / ipcio / ipcbuf in
// src/dada_handler.hh).  It provides just what the `beam` reader needs -- a header block, a ring
// of `nbufs` data blocks of `bufsz` bytes in POSIX shared memory keyed like a dada key,
// open/close of the next full block for reading, end-of-data -- plus the writer side used by
// tools `ringdb` (cf. `dada_db -k KEY -n N -b BYTES`) and `junkdb` (cf. `dada_junkdb -r RATE -t SEC`).
// Swap DadaHandler's three calls for dada_hdu_* / ipcio_* to run against the real PSRDADA.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace beam {

struct RingControl {
    uint64_t magic;
    uint64_t nbufs, bufsz, header_size;
    std::atomic<uint64_t> written;      // blocks completed by the writer
    std::atomic<uint64_t> released;     // blocks given back by the reader
    std::atomic<uint32_t> eod;          // writer is finished
    std::atomic<uint32_t> header_valid;
    uint64_t bytes[256];                // payload bytes of each ring block
    char header[4096];
};

class DadaRing {
  public:
    static constexpr uint64_t kMagic = 0x44534142465f5231ull;   // "DSABF_R1"
    static std::string shm_name(unsigned key) { char b[64]; snprintf(b, sizeof b, "/dsabf_dada_%08x", key); return b; }

    static bool create(unsigned key, uint64_t nbufs, uint64_t bufsz)
    {
        if (nbufs < 1 || nbufs > 256) return false;
        const std::string name = shm_name(key);
        shm_unlink(name.c_str());
        int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) { perror("shm_open"); return false; }
        const size_t total = data_offset() + nbufs * bufsz;
        if (ftruncate(fd, (off_t)total) != 0) { perror("ftruncate"); close(fd); return false; }
        void *p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) { perror("mmap"); return false; }
        RingControl *c = new (p) RingControl();
        c->nbufs = nbufs; c->bufsz = bufsz; c->header_size = sizeof(c->header);
        c->written = 0; c->released = 0; c->eod = 0; c->header_valid = 0;
        c->magic = kMagic;
        munmap(p, total);
        return true;
    }
    static void destroy(unsigned key) { shm_unlink(shm_name(key).c_str()); }

    bool connect(unsigned key)
    {
        int fd = shm_open(shm_name(key).c_str(), O_RDWR, 0600);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); return false; }
        total_ = (size_t)st.st_size;
        void *p = mmap(nullptr, total_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (p == MAP_FAILED) return false;
        ctl_ = (RingControl *)p;
        if (ctl_->magic != kMagic) { munmap(p, total_); ctl_ = nullptr; return false; }
        return true;
    }
    void disconnect() { if (ctl_) munmap(ctl_, total_); ctl_ = nullptr; }

    uint64_t nbufs() const { return ctl_->nbufs; }
    uint64_t bufsz() const { return ctl_->bufsz; }
    char *buffer(uint64_t i) const { return (char *)ctl_ + data_offset() + i * ctl_->bufsz; }

    // ---- writer ----
    void write_header(const char *h, size_t n) { memcpy(ctl_->header, h, n < sizeof(ctl_->header) ? n : sizeof(ctl_->header)); ctl_->header_valid = 1; }
    char *open_block_write()
    {
        while (ctl_->written.load() - ctl_->released.load() >= ctl_->nbufs) usleep(50);
        return buffer(ctl_->written.load() % ctl_->nbufs);
    }
    void close_block_write(uint64_t bytes) { ctl_->bytes[ctl_->written.load() % ctl_->nbufs] = bytes; ctl_->written.fetch_add(1); }
    void mark_eod() { ctl_->eod = 1; }

    // ---- reader ----
    const char *read_header(uint64_t *size)
    {
        while (!ctl_->header_valid.load()) usleep(100);
        *size = ctl_->header_size;
        return ctl_->header;
    }
    // blocks until a full block is available; at end of data returns a valid pointer with *bytes = 0
    char *open_block_read(uint64_t *bytes, uint64_t *block_id)
    {
        for (;;) {
            const uint64_t r = open_pos_;
            if (r < ctl_->written.load()) { *bytes = ctl_->bytes[r % ctl_->nbufs]; *block_id = r; open_pos_++; return buffer(r % ctl_->nbufs); }
            if (ctl_->eod.load()) { *bytes = 0; *block_id = r; return buffer(r % ctl_->nbufs); }
            usleep(20);
        }
    }
    // PSRDADA semantics: the (oldest open) block is returned to the writer at close.  The caller
    // must not close before its asynchronous copy out of the block has finished (the reference
    // does, beamformer.cu:389-401; DadaHandler::close_when_copied fixes that).
    void close_block_read() { if (rel_pos_ < open_pos_) { rel_pos_++; ctl_->released.store(rel_pos_); } }

  private:
    static size_t data_offset() { return (sizeof(RingControl) + 4095) & ~(size_t)4095; }
    RingControl *ctl_ = nullptr;
    size_t total_ = 0;
    uint64_t open_pos_ = 0, rel_pos_ = 0;
};

}  // namespace beam
