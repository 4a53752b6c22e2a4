// ring_tools.cc -- `ringdb` and `junkdb`: create / destroy a PSRDADA ring and feed it, standing in for
// PSRDADA's dada_db and dada_junkdb as the reference's makefile uses them
// (/root/reference/makefile:28-33: `dada_db -k baab -n 25 -b 134217728 -l -p`,
// `dada_junkdb -c 0 -z -k baab -r 4050 -t 25 config/correlator_header_dsaX.txt`).
// Written against the PSRDADA C API (ipcbuf_create / dada_hdu_* / ipcio_open_block_write ...): links with the
// shim library here, with -lpsrdada where PSRDADA is installed.
//   ringdb -k KEY [-n NBUFS] [-b BUFSZ] [-d]          create (or -d destroy) header ring (KEY+1) and data ring (KEY)
//   junkdb -k KEY [-r MB_PER_S] [-t SECONDS] [-n BLOCKS] [-g] [header_file]
//          writes blocks of fixed-seed pseudo-random bytes (splitmix64, seed 0xD5A; every byte value
//          is legal input) at the requested rate; -r 0 = as fast as possible; -n caps the block count;
//          -g writes generator-style constant blocks (0x70) instead.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <thread>
#include <vector>
#include <unistd.h>

#include "dada_def.h"
#include "dada_hdu.h"
#include "ipcbuf.h"
#include "ipcio.h"
#include "multilog.h"

static uint64_t splitmix64(uint64_t &s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv)
{
    const bool is_junk = strstr(argv[0], "junkdb") != nullptr;
    unsigned key = DADA_DEFAULT_BLOCK_KEY;
    uint64_t nbufs = 25, bufsz = 134217728;
    double rate = 4050, secs = 25;
    long max_blocks = -1;
    bool destroy = false, constant = false;
    int arg;
    while ((arg = getopt(argc, argv, "k:n:b:dr:t:c:zglp")) != -1) {
        switch (arg) {
        case 'k': sscanf(optarg, "%x", &key); break;
        case 'n': if (is_junk) max_blocks = atol(optarg); else nbufs = strtoull(optarg, nullptr, 10); break;
        case 'b': bufsz = strtoull(optarg, nullptr, 10); break;
        case 'd': destroy = true; break;
        case 'r': rate = atof(optarg); break;
        case 't': secs = atof(optarg); break;
        case 'g': constant = true; break;
        default: break;   // -c core, -z, -l, -p accepted and ignored
        }
    }
    if (!is_junk) {
        ipcbuf_t header = IPCBUF_INIT;
        ipcio_t data = IPCIO_INIT;
        if (destroy) {
            int rc = 0;
            if (ipcio_connect(&data, (key_t)key) == 0) ipcio_destroy(&data); else rc = 1;
            if (ipcbuf_connect(&header, (key_t)(key + 1)) == 0) ipcbuf_destroy(&header); else rc = 1;
            printf("destroyed ring %x\n", key);
            return rc;
        }
        if (ipcbuf_create(&header, (key_t)(key + 1), 8, DADA_DEFAULT_HEADER_SIZE, 1) < 0 || ipcio_create(&data, (key_t)key, nbufs, bufsz, 1) < 0) {
            fprintf(stderr, "ringdb: could not create ring %x\n", key);
            return 1;
        }
        printf("created ring %x: %llu blocks of %llu bytes\n", key, (unsigned long long)nbufs, (unsigned long long)bufsz);
        ipcbuf_disconnect(&header);
        ipcio_disconnect(&data);
        return 0;
    }
    multilog_t *log = multilog_open("junkdb", 0);
    multilog_add(log, stderr);
    dada_hdu_t *hdu = dada_hdu_create(log);
    dada_hdu_set_key(hdu, (key_t)key);
    if (dada_hdu_connect(hdu) < 0 || dada_hdu_lock_write(hdu) < 0) { fprintf(stderr, "junkdb: could not connect to ring %x\n", key); return 1; }
    {
        const uint64_t hsz = ipcbuf_get_bufsz(hdu->header_block);
        char *hb = ipcbuf_get_next_write(hdu->header_block);
        memset(hb, 0, hsz);
        if (optind < argc) { std::ifstream f(argv[optind]); f.read(hb, (std::streamsize)hsz - 1); }
        ipcbuf_mark_filled(hdu->header_block, hsz);
    }
    const uint64_t block = ipcbuf_get_bufsz((ipcbuf_t *)hdu->data_block);
    // one block of junk, generated once; each written block is a rotation of it
    std::vector<char> junk(block);
    uint64_t seed = 0xD5A;
    if (constant) memset(junk.data(), 0x70, junk.size());
    else for (size_t i = 0; i + 8 <= junk.size(); i += 8) { uint64_t v = splitmix64(seed); memcpy(&junk[i], &v, 8); }
    const auto t0 = std::chrono::steady_clock::now();
    long blocks = 0;
    for (;;) {
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if ((max_blocks >= 0 && blocks >= max_blocks) || (max_blocks < 0 && el >= secs)) break;
        if (rate > 0) {
            const double due = (double)(blocks + 1) * block / (rate * 1e6);
            if (due > el) std::this_thread::sleep_for(std::chrono::duration<double>(due - el));
        }
        uint64_t id = 0;
        char *b = ipcio_open_block_write(hdu->data_block, &id);
        if (!b) { fprintf(stderr, "junkdb: could not open a block for writing\n"); return 1; }
        const size_t rot = (size_t)(blocks % 16) * 4096;
        memcpy(b, junk.data() + rot, junk.size() - rot);
        memcpy(b + junk.size() - rot, junk.data(), rot);
        ipcio_close_block_write(hdu->data_block, block);
        blocks++;
    }
    dada_hdu_unlock_write(hdu);          // closes the data block for writing: raises end of data
    dada_hdu_destroy(hdu);
    multilog_close(log);
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("junkdb: wrote %ld blocks (%.1f MB) in %.2f s = %.1f MB/s\n", blocks, blocks * (double)block / 1e6, el,
           blocks * (double)block / 1e6 / el);
    return 0;
}
