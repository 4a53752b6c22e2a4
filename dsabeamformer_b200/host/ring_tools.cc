// ring_tools.cc -- `ringdb` and `junkdb`: create / destroy the shared-memory ring and feed it,
// standing in for PSRDADA's dada_db and dada_junkdb as the reference's makefile uses them
// (/root/reference/makefile:28-33: `dada_db -k baab -n 25 -b 134217728 -l -p`,
// `dada_junkdb -c 0 -z -k baab -r 4050 -t 25 config/correlator_header_dsaX.txt`).
//   ringdb -k KEY [-n NBUFS] [-b BUFSZ] [-d]          create (or -d destroy) the ring
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

#include "dada_ring.hh"

using namespace beam;

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
    unsigned key = 0xdada;
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
        if (destroy) { DadaRing::destroy(key); printf("destroyed ring %x\n", key); return 0; }
        if (!DadaRing::create(key, nbufs, bufsz)) { fprintf(stderr, "ringdb: could not create ring %x\n", key); return 1; }
        printf("created ring %x: %llu blocks of %llu bytes\n", key, (unsigned long long)nbufs, (unsigned long long)bufsz);
        return 0;
    }
    DadaRing ring;
    if (!ring.connect(key)) { fprintf(stderr, "junkdb: could not connect to ring %x\n", key); return 1; }
    std::vector<char> header(4096, 0);
    if (optind < argc) { std::ifstream f(argv[optind]); f.read(header.data(), 4096); }
    ring.write_header(header.data(), header.size());
    // one block of junk, generated once; each written block is a rotation of it
    std::vector<char> junk(ring.bufsz());
    uint64_t seed = 0xD5A;
    if (constant) memset(junk.data(), 0x70, junk.size());
    else for (size_t i = 0; i + 8 <= junk.size(); i += 8) { uint64_t v = splitmix64(seed); memcpy(&junk[i], &v, 8); }
    const auto t0 = std::chrono::steady_clock::now();
    long blocks = 0;
    for (;;) {
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if ((max_blocks >= 0 && blocks >= max_blocks) || (max_blocks < 0 && el >= secs)) break;
        if (rate > 0) {
            const double due = (double)(blocks + 1) * ring.bufsz() / (rate * 1e6);
            if (due > el) std::this_thread::sleep_for(std::chrono::duration<double>(due - el));
        }
        char *b = ring.open_block_write();
        const size_t rot = (size_t)(blocks % 16) * 4096;
        memcpy(b, junk.data() + rot, junk.size() - rot);
        memcpy(b + junk.size() - rot, junk.data(), rot);
        ring.close_block_write(ring.bufsz());
        blocks++;
    }
    ring.mark_eod();
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("junkdb: wrote %ld blocks (%.1f MB) in %.2f s = %.1f MB/s\n", blocks, blocks * (double)ring.bufsz() / 1e6, el,
           blocks * (double)ring.bufsz() / 1e6 / el);
    return 0;
}
