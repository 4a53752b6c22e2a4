// psrdada_shim.cc -- libpsrdada_shim: the PSRDADA calls declared in psrdada_compat/*.h over POSIX shared memory.
//
// TEST DOUBLE for the un-vendored PSRDADA library the reference links with -lpsrdada
// (/root/reference/makefile:5-6,13).  Symbols, argument order and the struct members the reference touches are
// PSRDADA's, so host code written against these headers (dada_handler.hh, ring_tools.cc, beam's output sink) and
// the reference's own dada_handler.hh link against either this file or the real library without source changes.
// Scope: one writer and one reader per ring, polling instead of SysV semaphores, end of data = a short (possibly
// empty) last buffer.  One ring = one shm object "/dsabf_dada_<key>": a page of ipcsync_t, then nbufs * bufsz bytes.
#include "psrdada_compat/ascii_header.h"
#include "psrdada_compat/dada_affinity.h"
#include "psrdada_compat/dada_def.h"
#include "psrdada_compat/dada_hdu.h"

#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

constexpr uint64_t kMagic = 0x44534142465f5232ull;   // "DSABF_R2"
constexpr size_t kCtlBytes = (sizeof(ipcsync_t) + 4095) & ~(size_t)4095;

std::string shm_name(key_t key)
{
    char b[64];
    snprintf(b, sizeof b, "/dsabf_dada_%08x", (unsigned)key);
    return b;
}
uint64_t ld(const volatile uint64_t *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
void st(volatile uint64_t *p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

int map_ring(ipcbuf_t *id, int fd, size_t total)
{
    void *p = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) { perror("ipcbuf: mmap"); return -1; }
    id->map_base = p;
    id->map_bytes = total;
    id->sync = (ipcsync_t *)p;
    return 0;
}
int index_buffers(ipcbuf_t *id)
{
    const uint64_t n = id->sync->nbufs;
    id->buffer = (char **)malloc(sizeof(char *) * n);
    if (!id->buffer) return -1;
    for (uint64_t i = 0; i < n; i++) id->buffer[i] = (char *)id->map_base + kCtlBytes + i * id->sync->bufsz;
    return 0;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------ ipcbuf --
int ipcbuf_create(ipcbuf_t *id, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned n_readers)
{
    if (!id || nbufs < 1 || nbufs > IPCBUF_MAX_BUFS || bufsz < 1 || n_readers != 1) {
        fprintf(stderr, "ipcbuf_create: invalid argument (1..%d buffers, exactly 1 reader)\n", IPCBUF_MAX_BUFS);
        return -1;
    }
    const std::string name = shm_name(key);
    shm_unlink(name.c_str());
    int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) { perror("ipcbuf_create: shm_open"); return -1; }
    const size_t total = kCtlBytes + (size_t)nbufs * bufsz;
    if (ftruncate(fd, (off_t)total) != 0) { perror("ipcbuf_create: ftruncate"); close(fd); shm_unlink(name.c_str()); return -1; }
    const int rc = map_ring(id, fd, total);
    close(fd);
    if (rc) return -1;
    memset(id->sync, 0, sizeof(ipcsync_t));
    id->sync->nbufs = nbufs;
    id->sync->bufsz = bufsz;
    st(&id->sync->magic, kMagic);
    id->key = key;
    id->created = 1;
    id->state = 1;
    return index_buffers(id);
}

int ipcbuf_connect(ipcbuf_t *id, key_t key)
{
    if (!id) return -1;
    int fd = shm_open(shm_name(key).c_str(), O_RDWR, 0600);
    if (fd < 0) { fprintf(stderr, "ipcbuf_connect: no ring with key %x\n", (unsigned)key); return -1; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < kCtlBytes) { close(fd); return -1; }
    const int rc = map_ring(id, fd, (size_t)sb.st_size);
    close(fd);
    if (rc) return -1;
    if (ld(&id->sync->magic) != kMagic || kCtlBytes + id->sync->nbufs * id->sync->bufsz > id->map_bytes) {
        fprintf(stderr, "ipcbuf_connect: key %x is not a ring of this library\n", (unsigned)key);
        munmap(id->map_base, id->map_bytes);
        id->sync = nullptr;
        return -1;
    }
    id->key = key;
    id->created = 0;
    id->state = 1;
    return index_buffers(id);
}

int ipcbuf_disconnect(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    if (id->locked) munlock(id->map_base, id->map_bytes);
    munmap(id->map_base, id->map_bytes);
    free(id->buffer);
    *id = ipcbuf_t IPCBUF_INIT;
    return 0;
}

int ipcbuf_destroy(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    const key_t key = id->key;
    ipcbuf_disconnect(id);
    return shm_unlink(shm_name(key).c_str()) == 0 ? 0 : -1;
}

int ipcbuf_lock_write(ipcbuf_t *id)
{
    if (!id || !id->sync || id->state != 1) return -1;
    if (__atomic_fetch_add(&id->sync->n_writers, 1u, __ATOMIC_ACQ_REL) != 0) {
        __atomic_fetch_sub(&id->sync->n_writers, 1u, __ATOMIC_ACQ_REL);
        fprintf(stderr, "ipcbuf_lock_write: ring %x already has a writer\n", (unsigned)id->key);
        return -1;
    }
    id->state = 3;
    return 0;
}
int ipcbuf_unlock_write(ipcbuf_t *id)
{
    if (!id || !id->sync || id->state != 3) return -1;
    __atomic_fetch_sub(&id->sync->n_writers, 1u, __ATOMIC_ACQ_REL);
    id->state = 1;
    return 0;
}
int ipcbuf_lock_read(ipcbuf_t *id)
{
    if (!id || !id->sync || id->state != 1) return -1;
    if (__atomic_fetch_add(&id->sync->n_readers, 1u, __ATOMIC_ACQ_REL) != 0) {
        __atomic_fetch_sub(&id->sync->n_readers, 1u, __ATOMIC_ACQ_REL);
        fprintf(stderr, "ipcbuf_lock_read: ring %x already has a reader\n", (unsigned)id->key);
        return -1;
    }
    id->state = 2;
    return 0;
}
int ipcbuf_unlock_read(ipcbuf_t *id)
{
    if (!id || !id->sync || id->state != 2) return -1;
    __atomic_fetch_sub(&id->sync->n_readers, 1u, __ATOMIC_ACQ_REL);
    id->state = 1;
    return 0;
}

char *ipcbuf_get_next_write(ipcbuf_t *id)
{
    if (!id || !id->sync) return nullptr;
    ipcsync_t *s = id->sync;
    while (ld(&s->w_count) - ld(&s->r_count) >= s->nbufs) usleep(50);
    return id->buffer[ld(&s->w_count) % s->nbufs];
}
int ipcbuf_mark_filled(ipcbuf_t *id, uint64_t nbytes)
{
    if (!id || !id->sync || nbytes > id->sync->bufsz) return -1;
    ipcsync_t *s = id->sync;
    const uint64_t w = ld(&s->w_count);
    s->nbytes[w % s->nbufs] = nbytes;
    if (id->eod_pending || nbytes < s->bufsz) { st(&s->eod, w + 1); id->eod_pending = 0; }   // a short buffer ends the data
    st(&s->w_count, w + 1);
    return 0;
}
int ipcbuf_enable_eod(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    id->eod_pending = 1;
    return 0;
}

int ipcbuf_eod(ipcbuf_t *id)
{
    if (!id || !id->sync) return 1;
    const uint64_t eod = ld(&id->sync->eod);
    return eod != 0 && ld(&id->sync->r_count) >= eod;
}
char *ipcbuf_get_next_read(ipcbuf_t *id, uint64_t *bytes)
{
    if (!id || !id->sync) return nullptr;
    ipcsync_t *s = id->sync;
    for (;;) {
        const uint64_t r = ld(&s->r_count);
        if (r < ld(&s->w_count)) {
            if (bytes) *bytes = s->nbytes[r % s->nbufs];
            id->last_read_index = r;
            return id->buffer[r % s->nbufs];
        }
        if (ipcbuf_eod(id)) return nullptr;
        usleep(20);
    }
}
int ipcbuf_mark_cleared(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    ipcsync_t *s = id->sync;
    const uint64_t r = ld(&s->r_count);
    if (r >= ld(&s->w_count)) { fprintf(stderr, "ipcbuf_mark_cleared: no buffer is open for reading\n"); return -1; }
    st(&s->r_count, r + 1);
    return 0;
}

uint64_t ipcbuf_get_nbufs(ipcbuf_t *id) { return id && id->sync ? id->sync->nbufs : 0; }
uint64_t ipcbuf_get_bufsz(ipcbuf_t *id) { return id && id->sync ? id->sync->bufsz : 0; }
uint64_t ipcbuf_get_write_count(ipcbuf_t *id) { return id && id->sync ? ld(&id->sync->w_count) : 0; }
uint64_t ipcbuf_get_read_count(ipcbuf_t *id) { return id && id->sync ? ld(&id->sync->r_count) : 0; }
uint64_t ipcbuf_get_read_index(ipcbuf_t *id) { return id && id->sync ? id->last_read_index : 0; }
uint64_t ipcbuf_get_nfull(ipcbuf_t *id) { return id && id->sync ? ld(&id->sync->w_count) - ld(&id->sync->r_count) : 0; }

int ipcbuf_lock(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    if (mlock(id->map_base, id->map_bytes) != 0) {
        // RLIMIT_MEMLOCK in containers: cudaHostRegister pins the pages anyway, so report and carry on like dada_db without -l
        fprintf(stderr, "ipcbuf_lock: mlock failed (%s); buffers stay pageable until registered\n", strerror(errno));
        return 0;
    }
    id->locked = 1;
    return 0;
}
int ipcbuf_unlock(ipcbuf_t *id)
{
    if (!id || !id->sync) return -1;
    if (id->locked) munlock(id->map_base, id->map_bytes);
    id->locked = 0;
    return 0;
}
int ipcbuf_get_device(ipcbuf_t *) { return -1; }

// ------------------------------------------------------------------- ipcio --
int ipcio_create(ipcio_t *ipc, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned n_readers)
{
    if (!ipc) return -1;
    ipc->curbuf = nullptr; ipc->curbufsz = 0; ipc->bytes = 0; ipc->rdwrt = 0;
    return ipcbuf_create(&ipc->buf, key, nbufs, bufsz, n_readers);
}
int ipcio_connect(ipcio_t *ipc, key_t key)
{
    if (!ipc) return -1;
    ipc->curbuf = nullptr; ipc->curbufsz = 0; ipc->bytes = 0; ipc->rdwrt = 0;
    return ipcbuf_connect(&ipc->buf, key);
}
int ipcio_disconnect(ipcio_t *ipc) { return ipc ? ipcbuf_disconnect(&ipc->buf) : -1; }
int ipcio_destroy(ipcio_t *ipc) { return ipc ? ipcbuf_destroy(&ipc->buf) : -1; }

int ipcio_open(ipcio_t *ipc, char rdwrt)
{
    if (!ipc || ipc->rdwrt) return -1;
    int rc;
    if (rdwrt == 'R' || rdwrt == 'r') rc = ipcbuf_lock_read(&ipc->buf);
    else if (rdwrt == 'W' || rdwrt == 'w') rc = ipcbuf_lock_write(&ipc->buf);
    else { fprintf(stderr, "ipcio_open: invalid rdwrt = '%c'\n", rdwrt); return -1; }
    if (rc < 0) return -1;
    ipc->rdwrt = rdwrt;
    ipc->curbuf = nullptr;
    ipc->bytes = 0;
    return 0;
}
int ipcio_is_open(ipcio_t *ipc) { return ipc && ipc->rdwrt != 0; }

int ipcio_stop(ipcio_t *ipc)
{
    if (!ipc || (ipc->rdwrt != 'W' && ipc->rdwrt != 'w')) return -1;
    if (ld(&ipc->buf.sync->eod)) return 0;              // a short block already ended the data
    if (ipc->curbuf) { fprintf(stderr, "ipcio_stop: a block is still open for writing\n"); return -1; }
    if (!ipcbuf_get_next_write(&ipc->buf)) return -1;
    ipcbuf_enable_eod(&ipc->buf);
    return ipcbuf_mark_filled(&ipc->buf, 0);
}
int ipcio_close(ipcio_t *ipc)
{
    if (!ipc || !ipc->rdwrt) return -1;
    int rc = 0;
    if (ipc->rdwrt == 'W' || ipc->rdwrt == 'w') {
        rc = ipcio_stop(ipc);
        ipcbuf_unlock_write(&ipc->buf);
    } else {
        ipcbuf_unlock_read(&ipc->buf);
    }
    ipc->rdwrt = 0;
    ipc->curbuf = nullptr;
    return rc;
}

char *ipcio_open_block_read(ipcio_t *ipc, uint64_t *curbufsz, uint64_t *block_id)
{
    if (!ipc || (ipc->rdwrt != 'R' && ipc->rdwrt != 'r')) { fprintf(stderr, "ipcio_open_block_read: not open for reading\n"); return nullptr; }
    if (ipc->curbuf) { fprintf(stderr, "ipcio_open_block_read: ipc->curbuf != 0 (a block is already open)\n"); return nullptr; }
    if (ipcbuf_eod(&ipc->buf)) { fprintf(stderr, "ipcio_open_block_read: ipcbuf_eod true, returning null ptr\n"); return nullptr; }
    ipc->curbuf = ipcbuf_get_next_read(&ipc->buf, &ipc->curbufsz);
    if (!ipc->curbuf) return nullptr;
    if (block_id) *block_id = ipcbuf_get_read_index(&ipc->buf);
    if (curbufsz) *curbufsz = ipc->curbufsz;
    ipc->bytes = 0;
    return ipc->curbuf;
}
ssize_t ipcio_close_block_read(ipcio_t *ipc, uint64_t bytes)
{
    if (!ipc || (ipc->rdwrt != 'R' && ipc->rdwrt != 'r')) return -1;
    if (!ipc->curbuf) { fprintf(stderr, "ipcio_close_block_read: ipc->curbuf == 0 (no block is open)\n"); return -1; }
    if (bytes != ipc->curbufsz) { fprintf(stderr, "ipcio_close_block_read: block not completely read (%llu of %llu bytes)\n",
                                          (unsigned long long)bytes, (unsigned long long)ipc->curbufsz); return -1; }
    ipc->curbuf = nullptr;
    ipc->bytes = 0;
    return ipcbuf_mark_cleared(&ipc->buf) == 0 ? 0 : -1;
}
char *ipcio_open_block_write(ipcio_t *ipc, uint64_t *block_id)
{
    if (!ipc || (ipc->rdwrt != 'W' && ipc->rdwrt != 'w')) { fprintf(stderr, "ipcio_open_block_write: not open for writing\n"); return nullptr; }
    if (ipc->curbuf) { fprintf(stderr, "ipcio_open_block_write: a block is already open\n"); return nullptr; }
    ipc->curbuf = ipcbuf_get_next_write(&ipc->buf);
    if (block_id) *block_id = ipcbuf_get_write_count(&ipc->buf);
    return ipc->curbuf;
}
ssize_t ipcio_close_block_write(ipcio_t *ipc, uint64_t bytes)
{
    if (!ipc || !ipc->curbuf || (ipc->rdwrt != 'W' && ipc->rdwrt != 'w')) return -1;
    ipc->curbuf = nullptr;
    return ipcbuf_mark_filled(&ipc->buf, bytes) == 0 ? 0 : -1;
}

// ---------------------------------------------------------------- multilog --
multilog_t *multilog_open(const char *program_name, char syslog)
{
    multilog_t *m = (multilog_t *)calloc(1, sizeof(multilog_t));
    if (!m) return nullptr;
    m->name = strdup(program_name ? program_name : "");
    m->syslog = syslog;
    return m;
}
int multilog_close(multilog_t *m)
{
    if (!m) return -1;
    free(m->name);
    free(m);
    return 0;
}
int multilog_add(multilog_t *m, FILE *f)
{
    if (!m || !f || m->n_logs >= MULTILOG_MAX_STREAMS) return -1;
    m->logs[m->n_logs++] = f;
    return 0;
}
int multilog(multilog_t *m, int priority, const char *format, ...)
{
    if (!m || !format) return -1;
    for (unsigned i = 0; i < m->n_logs; i++) {
        va_list ap;
        va_start(ap, format);
        if (priority <= LOG_WARNING) fprintf(m->logs[i], priority <= LOG_ERR ? "ERR: " : "WARN: ");
        fprintf(m->logs[i], "[%s] ", m->name);
        vfprintf(m->logs[i], format, ap);
        fflush(m->logs[i]);
        va_end(ap);
    }
    return 0;
}

// ---------------------------------------------------------------- dada_hdu --
dada_hdu_t *dada_hdu_create(multilog_t *log)
{
    dada_hdu_t *h = (dada_hdu_t *)calloc(1, sizeof(dada_hdu_t));
    if (!h) return nullptr;
    h->log = log;
    dada_hdu_set_key(h, DADA_DEFAULT_BLOCK_KEY);
    return h;
}
void dada_hdu_set_key(dada_hdu_t *h, key_t key)
{
    h->data_block_key = key;
    h->header_block_key = key + 1;
}
int dada_hdu_connect(dada_hdu_t *h)
{
    if (!h || h->data_block) return -1;
    ipcbuf_t *hb = (ipcbuf_t *)calloc(1, sizeof(ipcbuf_t));
    ipcio_t *db = (ipcio_t *)calloc(1, sizeof(ipcio_t));
    if (!hb || !db) { free(hb); free(db); return -1; }
    if (ipcbuf_connect(hb, h->header_block_key) < 0) {
        if (h->log) multilog(h->log, LOG_ERR, "Failed to connect to Header Block key=%x\n", (unsigned)h->header_block_key);
        free(hb); free(db);
        return -1;
    }
    if (ipcio_connect(db, h->data_block_key) < 0) {
        if (h->log) multilog(h->log, LOG_ERR, "Failed to connect to Data Block key=%x\n", (unsigned)h->data_block_key);
        ipcbuf_disconnect(hb);
        free(hb); free(db);
        return -1;
    }
    h->header_block = hb;
    h->data_block = db;
    h->header_size = ipcbuf_get_bufsz(hb);
    return 0;
}
int dada_hdu_disconnect(dada_hdu_t *h)
{
    if (!h || !h->data_block) return -1;
    ipcio_disconnect(h->data_block);
    ipcbuf_disconnect(h->header_block);
    free(h->data_block);
    free(h->header_block);
    h->data_block = nullptr;
    h->header_block = nullptr;
    return 0;
}
int dada_hdu_lock_read(dada_hdu_t *h)
{
    if (!h || !h->data_block) return -1;
    if (ipcbuf_lock_read(h->header_block) < 0) return -1;
    if (ipcio_open(h->data_block, 'R') < 0) { ipcbuf_unlock_read(h->header_block); return -1; }
    return 0;
}
int dada_hdu_unlock_read(dada_hdu_t *h)
{
    if (!h || !h->data_block) return -1;
    int rc = 0;
    if (ipcio_is_open(h->data_block) && ipcio_close(h->data_block) < 0) rc = -1;
    if (ipcbuf_unlock_read(h->header_block) < 0) rc = -1;
    return rc;
}
int dada_hdu_lock_write(dada_hdu_t *h)
{
    if (!h || !h->data_block) return -1;
    if (ipcbuf_lock_write(h->header_block) < 0) return -1;
    if (ipcio_open(h->data_block, 'W') < 0) { ipcbuf_unlock_write(h->header_block); return -1; }
    return 0;
}
int dada_hdu_unlock_write(dada_hdu_t *h)
{
    if (!h || !h->data_block) return -1;
    int rc = 0;
    if (ipcio_is_open(h->data_block) && ipcio_close(h->data_block) < 0) rc = -1;     // raises end of data
    if (ipcbuf_unlock_write(h->header_block) < 0) rc = -1;
    return rc;
}
void dada_hdu_destroy(dada_hdu_t *h)
{
    if (!h) return;
    if (h->data_block) dada_hdu_disconnect(h);
    free(h);
}

// ----------------------------------------------------------- dada_affinity --
int dada_bind_thread_to_core(int core)
{
    cpu_set_t set;
    CPU_ZERO(&set);
    CPU_SET(core, &set);
    return pthread_setaffinity_np(pthread_self(), sizeof set, &set) == 0 ? 0 : -1;
}

// ------------------------------------------------------------ ascii_header --
static const char *find_keyword(const char *header, const char *keyword)
{
    const size_t n = strlen(keyword);
    for (const char *p = header; p && *p;) {
        if (strncmp(p, keyword, n) == 0 && (p[n] == ' ' || p[n] == '\t')) return p;
        p = strchr(p, '\n');
        if (p) p++;
    }
    return nullptr;
}
int ascii_header_get(const char *header, const char *keyword, const char *code, ...)
{
    const char *p = header ? find_keyword(header, keyword) : nullptr;
    if (!p) return -1;
    p += strlen(keyword);
    va_list ap;
    va_start(ap, code);
    const int n = vsscanf(p, code, ap);
    va_end(ap);
    return n;
}
int ascii_header_set(char *header, const char *keyword, const char *code, ...)
{
    if (!header || !keyword || !code) return -1;
    char value[256], line[384];
    va_list ap;
    va_start(ap, code);
    vsnprintf(value, sizeof value, code, ap);
    va_end(ap);
    snprintf(line, sizeof line, "%-12s %s\n", keyword, value);
    char *old = (char *)find_keyword(header, keyword);
    std::string h(header);
    if (old) {
        const size_t at = (size_t)(old - header);
        size_t end = h.find('\n', at);
        end = end == std::string::npos ? h.size() : end + 1;
        h.replace(at, end - at, line);
    } else {
        h += line;
    }
    if (h.size() + 1 > DADA_DEFAULT_HEADER_SIZE) return -1;
    memcpy(header, h.c_str(), h.size() + 1);
    return 0;
}

}  // extern "C"
