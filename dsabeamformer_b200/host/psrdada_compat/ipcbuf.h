/* ipcbuf.h -- PSRDADA-compatible ring of fixed-size buffers in shared memory.
 *
 * COMPATIBILITY HEADER.  PSRDADA (the un-vendored, un-pinned dependency of the reference:
 * /root/reference/makefile:5-6,13) is not installed here.  This directory declares the part of its public C
 * API that the reference's reader uses (/root/reference/src/dada_handler.hh:25-155) with PSRDADA's names,
 * argument order and struct members, and libpsrdada_shim implements it over POSIX shared memory, so that
 *   - host/dada_handler.hh and host/ring_tools.cc are written against the PSRDADA API and build unchanged
 *     against the real library (make PSRDADA_INC=<psrdada>/src PSRDADA_LIBS="-L<psrdada>/src/.libs -lpsrdada");
 *   - the reference's own observation build compiles and links against this shim (the beam_ref_obs checker).
 * The shim is a single-writer / single-reader TEST DOUBLE: no SysV semaphores, no multiple readers, no
 * start-of-data arithmetic.  Struct members that the reference touches directly are kept: ipcbuf_t::sync
 * (->nbufs, ->bufsz) and ipcbuf_t::buffer[] (dada_handler.hh:139-147). */
#ifndef DSABF_COMPAT_IPCBUF_H
#define DSABF_COMPAT_IPCBUF_H
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPCBUF_MAX_BUFS 512

/* lives in the shared segment (PSRDADA: ipcsync_t) */
typedef struct {
    uint64_t magic;
    uint64_t nbufs;               /* buffers in the ring */
    uint64_t bufsz;               /* bytes per buffer */
    volatile uint64_t w_count;    /* buffers marked filled by the writer */
    volatile uint64_t r_count;    /* buffers marked cleared by the reader */
    volatile uint64_t eod;        /* 1 + index (w_count value) of the buffer that ends the data, 0 = none yet */
    volatile uint32_t n_writers, n_readers;
    uint64_t nbytes[IPCBUF_MAX_BUFS];   /* payload of each buffer */
} ipcsync_t;

typedef struct {
    int state;                    /* 0 disconnected, 1 connected, 2 reader, 3 writer */
    ipcsync_t *sync;              /* control block in shared memory */
    char **buffer;                /* base address of every buffer in this process */
    void *map_base;               /* shim: the whole mapping */
    uint64_t map_bytes;
    key_t key;
    int created;                  /* shim: this process created the segment */
    int locked;                   /* mlock()ed by ipcbuf_lock */
    int eod_pending;              /* writer: the next buffer marked filled ends the data */
    uint64_t last_read_index;     /* reader: index of the buffer last handed out */
} ipcbuf_t;

#define IPCBUF_INIT { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }

int ipcbuf_create(ipcbuf_t *, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned n_readers);
int ipcbuf_connect(ipcbuf_t *, key_t key);
int ipcbuf_disconnect(ipcbuf_t *);
int ipcbuf_destroy(ipcbuf_t *);

int ipcbuf_lock_write(ipcbuf_t *);
int ipcbuf_unlock_write(ipcbuf_t *);
int ipcbuf_lock_read(ipcbuf_t *);
int ipcbuf_unlock_read(ipcbuf_t *);

char *ipcbuf_get_next_write(ipcbuf_t *);                 /* blocks until a buffer is clear */
int ipcbuf_mark_filled(ipcbuf_t *, uint64_t nbytes);
int ipcbuf_enable_eod(ipcbuf_t *);                       /* the next buffer marked filled ends the data */

char *ipcbuf_get_next_read(ipcbuf_t *, uint64_t *bytes); /* blocks until a buffer is full; NULL after end of data */
int ipcbuf_mark_cleared(ipcbuf_t *);
int ipcbuf_eod(ipcbuf_t *);                              /* the reader has consumed the last buffer */

uint64_t ipcbuf_get_nbufs(ipcbuf_t *);
uint64_t ipcbuf_get_bufsz(ipcbuf_t *);
uint64_t ipcbuf_get_write_count(ipcbuf_t *);
uint64_t ipcbuf_get_read_count(ipcbuf_t *);
uint64_t ipcbuf_get_read_index(ipcbuf_t *);              /* index of the buffer last returned by get_next_read */
uint64_t ipcbuf_get_nfull(ipcbuf_t *);

int ipcbuf_lock(ipcbuf_t *);                             /* lock the buffers into RAM (dada_db -l) */
int ipcbuf_unlock(ipcbuf_t *);
int ipcbuf_get_device(ipcbuf_t *);                       /* -1: host memory (the shim never places buffers on a GPU) */

#ifdef __cplusplus
}
#endif
#endif
