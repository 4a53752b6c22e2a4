/* ipcio.h -- PSRDADA-compatible block I/O on top of ipcbuf (compatibility header, see ipcbuf.h).
 * The reference uses ipcio_open_block_read / ipcio_close_block_read (dada_handler.hh:92-98) and casts
 * ipcio_t* to ipcbuf_t* (:83,129), so `buf` must be the first member, as in PSRDADA. */
#ifndef DSABF_COMPAT_IPCIO_H
#define DSABF_COMPAT_IPCIO_H
#include "ipcbuf.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    ipcbuf_t buf;          /* must stay first */
    char *curbuf;          /* the open block, NULL if none */
    uint64_t curbufsz;
    uint64_t bytes;
    char rdwrt;            /* 0 closed, 'R'/'r' reading, 'W'/'w' writing */
} ipcio_t;

#define IPCIO_INIT { IPCBUF_INIT, 0, 0, 0, 0 }

int ipcio_create(ipcio_t *, key_t key, uint64_t nbufs, uint64_t bufsz, unsigned n_readers);
int ipcio_connect(ipcio_t *, key_t key);
int ipcio_disconnect(ipcio_t *);
int ipcio_destroy(ipcio_t *);

int ipcio_open(ipcio_t *, char rdwrt);
int ipcio_is_open(ipcio_t *);
int ipcio_stop(ipcio_t *);            /* writer: mark the end of data (an empty last block) */
int ipcio_close(ipcio_t *);           /* writer: ipcio_stop if still open for writing; then closed */

/* Only one block may be open at a time (PSRDADA returns NULL and complains otherwise). */
char *ipcio_open_block_read(ipcio_t *, uint64_t *bytes, uint64_t *block_id);
ssize_t ipcio_close_block_read(ipcio_t *, uint64_t bytes);
char *ipcio_open_block_write(ipcio_t *, uint64_t *block_id);
ssize_t ipcio_close_block_write(ipcio_t *, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
