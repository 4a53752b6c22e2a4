/* dada_hdu.h -- PSRDADA-compatible "header + data unit": a header ring (key + 1) and a data ring (key)
 * (compatibility header, see ipcbuf.h).  Call sites in the reference: dada_handler.hh:30-43,118-123. */
#ifndef DSABF_COMPAT_DADA_HDU_H
#define DSABF_COMPAT_DADA_HDU_H
#include "ipcio.h"
#include "multilog.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dada_hdu {
    multilog_t *log;
    ipcio_t *data_block;
    ipcbuf_t *header_block;
    char *header;
    uint64_t header_size;
    key_t data_block_key;
    key_t header_block_key;
} dada_hdu_t;

dada_hdu_t *dada_hdu_create(multilog_t *log);
void dada_hdu_set_key(dada_hdu_t *, key_t key);
int dada_hdu_connect(dada_hdu_t *);
int dada_hdu_disconnect(dada_hdu_t *);
int dada_hdu_lock_read(dada_hdu_t *);
int dada_hdu_unlock_read(dada_hdu_t *);
int dada_hdu_lock_write(dada_hdu_t *);
int dada_hdu_unlock_write(dada_hdu_t *);
void dada_hdu_destroy(dada_hdu_t *);

#ifdef __cplusplus
}
#endif
#endif
