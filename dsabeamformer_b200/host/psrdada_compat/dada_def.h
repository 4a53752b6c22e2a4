/* dada_def.h -- PSRDADA-compatible defaults (compatibility header, see ipcbuf.h). */
#ifndef DSABF_COMPAT_DADA_DEF_H
#define DSABF_COMPAT_DADA_DEF_H
#define DADA_DEFAULT_BLOCK_KEY 0x0000dada
#define DADA_DEFAULT_BLOCK_SIZE 524288
#define DADA_DEFAULT_BLOCK_NUM 4
#define DADA_DEFAULT_HEADER_SIZE 4096
#define DADA_TIMESTR "%Y-%m-%d-%H:%M:%S"
#endif
