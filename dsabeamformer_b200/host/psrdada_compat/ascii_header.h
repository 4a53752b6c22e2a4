/* ascii_header.h -- PSRDADA-compatible "KEYWORD value" header access (compatibility header, see ipcbuf.h).
 * The reference only includes it (beamformer.hh:35); beam's output sink uses ascii_header_set. */
#ifndef DSABF_COMPAT_ASCII_HEADER_H
#define DSABF_COMPAT_ASCII_HEADER_H
#ifdef __cplusplus
extern "C" {
#endif
/* code is a printf/scanf conversion such as "%d" or "%s"; header is a NUL-terminated buffer of
 * DADA_DEFAULT_HEADER_SIZE bytes.  Return: set 0 / -1; get: number of items parsed or -1. */
int ascii_header_set(char *header, const char *keyword, const char *code, ...);
int ascii_header_get(const char *header, const char *keyword, const char *code, ...);
#ifdef __cplusplus
}
#endif
#endif
