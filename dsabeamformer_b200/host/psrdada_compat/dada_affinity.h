/* dada_affinity.h -- PSRDADA-compatible thread binding (compatibility header, see ipcbuf.h).
 * Reference call site: dada_handler.hh:50-55. */
#ifndef DSABF_COMPAT_DADA_AFFINITY_H
#define DSABF_COMPAT_DADA_AFFINITY_H
#ifdef __cplusplus
extern "C" {
#endif
int dada_bind_thread_to_core(int core);
#ifdef __cplusplus
}
#endif
#endif
