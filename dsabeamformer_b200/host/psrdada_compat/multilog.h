/* multilog.h -- PSRDADA-compatible logger (compatibility header, see ipcbuf.h).
 * Used by the reference as multilog_open / multilog_add / multilog (dada_handler.hh:26-28,70,77). */
#ifndef DSABF_COMPAT_MULTILOG_H
#define DSABF_COMPAT_MULTILOG_H
#include <stdio.h>
#include <syslog.h>       /* LOG_INFO, LOG_ERR, LOG_WARNING: the priorities callers pass */

#ifdef __cplusplus
extern "C" {
#endif

#define MULTILOG_MAX_STREAMS 8
typedef struct {
    char *name;
    int syslog;
    FILE *logs[MULTILOG_MAX_STREAMS];
    unsigned n_logs;
} multilog_t;

multilog_t *multilog_open(const char *program_name, char syslog);
int multilog_close(multilog_t *);
int multilog_add(multilog_t *, FILE *);
int multilog(multilog_t *, int priority, const char *format, ...);

#ifdef __cplusplus
}
#endif
#endif
