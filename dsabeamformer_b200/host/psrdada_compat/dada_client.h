/* dada_client.h -- compatibility header (see ipcbuf.h).  The reference includes it (beamformer.hh:27) but
 * uses nothing of PSRDADA's client main-loop framework, and neither does this repo. */
#ifndef DSABF_COMPAT_DADA_CLIENT_H
#define DSABF_COMPAT_DADA_CLIENT_H
#include "dada_hdu.h"
#endif
