/* COMPILE-ONLY STUB of psrdada's ipcbuf.h (see multilog.h): call sites src/process_baseband.cu:172,199,209-219,309-310,807,832 */
#ifndef VF_STUB_IPCBUF_H
#define VF_STUB_IPCBUF_H
#include <stdint.h>
#include <sys/types.h>
typedef struct ipcbuf ipcbuf_t;
char *ipcbuf_get_next_read (ipcbuf_t *id, uint64_t *bytes);
int ipcbuf_mark_cleared (ipcbuf_t *id);
char *ipcbuf_get_next_write (ipcbuf_t *id);
int ipcbuf_mark_filled (ipcbuf_t *id, uint64_t nbytes);
uint64_t ipcbuf_get_nbufs (ipcbuf_t *id);
uint64_t ipcbuf_get_bufsz (ipcbuf_t *id);
uint64_t ipcbuf_get_nfull (ipcbuf_t *id);
int ipcbuf_eod (ipcbuf_t *id);
#endif
