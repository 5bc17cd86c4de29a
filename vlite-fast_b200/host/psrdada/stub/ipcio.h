/* COMPILE-ONLY STUB of psrdada's ipcio.h (see multilog.h): call sites src/process_baseband.cu:324,837,1038 */
#ifndef VF_STUB_IPCIO_H
#define VF_STUB_IPCIO_H
#include "ipcbuf.h"
typedef struct ipcio ipcio_t;      /* begins with an ipcbuf_t: psrdada code casts ipcio_t* to ipcbuf_t* (:214,308) */
ssize_t ipcio_read (ipcio_t *ipc, char *ptr, size_t bytes);
ssize_t ipcio_write (ipcio_t *ipc, char *ptr, size_t bytes);
char *ipcio_open_block_read (ipcio_t *ipc, uint64_t *curbufsz, uint64_t *block_id);
ssize_t ipcio_close_block_read (ipcio_t *ipc, uint64_t bytes);
char *ipcio_open_block_write (ipcio_t *ipc, uint64_t *block_id);
ssize_t ipcio_close_block_write (ipcio_t *ipc, uint64_t bytes);
#endif
