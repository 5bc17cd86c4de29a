/* COMPILE-ONLY STUB of psrdada's multilog.h: the prototypes process_baseband uses (src/process_baseband.cu:505-530),
 * so that vf_ring_psrdada.c can be type-checked where psrdada is not installed.  Never linked. */
#ifndef VF_STUB_MULTILOG_H
#define VF_STUB_MULTILOG_H
#include <stdio.h>
#include <syslog.h>
typedef struct multilog multilog_t;
multilog_t *multilog_open (const char *program_name, char syslog);
int multilog_close (multilog_t *m);
int multilog_add (multilog_t *m, FILE *fptr);
int multilog (multilog_t *m, int priority, const char *format, ...);
#endif
