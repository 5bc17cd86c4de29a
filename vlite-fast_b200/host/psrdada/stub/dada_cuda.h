/* COMPILE-ONLY STUB of psrdada's dada_cuda.h: page-locking of a ring's data blocks for DMA */
#ifndef VF_STUB_DADA_CUDA_H
#define VF_STUB_DADA_CUDA_H
#include "dada_hdu.h"
int dada_cuda_dbregister (dada_hdu_t *hdu);
int dada_cuda_dbunregister (dada_hdu_t *hdu);
#endif
