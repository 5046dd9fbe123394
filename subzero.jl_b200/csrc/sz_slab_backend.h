// sz_slab_backend.h — what the slab host logic (sz_slab.cpp) needs from a rank's handle beyond the public C ABI.
//
// sz_slab.cpp is compiled twice: into the CUDA library, where these calls are implemented in sz_api.cu on top of the
// peer-memory push / unpack kernels (sz_kernels_fp.cu), and — with -DSZ_ORACLE_BUILD — into the oracle's test library,
// where sz_slab.cpp itself implements them through szo_halo_pack / szo_halo_unpack and host memory.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/subzero_b200.h"

// Where a partner writes its records for this rank: process id, device, base pointer and cudaIpc handle of the
// receive arena, offsets of the two staging halves and of the two flags, expected message size.  Opaque to sz_slab.cpp.
struct SlabWire {
    unsigned char b[192];
};

// exchange lists of one rank against its partners (local 0-based floe indices, CSR over the partners)
struct SlabLists {
    int32_t rank, n_partners;
    const int32_t *partner_rank;  // [n_partners] ascending
    const int64_t *send_off, *send_idx;
    const int64_t *recv_off, *recv_idx;
    const uint8_t *owned;  // [n] of the local list
    double period_x, period_y;
};

#ifndef SZ_ORACLE_BUILD
// first thing of a (re)build: wait for the rank's stream; final != 0 (slab destroyed): also unmap the partners' arenas
int32_t szb_release_peers(sz_handle *h, int32_t final);
// after the local list was uploaded: register the lists, allocate this rank's receive arena; wire_out[p] tells partner
// p where to write
int32_t szb_configure(sz_handle *h, const SlabLists *lists, SlabWire *wire_out);
// peer_wire[p] = what partner p exported for this rank; ends with the first publication (epoch 1)
int32_t szb_connect(sz_handle *h, const SlabWire *peer_wire);
// One timestep in three phases, so that one host thread can drive several ranks: on ALL ranks the uploads and (host
// arrays) the publication of the uploaded boundary floes, then on all ranks the kernels — the first of which waits for
// the neighbours' records —, then wait on all ranks.
int32_t szb_step_publish(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out, int32_t host_mode);
int32_t szb_step_begin(sz_handle *h);
int32_t szb_step_end(sz_handle *h, int32_t host_mode);
double szb_max_displacement(sz_handle *h);
// halo copies := owners' current state, without stepping: publish (if the current epoch is not out yet) on every rank,
// then consume on every rank
int32_t szb_refresh_publish(sz_handle *h);
int32_t szb_refresh_consume(sz_handle *h);
// sz_upload_floes with Monte-Carlo points that are already resident (see upload_floes_impl in sz_api.cu)
int32_t szb_upload_floes_resident_mc(sz_handle *h, const sz_floe_soa *s, const int64_t *mc_src, int64_t n_extra);
// Monte-Carlo points [off, off + n) of the resident array -> host (migrants of a rebuild)
int32_t szb_fetch_mc(sz_handle *h, int64_t off, int64_t n, double *x, double *y);
int32_t szb_mc_offsets(sz_handle *h, int64_t *off /* [n_init + 1] */);
// page-locked host staging of a rebuild
int32_t szb_host_alloc(size_t bytes, void **out);
void szb_host_free(void *p);
#endif
