/*
 * shard.h -- z-slab decomposition of ONE world over several engine instances behind the public C API
 * (reference dependency update.c:110-111).  Internal header (not installed).
 *
 * An alis_shards is what the bridge talks to instead of a single alisb_engine:
 *   ALIS_SHARDS_SPLIT    n engines, one per device, each owning a contiguous range of z planes of the
 *                        world; every descriptor of the engine ABI (tables, material blocks, ADE boxes,
 *                        sources, TFSF strips, boundary copies) is clipped to the slab, the one plane
 *                        of tangential components per interface and half step travels with
 *                        alisb_halo_exchange
 *   ALIS_SHARDS_REPLICA  n identical engines, one per device: the auxiliary (incident-field) worlds
 *                        of a TFSF source, which every slab of the main world steps for itself
 * With n = 1 every call is forwarded unchanged.  The layer only uses the engine C ABI
 * (include/alis_b200.h), so it runs on the CUDA engine and on the CPU oracle alike.
 */
#ifndef ALIS_SHARD_H
#define ALIS_SHARD_H

#include "alis_b200.h"

enum { ALIS_SHARDS_SPLIT = 0, ALIS_SHARDS_REPLICA = 1 };

typedef struct alis_shards alis_shards;

/* devices: CUDA ordinals, one engine each (n >= 1).  A world too thin for n slabs gets fewer. */
int  alis_shards_create(const alisb_grid *g, const int *devices, int n, int mode, alis_shards **out);
void alis_shards_destroy(alis_shards *S);
int  alis_shards_count(const alis_shards *S);
int  alis_shards_mode(const alis_shards *S);
alisb_engine *alis_shards_engine(alis_shards *S, int s);
void alis_shards_range(const alis_shards *S, int s, int *c0, int *c1);    /* global planes owned by slab s */

/* same meaning and arguments as the alisb_* calls of the same name (global descriptors in, clipped ones out) */
int  alis_shards_set_axis_tables(alis_shards *S, int set, int axis, const float *dtd, const float *P, const float *P1, const float *P2);
int  alis_shards_set_spacing(alis_shards *S, int set, int axis, const float *ds);
int  alis_shards_set_inv_eps(alis_shards *S, const float *cx, const float *cy, const float *cz);
int  alis_shards_clear_dispersive(alis_shards *S);
int  alis_shards_add_dispersive(alis_shards *S, const alisb_dispersive *d);
int  alis_shards_clear_sources(alis_shards *S);
int  alis_shards_add_source(alis_shards *S, const alisb_source *d);
int  alis_shards_add_tfsf(alis_shards *S, const alisb_tfsf *t, alis_shards *aux);   /* t->aux is ignored: slab s uses aux's engine s */
int  alis_shards_set_copies(alis_shards *S, int set, const alisb_copy *ops, int n);
int  alis_shards_set_time(alis_shards *S, uint64_t n, float t, float tE, float tH);
int  alis_shards_get_time(alis_shards *S, uint64_t *n, float *t, float *tE, float *tH);
int  alis_shards_step_E(alis_shards *S);
int  alis_shards_step_H(alis_shards *S);
int  alis_shards_reset_fields(alis_shards *S);
int  alis_shards_read_box(alis_shards *S, int arr, const int32_t ids[6], const int32_t box[6], float *host);
int  alis_shards_write_box(alis_shards *S, int arr, const int32_t ids[6], const int32_t box[6], const float *host);

/* running DFT (alisb_phaser_*) of a monitor box that may span several slabs */
typedef struct alis_phasers alis_phasers;
int  alis_phasers_create(alis_shards *S, const int32_t box[6], int face_mask, int nfreq, alis_phasers **out);
void alis_phasers_destroy(alis_phasers *P);
int  alis_phasers_accumulate(alis_phasers *P, const float *expE, const float *expH);
int  alis_phasers_read(alis_phasers *P, int freq, int face, int which, float *host_complex);
int  alis_phasers_surface_sums(alis_phasers *P, int freq, const alisb_surface_job *jobs, int njobs, const float *weights, size_t nweights, double *out);

/* devices named by the environment: ALIS_DEVICES="0,1,2" (several slabs) or ALIS_DEVICE=k (one engine, default 0);
 * returns the count (<= cap) */
int  alis_shards_env_devices(int *devices, int cap);

#endif
