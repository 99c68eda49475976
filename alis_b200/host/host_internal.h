/*
 * host_internal.h -- glue shared by the host C layer (not installed).
 *
 * Fields live on the device.  The host structs keep
 *   W->b200_bridge : the engine + descriptor state (engine_bridge.c), created on first use
 *   W->b200_view   : bookkeeping of the lazily allocated HOST VIEWS W->E[m].x/y/z ... that the
 *                    monitors (fields.c, output.c, slices.c) read, refreshed per region on demand
 */
#ifndef ALIS_HOST_INTERNAL_H
#define ALIS_HOST_INTERNAL_H

#include <stdio.h>
#include <stdlib.h>
#include "alis_b200.h"
#include "engine_bridge.h"

/* allocation failure ends the program like the reference's MALLOC (alis_c.h:34) */
static inline void *alis_checked(void *p, size_t bytes)
{
	if (!p && bytes) { printf("Insufficient memory!\n"); exit(1); }
	return p;
}
#define xmalloc(bytes)        alis_checked(malloc(bytes), (bytes))
#define xcalloc(n, size)      alis_checked(calloc((n), (size)), (size_t)(n)*(size))
#define xrealloc(p, bytes)    alis_checked(realloc((p), (bytes)), (bytes))

enum { ALIS_DIRTY_TABLES = 1, ALIS_DIRTY_MATERIALS = 2, ALIS_DIRTY_SOURCES = 4, ALIS_DIRTY_COPIES = 8, ALIS_DIRTY_ALL = 15 };

void alis_host_attach(world W);                 /* called by createWorld / cloneWorld            */
void alis_host_detach(world W);                 /* frees bridge + views                          */
alis_bridge *alis_host_bridge(world W);         /* creates the engine on first use               */
void alis_host_invalidate(world W);             /* setup changed -> re-commit before next step   */
void alis_host_dirty(world W, int mask);
void alis_host_reset_fields(world W);           /* rebootWorld                                   */
void alis_host_touch(world W);                  /* device state advanced: host views are stale   */
unsigned long alis_view_version(world W);       /* number of such advances                       */
void alis_flux_cache_drop(world W);             /* output.c: forget the flux results kept for W  */

/* host views: make W->E[m].{x,y,z}, W->H[m].{x,y,z} valid on [i0-1..i1+1]x[j0-1..j1+1]x[k0-1..k1+1]
 * (clipped to the allocation), for every field set m, including TFSF aux worlds */
void alis_view_need(world W, int i0, int i1, int j0, int j1, int k0, int k1);
void alis_view_need_arr(world W, int set, int comp, int m, int i0, int i1, int j0, int j1, int k0, int k1);   /* one array only */
void alis_view_need_all(world W);
unsigned alis_field_arrays(world W, field F);   /* bit (m*2+set)*3+comp for every host view F reads */
void alis_view_need_for(world W, field F, int i0, int i1, int j0, int j1, int k0, int k1);   /* only the arrays F reads */
void alis_view_need_ade(world W);               /* W->E[m].Jx[n][p] ... host views                */
void alis_view_need_ade_box(world W, int i0, int i1, int j0, int j1, int k0, int k1);   /* ... over an index box only */
void alis_view_free(world W);
/* device twins of the field functions (fields.c) */
int    alis_host_monitors(void);                       /* env ALIS_HOST_MONITORS=1: host-view path only */
int    alis_field_device_fn(world W, field F, alisb_field_fn *d, int *take_log);   /* 1 when F has one   */
int    alis_device_region_ok(world W, const alisb_field_fn *d, const int box[6]);   /* inclusive INDEX box */
float *alis_eval_region(world W, field F, int i0, int i1, int j0, int j1, int k0, int k1);   /* dense block, k fastest; caller frees */
int    alis_device_absmax(world W, field F, const int box[6], float *out);          /* 0: use the host path */
int    alis_device_column_sums(world W, field F, const int box[6], float *cols);    /* 0: use the host path */
alis_phasers *alis_phaser_device(phaser P, int *freq);   /* device accumulators (of every z slab) + frequency index of P */


/* ALIS_PROFILE=1: inclusive wall time and call count of the public entry points, printed at exit (stderr).
 * Off (one predictable branch per call) otherwise. */
int    alis_prof_enabled(void);
double alis_prof_now(void);
void   alis_prof_add(const char *name, double seconds);
typedef struct { const char *name; double t0; } alis_prof_scope;
static inline void alis_prof_end(alis_prof_scope *s) { if (s->name) alis_prof_add(s->name, alis_prof_now() - s->t0); }
#define ALIS_PROF(label) alis_prof_scope alis_prof_scope_ __attribute__((cleanup(alis_prof_end))) = { alis_prof_enabled() ? (label) : 0, alis_prof_enabled() ? alis_prof_now() : 0 }

#endif
