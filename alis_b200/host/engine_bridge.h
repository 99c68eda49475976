/*
 * engine_bridge.h -- host-side glue between the ALiS public structs and the engine C ABI.
 * See engine_bridge.c.  Internal header (not installed).
 */
#ifndef ALIS_ENGINE_BRIDGE_H
#define ALIS_ENGINE_BRIDGE_H

#include "alis_b200.h"
#include "shard.h"

typedef struct alis_bridge alis_bridge;

alis_bridge  *alis_bridge_create(world W);           /* allocates the engine (fields start at 0) */
void          alis_bridge_destroy(alis_bridge *B);
void          alis_bridge_commit(alis_bridge *B);    /* upload coefficients/sources/copies       */
void          alis_bridge_invalidate(alis_bridge *B);/* setup changed: re-commit before next step */
void          alis_bridge_mark(alis_bridge *B, int dirty_mask);   /* 1 tables, 2 materials, 4 sources, 8 copies */
void          alis_bridge_step_E(alis_bridge *B);    /* updateE  (update.c:275-299)              */
void          alis_bridge_step_H(alis_bridge *B);    /* updateH  (update.c:302-341)              */
void          alis_bridge_chi2(alis_bridge *S, alis_bridge *P1, alis_bridge *P2, float alt);   /* nonLinearSHG/TWM (update.c:394-411) */
alisb_engine *alis_bridge_engine(alis_bridge *B);     /* the first engine (the only one unless ALIS_DEVICES names several) */
alis_shards  *alis_bridge_shards(alis_bridge *B);     /* all of them: what the world is stepped on (shard.h) */
int           alis_bridge_engines(alis_bridge *B);    /* > 1: the world is cut into z slabs, one engine per device (shard.c) */
int           alis_bridge_read_box(alis_bridge *B, int arr, const int32_t ids[6], const int32_t box[6], float *host);   /* alisb_read_box over all slabs */
int           alis_bridge_reset_fields(alis_bridge *B);
void          alis_bridge_read_field(alis_bridge *B, int set, int comp, int m, float *host);
void          alis_bridge_write_field(alis_bridge *B, int set, int comp, int m, const float *host);

/* provided by the including side (product host layer or test harness) */
int alis_source_target(world W, source S, int set, int *comp, int *m);   /* 0 on success */
int alis_bridge_device(void);                                            /* CUDA ordinal  */

#endif
