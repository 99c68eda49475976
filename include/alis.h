/* alis.h -- umbrella header of alis-b200 (drop-in for the reference's <alis.h>):
 * the public API plus the material / colour-map globals.  Include in ONE translation unit
 * (the globals are definitions); other units include alis_c.h. */
#ifndef ALIS_H
#define ALIS_H
#include "alis_api.h"
#include "alis_materials.h"
#endif
