/* alis_c.h -- compatibility name: the reference splits its API header (alis_c.h) from the
 * material table (alis_m.h); here the API lives in alis_api.h. */
#include "alis_api.h"
