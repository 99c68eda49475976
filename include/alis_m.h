/* alis_m.h -- compatibility name for the material table, see alis_materials.h */
#include "alis_materials.h"
