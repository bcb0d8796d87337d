/* s2kit/FST_semi_fly.h -- forwarding header: callers of the reference include "s2kit/FST_semi_fly.h" (reference
 * include/s2kit/FST_semi_fly.h); every prototype of the drop-in library lives in ../s2kit.h. */
#include "../s2kit.h"
