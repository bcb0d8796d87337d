/* s2kit/util.h -- forwarding header: callers of the reference include "s2kit/util.h" (reference
 * include/s2kit/util.h); every prototype of the drop-in library lives in ../s2kit.h. */
#include "../s2kit.h"
