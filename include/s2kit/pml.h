/* s2kit/pml.h -- forwarding header: callers of the reference include "s2kit/pml.h" (reference
 * include/s2kit/pml.h); every prototype of the drop-in library lives in ../s2kit.h. */
#include "../s2kit.h"
