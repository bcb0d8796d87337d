/* s2kit/chebyshev_nodes.h -- forwarding header: callers of the reference include "s2kit/chebyshev_nodes.h" (reference
 * include/s2kit/chebyshev_nodes.h); every prototype of the drop-in library lives in ../s2kit.h. */
#include "../s2kit.h"
