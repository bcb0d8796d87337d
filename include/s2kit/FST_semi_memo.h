/* s2kit/FST_semi_memo.h -- forwarding header: callers of the reference include "s2kit/FST_semi_memo.h" (reference
 * include/s2kit/FST_semi_memo.h); every prototype of the drop-in library lives in ../s2kit.h. */
#include "../s2kit.h"
