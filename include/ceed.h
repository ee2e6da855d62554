/* The reference includes <ceed.h> (/root/reference/elasticity.h); forward to the API header. */
#ifndef CEED_B200_CEED_TOP_H
#define CEED_B200_CEED_TOP_H
#include "ceed/ceed.h"
#endif
