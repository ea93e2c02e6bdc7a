/* matrix.h - the C Matrix API lives in mex.h of this stand-in (see there). */
#include "mex.h"
