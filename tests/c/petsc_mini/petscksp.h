/* TEST-ONLY: see petsc.h in this directory */
#include "petsc.h"
