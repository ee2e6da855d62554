/* TEST INFRASTRUCTURE ONLY -- see oracle/README.md.
 *
 * Minimal stand-in for the handful of <ceed.h> macros/typedefs that the reference's
 * single-source QFunction headers (/root/reference/qfunctions/*.h) need in order to
 * compile with a host C compiler.  libCEED itself is not vendored in the reference
 * (SURVEY.md section 8(c)); upstream defines these in ceed.h:
 *   CeedInt = int32, CeedScalar = double, CEED_Q_VLA = Q (host) and
 *   CEED_QFUNCTION(name) = a "<file>:<name>" locator string plus a static function.
 */
#ifndef ORACLE_REF_SHIM_H
#define ORACLE_REF_SHIM_H

#include <math.h>

typedef int CeedInt;
typedef double CeedScalar;

#define CEED_Q_VLA Q
#define CeedPragmaSIMD _Pragma("omp simd")
#define CEED_QFUNCTION(name) \
  static const char name##_loc[] = __FILE__ ":" #name; \
  static int name

#endif
