/* TEST INFRASTRUCTURE ONLY -- see oracle/README.md.
 *
 * Builds the REFERENCE's own QFunctions, from the sources where they lie under
 * /root/reference/qfunctions/ (never copied into this repo), into
 * oracle/_ref/libref_qf.so.  Compile with  -I/root/reference  (oracle/Makefile).
 *
 * Each exported ref_<Name> forwards to the reference's static QFunction <Name>
 * with the libCEED user-QFunction signature
 *   int f(void *ctx, CeedInt Q, const CeedScalar *const *in, CeedScalar *const *out)
 * and ref_<Name>_loc() returns the "<file>:<name>" locator the reference passes to
 * CeedQFunctionCreateInterior (src/setuplibceed.c:370,518,818).
 */
#include "ref_shim.h"

#include "qfunctions/common.h"
#include "qfunctions/linElas.h"
#include "qfunctions/hyperSS.h"
#include "qfunctions/hyperFS.h"
#include "qfunctions/constantForce.h"
#include "qfunctions/manufacturedForce.h"
#include "qfunctions/manufacturedTrue.h"

#define EXPORT_QF(name)                                                          \
  int ref_##name(void *ctx, CeedInt Q, const CeedScalar *const *in,             \
                 CeedScalar *const *out) {                                       \
    return name(ctx, Q, in, out);                                                \
  }                                                                              \
  const char *ref_##name##_loc(void) { return name##_loc; }

EXPORT_QF(SetupGeo)
EXPORT_QF(LinElasF)
EXPORT_QF(LinElasdF)
EXPORT_QF(LinElasEnergy)
EXPORT_QF(LinElasDiagnostic)
EXPORT_QF(HyperSSF)
EXPORT_QF(HyperSSdF)
EXPORT_QF(HyperSSEnergy)
EXPORT_QF(HyperSSDiagnostic)
EXPORT_QF(HyperFSF)
EXPORT_QF(HyperFSdF)
EXPORT_QF(HyperFSEnergy)
EXPORT_QF(HyperFSDiagnostic)
EXPORT_QF(SetupConstantForce)
EXPORT_QF(SetupMMSForce)
EXPORT_QF(MMSTrueSoln)
