/* TEST INFRASTRUCTURE ONLY: the BP QFunction headers of libceed_b200/qfunctions compiled as ordinary CPU user
 * QFunctions, so that the UNMODIFIED reference (oracle/_ref/lib/libceed.so, /cpu/self/...) can run exactly the
 * operators the b200 backend runs.  Built by oracle/Makefile against the reference's installed headers. */
#include <ceed/types.h>
#include <string.h>
#include "../libceed_b200/qfunctions/bp_apply.h"
#include "../libceed_b200/qfunctions/bp_geo.h"

typedef int (*BPQFunctionUser)(void *, const CeedInt, const CeedScalar *const *, CeedScalar *const *);

BPQFunctionUser bpqf_get(const char *name) {
  if (!strcmp(name, "BPSetupMassGeo")) return BPSetupMassGeo;
  if (!strcmp(name, "BPSetupDiffGeo")) return BPSetupDiffGeo;
  if (!strcmp(name, "BPMass")) return BPMass;
  if (!strcmp(name, "BPMass3")) return BPMass3;
  if (!strcmp(name, "BPDiff")) return BPDiff;
  if (!strcmp(name, "BPDiff3")) return BPDiff3;
  return 0;
}
